"""ctypes binding of ``libvajax_b200.so`` (include/vajax_b200.h).

The reference binds its native solvers with nanobind capsules registered as XLA FFI targets
(reference: vajax/sparse/umfpack_jax.py:40-107, vajax/sparse/umfpack_ffi.cpp:531-551).  JAX is
not installed in the build environment, so this is the "ctypes fallback for tests" named in
BASELINE.json's north_star; ``vajax_b200/ffi.py`` holds the FFI registration used when JAX is
present.  There is no CPU fallback: a missing library or a missing CUDA device raises.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VJX_LIB_PATH") or os.path.join(HERE, "libvajax_b200.so")  # override: kernel experiments
CSRC = os.path.join(HERE, "csrc")

NVCC_CMD = [
    "nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "--shared", "-Xcompiler", "-fPIC",
    os.path.join(CSRC, "vjx_engine.cu"), os.path.join(CSRC, "vjx_symbolic.cpp"), "-ldl",
]


def build_library(force: bool = False) -> str:
    """Compile the C-ABI library in-tree for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in ("vjx_engine.cu", "vjx_symbolic.cpp", "vjx_symbolic.h", "vjx_model_abi.h")]
    srcs.append(os.path.join(os.path.dirname(HERE), "include", "vajax_b200.h"))
    if not force and os.path.exists(LIB_PATH):
        if all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
            return LIB_PATH
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    r = subprocess.run(NVCC_CMD + ["-o", tmp], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libvajax_b200.so:\n" + r.stderr[-4000:])
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


class SourceDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("p", C.c_double * 8), ("pwl_offset", C.c_int), ("pwl_count", C.c_int)]


class ModelDesc(C.Structure):
    _fields_ = [
        ("plugin_path", C.c_char_p),
        ("n_dev", C.c_int), ("n_volt", C.c_int), ("n_dev_params", C.c_int), ("n_dev_cache", C.c_int),
        ("n_res", C.c_int), ("n_jac", C.c_int), ("n_lim", C.c_int),
        ("vnode1", C.c_void_p), ("vnode2", C.c_void_p),
        ("dev_params", C.c_void_p), ("dev_cache", C.c_void_p),
        ("shared_params", C.c_void_p), ("n_shared", C.c_int),
        ("shared_cache", C.c_void_p), ("n_shared_cache", C.c_int),
        ("simparam_defaults", C.c_void_p), ("n_simparams", C.c_int),
        ("res_idx", C.c_void_p), ("jac_row", C.c_void_p), ("jac_col", C.c_void_p),
    ]


class Options(C.Structure):
    _fields_ = [
        ("nr_damping", C.c_double), ("reltol", C.c_double), ("abstol", C.c_double), ("vntol", C.c_double),
        ("gmin", C.c_double), ("max_step", C.c_double),
        ("tran_itl", C.c_int),
        ("lte_ratio", C.c_double), ("redo_factor", C.c_double), ("lte_reltol", C.c_double),
        ("lte_abstol", C.c_double), ("min_dt", C.c_double), ("grow_factor", C.c_double), ("tran_ft", C.c_double),
        ("gshunt_init", C.c_double), ("gshunt_target", C.c_double),
        ("gshunt_steps", C.c_int), ("warmup_steps", C.c_int), ("max_order", C.c_int),
        ("max_consecutive_rejects", C.c_int),
        ("integration_method", C.c_int),
        ("pivot_tol", C.c_double),
        ("use_device_graph", C.c_int),
    ]


class CircuitDesc(C.Structure):
    _fields_ = [
        ("n_total", C.c_int), ("n_external", C.c_int), ("n_vsources", C.c_int), ("n_isources", C.c_int),
        ("nse", C.c_int),
        ("csr_indptr", C.c_void_p), ("csr_indices", C.c_void_p),
        ("n_models", C.c_int), ("models", C.POINTER(ModelDesc)),
        ("vs_node_p", C.c_void_p), ("vs_node_n", C.c_void_p),
        ("is_node_p", C.c_void_p), ("is_node_n", C.c_void_p),
        ("vsrc", C.POINTER(SourceDesc)), ("isrc", C.POINTER(SourceDesc)),
        ("pwl_times", C.c_void_p), ("pwl_values", C.c_void_p), ("n_pwl", C.c_int),
        ("noi_nodes", C.c_void_p), ("n_noi", C.c_int),
        ("internal_nodes", C.c_void_p), ("n_internal", C.c_int),
        ("opts", Options),
    ]


class VjxInitArgs(C.Structure):
    _fields_ = [("n_dev", C.c_int), ("vals", C.c_void_p), ("col_off", C.c_void_p), ("col_stride", C.c_void_p),
                ("cache", C.c_void_p), ("collapse", C.c_void_p)]


EXPORTS = [
    "vjx_last_error", "vjx_version", "vjx_device_count", "vjx_create", "vjx_destroy", "vjx_clear_cache",
    "vjx_update_model_data", "vjx_update_sources", "vjx_build_system", "vjx_lu_solve", "vjx_lu_analyze",
    "vjx_lu_stats", "vjx_lu_set_grid_div", "vjx_set_dense_regularization", "vjx_nr_solve", "vjx_nr_solve_async", "vjx_tran_begin", "vjx_tran_run", "vjx_tran_scalars", "vjx_tran_read",
    "vjx_tran_set_step_limit", "vjx_tran_set_save_list", "vjx_tran_run_ms", "vjx_tran_records", "vjx_tran_state", "vjx_time_phase", "vjx_lu_timeline", "vjx_launch_count",
    "vjx_sizes", "vjx_handle_id", "vjx_handle_from_id",
]

_lib: Optional[C.CDLL] = None


class VjxError(RuntimeError):
    pass


def is_available() -> bool:
    """Availability probe (reference: umfpack_jax.is_available, umfpack_jax.py:66-68)."""
    return os.path.exists(LIB_PATH)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VjxError(
            "libvajax_b200.so not built. Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.vjx_last_error.restype = C.c_char_p
    lib.vjx_launch_count.restype = C.c_longlong
    lib.vjx_launch_count.argtypes = [C.c_void_p]
    lib.vjx_destroy.restype = None
    lib.vjx_destroy.argtypes = [C.c_void_p]
    lib.vjx_create.argtypes = [C.POINTER(CircuitDesc), C.c_int, C.POINTER(C.c_void_p)]
    d, i, p = C.c_double, C.c_int, C.c_void_p
    lib.vjx_build_system.argtypes = [p, p, p, p, p, d, d, d, d, d, p, d, p, p, i, p, p, p, p, p, p, p]
    lib.vjx_lu_solve.argtypes = [p, p, p, p, p]
    lib.vjx_lu_analyze.argtypes = [p, p, p]
    lib.vjx_lu_stats.argtypes = [p, p]
    lib.vjx_lu_set_grid_div.argtypes = [p, i]
    lib.vjx_set_dense_regularization.argtypes = [p, d]
    lib.vjx_nr_solve.argtypes = [p, p, p, p, p, d, d, d, d, d, p, d, p, p, p, i, p, C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.POINTER(C.c_double), p, p, p, p, p, p]
    lib.vjx_nr_solve_async.argtypes = [p, p, p, p, p, d, d, d, d, d, p, d, p, p, p, i, p, p, p, p, p, p, p, p]
    lib.vjx_tran_begin.argtypes = [p, p, p, p, p, d, d, d, i, p]
    lib.vjx_tran_run.argtypes = [p, i, C.POINTER(C.c_int), p]
    lib.vjx_tran_scalars.argtypes = [p, p, p]
    lib.vjx_tran_set_step_limit.argtypes = [p, i, p]
    lib.vjx_tran_set_save_list.argtypes = [p, p, i]
    lib.vjx_tran_run_ms.argtypes = [p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.vjx_tran_read.argtypes = [p, i, i, p, p, p, p]
    lib.vjx_tran_records.argtypes = [p, i, i, p, p]
    lib.vjx_tran_state.argtypes = [p, p, p, p, p]
    lib.vjx_time_phase.argtypes = [p, i, i, C.POINTER(C.c_double), C.POINTER(C.c_int), p]
    lib.vjx_lu_timeline.argtypes = [p, i, p, p, p]
    lib.vjx_update_model_data.argtypes = [p, i, p, p, p, p]
    lib.vjx_update_sources.argtypes = [p, C.POINTER(SourceDesc), C.POINTER(SourceDesc)]
    lib.vjx_symbolic_probe.argtypes = [i, p, p, p, d, p, p, p]
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise VjxError(f"libvajax_b200 error {rc}: {load().vjx_last_error().decode()}")

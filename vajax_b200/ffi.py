"""JAX FFI registration of the library's custom-call handlers (BASELINE.json north_star; reference:
vajax/sparse/umfpack_jax.py:40-107 registers its UMFPACK handler the same way).

JAX / jaxlib are not installed in the build image, so nothing here runs in this repo's tests: the ctypes
binding in ``vajax_b200/lib.py`` is the tested path.  Where JAX is present:

    from vajax_b200 import ffi
    ffi.register()                                  # builds libvajax_b200_ffi.so on first use
    x = ffi.lu_solve(handle, csr_data, b)            # jittable; runs on the stream XLA provides

``handle`` is the ``ctypes.c_void_p`` of ``B200Engine.handle`` (or any ``vjx_create`` result); it travels
through the jitted function as the int64 attribute ``vjx_handle_id`` hands out.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import lib as vlib

HERE = os.path.dirname(os.path.abspath(__file__))
FFI_LIB_PATH = os.path.join(HERE, "libvajax_b200_ffi.so")
TARGETS = ("vjx_ffi_solve_f64", "vjx_ffi_build_system_f64", "vjx_ffi_nr_solve_f64")
_registered = False


def _jax():
    try:
        import jax  # noqa: F401
    except ImportError as e:   # pragma: no cover - no jax in the build image
        raise vlib.VjxError("vajax_b200.ffi needs JAX (jax.ffi); use the ctypes binding vajax_b200.lib without it") from e
    import jax

    return jax


def build_ffi(force: bool = False) -> str:
    """Compile csrc/vjx_ffi.cc against jaxlib's XLA FFI headers and link it to libvajax_b200.so."""
    jax = _jax()
    src = os.path.join(HERE, "csrc", "vjx_ffi.cc")
    if not force and os.path.exists(FFI_LIB_PATH) and os.path.getmtime(FFI_LIB_PATH) >= os.path.getmtime(src):
        return FFI_LIB_PATH
    vlib.build_library()
    cmd = ["nvcc", "-O2", "-std=c++17", "--shared", "-Xcompiler", "-fPIC", "-I", jax.ffi.include_dir(), src,
           "-L", HERE, "-l:libvajax_b200.so", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN", "-o", FFI_LIB_PATH]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise vlib.VjxError("building libvajax_b200_ffi.so failed:\n" + r.stderr[-4000:])
    return FFI_LIB_PATH


def register() -> None:
    """Register the three handlers as CUDA FFI targets (idempotent)."""
    global _registered
    if _registered:
        return
    jax = _jax()
    vlib.load()   # libvajax_b200.so first: the FFI library resolves the C-ABI symbols against it
    dll = C.CDLL(build_ffi())
    for name in TARGETS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(dll, name)), platform="CUDA")
    _registered = True


def handle_id(handle) -> int:
    lib = vlib.load()
    lib.vjx_handle_id.restype = C.c_longlong
    lib.vjx_handle_id.argtypes = [C.c_void_p]
    return int(lib.vjx_handle_id(handle))


def sizes(handle) -> dict:
    out = (C.c_int * 8)()
    lib = vlib.load()
    lib.vjx_sizes.argtypes = [C.c_void_p, C.c_void_p]
    vlib.check(lib.vjx_sizes(handle, out))
    keys = ("n_total", "n_external", "n_vsources", "n_isources", "n_unknowns", "n_aug", "nse", "total_limit_states")
    return dict(zip(keys, (int(v) for v in out)))


def lu_solve(handle, csr_data, b):
    """``linear_solve_fn(csr_data, b)`` as a jittable custom call (seam S1, solver_factories.py:985-1004)."""
    jax = _jax()
    import jax.numpy as jnp

    register()
    sz = sizes(handle)
    call = jax.ffi.ffi_call("vjx_ffi_solve_f64", jax.ShapeDtypeStruct((sz["n_aug"],), jnp.float64))
    return call(csr_data, b, handle=np.int64(handle_id(handle)))


def _opt(jnp, a):
    """None -> zero-length operand (the handlers read a zero-length buffer as NULL)."""
    return jnp.zeros((0,), dtype=jnp.float64) if a is None else a


def build_system(handle, X, vsource_vals, isource_vals, Q_prev, integ_c0, gmin=1e-12, gshunt=0.0, integ_c1=0.0,
                 integ_d1=0.0, dQdt_prev=None, integ_c2=0.0, Q_prev2=None, limit_state_in=None, nr_iteration=1):
    """Seam S2 as a jittable custom call: ``(csr_data, f, Q, I_vsource, limit_state_out, max_res_contrib)``, the
    reference's ``build_system`` contract (mna_builder.py:528-550,1010) with ``device_arrays`` inside the handle.
    The scalars are call attributes: they must be Python numbers at trace time (one traced call per coefficient set)."""
    jax = _jax()
    import jax.numpy as jnp

    register()
    sz = sizes(handle)
    f64 = jnp.float64
    out = (jax.ShapeDtypeStruct((sz["nse"],), f64), jax.ShapeDtypeStruct((sz["n_aug"],), f64),
           jax.ShapeDtypeStruct((sz["n_unknowns"],), f64), jax.ShapeDtypeStruct((max(sz["n_vsources"], 1),), f64),
           jax.ShapeDtypeStruct((max(sz["total_limit_states"], 1),), f64), jax.ShapeDtypeStruct((sz["n_unknowns"],), f64))
    call = jax.ffi.ffi_call("vjx_ffi_build_system_f64", out)
    return call(X, _opt(jnp, vsource_vals), _opt(jnp, isource_vals), Q_prev, _opt(jnp, dQdt_prev), _opt(jnp, Q_prev2),
                _opt(jnp, limit_state_in), handle=np.int64(handle_id(handle)), nr_iteration=np.int32(nr_iteration),
                integ_c0=float(integ_c0), gmin=float(gmin), gshunt=float(gshunt), integ_c1=float(integ_c1),
                integ_d1=float(integ_d1), integ_c2=float(integ_c2))


def nr_solve(handle, X_init, vsource_vals, isource_vals, Q_prev, integ_c0, gmin=1e-12, gshunt=0.0, integ_c1=0.0,
             integ_d1=0.0, dQdt_prev=None, integ_c2=0.0, Q_prev2=None, limit_state_in=None, res_tol_floor=None,
             max_iterations=-1):
    """Seam S3 as a jittable, stream-ordered custom call (``vjx_nr_solve_async``): the reference's 9 results
    ``(X, iterations, converged, max_f, Q, dQdt, I_vsource, limit_state_out, max_res_contrib)``
    (solver_factories.py:604-725); the device writes the three scalars, nothing synchronises."""
    jax = _jax()
    import jax.numpy as jnp

    register()
    sz = sizes(handle)
    f64 = jnp.float64
    nx, nu = sz["n_total"] + sz["n_vsources"], sz["n_unknowns"]
    out = (jax.ShapeDtypeStruct((nx,), f64), jax.ShapeDtypeStruct((4,), f64), jax.ShapeDtypeStruct((nu,), f64),
           jax.ShapeDtypeStruct((nu,), f64), jax.ShapeDtypeStruct((max(sz["n_vsources"], 1),), f64),
           jax.ShapeDtypeStruct((max(sz["total_limit_states"], 1),), f64), jax.ShapeDtypeStruct((nu,), f64))
    call = jax.ffi.ffi_call("vjx_ffi_nr_solve_f64", out)
    X, sc, Q, dQdt, Ivs, lim, mrc = call(
        X_init, _opt(jnp, vsource_vals), _opt(jnp, isource_vals), Q_prev, _opt(jnp, dQdt_prev), _opt(jnp, Q_prev2),
        _opt(jnp, limit_state_in), _opt(jnp, res_tol_floor), handle=np.int64(handle_id(handle)),
        max_iterations=np.int32(max_iterations), integ_c0=float(integ_c0), gmin=float(gmin), gshunt=float(gshunt),
        integ_c1=float(integ_c1), integ_d1=float(integ_d1), integ_c2=float(integ_c2))
    return (X, sc[0].astype(jnp.int32), sc[1] > 0.5, sc[2], Q, dQdt, Ivs[:sz["n_vsources"]],
            lim[:sz["total_limit_states"]], mrc)

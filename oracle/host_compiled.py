"""Compiled host evaluation of the model functions - the TIMING baseline of bench.py's CPU legs.

TEST / BENCH INFRASTRUCTURE, like everything under oracle/: only tests/ and bench.py's ``cpu_baseline`` /
``--impl reference`` legs import it.  It is NOT the parity checker - that stays the NumPy interpreter
(oracle/mir_interp.py), which shares no code with the product.  This module exists because the reference
does not interpret MIR either: it lowers the eval function to XLA and runs the compiled, multi-threaded CPU
code (``jax.vmap(eval_fn)`` under ``jit``: vajax/analysis/openvaf_models.py:997,
openvaf_jax/codegen/function_builder.py:828-1005).  A CPU leg that times an interpreter says nothing about the
reference's CPU path; this one times the same straight-line fp64 code compiled with ``g++ -O2`` and spread over
the host cores with OpenMP, one instance per loop iteration - the closest stand-in for XLA-CPU available in an
image without jax.

The statement-level translation MIR -> C is the one the CUDA emitter uses (``cuda_emitter._Body``: real control
flow, no FMA contraction), wrapped in a host loop with the interpreter's calling convention, so that
``eval_model`` here is a drop-in for ``mir_interp.eval_model`` (same arguments, same result dict).
``tests/test_host_compiled.py`` holds it to the interpreter on every fixture model (both agree to the last bits of
libm's exp / log / pow)."""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import re
import subprocess
import tempfile
import threading
from typing import Dict, Optional, Tuple

import numpy as np

from vajax_b200.codegen import cuda_emitter as em
from vajax_b200.mir.types import parse_function

from .mir_interp import SIMPARAM_DEFAULTS

BUILD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")
CXXFLAGS = ["-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c++17", "-w"]
_LOCK = threading.Lock()
_MODELS: Dict[Tuple[int, bool], "HostModel"] = {}
THREADS = int(os.environ.get("VJX_HOST_THREADS", "0")) or (os.cpu_count() or 1)

_HOST_SHIM = r"""
#include <math.h>
#include <string.h>
#include <stddef.h>
#include <omp.h>
#define __device__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
"""


def _host_prelude() -> str:
    """The emitter's helper functions (safe divide, NaN-propagating min / max, pnjlim, fetlim ...) as host code."""
    p = em.PRELUDE.replace("#include <cuda_runtime.h>", "").replace('#include "vjx_model_abi.h"', "")
    return _HOST_SHIM + p


def _mir_var(ref: str) -> str:
    return "v" + ref[4:] if ref and ref.startswith("mir_") else (ref or "")


class HostModel:
    """One model's eval function compiled for the host (cached on disk by source hash)."""

    def __init__(self, bundle: dict, use_limit: bool):
        self.bundle = bundle
        self.use_limit = bool(use_limit)
        fn = parse_function("eval", bundle["mir_data"], bundle.get("str_constants"))
        dae = bundle["dae_data"]
        self.params = list(fn.params)
        self.R, self.J = len(dae["residuals"]), len(dae["jacobian"])
        body = em._Body(fn, self.use_limit,
                        limit_in_expr="(LIN ? LIN[(size_t)d * nlin + ({idx} < nlin ? {idx} : nlin - 1)] : 0.0)")
        head = [f"{p} = (PC[{i}] >= 0) ? D[(size_t)PC[{i}] * n + d] : S[{i}];" for i, p in enumerate(fn.params)]
        pidx2val = {p[0]: f"v{p[1]}" for p in bundle["all_func_params"]}
        cache_map = bundle["init_mir_data"].get("cache_mapping", [])
        for ci, mp in enumerate(cache_map):
            var = pidx2val.get(mp["eval_param"], f"cached_{mp['eval_param']}")
            body.types.setdefault(var, "f")
            head.append(f"{var} = ({ci} < ccols) ? CV[(size_t)d * cstride + {ci}] : 0.0;")
        code = body.emit()

        def outv(ref: str) -> str:
            v = _mir_var(ref)
            return body.as_f(v) if body.defined(v) else "0.0"

        tail = []
        for i, r in enumerate(dae["residuals"]):
            tail.append(f"RR[(size_t)d * {self.R} + {i}] = {outv(r['resist_var'])}; RQ[(size_t)d * {self.R} + {i}] = {outv(r['react_var'])};")
        for j, e in enumerate(dae["jacobian"]):
            tail.append(f"JR[(size_t)d * {self.J} + {j}] = {outv(e['resist_var'])}; JQ[(size_t)d * {self.J} + {j}] = {outv(e['react_var'])};")
        # limiting right-hand side: sum over limited branches of J[row, hi] * (V_limited - V_raw)
        # (function_builder.py:1126-1212), accumulated in branch order like the interpreter
        self.n_lim = body.limit_count if self.use_limit else 0
        self.limit_traced = bool(body.limit_raw)
        if self.use_limit and self.n_lim:
            names = bundle["param_names"]
            pid2idx = {p: i for i, p in enumerate(fn.params)}

            def hi_node(vraw: str) -> Optional[str]:
                idx = pid2idx.get(vraw)
                if idx is None:
                    idx = next((oi for oi, vn in pidx2val.items() if vn == vraw), None)
                if idx is None or idx >= len(names):
                    return None
                m = re.match(r"^V\((\w+)(?:,(\w+))?\)$", names[idx])
                return m.group(1) if m else None

            jkey = {(e["row_node_name"], e["col_node_name"]): j for j, e in enumerate(dae["jacobian"])}
            for li in range(self.n_lim):
                vlim, vraw = body.limit_store.get(li), body.limit_raw.get(li)
                if vlim is not None and body.defined(vlim):
                    tail.append(f"LOUT[(size_t)d * {self.n_lim} + {li}] = {body.as_f(vlim)};")
                if vraw is None or vlim is None or not body.defined(vraw) or not body.defined(vlim):
                    continue
                hn = hi_node(vraw)
                if hn is None:
                    continue
                tail.append(f"{{ const double delta = {body.as_f(vlim)} - {body.as_f(vraw)};")
                for i, res in enumerate(dae["residuals"]):
                    j = jkey.get((res["node_name"], hn))
                    if j is None:
                        continue
                    e = dae["jacobian"][j]
                    rv, qv = _mir_var(e["resist_var"]), _mir_var(e["react_var"])
                    if body.defined(rv):
                        tail.append(f"  LR[(size_t)d * {self.R} + {i}] = LR[(size_t)d * {self.R} + {i}] + {body.as_f(rv)} * delta;")
                    if body.defined(qv):
                        tail.append(f"  LQ[(size_t)d * {self.R} + {i}] = LQ[(size_t)d * {self.R} + {i}] + {body.as_f(qv)} * delta;")
                tail.append("}")
        self.simparams = list(body.simparams)
        sp = [f"const double sp{k} = SP[{k}];" for k in range(len(self.simparams))]
        src = [_host_prelude(),
               "static inline __attribute__((always_inline)) void eval_one(const int d, const int n, const double* S, const int* PC, const double* D,",
               "    const double* CV, const size_t cstride, const int ccols, const double* SP, const double* LIN, const int nlin,",
               "    double* RR, double* RQ, double* JR, double* JQ, double* LR, double* LQ, double* LOUT) {"]
        src += ["  " + s for s in body.decls()] + ["  " + s for s in sp] + ["  " + s for s in head]
        src += ["  " + s for s in code] + ["  " + s for s in tail] + ["}"]
        src.append("""
extern "C" void host_eval(int n, const double* S, const int* PC, const double* D, const double* CV, size_t cstride, int ccols,
                          const double* SP, const double* LIN, int nlin, double* RR, double* RQ, double* JR, double* JQ,
                          double* LR, double* LQ, double* LOUT, int threads) {
  #pragma omp parallel for schedule(static) num_threads(threads)
  for (int d = 0; d < n; ++d) eval_one(d, n, S, PC, D, CV, cstride, ccols, SP, LIN, nlin, RR, RQ, JR, JQ, LR, LQ, LOUT);
}
""")
        self.source = "\n".join(src)
        self.lib = C.CDLL(self._compile())
        self.lib.host_eval.restype = None

    def _compile(self) -> str:
        h = hashlib.sha256((self.source + " ".join(CXXFLAGS)).encode()).hexdigest()[:20]
        os.makedirs(BUILD_DIR, exist_ok=True)
        so = os.path.join(BUILD_DIR, f"host_eval_{h}.so")
        if os.path.exists(so):
            return so
        fd, cpp = tempfile.mkstemp(prefix=f"host_eval_{h}.", suffix=".cpp", dir=BUILD_DIR)
        with os.fdopen(fd, "w") as f:
            f.write(self.source)
        tmp = cpp[:-4] + ".so.tmp"
        try:
            r = subprocess.run(["g++"] + CXXFLAGS + ["-o", tmp, cpp], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("g++ failed for the host eval function:\n" + r.stderr[-3000:])
            os.replace(tmp, so)
        finally:
            for t in (cpp, tmp):
                if os.path.exists(t):
                    os.unlink(t)
        return so

    def run(self, param_values: Dict[int, object], cache_values: np.ndarray, n: int, simparams: Dict[str, object],
            limit_state_in: Optional[np.ndarray], limit_param_map, threads: int) -> dict:
        lpm = limit_param_map or {}
        P = len(self.params)
        S = np.zeros(max(P, 1))
        PC = np.full(max(P, 1), -1, dtype=np.int32)
        cols = []
        for i in range(P):
            if i in lpm:
                kind, name = lpm[i]
                if kind == "prev_state" and limit_state_in is not None:
                    idx = int(name.split("_")[-1])
                    val = limit_state_in[:, min(idx, limit_state_in.shape[1] - 1)]
                elif kind == "enable_lim":
                    val = 1.0 if self.use_limit else 0.0
                else:
                    val = 0.0
            else:
                val = param_values.get(i, 0.0)
            a = np.asarray(val, dtype=np.float64)
            if a.ndim == 0:
                S[i] = float(a)
            else:
                PC[i] = len(cols)
                cols.append(np.broadcast_to(a, (n,)))
        D = np.ascontiguousarray(np.stack(cols)) if cols else np.zeros((1, max(n, 1)))
        CV = np.ascontiguousarray(cache_values, dtype=np.float64)
        if CV.ndim != 2 or CV.shape[1] == 0:
            CV = np.zeros((max(n, 1), 1))
            ccols = 0
        else:
            ccols = CV.shape[1]
        SP = np.array([float(simparams.get(nm, SIMPARAM_DEFAULTS.get(nm, 0.0))) for nm in self.simparams] or [0.0])
        if limit_state_in is not None and np.size(limit_state_in):
            LIN = np.ascontiguousarray(limit_state_in, dtype=np.float64)
            lin_p, nlin = LIN.ctypes.data_as(C.c_void_p), LIN.shape[1]
        else:
            LIN, lin_p, nlin = None, None, 0
        R, J = self.R, self.J
        RR, RQ, LR, LQ = (np.zeros((n, R)) for _ in range(4))
        JR, JQ = np.zeros((n, J)), np.zeros((n, J))
        LOUT = np.zeros((n, max(self.n_lim, 0)))
        lout = LOUT if LOUT.size else np.zeros(1)
        p = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
        if n > 0:
            self.lib.host_eval(C.c_int(n), p(S), p(PC), p(D), p(CV), C.c_size_t(CV.shape[1]), C.c_int(ccols), p(SP), lin_p,
                               C.c_int(nlin), p(RR), p(RQ), p(JR), p(JQ), p(LR), p(LQ), p(lout), C.c_int(max(1, threads)))
        return {"res_resist": RR, "res_react": RQ, "jac_resist": JR, "jac_react": JQ, "lim_rhs_resist": LR,
                "lim_rhs_react": LQ, "limit_state_out": LOUT, "n_limit_states": self.n_lim,
                "limit_traced": self.limit_traced, "simparams_used": list(self.simparams),
                "flops_executed": None, "transc_executed": None}


def get_model(bundle: dict, use_limit: bool) -> HostModel:
    key = (id(bundle), bool(use_limit))
    with _LOCK:
        m = _MODELS.get(key)
        if m is None:
            m = _MODELS[key] = HostModel(bundle, use_limit)
    return m


def eval_model(bundle: dict, param_values: Dict[int, object], cache_values: np.ndarray, n: int,
               simparams: Dict[str, object], limit_state_in: Optional[np.ndarray] = None,
               use_limit: bool = False, limit_param_map=None, threads: Optional[int] = None) -> dict:
    """Drop-in for ``mir_interp.eval_model`` (same arguments and result keys), compiled and multi-threaded."""
    return get_model(bundle, use_limit).run(param_values, cache_values, n, simparams, limit_state_in,
                                            limit_param_map, THREADS if threads is None else threads)


class installed:
    """Context manager: the oracle engine evaluates models through the compiled host code inside the block.

        with host_compiled.installed(threads=8):
            oracle_engine.run_transient(...)
    """

    def __init__(self, threads: Optional[int] = None):
        self.threads = THREADS if threads is None else int(threads)

    def __enter__(self):
        from . import mir_interp
        self._mod, self._orig = mir_interp, mir_interp.eval_model
        th = self.threads
        mir_interp.eval_model = lambda *a, **k: eval_model(*a, threads=th, **k)
        return self

    def __exit__(self, *exc):
        self._mod.eval_model = self._orig
        return False

"""CUDA-C emitter: lowers OpenVAF MIR to one fp64 ``__global__`` kernel per model.

This is the B200 counterpart of the reference's MIR -> Python/jnp source generator
(reference: openvaf_jax/codegen/function_builder.py:828-1005 ``build_with_cache_split``,
:605-692 ``build_split`` for init; opcode semantics openvaf_jax/codegen/instruction.py:156-703,
openvaf_jax/openvaf_ast/expressions.py:349-432).  Differences by design:

* The reference linearises the CFG (every block executed, PHI -> ``jnp.where`` through a
  dominator search, openvaf_jax/mir/ssa.py:212-255).  A GPU thread has real control flow, so
  blocks become labels, ``br`` becomes ``if/goto`` and PHIs are resolved by assignments on the
  incoming edges.  Both agree whenever the reference's PHI resolution is correct.
* One thread evaluates one instance.  Varying parameters and cache columns are
  structure-of-arrays ``(column, n_dev)`` so a warp's loads are coalesced; shared parameters /
  shared cache live in ``__constant__`` memory (operands come straight from the constant bank).
* Branch voltages are gathered from the node-voltage vector inside the kernel (the reference
  writes them into ``device_params`` first, vajax/analysis/mna_builder.py:915-918).
* Outputs are written exactly once, straight to their position in the slot-sorted contribution
  arrays consumed by the stamping kernels (``jr + c0*jc`` combined, NaN -> 0 as in
  vajax/analysis/mna_builder.py:953-966) - no (n_dev, J) intermediates.
* MIR loops are rejected (``Function.topo_order``) instead of copying the reference's float32
  loop-carried-state quirk.

Python's dynamic typing in the generated reference code (ints/bools/floats mix freely) is
reproduced with static type inference + Python-style promotion (bool -> int -> double),
floor-division and sign-of-divisor modulo.
"""

from __future__ import annotations

import hashlib
import os
import re
import subprocess
import tempfile
import threading
from typing import Dict, List, Optional, Tuple

from ..mir.types import ANALYSIS_TYPE_MAP, Function, Inst, classify_call, parse_function

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
PLUGIN_DIR = os.path.join(os.path.dirname(HERE), "_plugins")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-I", CSRC,
]

# SIMPARAM defaults (reference: vajax/__init__.py:191-208)
SIMPARAM_DEFAULTS = {
    "$analysis_type": 0.0, "gmin": 1e-12, "abstol": 1e-12, "vntol": 1e-6, "reltol": 1e-3,
    "tnom": 27.0, "scale": 1.0, "shrink": 0.0, "imax": 1.0, "$abstime": 0.0, "$mfactor": 1.0,
    "iniLim": 0.0, "iteration": 1.0, "gdev": 0.0, "sourceScaleFactor": 1.0,
}

ARITH = {"fadd": "+", "fsub": "-", "fmul": "*", "iadd": "+", "isub": "-", "imul": "*"}
CMP = {"feq": "==", "fne": "!=", "flt": "<", "fgt": ">", "fle": "<=", "fge": ">=",
       "ieq": "==", "ine": "!=", "ilt": "<", "igt": ">", "ile": "<=", "ige": ">=",
       "beq": "==", "bne": "!=", "seq": "==", "sne": "!="}
BITWISE = {"iand": "&", "ior": "|", "ixor": "^", "ishl": "<<", "ishr": ">>"}
UNARY_F = {"exp": "vjx_exp", "floor": "floor", "ceil": "ceil", "sin": "sin", "cos": "cos", "tan": "tan",
           "sinh": "sinh", "cosh": "cosh", "tanh": "tanh", "abs": "fabs", "fabs": "fabs",
           "asin": "asin", "acos": "acos", "atan": "atan", "asinh": "asinh", "acosh": "acosh",
           "atanh": "atanh"}
BINARY_F = {"pow": "vjx_pow", "atan2": "atan2", "hypot": "hypot", "fmin": "vjx_fmin", "fmax": "vjx_fmax"}
TRANSC = {"exp", "ln", "log", "clog2", "sqrt", "pow", "sin", "cos", "tan", "sinh", "cosh", "tanh",
          "asin", "acos", "atan", "asinh", "acosh", "atanh", "atan2", "hypot"}

PRELUDE = r"""
#include <cuda_runtime.h>
#include <math.h>
#include "vjx_model_abi.h"
#define VJX_INF (__longlong_as_double(0x7ff0000000000000LL))
#define VJX_NAN (__longlong_as_double(0x7ff8000000000000LL))
// jnp.maximum / jnp.minimum propagate NaN (C fmax/fmin do not)
__device__ __forceinline__ double vjx_fmax(double a, double b) { return (a != a || b != b) ? VJX_NAN : (a > b ? a : b); }
__device__ __forceinline__ double vjx_fmin(double a, double b) { return (a != a || b != b) ? VJX_NAN : (a < b ? a : b); }
// safe divide: where(b == 0, 0, a / where(b == 0, 1, b))   (expressions.py:349-374)
// VJX_MATH_FN: __forceinline__ by default; __noinline__ (VJX_EVAL_NOINLINE_MATH=1) turns the ~700 divisions /
// roots / logarithms / exponentials of a PSP103-weight model into calls of a few shared routines: a third of the
// kernel's code is their inlined expansions, and the kernel is bound by instruction fetch, not by the fp64 pipe
#ifndef VJX_MATH_FN
#define VJX_MATH_FN __forceinline__
#endif
__device__ VJX_MATH_FN double vjx_fdiv(double a, double b) { return b == 0.0 ? 0.0 : a / b; }
__device__ VJX_MATH_FN double vjx_sqrt(double a) { return sqrt(vjx_fmax(a, 0.0)); }
__device__ VJX_MATH_FN double vjx_ln(double a) { return log(vjx_fmax(a, 1e-300)); }
__device__ VJX_MATH_FN double vjx_exp(double a) { return exp(a); }
__device__ VJX_MATH_FN double vjx_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ double vjx_log10(double a) { return log10(vjx_fmax(a, 1e-300)); }
__device__ __forceinline__ double vjx_clog2(double a) { return ceil(log2(vjx_fmax(a, 1e-300))); }
// Python semantics of // and % (sign of the divisor)
__device__ __forceinline__ int vjx_ifloordiv(int a, int b) { if (b == 0) return 0; int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }
__device__ __forceinline__ int vjx_imod(int a, int b) { if (b == 0) return 0; int r = a % b; if (r != 0 && ((r < 0) != (b < 0))) r += b; return r; }
__device__ __forceinline__ double vjx_ffloordiv(double a, double b) { return floor(a / b); }
__device__ __forceinline__ double vjx_fmod(double a, double b) { double r = fmod(a, b); if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b; return r; }
// SPICE limiting (reference: vajax/analysis/limiting.py:22-107)
__device__ __forceinline__ double vjx_pnjlim(double vnew, double vold, double vt, double vcrit) {
  double delta_v = vnew - vold;
  double arg = delta_v / vt;
  bool fwd = (vnew > vcrit) && (fabs(delta_v) > 2 * vt);
  double lim_pos = vold + vt * log(vjx_fmax(1.0 + arg, 1e-30));
  double lim_neg = vold - vt * log(vjx_fmax(1.0 - arg, 1e-30));
  double lim_forward = arg >= 0 ? lim_pos : lim_neg;
  double lim_reverse = vnew > 0 ? vt * log(vjx_fmax(vnew / vt, 1e-30)) : vcrit;
  double lim_fwd = vold > 0 ? lim_forward : lim_reverse;
  double neg_clamp = vold > 0 ? -vold - 1.0 : 2.0 * vold - 1.0;
  double lim_negv = vjx_fmax(vnew, neg_clamp);
  return fwd ? lim_fwd : (vnew < 0 ? lim_negv : vnew);
}
__device__ __forceinline__ double vjx_fetlim(double vnew, double vold, double vto) {
  double max_step = (vold >= vto) ? (2 * fabs(vold - vto) + 2) : 0.5;
  double delta = vnew - vold;
  double c = delta < -max_step ? -max_step : (delta > max_step ? max_step : delta);
  return vold + c;
}
__device__ __forceinline__ double vjx_nan0(double v) { return (v != v) ? 0.0 : v; }
"""


def _flit(x: float) -> str:
    if x != x:
        return "VJX_NAN"
    if x == float("inf"):
        return "VJX_INF"
    if x == float("-inf"):
        return "(-VJX_INF)"
    s = repr(float(x))
    if "e" not in s and "." not in s and "n" not in s:
        s += ".0"
    return s


class _Body:
    """Emits the body of one MIR function with real control flow."""

    def __init__(self, fn: Function, use_limit: bool, limit_in_expr: str = "0.0"):
        self.fn = fn
        self.use_limit = use_limit
        self.types: Dict[str, str] = {}
        self.lines: List[str] = []
        self.simparams: List[str] = ["$analysis_type"]
        self.limit_count = 0
        self.limit_store: Dict[int, str] = {}
        self.builtin_raw: Dict[str, str] = {}
        self.limit_raw: Dict[int, str] = {}
        self.flops = {"arith": 0, "div": 0, "transc": 0, "insts": 0}
        self.warnings: List[str] = []
        self.limit_in_expr = limit_in_expr
        for p in fn.params:
            self.types[p] = "f"

    # ---- operand helpers
    def ty(self, v: str) -> str:
        if v in self.fn.fconst:
            return "f"
        if v in self.fn.bconst:
            return "b"
        if v in self.fn.iconst:
            return "i"
        return self.types.get(v, "f")

    def defined(self, v: str) -> bool:
        return v in self.fn.fconst or v in self.fn.bconst or v in self.fn.iconst or v in self.types

    def ex(self, v: str) -> str:
        """Operand expression in its native type; undefined operands -> 0.0
        (reference: openvaf_jax/codegen/context.py:278-280)."""
        if v in self.fn.fconst:
            return _flit(self.fn.fconst[v])
        if v in self.fn.bconst:
            return "true" if self.fn.bconst[v] else "false"
        if v in self.fn.iconst:
            return str(int(self.fn.iconst[v]))
        if v in self.types:
            return v
        return "0.0"

    def as_f(self, v: str) -> str:
        t = self.ty(v)
        e = self.ex(v)
        if t == "f":
            return e
        if t == "b":
            return f"({e} ? 1.0 : 0.0)"
        return f"((double){e})"

    def as_b(self, v: str) -> str:
        t = self.ty(v)
        e = self.ex(v)
        if t == "b":
            return e
        if t == "i":
            return f"({e} != 0)"
        return f"({e} != 0.0)"

    def as_i(self, v: str) -> str:
        t = self.ty(v)
        e = self.ex(v)
        if t == "i":
            return e
        if t == "b":
            return f"((int){e})"
        return f"((int){e})"

    def num(self, v: str, t: str) -> str:
        return self.as_f(v) if t == "f" else self.as_i(v)

    @staticmethod
    def promote(*ts: str) -> str:
        if "f" in ts:
            return "f"
        return "i"

    # ---- simparams / limits
    def simparam_index(self, name: str) -> int:
        if name not in self.simparams:
            self.simparams.append(name)
        return self.simparams.index(name)

    @staticmethod
    def _lim_index(func_name: str) -> int:
        m = re.search(r"LimState\((\d+)\)", func_name)
        if m:
            return int(m.group(1))
        m = re.search(r"(\d+)", func_name)
        return int(m.group(1)) if m else 0

    def _register_limit(self, idx: int):
        # limit registry assigns indices in first-use order of the *names* lim_state<idx>
        # (reference: context.py:161-179); with one state per name the array index equals the
        # registration order.  We keep a map name-index -> array index.
        if not hasattr(self, "_limreg"):
            self._limreg: Dict[int, int] = {}
        if idx not in self._limreg:
            self._limreg[idx] = self.limit_count
            self.limit_count += 1
        return self._limreg[idx]

    # ---- instruction translation
    def translate(self, ins: Inst) -> Optional[Tuple[str, str]]:
        """Return (type, expression) for a value-producing instruction."""
        op = ins.opcode
        o = ins.operands
        self.flops["insts"] += 1
        if op in ARITH:
            t = self.promote(self.ty(o[0]), self.ty(o[1]))
            if t == "f":
                self.flops["arith"] += 1
            return t, f"{self.num(o[0], t)} {ARITH[op]} {self.num(o[1], t)}"
        if op == "fdiv":
            self.flops["div"] += 1
            return "f", f"vjx_fdiv({self.as_f(o[0])}, {self.as_f(o[1])})"
        if op == "frem" or op == "irem":
            t = self.promote(self.ty(o[0]), self.ty(o[1]))
            f = "vjx_fmod" if t == "f" else "vjx_imod"
            return t, f"{f}({self.num(o[0], t)}, {self.num(o[1], t)})"
        if op == "idiv":
            t = self.promote(self.ty(o[0]), self.ty(o[1]))
            f = "vjx_ffloordiv" if t == "f" else "vjx_ifloordiv"
            return t, f"{f}({self.num(o[0], t)}, {self.num(o[1], t)})"
        if op in BITWISE:
            return "i", f"{self.as_i(o[0])} {BITWISE[op]} {self.as_i(o[1])}"
        if op in CMP:
            ta, tb = self.ty(o[0]), self.ty(o[1])
            if ta == "b" and tb == "b":
                return "b", f"{self.ex(o[0])} {CMP[op]} {self.ex(o[1])}"
            t = self.promote(ta, tb)
            return "b", f"{self.num(o[0], t)} {CMP[op]} {self.num(o[1], t)}"
        if op == "fneg":
            self.flops["arith"] += 1
            t = self.promote(self.ty(o[0]))
            return t, f"-{self.num(o[0], t)}"
        if op == "ineg":
            t = self.promote(self.ty(o[0]))
            return t, f"-{self.num(o[0], t)}"
        if op == "bnot":
            return "b", f"!{self.as_b(o[0])}"
        if op == "inot":
            return "i", f"~{self.as_i(o[0])}"
        if op == "sqrt":
            self.flops["transc"] += 1
            return "f", f"vjx_sqrt({self.as_f(o[0])})"
        if op == "ln":
            self.flops["transc"] += 1
            return "f", f"vjx_ln({self.as_f(o[0])})"
        if op == "log":
            self.flops["transc"] += 1
            return "f", f"vjx_log10({self.as_f(o[0])})"
        if op == "clog2":
            self.flops["transc"] += 1
            return "f", f"vjx_clog2({self.as_f(o[0])})"
        if op in UNARY_F:
            if op in TRANSC:
                self.flops["transc"] += 1
            return "f", f"{UNARY_F[op]}({self.as_f(o[0])})"
        if op in BINARY_F and len(o) >= 2:
            if op in TRANSC:
                self.flops["transc"] += 1
            return "f", f"{BINARY_F[op]}({self.as_f(o[0])}, {self.as_f(o[1])})"
        if op == "ifcast":
            return "f", self.as_f(o[0])
        if op == "ficast":
            return "i", f"((int)floor({self.as_f(o[0])}))"
        if op == "fbcast":
            return "b", f"({self.as_f(o[0])} != 0.0)"
        if op == "ibcast":
            t = self.promote(self.ty(o[0]))
            return "b", f"({self.num(o[0], t)} != 0)"
        if op == "bfcast":
            return "f", f"({self.as_b(o[0])} ? 1.0 : 0.0)"
        if op == "bicast":
            return "i", self.as_i(o[0])
        if op in ("optbarrier", "copy"):
            if not o:
                return "f", "0.0"
            return self.ty(o[0]), self.ex(o[0])
        if op == "select":
            if len(o) < 3:
                return "f", "0.0"
            ta, tb = self.ty(o[1]), self.ty(o[2])
            if ta == "b" and tb == "b":
                return "b", f"({self.as_b(o[0])} ? {self.ex(o[1])} : {self.ex(o[2])})"
            t = self.promote(ta, tb)
            return t, f"({self.as_b(o[0])} ? {self.num(o[1], t)} : {self.num(o[2], t)})"
        if op == "call":
            return self.translate_call(ins)
        raise ValueError(f"Unknown opcode: {ins.opcode}")

    def translate_call(self, ins: Inst) -> Tuple[str, str]:
        kind = classify_call(ins.func_name)
        o = ins.operands
        fl = ins.func_name.lower()
        if kind == "simparam":
            if o and o[0] in self.fn.sconst:
                idx = self.simparam_index(self.fn.sconst[o[0]])
                return "f", f"sp{idx}"
            return "f", "0.0"
        if kind == "analysis":
            if o and o[0] in self.fn.sconst:
                code = ANALYSIS_TYPE_MAP.get(self.fn.sconst[o[0]].lower(), 0)
                return "b", f"(sp0 == {float(code)})"
            return "b", "false"
        if kind == "passthrough":
            return (self.ty(o[0]), self.ex(o[0])) if o else ("f", "0.0")
        if kind == "abs":
            return ("f", f"fabs({self.as_f(o[0])})") if o else ("f", "0.0")
        if kind in ("min", "max") and len(o) >= 2:
            return "f", f"vjx_f{kind}({self.as_f(o[0])}, {self.as_f(o[1])})"
        if kind == "storelimit":
            if self.use_limit:
                idx = self._lim_index(ins.func_name)
                ai = self._register_limit(idx)
                if not o:
                    return "f", "0.0"
                self.limit_store[ai] = o[0]
                raw = self.builtin_raw.get(o[0])
                if raw:
                    self.limit_raw[ai] = raw
                return self.ty(o[0]), self.ex(o[0])
            return (self.ty(o[0]), self.ex(o[0])) if o else ("f", "0.0")
        if kind == "loadlimit":
            if self.use_limit:
                ai = self._register_limit(self._lim_index(ins.func_name))
                return "f", self.limit_in_expr.format(idx=ai)
            return (self.ty(o[0]), self.ex(o[0])) if o else ("f", "0.0")
        if kind == "builtinlimit":
            if not o:
                return "f", "0.0"
            if self.use_limit:
                name = "pnjlim"
                for known in ("pnjlim", "fetlim", "limexp", "limvds"):
                    if known in fl:
                        name = known
                        break
                if name not in ("pnjlim", "fetlim"):
                    raise KeyError(f"limit function '{name}' not provided (reference limit_funcs has pnjlim, fetlim)")
                if ins.result:
                    self.builtin_raw[ins.result] = o[0]
                args = ", ".join(self.as_f(x) for x in (o[:4] if name == "pnjlim" else o[:3]))
                self.flops["transc"] += 2
                return "f", f"vjx_{name}({args})"
            return self.ty(o[0]), self.ex(o[0])
        if kind == "zero" and fl and not any(
            x in fl for x in ("noise", "collapse", "set_", "setretflag", "simparam_str", "discontinuity",
                              "display", "strobe", "write", "debug", "ddx")):
            self.warnings.append(f"Unknown MIR function '{ins.func_name}' - returning zero")
        return "f", "0.0"

    # ---- whole body
    def emit(self, indent: str = "  ", lockstep: int = 0) -> List[str]:
        fn = self.fn
        order = fn.topo_order()
        reachable = set(order)
        # pass 1: type inference + expression generation in topological order
        body: Dict[str, List[str]] = {b: [] for b in order}
        phi_assign: Dict[str, List[str]] = {b: [] for b in order}
        phis: List[Tuple[Inst, str]] = []
        for b in order:
            for ins in fn.blocks[b]:
                if ins.opcode == "phi":
                    ts = [self.ty(v) for (_p, v) in ins.phi] or ["f"]
                    t = "b" if all(x == "b" for x in ts) else self.promote(*ts)
                    self.types[ins.result] = t
                    phis.append((ins, t))
                    continue
                if ins.opcode in ("br", "jmp", "exit"):
                    continue
                t, e = self.translate(ins)
                if ins.result:
                    self.types[ins.result] = t
                    body[b].append(f"{ins.result} = {e};")
                else:
                    body[b].append(f"(void)({e});")
        for ins, t in phis:
            for pred, val in ins.phi:
                if pred in reachable:
                    if t == "b":
                        e = self.as_b(val)
                    else:
                        e = self.num(val, t)
                    phi_assign[pred].append(f"{ins.result} = {e};")
        # Lock-step barriers (lockstep = M > 0): the body is straight-line code several times the size of the
        # instruction cache, executed once per thread, so warps of a CTA that drift apart each fetch it from L2 on
        # their own.  A block through which every entry -> exit path passes (no edge of the acyclic CFG jumps over
        # it) is executed exactly once by every thread, so a CTA barrier is legal there and inside it; one every
        # ~M statements keeps the CTA's warps on the same cache lines.  The kernel then must not return early
        # (emit_eval_source clamps the instance index and predicates the stores instead).
        cut = set()
        if lockstep > 0:
            pos = {b: i for i, b in enumerate(order)}
            reach = 0
            for i, b in enumerate(order):
                if reach <= i:
                    cut.add(b)
                su = fn._succ_of(b)
                reach = max([reach] + [pos[x] for x in su] + ([len(order)] if not su else []))
        out: List[str] = []
        since = 0
        self.n_barriers = 0
        for b in order:
            out.append(f"L_{b}: ;")
            if b in cut:
                for st in body[b]:
                    if since >= lockstep:
                        out.append(indent + "__syncthreads();")
                        self.n_barriers += 1
                        since = 0
                    out.append(indent + st)
                    since += 1
            else:
                out.extend(indent + s for s in body[b])
                since += len(body[b])
            out.extend(indent + s for s in phi_assign[b])
            term = fn.blocks[b][-1] if fn.blocks[b] else None
            if term is None or term.opcode == "exit":
                out.append(indent + "goto L_exit;")
            elif term.opcode == "br":
                out.append(indent + f"if ({self.as_b(term.cond)}) goto L_{term.true_block}; else goto L_{term.false_block};")
            elif term.opcode == "jmp":
                out.append(indent + f"goto L_{term.target};")
            else:
                out.append(indent + "goto L_exit;")
        out.append("L_exit: ;")
        return out

    def decls(self, skip=()) -> List[str]:
        ctype = {"f": "double", "i": "int", "b": "bool"}
        groups: Dict[str, List[str]] = {"f": [], "i": [], "b": []}
        for v, t in self.types.items():
            if v in skip:
                continue
            groups[t].append(v)
        out = []
        for t, vs in groups.items():
            for k in range(0, len(vs), 12):
                init = "0.0" if t == "f" else ("0" if t == "i" else "false")
                out.append(f"{ctype[t]} " + ", ".join(f"{v} = {init}" for v in vs[k:k + 12]) + ";")
        return out


def _mir_var(ref: str) -> str:
    if ref and ref.startswith("mir_"):
        return "v" + ref[4:]
    return ref or ""


# ---- output slicing -------------------------------------------------------------------------------
# One thread per instance walks the whole MIR: a serial chain of ~4e4 SASS instructions for a PSP103-weight model,
# so a c6288-sized circuit (10 k instances = 0.5 warps per scheduler) is bound by that chain's latency.  The chain
# is the value path plus one derivative path per Jacobian column; the columns are independent of each other.
# With K slices the same body is instantiated K times (template parameter), slice k stores only its own group
# of outputs and the compiler's dead-code elimination removes what the group does not need.  blockIdx.y picks the
# slice (uniform per CTA), so K x more warps each walk a shorter chain.  Every output is still computed by exactly
# the same sequence of fp64 operations (no contraction, no reassociation): results are bit-identical for any K.
SLICE_WEIGHT = {"fdiv": 20, "exp": 60, "ln": 60, "log": 60, "sqrt": 25, "pow": 150, "sin": 60, "cos": 60, "tan": 80,
                "sinh": 80, "cosh": 80, "tanh": 80, "atan": 60, "atan2": 80, "hypot": 40, "asin": 80, "acos": 80,
                "asinh": 80, "acosh": 80, "atanh": 80, "clog2": 60, "call": 40}


SLICE_MIN_COST = 3000   # weighted instruction count below which a model is never sliced


def plan_slices(fn: Function, dae: dict, n_slices: int) -> List[dict]:
    """Group the outputs (Jacobian entries by column node; all residuals together) into ``n_slices`` groups with a
    small maximum backward-slice cost.  Returns [{"jac": [entry ids], "res": bool, "cost": int}]."""
    R, J = len(dae["residuals"]), len(dae["jacobian"])
    if n_slices <= 1:
        return [{"jac": list(range(J)), "res": True, "cost": 0}]
    defs: Dict[str, Inst] = {}
    for insts in fn.blocks.values():
        for ins in insts:
            if ins.result:
                defs[ins.result] = ins

    def back(vars_) -> set:
        seen, st = set(), [v for v in vars_ if v in defs]
        while st:
            v = st.pop()
            if v in seen:
                continue
            seen.add(v)
            ins = defs[v]
            for o in list(ins.operands) + [x for _, x in ins.phi]:
                if o in defs and o not in seen:
                    st.append(o)
        return seen

    def cost(sl: set) -> int:
        return sum(SLICE_WEIGHT.get(defs[v].opcode, 1) for v in sl)

    if cost(set(defs)) < SLICE_MIN_COST:   # a resistor / diode is not worth K x the CTAs
        return [{"jac": list(range(J)), "res": True, "cost": cost(set(defs))}]
    ctrl = back([ins.cond for insts in fn.blocks.values() for ins in insts if ins.opcode == "br" and ins.cond])
    items = []   # (set, jac ids, is_res)
    res_vars = [_mir_var(r[k]) for r in dae["residuals"] for k in ("resist_var", "react_var")]
    items.append((back(res_vars) | ctrl, [], True))
    cols: Dict[str, List[int]] = {}
    for j, e in enumerate(dae["jacobian"]):
        cols.setdefault(e["col_node_name"], []).append(j)
    for _c, js in cols.items():
        vs = [_mir_var(dae["jacobian"][j][k]) for j in js for k in ("resist_var", "react_var")]
        items.append((back(vs) | ctrl, js, False))
    items.sort(key=lambda it: -cost(it[0]))
    K = min(n_slices, len(items))
    bins = [{"set": set(), "jac": [], "res": False} for _ in range(K)]
    for sl, js, is_res in items:
        # the bin whose cost after taking this item is smallest
        best = min(range(K), key=lambda b: cost(bins[b]["set"] | sl))
        bins[best]["set"] |= sl
        bins[best]["jac"].extend(js)
        bins[best]["res"] = bins[best]["res"] or is_res
    out = [{"jac": sorted(b["jac"]), "res": b["res"], "cost": cost(b["set"])} for b in bins if b["jac"] or b["res"]]
    assert sorted(j for b in out for j in b["jac"]) == list(range(J)) and sum(1 for b in out if b["res"]) == 1
    return out


def default_slices(n_dev: int) -> int:
    """Slices for a model with ``n_dev`` instances: about two warps per SM sub-partition of a B200 (148 x 4) in
    total, none when the instance count alone fills the machine.  Measured at c6288 size (314 warps of instances,
    profiles/r2_eval_variants.txt): 1 slice 125 us, 3 slices 82 us, 6 slices 104 us - the shorter chain of a slice
    is paid for with more warps sharing an SM's instruction fetch.  VJX_EVAL_SLICES overrides."""
    env = os.environ.get("VJX_EVAL_SLICES")
    if env:
        return max(1, int(env))
    warps = max(1, (n_dev + 31) // 32)
    return int(max(1, min(6, (2 * 592) // warps)))



def compute_eval_layout(bundle: dict, shared_indices, varying_indices, shared_cache_indices,
                        varying_cache_indices, limit_param_map=None) -> dict:
    """Build the emitter's parameter-source tables from the host-side split
    (reference: function_builder.py:867-900 idx_mapping / cache_idx_mapping).  Voltage params
    are split out of ``varying_indices`` because the kernel gathers them from ``V``."""
    kinds = bundle["param_kinds"]
    param_src: Dict[int, tuple] = {}
    for pos, idx in enumerate(shared_indices):
        param_src[idx] = ("shared", pos)
    volt = [i for i in varying_indices if i < len(kinds) and kinds[i] == "voltage"]
    nonvolt = [i for i in varying_indices if not (i < len(kinds) and kinds[i] == "voltage")]
    for k, idx in enumerate(volt):
        param_src[idx] = ("voltage", k)
    for pos, idx in enumerate(nonvolt):
        param_src[idx] = ("device", pos)
    for idx, (kind, name) in (limit_param_map or {}).items():
        if kind == "prev_state":
            param_src[idx] = ("prev_state", int(name.split("_")[-1]))
        else:
            param_src[idx] = (kind, 0)
    cache_src: Dict[int, tuple] = {}
    for pos, idx in enumerate(shared_cache_indices):
        cache_src[idx] = ("shared_cache", pos)
    for pos, idx in enumerate(varying_cache_indices):
        cache_src[idx] = ("device_cache", pos)
    return {
        "param_src": param_src, "cache_src": cache_src,
        "n_shared": len(shared_indices), "n_volt": len(volt), "n_dev_params": len(nonvolt),
        "n_shared_cache": len(shared_cache_indices), "n_dev_cache": len(varying_cache_indices),
        "volt_param_indices": volt, "dev_param_indices": nonvolt,
    }


def default_min_blocks(n_dev: int) -> int:
    """Resident CTAs per SM asked of ptxas (``__launch_bounds__`` second argument; 0 = unconstrained).  A model
    kernel that fills the machine several times over (mul64: 2,082 CTAs) is bound by how many warps an SM
    sub-partition can switch between: 5 CTAs of 128 threads (102 registers, spills into L1) beat 2 CTAs at 255
    registers - mul64, psp103w: 2 CTAs 1.24 ms, 3: 1.27, 4: 1.17, **5: 0.99**, 6: 1.16, 8: 1.04
    (profiles/r2_eval_variants.txt).  VJX_EVAL_MINBLOCKS overrides."""
    env = os.environ.get("VJX_EVAL_MINBLOCKS")
    if env:
        return max(0, int(env))
    return 5 if (n_dev + 127) // 128 > 4 * 148 else 0


def emit_eval_source(bundle: dict, layout: dict, use_limit_functions: bool = False, n_slices: int = 1,
                     min_blocks: int = 0, lockstep: Optional[int] = None) -> Tuple[str, dict]:
    fn = parse_function("eval", bundle["mir_data"], bundle.get("str_constants"))
    dae = bundle["dae_data"]
    param_idx_to_val = {p[0]: f"v{p[1]}" for p in bundle["all_func_params"]}
    R = len(dae["residuals"])
    J = len(dae["jacobian"])
    body = _Body(fn, use_limit_functions, limit_in_expr="(a.limit_in ? a.limit_in[(size_t)d * NLIM + {idx}] : 0.0)")

    head: List[str] = []
    # ---- parameter mapping (function_builder.py:1007-1058)
    for i, p in enumerate(fn.params):
        src = layout["param_src"].get(i)
        if src is None:
            head.append(f"{p} = 0.0;")
            continue
        k, v = src
        if k == "shared":
            head.append(f"{p} = C.shared[{v}];")
        elif k == "device":
            head.append(f"{p} = a.dev_params[(size_t){v} * n + d];")
        elif k == "voltage":
            head.append(f"{p} = a.V[a.vnode1[(size_t){v} * n + d]] - a.V[a.vnode2[(size_t){v} * n + d]];")
        elif k == "prev_state":
            head.append(f"{p} = (a.limit_in ? a.limit_in[(size_t)d * NLIM + {v}] : 0.0);")
        elif k == "enable_lim":
            head.append(f"{p} = {1.0 if use_limit_functions else 0.0};")
        else:  # new_state, enable_integration
            head.append(f"{p} = 0.0;")
    # ---- cache mapping (function_builder.py:1060-1090)
    cache_map = bundle["init_mir_data"].get("cache_mapping", [])
    for ci, mp in enumerate(cache_map):
        var = param_idx_to_val.get(mp["eval_param"], f"cached_{mp['eval_param']}")
        body.types.setdefault(var, "f")
        src = layout["cache_src"].get(ci)
        if src is None:
            head.append(f"{var} = 0.0;")
        elif src[0] == "shared_cache":
            head.append(f"{var} = C.shared_cache[{src[1]}];")
        else:
            head.append(f"{var} = a.dev_cache[(size_t){src[1]} * n + d];")

    lockstep = int(os.environ.get("VJX_EVAL_LOCKSTEP", "0")) if lockstep is None else int(lockstep)
    code = body.emit(lockstep=lockstep)

    def outv(ref: str) -> str:
        v = _mir_var(ref)
        return body.as_f(v) if body.defined(v) else "0.0"

    tail: List[str] = []
    res_r = [outv(r["resist_var"]) for r in dae["residuals"]]
    res_q = [outv(r["react_var"]) for r in dae["residuals"]]
    jac_r = [outv(e["resist_var"]) for e in dae["jacobian"]]
    jac_q = [outv(e["react_var"]) for e in dae["jacobian"]]

    # ---- lim_rhs (function_builder.py:1126-1212): sum_k J[i, hi_k] * (V_lim_k - V_raw_k)
    lim_r = ["0.0"] * R
    lim_q = ["0.0"] * R
    limited = []
    if use_limit_functions:
        param_id_to_idx = {p: i for i, p in enumerate(fn.params)}
        names = bundle["param_names"]

        def hi_node(vraw: str) -> Optional[str]:
            idx = param_id_to_idx.get(vraw)
            if idx is None:
                for oi, vn in param_idx_to_val.items():
                    if vn == vraw:
                        idx = oi
                        break
            if idx is None or idx >= len(names):
                return None
            m = re.match(r"^V\((\w+)(?:,(\w+))?\)$", names[idx])
            return m.group(1) if m else None

        for li in range(body.limit_count):
            vraw = body.limit_raw.get(li)
            vlim = body.limit_store.get(li)
            if vraw is None or vlim is None or not body.defined(vraw) or not body.defined(vlim):
                continue
            hn = hi_node(vraw)
            if hn is None:
                continue
            limited.append((li, vraw, vlim, hn))
        if limited:
            jkey = {(e["row_node_name"], e["col_node_name"]): j for j, e in enumerate(dae["jacobian"])}
            for bi, (_li, vraw, vlim, _hn) in enumerate(limited):
                tail.append(f"const double lim_delta_{bi} = {body.as_f(vlim)} - {body.as_f(vraw)};")
            for i, res in enumerate(dae["residuals"]):
                tr, tq = [], []
                for bi, (_li, _vr, _vl, hn) in enumerate(limited):
                    j = jkey.get((res["node_name"], hn))
                    if j is None:
                        continue
                    e = dae["jacobian"][j]
                    rv, qv = _mir_var(e["resist_var"]), _mir_var(e["react_var"])
                    if body.defined(rv):
                        tr.append(f"{body.as_f(rv)} * lim_delta_{bi}")
                    if body.defined(qv):
                        tq.append(f"{body.as_f(qv)} * lim_delta_{bi}")
                if tr:
                    lim_r[i] = " + ".join(tr) if len(tr) == 1 else "(" * (len(tr) - 1) + tr[0] + "".join(f" + {t})" for t in tr[1:])
                if tq:
                    lim_q[i] = " + ".join(tq) if len(tq) == 1 else "(" * (len(tq) - 1) + tq[0] + "".join(f" + {t})" for t in tq[1:])
    n_lim = body.limit_count if use_limit_functions else 0
    if use_limit_functions and n_lim > 0 and not body.limit_raw:
        # inline limiting without BuiltinLimit: the reference disables limiting for the model
        # (openvaf_models.py:972-990) - signal it to the caller.
        src2, meta2 = emit_eval_source(bundle, layout, use_limit_functions=False, n_slices=n_slices, min_blocks=min_blocks, lockstep=lockstep)
        meta2["limiting_disabled"] = True
        return src2, meta2

    slices = plan_slices(fn, dae, n_slices)
    K = len(slices)
    # The destination indices are read through the non-coherent path in groups, ahead of the stores that use them:
    # emitted as "load index, test, store" per output the compiler keeps that order (a store through jsorted may
    # alias the next jdest load as far as it knows), and the tail becomes R + J sequential memory round trips -
    # 39 % of the kernel's time at mul64 size (profiles/r2_ncu_eval_kernel_mul64_psp103w.txt).
    GROUP = 16
    tail.append("const double c0 = a.ctl[VJX_CTL_C0];")

    def emit_group(items, indent="  "):
        for g0 in range(0, len(items), GROUP):
            grp = items[g0:g0 + GROUP]
            tail.append(indent + "{")
            for kind, e in grp:
                arr = "jdest" if kind == "j" else "vdest"
                tail.append(f"{indent}  const int p{kind}{e} = {'live ? ' if lockstep > 0 else ''}__ldg(a.{arr} + (size_t){e} * n + d){' : -1' if lockstep > 0 else ''};")
            for kind, e in grp:
                if kind == "j":
                    tail.append(f"{indent}  if (pj{e} >= 0) a.jsorted[pj{e}] = vjx_nan0({jac_r[e]} + c0 * {jac_q[e]});")
                else:
                    tail.append(
                        f"{indent}  if (pr{e} >= 0) {{ double2* q = reinterpret_cast<double2*>(a.vsorted + (size_t)pr{e} * 4); "
                        f"q[0] = make_double2(vjx_nan0({res_r[e]}), vjx_nan0({res_q[e]})); q[1] = make_double2(vjx_nan0({lim_r[e]}), vjx_nan0({lim_q[e]})); }}")
            tail.append(indent + "}")

    # S = -1: every output; S = k >= 0: the outputs of slice k; S = -2: the residual-side outputs only (the build at the
    # converged X that nr_solve ends with needs Q, the currents and the limit states, no Jacobian:
    # solver_factories.py:689-713) - the derivative paths fall to dead-code elimination
    for k, sl in enumerate(slices):
        if sl["jac"]:
            tail.append(f"if constexpr (S == -1 || S == {k}) {{")
            emit_group([("j", e) for e in sl["jac"]])
            tail.append("}")
        if sl["res"]:
            tail.append(f"if constexpr (S == -1 || S == -2 || S == {k}) {{")
            emit_group([("r", r) for r in range(R)])
            for li in range(n_lim):
                if li not in body.limit_store:
                    raise ValueError(f"Limit state {li} was registered (LoadLimit) but no StoreLimit found.")
                tail.append(f"  if ({'live && ' if lockstep > 0 else ''}a.limit_out) a.limit_out[(size_t)d * NLIM + {li}] = {body.as_f(body.limit_store[li])};")
            tail.append("}")

    # ---- simparams prologue
    sp_lines = []
    for k, name in enumerate(body.simparams):
        if name == "$analysis_type":
            sp_lines.append(f"const double sp{k} = (a.ctl[VJX_CTL_C0] > 0.0) ? 2.0 : 0.0;")
        elif name == "gmin":
            sp_lines.append(f"const double sp{k} = a.ctl[VJX_CTL_GMIN];")
        elif name == "iniLim":
            sp_lines.append(f"const double sp{k} = (a.ctl[VJX_CTL_ITER] == 0.0) ? 1.0 : 0.0;")
        elif name == "iteration":
            sp_lines.append(f"const double sp{k} = a.ctl[VJX_CTL_ITER];")
        else:
            sp_lines.append(f"const double sp{k} = C.simparam_defaults[{k}];")

    ns, nsc, nsp = max(1, layout["n_shared"]), max(1, layout["n_shared_cache"]), max(1, len(body.simparams))
    src: List[str] = (["#define VJX_MATH_FN __noinline__"] if os.environ.get("VJX_EVAL_NOINLINE_MATH") == "1" else []) + [PRELUDE]
    src.append(f"#define NLIM {max(1, n_lim)}")
    src.append(f"struct VjxConst {{ double shared[{ns}]; double shared_cache[{nsc}]; double simparam_defaults[{nsp}]; }};")
    src.append("__constant__ VjxConst C;")
    # the body once, as a template: S < 0 stores every output, S = k only the outputs of slice k
    src.append("template <int S> __device__ __forceinline__ void vjx_eval_body(const VjxEvalArgs& a, const int d, const int n, const bool live) {")
    src.extend("  " + s for s in body.decls())
    src.extend("  " + s for s in sp_lines)
    src.extend("  " + s for s in head)
    src.extend("  " + s for s in code)
    src.extend("  " + s for s in tail)
    src.append("}")
    tb = int(os.environ.get("VJX_EVAL_BLOCK", "128"))
    lb = f"__launch_bounds__({tb}, {int(min_blocks)})" if min_blocks else f"__launch_bounds__({tb})"
    src.append(f"extern \"C\" __global__ void {lb} vjx_eval_kernel(VjxEvalArgs a) {{")
    entry = (["  const int d0 = blockIdx.x * blockDim.x + threadIdx.x;", "  const int n = a.n_dev;",
              "  const bool live = d0 < n;", "  const int d = live ? d0 : n - 1;   // no early return: the body holds CTA barriers"]
             if lockstep > 0 else
             ["  const int d = blockIdx.x * blockDim.x + threadIdx.x;", "  const int n = a.n_dev;", "  if (d >= n) return;",
              "  const bool live = true;"])
    src.extend(entry)
    if K == 1:
        src.append("  vjx_eval_body<-1>(a, d, n, live);")
    else:
        src.append("  switch (blockIdx.y) {")
        for k in range(K):
            src.append(f"    case {k}: vjx_eval_body<{k}>(a, d, n, live); break;")
        src.append("  }")
    src.append("}")
    src.append(f"extern \"C\" __global__ void {lb} vjx_eval_kernel_res(VjxEvalArgs a) {{")
    src.extend(entry)
    src.append("  vjx_eval_body<-2>(a, d, n, live);")
    src.append("}")
    src.append(f"""
extern "C" int vjx_model_eval(const VjxEvalArgs* a, cudaStream_t s) {{
  if (a->n_dev <= 0) return 0;
  if (a->jsorted == nullptr)   // residual-side outputs only
    vjx_eval_kernel_res<<<dim3((a->n_dev + {tb - 1}) / {tb}, 1), {tb}, 0, s>>>(*a);
  else
    vjx_eval_kernel<<<dim3((a->n_dev + {tb - 1}) / {tb}, {K}), {tb}, 0, s>>>(*a);
  return (int)cudaGetLastError();
}}
extern "C" int vjx_model_set_const(const double* sh, int ns, const double* sc, int nsc,
                                   const double* sp, int nsp, cudaStream_t s) {{
  if (ns > {ns} || nsc > {nsc} || nsp > {nsp}) return -1;
  cudaError_t e = cudaSuccess;
  if (ns > 0) e = cudaMemcpyToSymbolAsync(C, sh, sizeof(double) * ns, offsetof(VjxConst, shared), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && nsc > 0) e = cudaMemcpyToSymbolAsync(C, sc, sizeof(double) * nsc, offsetof(VjxConst, shared_cache), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && nsp > 0) e = cudaMemcpyToSymbolAsync(C, sp, sizeof(double) * nsp, offsetof(VjxConst, simparam_defaults), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  return (int)e;
}}
extern "C" int vjx_model_abi_version() {{ return 1; }}
""")
    meta = {
        "n_residuals": R, "n_jacobian": J, "n_limit_states": n_lim,
        "simparams_used": list(body.simparams),
        "simparam_indices": {n: i for i, n in enumerate(body.simparams)},
        "simparam_defaults": [SIMPARAM_DEFAULTS.get(n, 0.0) for n in body.simparams],
        "limit_metadata": {"limit_count": n_lim, "raw_operands": dict(body.limit_raw),
                           "store_operands": dict(body.limit_store), "limited_branches": limited},
        "flops_static": dict(body.flops),
        "n_slices": K, "slice_costs": [sl["cost"] for sl in slices], "lockstep": lockstep, "n_barriers": getattr(body, "n_barriers", 0),
        "warnings": list(body.warnings),
        "use_limit_functions": bool(use_limit_functions),
        "node_names": [r["node_name"] for r in dae["residuals"]],
        "jacobian_keys": [(e["row_node_name"], e["col_node_name"]) for e in dae["jacobian"]],
    }
    return "\n".join(src), meta


def emit_init_source(bundle: dict) -> Tuple[str, dict]:
    """Init kernel: params -> cache columns + collapse decisions
    (reference: function_builder.py:520-585 build_simple / :605-692 build_split)."""
    fn = parse_function("init", bundle["init_mir_data"], bundle.get("str_constants"))
    body = _Body(fn, False)
    head = []
    for i, p in enumerate(fn.params):
        head.append(f"{p} = a.vals[a.col_off[{i}] + (size_t)d * a.col_stride[{i}]];")
    cache_map = bundle["init_mir_data"].get("cache_mapping", [])
    for mp in cache_map:  # pre-initialise cache vars (function_builder.py:736-757)
        pass
    code = body.emit()
    tail = []
    for k, mp in enumerate(cache_map):
        v = mp["init_value"]
        e = body.as_f(v) if body.defined(v) else "0.0"
        tail.append(f"a.cache[(size_t){k} * n + d] = {e};")
    outs = list(bundle.get("collapse_decision_outputs", []))
    for k, (_pair, vn) in enumerate(outs):
        neg = vn.startswith("!")
        v = vn[1:] if neg else vn
        if body.defined(v):
            e = f"(!{body.as_b(v)})" if neg else body.as_f(v)
            if neg:
                e = f"({e} ? 1.0f : 0.0f)"
            else:
                e = f"((float)({e}))"
        else:
            e = "1.0f" if neg else "0.0f"
        tail.append(f"if (a.collapse) a.collapse[(size_t){k} * n + d] = {e};")
    src = [PRELUDE]
    src.append("extern \"C\" __global__ void __launch_bounds__(128) vjx_init_kernel(VjxInitArgs a) {")
    src.append("  const int d = blockIdx.x * blockDim.x + threadIdx.x;")
    src.append("  const int n = a.n_dev;")
    src.append("  if (d >= n) return;")
    src.extend("  " + s for s in body.decls())
    src.extend("  " + s for s in head)
    src.extend("  " + s for s in code)
    src.extend("  " + s for s in tail)
    src.append("}")
    src.append("""
extern "C" int vjx_model_init(const VjxInitArgs* a, cudaStream_t s) {
  if (a->n_dev <= 0) return 0;
  vjx_init_kernel<<<(a->n_dev + 127) / 128, 128, 0, s>>>(*a);
  return (int)cudaGetLastError();
}
extern "C" int vjx_model_abi_version() { return 1; }
""")
    meta = {"cache_size": len(cache_map), "n_collapse_outputs": len(outs),
            "n_init_params": len(fn.params), "flops_static": dict(body.flops)}
    return "\n".join(src), meta


def model_fmad() -> bool:
    """FMA contraction inside model kernels. Off by default so that mul/add chains round
    exactly like the oracle's numpy arithmetic (and XLA CPU, which does not contract)."""
    return os.environ.get("VJX_MODEL_FMAD", "0") == "1"


_COMPILE_LOCKS: Dict[str, "threading.Lock"] = {}
_COMPILE_LOCKS_MU = threading.Lock()


def compile_plugin(source: str, tag: str = "model") -> str:
    """Compile emitted CUDA-C into an in-tree shared object, cached by a SHA-256 of
    source + flags (the reference keys its exec cache the same way, openvaf_jax/cache.py:22-106).
    Safe against concurrent callers: threads of one process (``run_corners(concurrent=k)`` emits the same
    source from every engine on a cold cache) serialise on a per-hash lock, and every compilation writes to
    its own temporary files (``mkstemp``) before the atomic rename, so other processes never see a
    half-written plugin."""
    flags = list(NVCC_FLAGS) + ["-fmad=" + ("true" if model_fmad() else "false")]
    h = hashlib.sha256((source + "\0" + " ".join(flags)).encode()).hexdigest()[:20]
    os.makedirs(PLUGIN_DIR, exist_ok=True)
    so = os.path.join(PLUGIN_DIR, f"{tag}_{h}.so")
    with _COMPILE_LOCKS_MU:
        lock = _COMPILE_LOCKS.setdefault(h, threading.Lock())
    with lock:
        if os.path.exists(so):
            return so
        fd, cu_tmp = tempfile.mkstemp(prefix=f"{tag}_{h}.", suffix=".cu", dir=PLUGIN_DIR)
        with os.fdopen(fd, "w") as f:
            f.write(source)
        fd, so_tmp = tempfile.mkstemp(prefix=f"{tag}_{h}.", suffix=".so.tmp", dir=PLUGIN_DIR)
        os.close(fd)
        try:
            r = subprocess.run(["nvcc"] + flags + ["-o", so_tmp, cu_tmp], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {tag}_{h}:\n{r.stderr[-4000:]}")
            os.chmod(so_tmp, 0o755)
            os.replace(cu_tmp, os.path.join(PLUGIN_DIR, f"{tag}_{h}.cu"))   # kept for -lineinfo / inspection
            os.replace(so_tmp, so)
        finally:
            for t in (cu_tmp, so_tmp):
                if os.path.exists(t):
                    os.unlink(t)
    return so

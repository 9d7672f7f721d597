"""ORACLE (test infrastructure, not product): vectorised NumPy interpreter of OpenVAF MIR.

Restates the semantics of the reference's MIR -> JAX translation:
  * opcode semantics: openvaf_jax/codegen/instruction.py:156-468
    (safe divide openvaf_jax/openvaf_ast/expressions.py:349-374, clamped sqrt :376-381,
    clamped ln/log10/clog2 :383-410, ficast = int32(floor(x)) :428-432)
  * callbacks: openvaf_jax/codegen/instruction.py:469-859
  * parameter / cache mapping and outputs: openvaf_jax/codegen/function_builder.py:1007-1124
  * lim_rhs: function_builder.py:1126-1212; limit_state_out :1360-1396
  * init: function_builder.py:520-604, 736-811

The reference executes every block and turns PHIs into ``jnp.where``; here every block is
executed over all instances too, but PHIs select on the *actual incoming edge* masks, which is
the MIR's own (branching) semantics.  Values that are only defined in a block an instance did
not execute read as 0.0 (the reference pre-initialises outputs to 0.0,
function_builder.py:1398-1448).  MIR loops are not supported (SURVEY.md §8a footnote).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this package.
"""

from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

import numpy as np

ANALYSIS_TYPE_MAP = {"dc": 0, "static": 0, "ac": 1, "tran": 2, "transient": 2, "noise": 3, "nodeset": 4}

SIMPARAM_DEFAULTS = {  # reference: vajax/__init__.py:191-208
    "$analysis_type": 0.0, "gmin": 1e-12, "abstol": 1e-12, "vntol": 1e-6, "reltol": 1e-3,
    "tnom": 27.0, "scale": 1.0, "shrink": 0.0, "imax": 1.0, "$abstime": 0.0, "$mfactor": 1.0,
    "iniLim": 0.0, "iteration": 1.0, "gdev": 0.0, "sourceScaleFactor": 1.0,
}


def pnjlim(vnew, vold, vt=0.026, vcrit=0.6):
    """reference: vajax/analysis/limiting.py:22-81"""
    with np.errstate(all="ignore"):
        delta_v = vnew - vold
        arg = delta_v / vt
        fwd = (vnew > vcrit) & (np.abs(delta_v) > 2 * vt)
        lim_pos = vold + vt * np.log(np.maximum(1.0 + arg, 1e-30))
        lim_neg = vold - vt * np.log(np.maximum(1.0 - arg, 1e-30))
        lim_forward = np.where(arg >= 0, lim_pos, lim_neg)
        lim_reverse = np.where(vnew > 0, vt * np.log(np.maximum(vnew / vt, 1e-30)), vcrit)
        lim_fwd = np.where(vold > 0, lim_forward, lim_reverse)
        neg_clamp = np.where(vold > 0, -vold - 1.0, 2.0 * vold - 1.0)
        lim_negv = np.maximum(vnew, neg_clamp)
        return np.where(fwd, lim_fwd, np.where(vnew < 0, lim_negv, vnew))


def fetlim(vnew, vold, vto=0.5):
    """reference: vajax/analysis/limiting.py:84-107"""
    max_step = np.where(vold >= vto, 2 * np.abs(vold - vto) + 2, 0.5)
    delta = vnew - vold
    return vold + np.clip(delta, -max_step, max_step)


LIMIT_FUNCS = {"pnjlim": pnjlim, "fetlim": fetlim}


def _resolve_func(inst: dict, decls: dict) -> str:
    fn = inst.get("func_name")
    if not fn:
        ref = inst.get("func_ref")
        if ref in decls:
            fn = decls[ref].get("name")
    return fn or ""


class _Interp:
    def __init__(self, mir: dict, str_constants: Dict[str, str], n: int, use_limit: bool,
                 simparams: Dict[str, object], limit_state_in: Optional[np.ndarray]):
        self.mir = mir
        self.n = n
        self.sconst = dict(str_constants or {})
        self.use_limit = use_limit
        self.simparams = simparams
        self.limit_state_in = limit_state_in
        self.env: Dict[str, np.ndarray] = {}
        self.defblock: Dict[str, str] = {}
        self.active: Dict[str, np.ndarray] = {}
        self.edge: Dict[Tuple[str, str], np.ndarray] = {}
        self.limit_reg: Dict[int, int] = {}
        self.limit_store: Dict[int, str] = {}
        self.builtin_raw: Dict[str, str] = {}
        self.limit_raw: Dict[int, str] = {}
        self.simparams_used: List[str] = ["$analysis_type"]
        self.flops = np.zeros(n, dtype=np.int64)  # executed-path fp op count per instance
        self.transc = np.zeros(n, dtype=np.int64)
        for k, v in {"v3": 0.0, "v6": 1.0, "v7": -1.0}.items():
            self.env[k] = np.float64(v)
        for k, v in {"v4": 0, "v5": 1}.items():
            self.env[k] = np.int32(v)
        self.env["v1"] = np.bool_(False)
        self.env["v2"] = np.bool_(True)
        for k, v in mir.get("constants", {}).items():
            self.env[k] = np.float64(v)
        for k, v in mir.get("int_constants", {}).items():
            self.env[k] = np.int32(v)
        for k, v in mir.get("bool_constants", {}).items():
            self.env[k] = np.bool_(v)

    def get(self, v: str):
        return self.env.get(v, np.float64(0.0))  # undefined -> 0.0 (context.py:278-280)

    @staticmethod
    def _num(x):
        x = np.asarray(x)
        return x.astype(np.int32) if x.dtype == np.bool_ else x

    def _limidx(self, fn: str) -> int:
        m = re.search(r"LimState\((\d+)\)", fn)
        if m:
            return int(m.group(1))
        m = re.search(r"(\d+)", fn)
        return int(m.group(1)) if m else 0

    def _reg(self, idx: int) -> int:
        if idx not in self.limit_reg:
            self.limit_reg[idx] = len(self.limit_reg)
        return self.limit_reg[idx]

    def call(self, inst: dict, fn: str):
        fl = fn.lower()
        o = inst.get("operands", [])
        zero = np.float64(0.0)
        if not fl:
            return zero
        if "simparam" in fl:
            if o and o[0] in self.sconst:
                name = self.sconst[o[0]]
                if name not in self.simparams_used:
                    self.simparams_used.append(name)
                return np.asarray(self.simparams.get(name, SIMPARAM_DEFAULTS.get(name, 0.0)), dtype=np.float64)
            return zero
        if "discontinuity" in fl:
            return zero
        if any(x in fl for x in ("display", "strobe", "write", "debug")):
            return zero
        if "analysis" in fl:
            if o and o[0] in self.sconst:
                code = ANALYSIS_TYPE_MAP.get(self.sconst[o[0]].lower(), 0)
                return np.asarray(self.simparams.get("$analysis_type", 0.0)) == code
            return np.bool_(False)
        if "ddt" in fl:
            return self.get(o[0]) if o else zero
        if "ddx" in fl:
            return zero
        if any(x in fl for x in ("white_noise", "flicker_noise", "noise_table")):
            return zero
        if fl == "abs":
            return np.abs(self.get(o[0])) if o else zero
        if fl == "min" and len(o) >= 2:
            return np.minimum(self.get(o[0]), self.get(o[1]))
        if fl == "max" and len(o) >= 2:
            return np.maximum(self.get(o[0]), self.get(o[1]))
        if "storelimit" in fl:
            if self.use_limit:
                ai = self._reg(self._limidx(fn))
                if not o:
                    return zero
                self.limit_store[ai] = o[0]
                raw = self.builtin_raw.get(o[0])
                if raw:
                    self.limit_raw[ai] = raw
                return self.get(o[0])
            return self.get(o[0]) if o else zero
        if "loadlimit" in fl:
            if self.use_limit:
                ai = self._reg(self._limidx(fn))
                return self.limit_state_in[:, ai]
            return self.get(o[0]) if o else zero
        if "$limit" in fl or "builtinlimit" in fl:
            if not o:
                return zero
            if self.use_limit:
                name = "pnjlim"
                for known in ("pnjlim", "fetlim", "limexp", "limvds"):
                    if known in fl:
                        name = known
                        break
                if inst.get("result"):
                    self.builtin_raw[inst["result"]] = o[0]
                args = [np.asarray(self.get(x), dtype=np.float64) for x in o]
                self.transc += 2
                return LIMIT_FUNCS[name](*args)
            return self.get(o[0])
        return zero

    def run(self):
        mir = self.mir
        decls = dict(mir.get("function_decls", {}))
        by_block: Dict[str, List[dict]] = {b: [] for b in mir.get("blocks", {})}
        for ins in mir.get("instructions", []):
            by_block.setdefault(ins["block"], []).append(ins)
        # topological order (DFS reverse post-order)
        def succ(b):
            if not by_block[b]:
                return []
            t = by_block[b][-1]
            op = t["opcode"].lower()
            if op == "br":
                return [t["true_block"], t["false_block"]]
            if op == "jmp":
                return [t.get("destination") or t.get("target_block")]
            return []
        entry = "block0" if "block0" in by_block else next(iter(by_block))
        order, seen = [], {}
        stack = [(entry, iter(succ(entry)))]
        seen[entry] = 1
        while stack:
            node, it = stack[-1]
            adv = False
            for s in it:
                if s not in seen:
                    seen[s] = 1
                    stack.append((s, iter(succ(s))))
                    adv = True
                    break
                if seen[s] == 1:
                    raise NotImplementedError("MIR loops are not supported")
            if not adv:
                seen[node] = 2
                order.append(node)
                stack.pop()
        order.reverse()
        n = self.n
        all_true = np.ones(n, dtype=bool)
        with np.errstate(all="ignore"):
            for b in order:
                if b == entry:
                    act = all_true
                else:
                    act = np.zeros(n, dtype=bool)
                    for (p, q), m in self.edge.items():
                        if q == b:
                            act = act | m
                self.active[b] = act
                for ins in by_block[b]:
                    op = ins["opcode"].lower()
                    res = ins.get("result")
                    if op == "phi":
                        val = None
                        for e in ins.get("phi_operands", []):
                            m = self.edge.get((e["block"], b))
                            if m is None:
                                continue
                            x = self.get(e["value"])
                            if val is None:
                                val = np.where(m, x, np.zeros((), dtype=np.asarray(x).dtype))
                            else:
                                val = np.where(m, x, val)
                        self.env[res] = val if val is not None else np.float64(0.0)
                        self.defblock[res] = b
                        continue
                    if op == "br":
                        c = np.asarray(self.get(ins["condition"]))
                        c = (c != 0) if c.dtype != np.bool_ else c
                        c = np.broadcast_to(c, (n,))
                        t, f = ins["true_block"], ins["false_block"]
                        self.edge[(b, t)] = self.edge.get((b, t), np.zeros(n, bool)) | (act & c)
                        self.edge[(b, f)] = self.edge.get((b, f), np.zeros(n, bool)) | (act & ~c)
                        continue
                    if op == "jmp":
                        d = ins.get("destination") or ins.get("target_block")
                        self.edge[(b, d)] = self.edge.get((b, d), np.zeros(n, bool)) | act
                        continue
                    if op == "exit":
                        continue
                    v = self.exec(ins, op, decls, act)
                    if res:
                        self.env[res] = v
                        self.defblock[res] = b

    def exec(self, ins, op, decls, act):
        o = ins.get("operands", [])
        g = self.get
        N = self._num
        if op in ("fadd", "fsub", "fmul", "iadd", "isub", "imul"):
            a, b = N(g(o[0])), N(g(o[1]))
            if a.dtype.kind == "f" or b.dtype.kind == "f":
                self.flops += act
            return a + b if op[1:] == "add" else (a - b if op[1:] == "sub" else a * b)
        if op == "fdiv":
            a, b = np.asarray(g(o[0]), dtype=np.float64), np.asarray(g(o[1]), dtype=np.float64)
            self.flops += act
            return np.where(b == 0.0, 0.0, a / np.where(b == 0.0, 1.0, b))
        if op in ("frem", "irem"):
            a, b = N(g(o[0])), N(g(o[1]))
            return np.mod(a, b)
        if op == "idiv":
            a, b = N(g(o[0])), N(g(o[1]))
            return np.floor_divide(a, b)
        if op in ("iand", "ior", "ixor", "ishl", "ishr"):
            a, b = N(g(o[0])).astype(np.int32), N(g(o[1])).astype(np.int32)
            return {"iand": a & b, "ior": a | b, "ixor": a ^ b, "ishl": a << b, "ishr": a >> b}[op]
        if op in ("feq", "ieq", "beq", "seq"):
            return np.asarray(g(o[0])) == np.asarray(g(o[1]))
        if op in ("fne", "ine", "bne", "sne"):
            return np.asarray(g(o[0])) != np.asarray(g(o[1]))
        if op in ("flt", "ilt"):
            return N(g(o[0])) < N(g(o[1]))
        if op in ("fgt", "igt"):
            return N(g(o[0])) > N(g(o[1]))
        if op in ("fle", "ile"):
            return N(g(o[0])) <= N(g(o[1]))
        if op in ("fge", "ige"):
            return N(g(o[0])) >= N(g(o[1]))
        if op in ("fneg", "ineg"):
            a = N(g(o[0]))
            if a.dtype.kind == "f":
                self.flops += act
            return -a
        if op == "bnot":
            a = np.asarray(g(o[0]))
            return ~a if a.dtype == np.bool_ else (a == 0)
        if op == "inot":
            return ~N(g(o[0])).astype(np.int32)
        f64 = lambda x: np.asarray(g(x), dtype=np.float64)  # noqa: E731
        if op == "sqrt":
            self.transc += act
            return np.sqrt(np.maximum(f64(o[0]), 0.0))
        if op == "ln":
            self.transc += act
            return np.log(np.maximum(f64(o[0]), 1e-300))
        if op == "log":
            self.transc += act
            return np.log10(np.maximum(f64(o[0]), 1e-300))
        if op == "clog2":
            self.transc += act
            return np.ceil(np.log2(np.maximum(f64(o[0]), 1e-300)))
        un = {"exp": np.exp, "floor": np.floor, "ceil": np.ceil, "sin": np.sin, "cos": np.cos, "tan": np.tan,
              "sinh": np.sinh, "cosh": np.cosh, "tanh": np.tanh, "abs": np.abs, "fabs": np.abs,
              "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan, "asinh": np.arcsinh,
              "acosh": np.arccosh, "atanh": np.arctanh}
        if op in un and not (op == "hypot"):
            if op not in ("floor", "ceil", "abs", "fabs"):
                self.transc += act
            return un[op](f64(o[0]))
        bi = {"pow": np.power, "atan2": np.arctan2, "hypot": np.hypot, "fmin": np.minimum, "fmax": np.maximum}
        if op in bi and len(o) >= 2:
            if op in ("pow", "atan2", "hypot"):
                self.transc += act
            return bi[op](f64(o[0]), f64(o[1]))
        if op == "ifcast":
            return f64(o[0])
        if op == "ficast":
            return np.floor(f64(o[0])).astype(np.int32)
        if op == "fbcast":
            return f64(o[0]) != 0.0
        if op == "ibcast":
            return N(g(o[0])) != 0
        if op == "bfcast":
            a = np.asarray(g(o[0]))
            a = a if a.dtype == np.bool_ else (a != 0)
            return np.where(a, 1.0, 0.0)
        if op == "bicast":
            return N(g(o[0])).astype(np.int32)
        if op in ("optbarrier", "copy"):
            return g(o[0]) if o else np.float64(0.0)
        if op == "select":
            if len(o) < 3:
                return np.float64(0.0)
            c = np.asarray(g(o[0]))
            c = c if c.dtype == np.bool_ else (c != 0)
            a, b = np.asarray(g(o[1])), np.asarray(g(o[2]))
            if not (a.dtype == np.bool_ and b.dtype == np.bool_):
                a, b = N(a), N(b)
            return np.where(c, a, b)
        if op == "call":
            return self.call(ins, _resolve_func(ins, decls))
        raise ValueError(f"Unknown opcode: {op}")

    def out(self, v: str) -> np.ndarray:
        """Value of an output reference; 0.0 where its defining block was not executed."""
        if v not in self.env:
            return np.zeros(self.n)
        x = np.asarray(self.env[v])
        x = x.astype(np.float64) if x.dtype != np.float64 else x
        x = np.broadcast_to(x, (self.n,))
        b = self.defblock.get(v)
        if b is not None:
            x = np.where(self.active[b], x, 0.0)
        return np.array(x, dtype=np.float64)


def _mir_var(ref: str) -> str:
    return "v" + ref[4:] if ref and ref.startswith("mir_") else (ref or "")


def eval_model(bundle: dict, param_values: Dict[int, object], cache_values: np.ndarray, n: int,
               simparams: Dict[str, object], limit_state_in: Optional[np.ndarray] = None,
               use_limit: bool = False, limit_param_map: Optional[Dict[int, Tuple[str, str]]] = None) -> dict:
    """Evaluate ``n`` instances.  ``param_values[i]`` (scalar or (n,) array) is the value of eval
    parameter ``i``; parameters absent from the dict read 0 (function_builder.py:1052-1054).
    ``cache_values`` is (n, cache_size) in cache_mapping order."""
    mir = bundle["mir_data"]
    dae = bundle["dae_data"]
    it = _Interp(mir, bundle.get("str_constants", {}), n, use_limit, simparams, limit_state_in)
    lpm = limit_param_map or {}
    for i, p in enumerate(mir["params"]):
        if i in lpm:
            kind, name = lpm[i]
            if kind == "prev_state":
                idx = int(name.split("_")[-1])
                # jnp indexing clamps an out-of-range static index (the dummy state is (n, 1) when limiting is off)
                val = limit_state_in[:, min(idx, limit_state_in.shape[1] - 1)] if limit_state_in is not None else 0.0
            elif kind == "enable_lim":
                val = 1.0 if use_limit else 0.0
            else:
                val = 0.0
        else:
            val = param_values.get(i, 0.0)
        it.env[p] = np.asarray(val, dtype=np.float64)
    pidx2val = {p[0]: f"v{p[1]}" for p in bundle["all_func_params"]}
    for ci, mp in enumerate(bundle["init_mir_data"].get("cache_mapping", [])):
        var = pidx2val.get(mp["eval_param"], f"cached_{mp['eval_param']}")
        it.env[var] = np.asarray(cache_values[:, ci], dtype=np.float64) if cache_values.shape[1] > ci else np.float64(0.0)
    it.run()
    global _LAST
    _LAST = it   # (kept for the bisection helper in tests/golden)
    R, J = len(dae["residuals"]), len(dae["jacobian"])
    res_r = np.stack([it.out(_mir_var(r["resist_var"])) for r in dae["residuals"]], axis=1) if R else np.zeros((n, 0))
    res_q = np.stack([it.out(_mir_var(r["react_var"])) for r in dae["residuals"]], axis=1) if R else np.zeros((n, 0))
    jac_r = np.stack([it.out(_mir_var(e["resist_var"])) for e in dae["jacobian"]], axis=1) if J else np.zeros((n, 0))
    jac_q = np.stack([it.out(_mir_var(e["react_var"])) for e in dae["jacobian"]], axis=1) if J else np.zeros((n, 0))
    lim_r = np.zeros((n, R))
    lim_q = np.zeros((n, R))
    n_lim = len(it.limit_reg) if use_limit else 0
    limit_out = np.zeros((n, max(n_lim, 0)))
    if use_limit and n_lim:
        names = bundle["param_names"]
        pid2idx = {p: i for i, p in enumerate(mir["params"])}

        def hi_node(vraw):
            idx = pid2idx.get(vraw)
            if idx is None:
                for oi, vn in pidx2val.items():
                    if vn == vraw:
                        idx = oi
                        break
            if idx is None or idx >= len(names):
                return None
            m = re.match(r"^V\((\w+)(?:,(\w+))?\)$", names[idx])
            return m.group(1) if m else None

        jkey = {(e["row_node_name"], e["col_node_name"]): j for j, e in enumerate(dae["jacobian"])}
        for li in range(n_lim):
            vraw, vlim = it.limit_raw.get(li), it.limit_store.get(li)
            if vlim is not None:
                limit_out[:, li] = it.out(vlim)
            if vraw is None or vlim is None:
                continue
            hn = hi_node(vraw)
            if hn is None:
                continue
            delta = it.out(vlim) - it.out(vraw)
            for i, res in enumerate(dae["residuals"]):
                j = jkey.get((res["node_name"], hn))
                if j is None:
                    continue
                e = dae["jacobian"][j]
                rv, qv = _mir_var(e["resist_var"]), _mir_var(e["react_var"])
                if rv in it.env:
                    lim_r[:, i] = lim_r[:, i] + it.out(rv) * delta
                if qv in it.env:
                    lim_q[:, i] = lim_q[:, i] + it.out(qv) * delta
    return {
        "res_resist": res_r, "res_react": res_q, "jac_resist": jac_r, "jac_react": jac_q,
        "lim_rhs_resist": lim_r, "lim_rhs_react": lim_q, "limit_state_out": limit_out,
        "n_limit_states": n_lim, "limit_traced": bool(it.limit_raw), "simparams_used": it.simparams_used,
        "flops_executed": it.flops, "transc_executed": it.transc,
    }


def init_model(bundle: dict, init_values: List[object], n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Run the init MIR: returns (cache (n, cache_size), collapse_decisions (n, n_outputs) f32)."""
    mir = bundle["init_mir_data"]
    it = _Interp(mir, bundle.get("str_constants", {}), n, False, {}, None)
    for i, p in enumerate(mir["params"]):
        it.env[p] = np.asarray(init_values[i] if i < len(init_values) else 0.0, dtype=np.float64)
    it.run()
    cm = mir.get("cache_mapping", [])
    cache = np.stack([it.out(m["init_value"]) for m in cm], axis=1) if cm else np.zeros((n, 0))
    outs = list(bundle.get("collapse_decision_outputs", []))
    col = np.zeros((n, len(outs)), dtype=np.float32)
    for k, (_pair, vn) in enumerate(outs):
        neg = vn.startswith("!")
        v = vn[1:] if neg else vn
        if v in it.env:
            x = np.broadcast_to(np.asarray(it.env[v]), (n,))
            if neg:
                x = ~x if x.dtype == np.bool_ else (x == 0)
            col[:, k] = x.astype(np.float32)
        else:
            col[:, k] = 1.0 if neg else 0.0
    return cache, col

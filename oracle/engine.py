"""ORACLE (test infrastructure, not product): NumPy/SciPy fp64 restatement of VAJAX's transient
Newton-Raphson hot path.  Every function cites the reference lines it follows.

Pinned by: tests/test_oracle_known_answers.py (the reference's own known-answer tests:
tests/test_fori_loop_nr.py, tests/test_integration_methods.py, tests/test_adaptive_timestep.py)
and tests/golden/*.json produced by running the reference's own modules in the build
container (tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (vajax_b200/) never does.
"""

from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from . import mir_interp
from .mir_interp import SIMPARAM_DEFAULTS

U32_MAX = 4294967295


# --------------------------------------------------------------------------------------
# Sources (reference: vajax/analysis/sources.py:26-187, build_source_fn :190-296)
# --------------------------------------------------------------------------------------
def _num(x, default=0.0):
    return float(default if x is None else x)


def make_source(params: dict) -> Callable[[float], float]:
    st = str(params.get("type", "dc")).lower()
    g = params.get
    if st in ("dc", "0", "0.0", ""):
        dc = g("dc", 0)
        return lambda t, v=dc: float(v)
    if st == "pulse":
        v0, v1 = _num(g("val0", 0)), _num(g("val1", 1))
        r, f = _num(g("rise", 1e-9)), _num(g("fall", 1e-9))
        w, p, d = _num(g("width", 1e-6)), _num(g("period", 2e-6)), _num(g("delay", 0))

        def pulse(t):
            tp = np.mod(t - d, p)
            rising = v0 + (v1 - v0) * tp / r
            falling = v1 - (v1 - v0) * (tp - r - w) / f
            if t < d:
                return v0
            if tp < r:
                return float(rising)
            if tp < r + w:
                return v1
            if tp < r + w + f:
                return float(falling)
            return v0

        return pulse
    if st == "sine":
        dc, a = _num(g("sinedc", 0)), _num(g("ampl", 1))
        fr, ph = _num(g("freq", 1e3)), _num(g("phase", 0))
        return lambda t: float(dc + a * np.sin(2 * np.pi * fr * t + ph))
    if st == "pwl":
        wave = g("wave", [0, 0])
        if isinstance(wave, str):
            import ast

            wave = ast.literal_eval(wave)
        wa = np.asarray(wave, dtype=np.float64)
        times = wa[0::2] * _num(g("stretch", 1))
        vals = wa[1::2] * _num(g("scale", 1)) + _num(g("offset", 0))
        per = _num(g("pwlperiod", 0))

        def pwl(t):
            if per > 0:
                t = np.mod(t, per)
            return float(np.interp(t, times, vals))

        return pwl
    if st == "exp":
        v0, v1, d = _num(g("val0", 0)), _num(g("val1", 1)), _num(g("delay", 0))
        td, t1, t2 = _num(g("td2", 1e-9)), _num(g("tau1", 1e-10)), _num(g("tau2", 1e-10))

        def expf(t):
            tf = d + td
            rising = v0 + (v1 - v0) * (1 - np.exp(-(t - d) / t1))
            raf = v0 + (v1 - v0) * (1 - np.exp(-td / t1))
            falling = raf + (v0 - raf) * (1 - np.exp(-(t - tf) / t2))
            return float(v0 if t < d else (rising if t < tf else falling))

        return expf
    if st in ("am", "fm"):
        dc, a, fr = _num(g("sinedc", 0)), _num(g("ampl", 1)), _num(g("freq", 1e3))
        ph = _num(g("phase", g("sinephase", 0)))
        d, mf = _num(g("delay", 0)), _num(g("modfreq", 1e3))
        mph, mi = _num(g("modphase", 0)), _num(g("modindex", 0.5))
        if st == "am":
            def am(t):
                initial = dc + a * np.sin(ph * np.pi / 180) * (1 + mi * np.sin(mph * np.pi / 180))
                carrier = np.sin(2 * np.pi * fr * (t - d) + ph * np.pi / 180)
                mod = 1 + mi * np.sin(2 * np.pi * mf * (t - d) + mph * np.pi / 180)
                return float(initial if t < d else dc + a * carrier * mod)

            return am

        def fm(t):
            initial = dc + a * np.sin(ph * np.pi / 180 + mi * np.sin(mph * np.pi / 180))
            cph = 2 * np.pi * fr * (t - d) + ph * np.pi / 180
            mc = mi * np.sin(2 * np.pi * mf * (t - d) + mph * np.pi / 180)
            return float(initial if t < d else dc + a * np.sin(cph + mc))

        return fm
    raise ValueError(f"Unknown source type '{st}'")


# --------------------------------------------------------------------------------------
# Integration coefficients (reference: vajax/analysis/integration.py:88-137 and the in-loop
# forms in vajax/analysis/transient/full_mna.py:1851-1891)
# --------------------------------------------------------------------------------------
def compute_coefficients(method: str, dt: float) -> dict:
    inv_dt = 1.0 / dt
    if method == "be":
        return dict(c0=inv_dt, c1=-inv_dt, c2=0.0, d1=0.0, error_coeff=0.5)
    if method == "trap":
        return dict(c0=2.0 * inv_dt, c1=-2.0 * inv_dt, c2=0.0, d1=-1.0, error_coeff=1.0 / 12.0)
    if method in ("gear2", "bdf2"):
        return dict(c0=1.5 * inv_dt, c1=-2.0 * inv_dt, c2=0.5 * inv_dt, d1=0.0, error_coeff=2.0 / 9.0)
    raise ValueError(method)


def step_coefficients(method: str, dt_cur: float, dt_prev: float) -> Tuple[float, float, float, float, float]:
    """(c0, c1, d1, c2, error_coeff_integ) exactly as body_fn computes them
    (reference: full_mna.py:1851-1891)."""
    inv_dt = 1.0 / dt_cur
    if method == "be":
        return inv_dt, -inv_dt, 0.0, 0.0, -0.5
    if method == "trap":
        return 2.0 * inv_dt, -2.0 * inv_dt, -1.0, 0.0, 1.0 / 24.0
    omega = dt_cur / dt_prev if dt_prev > 0 else 1.0
    op1 = 1.0 + omega
    return (1.0 + 2.0 * omega) / (op1 * dt_cur), -op1 / dt_cur, 0.0, (omega * omega) / (op1 * dt_cur), -2.0 / 9.0


# --------------------------------------------------------------------------------------
# Predictor / LTE (reference: vajax/analysis/transient/adaptive.py:94-218, :221-339)
# --------------------------------------------------------------------------------------
def predict_voltage(V_history, dt_history, history_count, new_dt, max_order):
    order = max(min(history_count - 1, max_order), 0)
    cumsum_dt = np.cumsum(dt_history)
    tau = -(new_dt + np.concatenate([[0.0], cumsum_dt[:-1]])) / new_dt
    a0 = np.array([1.0, 0.0, 0.0])
    err0 = 1.0
    ratio1 = new_dt / max(dt_history[0], 1e-30)
    a1 = np.array([1.0 + ratio1, -ratio1, 0.0])
    err1 = -ratio1 * (1 + ratio1) / 2
    tau0, tau1, tau2 = -1.0, tau[1], tau[2]
    denom0 = (tau0 - tau1) * (tau0 - tau2)
    denom1 = (tau1 - tau0) * (tau1 - tau2)
    denom2 = (tau2 - tau0) * (tau2 - tau1)
    denom0 = 1e-30 if abs(denom0) < 1e-30 else denom0
    denom1 = 1e-30 if abs(denom1) < 1e-30 else denom1
    denom2 = 1e-30 if abs(denom2) < 1e-30 else denom2
    l0 = (0 - tau1) * (0 - tau2) / denom0
    l1 = (0 - tau0) * (0 - tau2) / denom1
    l2 = (0 - tau0) * (0 - tau1) / denom2
    a2 = np.array([l0, l1, l2])
    err2 = -tau0 * tau1 * tau2 / 6
    a = a0 if order == 0 else (a1 if order == 1 else a2)
    err = err0 if order == 0 else (err1 if order == 1 else err2)
    a_masked = np.where(np.arange(3) <= order, a, 0.0)
    # einsum("i,ij->j") accumulates i = 0,1,2 in order
    V_pred = a_masked[0] * V_history[0] + a_masked[1] * V_history[1] + a_masked[2] * V_history[2]
    return V_pred, err, a_masked


def compute_lte_timestep(V_new, V_pred, pred_err_coeff, dt, history_count, cfg, error_coeff_integ, V_max_historic):
    lte = (V_new - V_pred) * (error_coeff_integ / (error_coeff_integ - pred_err_coeff))
    V_current = np.maximum(np.abs(V_new), np.abs(V_pred))
    V_scale = np.maximum(V_max_historic, V_current) if V_max_historic is not None else V_current
    tol = cfg.reltol * V_scale + cfg.abstol
    lte_norm = np.max(np.abs(lte) / tol)
    order = min(history_count - 1, cfg.max_order)
    factor = np.power(cfg.lte_ratio / max(lte_norm, 1e-10), 1.0 / (order + 1))
    factor = min(max(factor, 0.1), cfg.grow_factor)
    return dt * factor, lte_norm


@dataclass
class AdaptiveConfig:
    """reference: vajax/analysis/transient/adaptive.py:37-91"""
    lte_ratio: float = 3.5
    redo_factor: float = 2.5
    reltol: float = 1e-3
    abstol: float = 1e-12
    min_dt: float = 1e-15
    max_dt: float = float("inf")
    tran_minpts: int = 50
    warmup_steps: int = 2
    max_order: int = 2
    grow_factor: float = 2.0
    nr_convtol: float = 1.0
    gshunt_init: float = 0.0
    gshunt_steps: int = 5
    gshunt_target: float = 0.0
    tran_fs: float = 0.25
    tran_ft: float = 0.25
    integration_method: str = "trap"
    max_consecutive_rejects: int = 5


@dataclass
class SimOptions:
    """reference: vajax/analysis/options.py:34-190 (hot-path subset)"""
    nr_damping: float = 1.0
    tran_itl: int = 20
    reltol: float = 1e-3
    abstol: float = 1e-12
    vntol: float = 1e-6
    gmin: float = 1e-12
    temp: float = 27.0
    icmode: str = "op"


# --------------------------------------------------------------------------------------
# Node collapse (reference: vajax/analysis/node_setup.py:15-156)
# --------------------------------------------------------------------------------------
def compute_collapse_roots(pairs, n_nodes):
    parent = list(range(n_nodes))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for a, b in pairs:
        if b != U32_MAX and a < n_nodes and b < n_nodes:
            px, py = find(a), find(b)
            if px != py:
                if py < 4:
                    parent[px] = py
                elif px < 4:
                    parent[py] = px
                else:
                    parent[py] = px
    return {i: find(i) for i in range(n_nodes)}


def setup_internal_nodes(devices, num_nodes, bundles, collapse_decisions):
    next_internal = num_nodes
    out: Dict[str, Dict[str, int]] = {}
    for dev in devices:
        if not dev.get("is_openvaf"):
            continue
        b = bundles.get(dev["model"])
        if not b:
            continue
        model_nodes = b["nodes"]
        roots = compute_collapse_roots(collapse_decisions.get(dev["name"], []), len(model_nodes))
        ext = dev["nodes"]
        n_ext = len(ext)
        root_to_circ: Dict[int, int] = {}
        imap = {}
        for i in range(n_ext, len(model_nodes)):
            root = roots.get(i, i)
            if root < n_ext:
                imap[model_nodes[i]] = ext[root]
            else:
                if root not in root_to_circ:
                    root_to_circ[root] = next_internal
                    next_internal += 1
                imap[model_nodes[i]] = root_to_circ[root]
        out[dev["name"]] = imap
    return next_internal, out


# node-name resolution table (reference: vajax/analysis/engine.py:620-692)
_INTERNAL_NAME_MAP = {
    "NOI": "node4", "GP": "node5", "SI": "node6", "DI": "node7", "BP": "node8", "BI": "node9",
    "BS": "node10", "BD": "node11", "G": "node1", "D": "node0", "S": "node2", "B": "node3",
    "A": "node0", "C": "node1", "CI": "node2", "a": "node0", "c": "node1", "a_int": "node2",
}
_SIMPLE_2TERM = {"A": "node0", "B": "node1"}


def parse_voltage_param(name, node_map, ground=0):
    m = re.match(r"V\(([^,)]+)(?:,([^)]+))?\)", name)
    if not m:
        return (ground, ground)

    def resolve(nm):
        if nm in _INTERNAL_NAME_MAP:
            mp = _INTERNAL_NAME_MAP[nm]
            if mp in node_map:
                return node_map[mp]
            if nm in _SIMPLE_2TERM and _SIMPLE_2TERM[nm] in node_map:
                return node_map[_SIMPLE_2TERM[nm]]
        elif nm in _SIMPLE_2TERM:
            if _SIMPLE_2TERM[nm] in node_map:
                return node_map[_SIMPLE_2TERM[nm]]
        return node_map.get(nm, node_map.get(nm.lower(), ground))

    n1 = resolve(m.group(1).strip())
    n2 = resolve(m.group(2).strip()) if m.group(2) else ground
    return (n1, n2)


# --------------------------------------------------------------------------------------
# Engine
# --------------------------------------------------------------------------------------
class OracleEngine:
    """CPU restatement of CircuitEngine.prepare()/run_transient() for the hot path only
    (reference: vajax/analysis/engine.py:824-1054, 1054-1175)."""

    def __init__(self, circuit: dict, bundles: Dict[str, dict], use_device_limiting: bool = False):
        self.circuit = circuit
        self.devices = circuit["devices"]
        self.num_nodes = int(circuit["num_nodes"])
        self.node_names = dict(circuit.get("node_names", {}))
        self.bundles = bundles
        self.use_device_limiting = use_device_limiting
        ap = dict(circuit.get("analysis", {}))
        self.analysis_params = ap
        self.options = SimOptions()
        for k in ("nr_damping", "tran_itl", "reltol", "abstol", "vntol", "gmin", "temp", "icmode"):
            if k in ap:
                setattr(self.options, k, type(getattr(self.options, k))(ap[k]))
        self.temperature_k = float(circuit.get("temperature_k", self.options.temp + 273.15))
        self._setup_done = False

    # ---- early collapse decisions (reference: openvaf_models.py:556-670)
    def _collapse_decisions(self):
        out: Dict[str, List[Tuple[int, int]]] = {}
        by_type: Dict[str, List[dict]] = {}
        for dev in self.devices:
            if dev.get("is_openvaf"):
                by_type.setdefault(dev["model"], []).append(dev)
        for mt, devs in by_type.items():
            b = self.bundles[mt]
            pairs = [tuple(p) for p in b["collapsible_pairs"]]
            names = [p.lower() for p in b["init_param_names"]]
            defaults = {k.lower(): v for k, v in b["param_defaults"].items()}
            if len(names) == 0 or len(pairs) == 0:
                for dev in devs:
                    out[dev["name"]] = list(pairs)
                continue
            groups: Dict[tuple, List[dict]] = {}
            for dev in devs:
                dp = dev.get("params", {})
                key = tuple(float(dp[n]) if n in dp else float(defaults.get(n, 0.0)) for n in names)
                groups.setdefault(key, []).append(dev)
            for key, gd in groups.items():
                _, col = mir_interp.init_model(b, list(key), 1)
                sel = [pairs[i] for i in range(len(pairs)) if i < col.shape[1] and col[0, i] > 0.5]
                for dev in gd:
                    out[dev["name"]] = sel
        return out

    # ---- prepare_static_inputs (reference: openvaf_models.py:668-1056)
    def _prepare_model(self, mt: str, devs: List[dict], internal_nodes) -> dict:
        b = self.bundles[mt]
        names, kinds = b["param_names"], b["param_kinds"]
        model_nodes = b["nodes"]
        n_dev, n_params = len(devs), len(names)
        voltage_indices = [i for i, k in enumerate(kinds) if k == "voltage"]
        col_values: Dict[int, Any] = {}
        varying = set(voltage_indices)
        param_to_cols: Dict[str, List[int]] = {}
        given_to_cols: Dict[str, List[int]] = {}
        limit_param_map: Dict[int, Tuple[str, str]] = {}
        defaults = b["param_defaults"]
        for idx, (nm, kd) in enumerate(zip(names, kinds)):
            nl = nm.lower()
            if kd == "param":
                param_to_cols.setdefault(nl, []).append(idx)
            elif kd == "param_given":
                given_to_cols.setdefault(nl, []).append(idx)
            elif kd == "temperature":
                col_values[idx] = self.temperature_k
            elif kd == "sysfun" and nl == "mfactor":
                col_values[idx] = 1.0
            elif kd in ("prev_state", "enable_lim", "new_state", "enable_integration"):
                limit_param_map[idx] = (kd, nm)
        all_dev_params = [d["params"] for d in devs]
        all_unique = set()
        for p in all_dev_params:
            all_unique.update(k.lower() for k in p.keys())
        for pn in all_unique:
            if pn in param_to_cols:
                vals = np.array([float(p.get(pn, p.get(pn.upper(), 0.0))) for p in all_dev_params])
                if np.all(vals == vals[0]):
                    for c in param_to_cols[pn]:
                        col_values[c] = float(vals[0])
                else:
                    for c in param_to_cols[pn]:
                        col_values[c] = vals
                        varying.add(c)
            if pn in given_to_cols:
                for c in given_to_cols[pn]:
                    col_values[c] = 1.0
        for pn, cols in param_to_cols.items():
            if pn not in all_unique:
                if pn in ("tnom", "tref", "tr"):
                    dflt = 27.0
                elif pn in ("nf", "mult", "ns", "nd"):
                    dflt = 1.0
                else:
                    dflt = defaults.get(pn, 0.0)
                for c in cols:
                    col_values[c] = dflt
        # device contexts
        dae = b["dae_data"]
        va_terms, va_int = dae["terminals"], dae["internal_nodes"]
        ctxs = []
        for dev in devs:
            ext = dev["nodes"]
            inodes = internal_nodes.get(dev["name"], {})
            nm: Dict[str, int] = {}
            for i in range(len(ext)):
                nm[model_nodes[i]] = ext[i]
            for k, v in inodes.items():
                nm[k] = v
            for i, vn in enumerate(va_terms):
                if i < len(ext):
                    nm[vn] = ext[i]
            nt = len(va_terms)
            for i, vn in enumerate(va_int):
                key = f"node{nt + i}"
                if key in inodes:
                    nm[vn] = inodes[key]
            pairs = [parse_voltage_param(names[i], nm, 0) for i in voltage_indices]
            ctxs.append({"name": dev["name"], "node_map": nm, "voltage_node_pairs": pairs})
        shared_idx = [c for c in range(n_params) if c not in limit_param_map and c not in varying]
        varying_idx = [c for c in range(n_params) if c not in limit_param_map and c in varying]
        shared_vals = []
        for c in shared_idx:
            v = col_values.get(c, 0.0)
            shared_vals.append(float(v[0]) if isinstance(v, np.ndarray) else float(v))
        dev_cols = []
        for c in varying_idx:
            v = col_values.get(c)
            if v is None:
                dev_cols.append(np.zeros(n_dev))
            elif isinstance(v, np.ndarray):
                dev_cols.append(v.astype(np.float64))
            else:
                dev_cols.append(np.full(n_dev, float(v)))
        device_params = np.column_stack(dev_cols) if dev_cols else np.zeros((n_dev, 0))
        # init -> cache (openvaf_models.py:886-921)
        eval_name_to_idx = {n.lower(): i for i, n in enumerate(names)}
        init_to_eval = [eval_name_to_idx.get(n.lower(), -1) for n in b["init_param_names"]]
        s_pos = {c: k for k, c in enumerate(shared_idx)}
        v_pos = {c: k for k, c in enumerate(varying_idx)}
        init_vals = []
        for ei in init_to_eval:
            if ei in s_pos:
                init_vals.append(shared_vals[s_pos[ei]])
            elif ei in v_pos:
                init_vals.append(device_params[:, v_pos[ei]])
            else:
                init_vals.append(0.0)
        cache, collapse = mir_interp.init_model(b, init_vals, n_dev)
        ncc = cache.shape[1]
        if ncc > 0 and n_dev > 1:
            const_mask = np.all(cache == cache[0:1, :], axis=0)
            shared_cache_idx = [int(i) for i in np.where(const_mask)[0]]
            varying_cache_idx = [int(i) for i in np.where(~const_mask)[0]]
        else:
            shared_cache_idx, varying_cache_idx = [], list(range(ncc))
        vpos = [v_pos[v] for v in voltage_indices if v in v_pos]
        # probe limiting support (openvaf_models.py:972-990)
        use_lim = self.use_device_limiting
        n_lim = 0
        if use_lim:
            probe = mir_interp.eval_model(
                b, {}, np.zeros((1, ncc)), 1, {}, np.zeros((1, 8)), True, limit_param_map)
            n_lim = probe["n_limit_states"]
            if n_lim > 0 and not probe["limit_traced"]:
                use_lim, n_lim = False, 0
        R, J = len(dae["residuals"]), len(dae["jacobian"])
        node_names = [r["node_name"] for r in dae["residuals"]]
        jkeys = [(e["row_node_name"], e["col_node_name"]) for e in dae["jacobian"]]
        # build_stamp_index_mapping (mna_builder.py:44-118)
        res_idx = np.full((n_dev, R), -1, dtype=np.int32)
        jr = np.full((n_dev, J), -1, dtype=np.int32)
        jc = np.full((n_dev, J), -1, dtype=np.int32)
        for d, ctx in enumerate(ctxs):
            nm = ctx["node_map"]
            for r, name in enumerate(node_names):
                ni = nm.get(name)
                if ni is not None and ni != 0 and ni > 0:
                    res_idx[d, r] = ni - 1
            for j, (rn, cn) in enumerate(jkeys):
                ri, ci = nm.get(rn), nm.get(cn)
                if ri is not None and ci is not None and ri > 0 and ci > 0:
                    jr[d, j] = ri - 1
                    jc[d, j] = ci - 1
        vn1 = np.array([[p[0] for p in c["voltage_node_pairs"]] for c in ctxs], dtype=np.int32).reshape(n_dev, len(voltage_indices))
        vn2 = np.array([[p[1] for p in c["voltage_node_pairs"]] for c in ctxs], dtype=np.int32).reshape(n_dev, len(voltage_indices))
        return dict(
            bundle=b, n_dev=n_dev, shared_indices=shared_idx, varying_indices=varying_idx,
            shared_params=np.array(shared_vals), device_params=device_params,
            voltage_positions=np.array(vpos, dtype=np.int32), voltage_indices=voltage_indices,
            cache=cache, shared_cache_indices=shared_cache_idx, varying_cache_indices=varying_cache_idx,
            limit_param_map=limit_param_map, use_device_limiting=use_lim, num_limit_states=n_lim,
            res_indices=res_idx, jac_row_indices=jr, jac_col_indices=jc, vnode1=vn1, vnode2=vn2,
            R=R, J=J, contexts=ctxs,
        )

    # ---- _build_transient_setup + _ensure_full_mna_solver (engine.py:1054-1175, full_mna.py:344-560)
    def setup(self):
        if self._setup_done:
            return
        self.collapse = self._collapse_decisions()
        self.n_total, self.device_internal_nodes = setup_internal_nodes(
            self.devices, self.num_nodes, self.bundles, self.collapse)
        self.n_unknowns = self.n_total - 1
        self.n_external = self.num_nodes
        by_type: Dict[str, List[dict]] = {}
        self.vsources, self.isources = [], []
        for dev in self.devices:
            if dev.get("is_openvaf"):
                by_type.setdefault(dev["model"], []).append(dev)
            elif dev["model"] == "vsource":
                self.vsources.append(dev)
            elif dev["model"] == "isource":
                self.isources.append(dev)
        self.models = {mt: self._prepare_model(mt, devs, self.device_internal_nodes) for mt, devs in by_type.items()}
        self.n_vs = len(self.vsources)
        self.n_aug = self.n_unknowns + self.n_vs
        self.vs_p = np.array([d["nodes"][0] for d in self.vsources], dtype=np.int32)
        self.vs_n = np.array([d["nodes"][1] for d in self.vsources], dtype=np.int32)
        self.is_f_idx = np.array(
            [[(d["nodes"][0] - 1) if d["nodes"][0] != 0 else -1, (d["nodes"][1] - 1) if d["nodes"][1] != 0 else -1]
             for d in self.isources], dtype=np.int32).reshape(len(self.isources), 2)
        self.vsrc_fns = [make_source(d["params"]) for d in self.vsources]
        self.isrc_fns = [make_source(d["params"]) for d in self.isources]
        # limit state offsets (mna_builder.py:386-402)
        self.total_limit_states = 0
        self.limit_offsets = {}
        for mt, m in self.models.items():
            if m["use_device_limiting"] and m["num_limit_states"] > 0:
                self.limit_offsets[mt] = (self.total_limit_states, m["n_dev"], m["num_limit_states"])
                self.total_limit_states += m["n_dev"] * m["num_limit_states"]
        # NOI + internal device node indices (full_mna.py:360-374)
        noi, internal = [], []
        for _dn, inodes in self.device_internal_nodes.items():
            for nname, nidx in inodes.items():
                internal.append(nidx)
                if nname == "node4":
                    noi.append(nidx)
        self.noi_indices = np.array(noi, dtype=np.int64)
        self.internal_indices = np.array(sorted(set(internal)), dtype=np.int64)
        # CSR pattern discovery (full_mna.py:423-513): trial COO = valid device entries, masked
        # entries at (0,0), vsource B/B^T (masked at (0,0)), all node diagonals
        n_aug = self.n_aug
        rows, cols = [], []
        for m in self.models.values():
            r, c = m["jac_row_indices"].ravel(), m["jac_col_indices"].ravel()
            valid = (r >= 0) & (c >= 0)
            rr = np.where(valid, r, 0)
            cc = np.where(valid, c, 0)
            rows += [rr, rr]
            cols += [cc, cc]
        if self.n_vs > 0:
            br = np.arange(self.n_vs)
            vp, vn_ = self.vs_p > 0, self.vs_n > 0
            rows += [np.where(vp, self.vs_p - 1, 0), np.where(vn_, self.vs_n - 1, 0),
                     np.where(vp, self.n_unknowns + br, 0), np.where(vn_, self.n_unknowns + br, 0)]
            cols += [np.where(vp, self.n_unknowns + br, 0), np.where(vn_, self.n_unknowns + br, 0),
                     np.where(vp, self.vs_p - 1, 0), np.where(vn_, self.vs_n - 1, 0)]
        di = np.arange(self.n_unknowns)
        rows.append(di)
        cols.append(di)
        lin = np.concatenate([np.asarray(x, dtype=np.int64) for x in rows]) * n_aug + np.concatenate(
            [np.asarray(x, dtype=np.int64) for x in cols])
        self.unique_linear = np.unique(lin)
        self.nse = len(self.unique_linear)
        urows = (self.unique_linear // n_aug).astype(np.int32)
        self.csr_indices = (self.unique_linear % n_aug).astype(np.int32)
        self.csr_indptr = np.zeros(n_aug + 1, dtype=np.int32)
        self.csr_indptr[1:] = np.cumsum(np.bincount(urows, minlength=n_aug))
        # stamp positions (mna_builder.py:126-255)
        for m in self.models.values():
            m["csr_pos"] = self._positions(m["jac_row_indices"], m["jac_col_indices"])
        self.diag_pos = self._positions(di, di)
        assert np.all(self.diag_pos >= 0)
        if self.n_vs > 0:
            br = np.arange(self.n_vs)
            vp, vn_ = self.vs_p > 0, self.vs_n > 0
            f = lambda r, c, v: np.where(v, self._positions(np.where(v, r, 0), np.where(v, c, 0)), -1)  # noqa: E731
            self.vs_pos = dict(
                b_p=f(self.vs_p - 1, self.n_unknowns + br, vp), b_n=f(self.vs_n - 1, self.n_unknowns + br, vn_),
                bt_p=f(self.n_unknowns + br, self.vs_p - 1, vp), bt_n=f(self.n_unknowns + br, self.vs_n - 1, vn_),
                valid_p=vp, valid_n=vn_)
        # NOI masks (solver_factories.py:37-137)
        nu = self.n_unknowns
        # masks are built on the n_unknowns node equations (negative indices wrap there, exactly
        # like jnp's .at[]: an internal node collapsed onto ground gives index -1 = last node)
        # and then extended with True for the vsource equations (solver_factories.py:1018-1036)
        conv = np.ones(nu, dtype=bool)
        if len(self.internal_indices):
            conv[self.internal_indices - 1] = False
        rmask = np.ones(nu, dtype=bool)
        self.noi_res_idx = np.array(sorted(set(int(x) for x in self.noi_indices - 1)), dtype=np.int64)
        if len(self.noi_res_idx):
            rmask[self.noi_res_idx] = False
            rowid = np.repeat(np.arange(n_aug), np.diff(self.csr_indptr))
            isnoi = np.zeros(n_aug, dtype=bool)
            isnoi[self.noi_res_idx] = True
            self.noi_row_mask = np.where(isnoi[rowid])[0]
            self.noi_col_mask = np.where(isnoi[self.csr_indices] & (rowid < nu))[0]
            self.noi_diag = np.where(isnoi[rowid] & (rowid == self.csr_indices))[0]
        self.residual_mask = np.concatenate([rmask, np.ones(self.n_vs, dtype=bool)])
        if len(self.internal_indices):
            self.residual_conv_mask = np.concatenate([conv, np.ones(self.n_vs, dtype=bool)])
        else:
            self.residual_conv_mask = self.residual_mask.copy()
        self._setup_done = True

    def _positions(self, r, c):
        r, c = np.asarray(r), np.asarray(c)
        valid = (r >= 0) & (c >= 0)
        lin = r.astype(np.int64) * self.n_aug + c.astype(np.int64)
        pos = np.clip(np.searchsorted(self.unique_linear, lin.ravel()), 0, self.nse - 1)
        ok = (self.unique_linear[pos] == lin.ravel()) & valid.ravel()
        return np.where(ok, pos, -1).astype(np.int32).reshape(r.shape)

    # ---- build_system, CSR-direct semantics (mna_builder.py:759-1076)
    def build_system(self, X, vsource_vals, isource_vals, Q_prev, c0, gmin_arg=1e-12, gshunt=0.0,
                     c1=0.0, d1=0.0, dQdt_prev=None, c2=0.0, Q_prev2=None, limit_state_in=None,
                     nr_iteration=1, return_parts=False):
        self.setup()
        nu, n_total = self.n_unknowns, self.n_total
        V = X[:n_total]
        I_branch = X[n_total:]
        csr = np.zeros(self.nse)
        f_resist, Q = np.zeros(nu), np.zeros(nu)
        maxres, lim_r, lim_q = np.zeros(nu), np.zeros(nu), np.zeros(nu)
        limit_out = np.zeros(self.total_limit_states)

        def stamp(vec, idx, val):
            valid = idx >= 0
            v = np.where(valid, val, 0.0)
            v = np.where(np.isnan(v), 0.0, v)
            np.add.at(vec, np.where(valid, idx, 0), v)

        def stamp_max(vec, idx, val):
            valid = idx >= 0
            v = np.where(valid, np.abs(val), 0.0)
            v = np.where(np.isnan(v), 0.0, v)
            np.maximum.at(vec, np.where(valid, idx, 0), v)

        if len(self.isources) and len(isource_vals):
            fv = (np.asarray(isource_vals)[:, None] * np.array([1.0, -1.0])[None, :]).ravel()
            fi = self.is_f_idx.ravel()
            stamp(f_resist, fi, fv)
            stamp_max(maxres, fi, fv)
        parts = {}
        for mt, m in self.models.items():
            n_dev = m["n_dev"]
            vu = V[m["vnode1"]] - V[m["vnode2"]]
            dp = m["device_params"].copy()
            if dp.shape[1]:
                dp[:, m["voltage_positions"]] = vu
            pv: Dict[int, Any] = {}
            for k, c in enumerate(m["shared_indices"]):
                pv[c] = m["shared_params"][k]
            for k, c in enumerate(m["varying_indices"]):
                pv[c] = dp[:, k]
            simp = dict(SIMPARAM_DEFAULTS)
            simp["$analysis_type"] = 2.0 if c0 > 0 else 0.0
            simp["gmin"] = gmin_arg
            simp["iniLim"] = 1.0 if nr_iteration == 0 else 0.0
            simp["iteration"] = float(nr_iteration)
            n_lim = max(1, m["num_limit_states"])
            if m["use_device_limiting"] and mt in self.limit_offsets and limit_state_in is not None:
                off, _, nl = self.limit_offsets[mt]
                lsi = limit_state_in[off:off + n_dev * nl].reshape(n_dev, nl)
            else:
                lsi = np.zeros((n_dev, n_lim))
            out = mir_interp.eval_model(m["bundle"], pv, m["cache"], n_dev, simp, lsi,
                                        m["use_device_limiting"], m["limit_param_map"])
            if m["use_device_limiting"] and mt in self.limit_offsets:
                off, _, nl = self.limit_offsets[mt]
                limit_out[off:off + n_dev * nl] = out["limit_state_out"][:, :nl].ravel()
            comb = out["jac_resist"] + c0 * out["jac_react"]
            pos = m["csr_pos"]
            vj = pos >= 0
            jv = np.where(vj, comb, 0.0)
            jv = np.where(np.isnan(jv), 0.0, jv)
            np.add.at(csr, np.where(vj, pos, 0).ravel(), jv.ravel())
            ri = m["res_indices"].ravel()
            stamp(f_resist, ri, out["res_resist"].ravel())
            stamp_max(maxres, ri, out["res_resist"].ravel())
            stamp(Q, ri, out["res_react"].ravel())
            stamp(lim_r, ri, out["lim_rhs_resist"].ravel())
            stamp(lim_q, ri, out["lim_rhs_react"].ravel())
            parts[mt] = out
        np.add.at(csr, self.diag_pos, self.options.gmin + gshunt)
        if self.n_vs > 0:
            vp = self.vs_pos
            np.add.at(csr, np.where(vp["valid_p"], vp["b_p"], 0), np.where(vp["valid_p"], 1.0, 0.0))
            np.add.at(csr, np.where(vp["valid_n"], vp["b_n"], 0), np.where(vp["valid_n"], -1.0, 0.0))
            np.add.at(csr, np.where(vp["valid_p"], vp["bt_p"], 0), np.where(vp["valid_p"], 1.0, 0.0))
            np.add.at(csr, np.where(vp["valid_n"], vp["bt_n"], 0), np.where(vp["valid_n"], -1.0, 0.0))
        # _build_residual (mna_builder.py:1012-1076; mna.py:489-596, 673-719)
        f_resist = f_resist - lim_r
        I_vs = np.where(self.vs_p > 0, -f_resist[np.maximum(self.vs_p - 1, 0)], 0.0) if self.n_vs else np.zeros(0)
        dq = dQdt_prev if dQdt_prev is not None else np.zeros(nu)
        q2 = Q_prev2 if Q_prev2 is not None else np.zeros(nu)
        if self.n_vs > 0:
            fb = np.zeros(nu)
            idx = np.concatenate([np.where(self.vs_p > 0, self.vs_p - 1, 0), np.where(self.vs_n > 0, self.vs_n - 1, 0)])
            val = np.concatenate([np.where(self.vs_p > 0, I_branch, 0.0), np.where(self.vs_n > 0, -I_branch, 0.0)])
            np.add.at(fb, idx, val)
            f_resist = f_resist + fb
        eff = self.options.gmin + gshunt
        f_node = f_resist - np.zeros_like(f_resist) + c0 * (Q - lim_q) + c1 * Q_prev + d1 * dq + c2 * q2
        f_node = f_node + eff * V[1:]
        f_branch = V[self.vs_p] - V[self.vs_n] - np.asarray(vsource_vals) if self.n_vs else np.zeros(0)
        f = np.concatenate([f_node, f_branch])
        if return_parts:
            return csr, f, Q, I_vs, limit_out, maxres, parts
        return csr, f, Q, I_vs, limit_out, maxres

    # ---- linear solve with NOI enforcement (solver_factories.py:1268-1276; dense :775-788)
    def linear_solve(self, csr, f):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla

        csr = csr.copy()
        fs = f.copy()
        if len(self.noi_res_idx):
            csr[self.noi_row_mask] = 0.0
            csr[self.noi_col_mask] = 0.0
            csr[self.noi_diag] = 1.0
            fs[self.noi_res_idx] = 0.0
        A = sp.csr_matrix((csr, self.csr_indices, self.csr_indptr), shape=(self.n_aug, self.n_aug))
        if getattr(self, "linear_solver", "sparse") == "dense":
            # A12, make_dense_full_mna_solver.linear_solve (solver_factories.py:784-788): J + 1e-14 I, LAPACK solve.
            # With this the oracle reproduces the reference's DENSE-path transients bit for bit (the ring: 67 accepted
            # steps, against 64 on the sparse path - the regulariser, not rounding, is the whole difference).
            return np.linalg.solve(A.toarray() + 1e-14 * np.eye(self.n_aug), -fs)
        if self.n_aug <= 400:
            return np.linalg.solve(A.toarray(), -fs)
        if getattr(self, "reuse_ordering", False):
            return self._solve_cached_ordering(A, -fs)
        return spla.splu(A.tocsc()).solve(-fs)

    def _solve_cached_ordering(self, A, b):
        """Timing legs only (bench.py sets ``reuse_ordering``; the parity checks never do): the fill-reducing
        ordering is computed once per pattern and re-used, as the reference's UMFPACK path caches its symbolic
        analysis by (n, nnz) and repeats only the numeric factorisation (vajax/sparse/umfpack_ffi.cpp:140-179,
        266-295).  SciPy's SuperLU has no symbolic handle, so the cached object is the column permutation of a
        first MMD(A^T + A) factorisation, applied symmetrically before a NATURAL-order ``splu``
        (c6288: 134 ms -> 14 ms per factorisation: most of a default ``splu`` is the ordering, and the supernode
        relaxation / panel heuristics only cost time on a matrix this sparse)."""
        import scipy.sparse.linalg as spla

        perm = getattr(self, "_perm_cache", None)
        if perm is None:
            # perm_c maps a column to its position: the elimination ORDER is its inverse
            perm = self._perm_cache = np.argsort(np.asarray(spla.splu(A.tocsc(), permc_spec="MMD_AT_PLUS_A").perm_c))
        Ap = A[perm][:, perm].tocsc()
        x = np.empty_like(b)
        x[perm] = spla.splu(Ap, permc_spec="NATURAL", diag_pivot_thresh=0.1, relax=1, panel_size=1).solve(b[perm])
        return x

    # ---- nr_solve, while-loop mode (solver_factories.py:230-325, 530-725)
    def nr_solve(self, X_init, vsource_vals, isource_vals, Q_prev, c0, gmin=1e-12, gshunt=0.0, c1=0.0, d1=0.0,
                 dQdt_prev=None, c2=0.0, Q_prev2=None, limit_state_in=None, res_tol_floor=None,
                 max_iterations=None, trace=None):
        self.setup()
        o = self.options
        nu, n_total, n_vs = self.n_unknowns, self.n_total, self.n_vs
        max_it = o.tran_itl if max_iterations is None else max_iterations
        dq = dQdt_prev if dQdt_prev is not None else np.zeros(nu)
        q2 = Q_prev2 if Q_prev2 is not None else np.zeros(nu)
        ls = limit_state_in if limit_state_in is not None else np.zeros(self.total_limit_states)
        rtf = res_tol_floor if res_tol_floor is not None else np.full(nu, o.abstol)
        delta_abs_tol = np.concatenate([np.full(nu, o.vntol), np.full(n_vs, o.abstol)])
        X = np.array(X_init, dtype=np.float64)
        it, conv, max_f = 0, False, np.inf
        while (not conv) and it < max_it:
            csr, f, Q, _, ls_out, mrc = self.build_system(
                X, vsource_vals, isource_vals, Q_prev, c0, gmin, gshunt, c1, d1, dq, c2, q2, ls, it)
            res_tol = np.concatenate([np.maximum(mrc * o.reltol, rtf), np.full(n_vs, o.vntol)])
            f_check = np.where(self.residual_conv_mask, f, 0.0)
            max_f = np.max(np.abs(f_check)) if len(f_check) else 0.0
            res_conv = bool(np.all(np.abs(f_check) < res_tol))
            delta = self.linear_solve(csr, f)
            conv_delta = np.concatenate([delta[:nu] * o.nr_damping, delta[nu:]])
            X_ref = np.concatenate([X[1:n_total], X[n_total:]])
            tol = np.maximum(np.abs(X_ref) * o.reltol, delta_abs_tol)
            conv_delta = np.where(self.residual_mask, conv_delta, 0.0)
            delta_conv = (it == 0) or bool(np.all(np.abs(conv_delta) < tol))
            V_delta = delta[:nu]
            max_V = np.max(np.abs(V_delta)) if nu else 0.0
            V_scale = min(1.0, 1e30 / max(max_V, 1e-30))
            Xc = X.copy()
            Xc[1:n_total] += V_delta * V_scale * o.nr_damping
            Xc[n_total:] += delta[nu:]
            if len(self.noi_indices):
                Xc[self.noi_indices] = 0.0
            this_conv = res_conv and delta_conv
            if self.total_limit_states > 0:
                ld = np.max(np.abs(ls_out - ls))
                lref = max(np.max(np.abs(ls)) * o.reltol, o.vntol)
                this_conv = this_conv and (ld < lref) and (it >= 1)
            if trace is not None:
                trace.append(dict(it=it, X=X.copy(), csr=csr, f=f, delta=delta, max_f=max_f,
                                  res_conv=res_conv, delta_conv=delta_conv))
            X, conv, ls = Xc, this_conv, ls_out
            it += 1
        _, _, Qf, I_vs, _, mrc_f = self.build_system(
            X, vsource_vals, isource_vals, Q_prev, c0, gmin, gshunt, c1, d1, dq, c2, q2, ls, it)
        dQdt = c0 * Qf + c1 * Q_prev + d1 * dq + c2 * q2
        return X, it, conv, max_f, Qf, dQdt, I_vs, ls, mrc_f

    # ---- source arrays at time t (full_mna.py:1701-1760)
    def eval_sources(self, t):
        return (np.array([fn(t) for fn in self.vsrc_fns], dtype=np.float64),
                np.array([fn(t) for fn in self.isrc_fns], dtype=np.float64))

    def adaptive_config(self) -> AdaptiveConfig:
        """reference: engine.py:884-907"""
        ap = self.analysis_params
        kw = {}
        for src, dst, cast in [("tran_lteratio", "lte_ratio", float), ("tran_redofactor", "redo_factor", float),
                               ("nr_convtol", "nr_convtol", float), ("tran_gshunt", "gshunt_init", float),
                               ("reltol", "reltol", float), ("abstol", "abstol", float), ("tran_fs", "tran_fs", float),
                               ("tran_minpts", "tran_minpts", int), ("maxstep", "max_dt", float),
                               ("tran_method", "integration_method", str)]:
            if src in ap:
                kw[dst] = cast(ap[src])
        return AdaptiveConfig(**kw)

    def init_mid_rail(self):
        """reference: full_mna.py:694-746"""
        vdd = 0.0
        for d in self.devices:
            if d["model"] == "vsource":
                dc = d["params"].get("dc", 0.0)
                if dc > vdd:
                    vdd = dc
        V = np.full(self.n_total, vdd / 2.0)
        V[0] = 0.0
        for name, idx in self.node_names.items():
            nl = name.lower()
            if "vdd" in nl or "vcc" in nl:
                V[idx] = vdd
            elif nl in ("gnd", "vss", "0"):
                V[idx] = 0.0
        for d in self.devices:
            if d["model"] == "vsource" and len(d["nodes"]) >= 2:
                p, n = d["nodes"][0], d["nodes"][1]
                if n == 0 and p > 0:
                    V[p] = float(d["params"].get("dc", 0.0))
        for _dn, inodes in self.device_internal_nodes.items():
            if "node4" in inodes:
                V[inodes["node4"]] = 0.0
        return V

    # ---- run (full_mna.py:780-1196 + body_fn :1840-2194)
    def run_transient(self, t_stop=None, dt=None, max_steps=None, record=None):
        self.setup()
        ap = self.analysis_params
        t_stop = float(ap["stop"]) if t_stop is None else float(t_stop)
        dt = float(ap["step"]) if dt is None else float(dt)
        if max_steps is None:
            max_steps = int(t_stop / dt * 1.1) + 10  # engine.py:875
        cfg = self.adaptive_config()
        if cfg.tran_minpts > 0:
            cfg.max_dt = min(cfg.max_dt, t_stop / cfg.tran_minpts)
        if cfg.tran_fs != 1.0:
            dt = dt * cfg.tran_fs
        nu, n_total, n_vs, n_ext = self.n_unknowns, self.n_total, self.n_vs, self.n_external
        X0 = np.zeros(n_total + n_vs)
        X0[:n_total] = self.init_mid_rail()
        vs0, is0 = self.eval_sources(0.0)
        if self.options.icmode == "uic":
            X0[:n_total] = 0.0
            _, _, Q_init, I_dc, _, _ = self.build_system(
                X0, vs0, is0, np.zeros(nu), 0.0, self.options.gmin, 0.0, 0.0, 0.0, np.zeros(nu), 0.0,
                np.zeros(nu), np.zeros(self.total_limit_states), 1)
            dc_mrc = np.zeros(nu)
        else:
            X_dc, _, _, _, Q_dc, _, I_dc, _, dc_mrc = self.nr_solve(
                X0, vs0, is0, np.zeros(nu), 0.0, self.options.gmin, 0.0, 0.0, 0.0, None, 0.0, None, None, None)
            X0, Q_init = X_dc, Q_dc
        max_hist = cfg.max_order + 2
        Vh = np.zeros((max_hist, n_total))
        Vh[0] = X0[:n_total]
        dth = np.full(max_hist, dt)
        times = np.zeros(max_steps)
        V_out = np.zeros((max_steps, n_ext))
        V_out[0] = X0[:n_ext]
        I_out = np.zeros((max_steps, max(n_vs, 1)))
        if n_vs:
            I_out[0, :n_vs] = I_dc[:n_vs]
        s = dict(t=0.0, dt=dt, X=X0, Q_prev=Q_init, dQdt_prev=np.zeros(nu), Q_prev2=np.zeros(nu),
                 hist_count=1, step_idx=1, warmup=0, total_nr=0, rejected=0, min_dt_used=dt, max_dt_used=dt,
                 consec=0, Vmax=np.abs(X0[:n_total]), limit_state=np.zeros(self.total_limit_states),
                 hist_mrc=dc_mrc)
        method = cfg.integration_method
        while s["t"] < t_stop and s["step_idx"] < max_steps:
            t, dt_cur, X = s["t"], s["dt"], s["X"]
            with np.errstate(all="ignore"):
                c0, c1, d1, c2, eci = step_coefficients(method, dt_cur, dth[0])
            t_next = t + dt_cur
            vs, isv = self.eval_sources(t_next)
            can_predict = (s["warmup"] >= cfg.warmup_steps) and (s["hist_count"] >= 2)
            with np.errstate(all="ignore"):
                V_pred, pec, _ = predict_voltage(Vh, dth, s["hist_count"], dt_cur, cfg.max_order)
            X_init = X.copy()
            if can_predict:
                X_init[:n_total] = V_pred
            ramp = min(max(s["step_idx"] / cfg.gshunt_steps, 0.0), 1.0) if cfg.gshunt_steps > 0 else 1.0
            gsh = cfg.gshunt_init + ramp * (cfg.gshunt_target - cfg.gshunt_init)
            rtf = np.maximum(s["hist_mrc"] * cfg.reltol, cfg.abstol)
            X_new, iters, conv, max_f, Q, dQdt, I_vs, ls_out, mrc = self.nr_solve(
                X_init, vs, isv, s["Q_prev"], c0, 1e-12, gsh, c1, d1, s["dQdt_prev"], c2, s["Q_prev2"],
                s["limit_state"], rtf)
            nr_failed = not conv
            at_min = dt_cur <= cfg.min_dt
            nr_fail_min = nr_failed and at_min
            V_new = X_new[:n_total]
            with np.errstate(all="ignore"):
                dt_lte, lte_norm = compute_lte_timestep(V_new, V_pred, pec, dt_cur, s["hist_count"], cfg, eci, s["Vmax"])
                lte_reject = bool(dt_cur / dt_lte > cfg.redo_factor) and can_predict and conv
            nr_reject = nr_failed and not at_min
            force = s["consec"] >= cfg.max_consecutive_rejects
            reject = (lte_reject or nr_reject) and not force
            accept = not reject
            consec = 0 if accept else (s["consec"] + 1 if lte_reject else s["consec"])
            dt_from_lte = dt_lte if can_predict else dt_cur
            dt_nr = max(dt_cur * cfg.tran_ft, cfg.min_dt)
            last_good = dth[0] if s["hist_count"] > 0 else cfg.min_dt * 16.0
            dt_rec = max(last_good / 8.0, cfg.min_dt * 2.0)
            new_dt = (dt_rec if at_min else dt_nr) if nr_failed else dt_from_lte
            new_dt = min(max(new_dt, cfg.min_dt), cfg.max_dt)
            new_dt = min(new_dt, t_stop - t_next)
            use_new = accept and not nr_fail_min
            if record is not None:
                record.append(dict(step=s["step_idx"], t_next=t_next, dt=dt_cur, nr_iters=iters, accept=accept,
                                   lte_reject=lte_reject, nr_failed=nr_failed, lte_norm=float(lte_norm),
                                   new_dt=new_dt, max_f=float(max_f), X=X_new.copy()))
            if accept:
                s["t"] = t_next
                times[s["step_idx"]] = t_next
                Xrec = X_new if use_new else X
                V_out[s["step_idx"]] = Xrec[:n_ext]
                if n_vs:
                    I_out[s["step_idx"], :n_vs] = 0.0 if nr_fail_min else I_vs[:n_vs]
                s["step_idx"] += 1
                s["min_dt_used"] = min(s["min_dt_used"], dt_cur)
                s["max_dt_used"] = max(s["max_dt_used"], dt_cur)
            if use_new:
                s["Q_prev2"] = s["Q_prev"]
                s["X"], s["Q_prev"], s["dQdt_prev"] = X_new, Q, dQdt
                Vh = np.roll(Vh, 1, axis=0)
                Vh[0] = V_new
                dth = np.roll(dth, 1)
                dth[0] = dt_cur
                s["hist_count"] = min(s["hist_count"] + 1, max_hist)
                s["warmup"] += 1
                s["Vmax"] = np.maximum(s["Vmax"], np.abs(V_new))
                s["hist_mrc"] = np.maximum(s["hist_mrc"], mrc)
                if self.total_limit_states:
                    s["limit_state"] = ls_out
            s["dt"] = new_dt
            s["total_nr"] += int(iters)
            s["rejected"] += 1 if reject else 0
            s["consec"] = consec
        n = s["step_idx"]
        stats = dict(n_steps=n, accepted_steps=n, rejected_steps=s["rejected"], total_nr_iterations=s["total_nr"],
                     min_dt_used=s["min_dt_used"], max_dt_used=s["max_dt_used"], I_out=I_out,
                     final_state=s, V_history=Vh, dt_history=dth)
        return times, V_out, stats

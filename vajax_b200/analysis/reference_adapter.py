"""From the reference's prepared objects to a live ``vjx_handle`` ("the reference-side adapter").

The reference's ``CircuitEngine.prepare()`` / ``FullMNAStrategy.ensure_setup()`` leave behind everything the hot path
needs (reference files, all under /root/reference):

* ``runner.devices / num_nodes / node_names / analysis_params / options``  (vajax/analysis/engine.py:421-618),
* ``runner._device_collapse_decisions``  (openvaf_models.py:551-670),
* ``runner._compiled_models[mt]``: ``translator`` (the MIR), ``shared_indices / varying_indices / shared_params /
  device_params / shared_cache / device_cache / shared_cache_indices / varying_cache_indices / limit_param_map /
  use_device_limiting``  (openvaf_models.py:839-1058),
* ``setup.static_inputs_cache[mt] = (voltage_indices, stamp_indices, voltage_node1, voltage_node2, cache, collapse)``
  (engine.py:1115-1156; ``TransientSetup`` transient/base.py:35-60),
* ``strategy._bcsr_indptr / _bcsr_indices``: the CSR pattern  (transient/full_mna.py:423-513).

``engine_from_reference`` feeds them into ``B200Engine`` unchanged - the parameter / cache split, the cache values and
the collapse decisions are the REFERENCE's, nothing is re-derived on the GPU - and cross-checks what the engine still
computes itself (node numbering -> voltage gather indices and stamp indices, CSR pattern) against the reference's
arrays, entry by entry.  ``make_b200_full_mna_solver`` returns an ``nr_solve`` with the reference's 15-in / 9-out
signature (solver_factories.py:604-725), the object ``_ensure_full_mna_solver`` (full_mna.py:589-684) caches.

Nothing here imports the reference: it only reads attributes of objects the caller already holds.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from .engine import B200Engine

BUNDLE_KEYS = ("param_names", "param_kinds", "nodes", "collapse_decision_outputs", "collapsible_pairs", "num_collapsible",
               "init_param_names", "init_param_kinds")


def bundle_from_translator(tr: Any) -> dict:
    """The MIR bundle (``OpenVAFToJAX.get_cache_data`` schema, openvaf_jax/__init__.py:196-222) of a live translator,
    whether it was built from a compiled module or ``from_cache``."""
    mod = tr.module
    b = {"mir_data": tr.mir_data, "init_mir_data": tr.init_mir_data, "dae_data": tr.dae_data, "str_constants": tr.str_constants}
    for k in BUNDLE_KEYS:
        v = getattr(mod, k)
        b[k] = v if isinstance(v, int) else list(v)
    b["all_func_params"] = list(mod.get_all_func_params())
    b["param_defaults"] = dict(mod.get_param_defaults())
    b["osdi_descriptor"] = mod.get_osdi_descriptor()
    b["model_name"] = getattr(mod, "name", "model")
    return b


def circuit_from_runner(runner: Any) -> dict:
    """The flat circuit dict of ``B200Engine`` from a prepared reference ``CircuitEngine``."""
    ap = dict(getattr(runner, "analysis_params", {}) or {})
    o = runner.options
    for k in ("step", "stop", "maxstep", "icmode", "tran_lteratio", "nr_convtol", "reltol", "abstol", "vntol", "tran_itl"):
        if k not in ap and hasattr(o, k) and getattr(o, k) is not None:
            ap[k] = getattr(o, k)
    if "tran_method" not in ap and hasattr(o, "tran_method"):
        tm = o.tran_method
        ap["tran_method"] = getattr(tm, "value", str(tm))
    temp = getattr(runner, "_simulation_temperature", None)
    return dict(num_nodes=int(runner.num_nodes), node_names=dict(runner.node_names), devices=runner.devices, analysis=ap,
                temperature_k=float(temp) if temp is not None else float(o.temp) + 273.15)


def model_data_from_reference(runner: Any, setup: Any) -> Dict[str, dict]:
    """Per model type: the reference's parameter / cache split and values, in the layout ``B200Engine`` consumes."""
    out: Dict[str, dict] = {}
    for mt, tup in setup.static_inputs_cache.items():
        comp = runner._compiled_models[mt]
        voltage_indices, stamp_indices, vnode1, vnode2, cache, _collapse = tup
        out[mt] = dict(
            shared_indices=[int(i) for i in comp["shared_indices"]], varying_indices=[int(i) for i in comp["varying_indices"]],
            shared_params=np.asarray(comp["shared_params"], dtype=np.float64),
            device_params=np.asarray(comp["device_params"], dtype=np.float64),
            cache=np.asarray(cache, dtype=np.float64),
            shared_cache_indices=[int(i) for i in comp["shared_cache_indices"]],
            varying_cache_indices=[int(i) for i in comp["varying_cache_indices"]],
            limit_param_map=dict(comp.get("limit_param_map") or {}),
            collapse_by_device=getattr(runner, "_device_collapse_decisions", None),
            # kept for the cross-check
            ref_voltage_indices=[int(i) for i in voltage_indices],
            ref_vnode1=np.asarray(vnode1, dtype=np.int32), ref_vnode2=np.asarray(vnode2, dtype=np.int32),
            ref_res=np.asarray(stamp_indices["res_indices"], dtype=np.int32),
            ref_jrow=np.asarray(stamp_indices["jac_row_indices"], dtype=np.int32),
            ref_jcol=np.asarray(stamp_indices["jac_col_indices"], dtype=np.int32),
        )
    return out


def _cross_check(eng: B200Engine, md: Dict[str, dict], setup: Any, strategy: Optional[Any]) -> None:
    if (eng.n_total, eng.n_unknowns) != (int(setup.n_total), int(setup.n_unknowns)):
        raise ValueError(f"node numbering differs from the reference: {(eng.n_total, eng.n_unknowns)} != "
                         f"{(int(setup.n_total), int(setup.n_unknowns))}")
    for mt, m in eng.models.items():
        g = md[mt]
        for ours, ref, what in ((m["vnode1"], g["ref_vnode1"], "voltage_node1"), (m["vnode2"], g["ref_vnode2"], "voltage_node2"),
                                (m["res_indices"], g["ref_res"], "res_indices"), (m["jac_row_indices"], g["ref_jrow"], "jac_row_indices"),
                                (m["jac_col_indices"], g["ref_jcol"], "jac_col_indices")):
            if ours.shape != ref.shape or not np.array_equal(ours, ref):
                raise ValueError(f"{mt}: {what} differs from the reference's static_inputs_cache")
        if not np.array_equal(m["shared_cache"], g["cache"][0, g["shared_cache_indices"]] if len(g["cache"]) else m["shared_cache"]):
            raise ValueError(f"{mt}: shared cache values differ")
    if strategy is not None and getattr(strategy, "_bcsr_indptr", None) is not None:
        ip, ix = np.asarray(strategy._bcsr_indptr), np.asarray(strategy._bcsr_indices)
        if not (np.array_equal(ip, eng.csr_indptr) and np.array_equal(ix, eng.csr_indices)):
            raise ValueError("CSR pattern differs from the reference's _bcsr_indptr / _bcsr_indices")


def engine_from_reference(runner: Any, setup: Any, strategy: Optional[Any] = None, *, host_only: bool = False,
                          **engine_kwargs) -> B200Engine:
    """``B200Engine`` on the reference's prepared state.  ``host_only=True`` stops before the GPU is touched (the arrays of
    ``vjx_circuit_desc`` are then in ``engine.describe()``); otherwise the returned engine owns a live handle."""
    bundles = {mt: bundle_from_translator(runner._compiled_models[mt]["translator"]) for mt in setup.static_inputs_cache}
    circ = circuit_from_runner(runner)
    lim = any(bool(runner._compiled_models[mt].get("use_device_limiting")) for mt in setup.static_inputs_cache)
    eng = B200Engine(circ, bundles, use_device_limiting=lim, **engine_kwargs)
    md = model_data_from_reference(runner, setup)
    eng._model_data = md
    if host_only:
        eng.describe()
    else:
        eng.prepare()
    _cross_check(eng, md, setup, strategy)
    return eng


def make_b200_full_mna_solver(runner: Any, setup: Any, strategy: Optional[Any] = None, **engine_kwargs):
    """What ``FullMNAStrategy._ensure_full_mna_solver`` (full_mna.py:589-684) would cache for the B200 path: an
    ``nr_solve`` with the reference's signature and result order (solver_factories.py:604-725)

        nr_solve(X_init, vsource_vals, isource_vals, Q_prev, integ_c0, device_arrays, gmin=1e-12, gshunt=0.0, integ_c1=0.0,
                 integ_d1=0.0, dQdt_prev=None, integ_c2=0.0, Q_prev2=None, limit_state_in=None, res_tol_floor=None)
          -> (X, iterations, converged, max_f, Q, dQdt, I_vsource, limit_state_out, max_res_contrib)

    ``device_arrays`` is accepted and ignored: the per-device arrays already live in the handle
    (``vjx_update_model_data`` replaces them for a new corner).  The engine is returned as ``nr_solve.engine``."""
    eng = engine_from_reference(runner, setup, strategy, **engine_kwargs)

    def nr_solve(X_init, vsource_vals, isource_vals, Q_prev, integ_c0, device_arrays=None, gmin=1e-12, gshunt=0.0,
                 integ_c1=0.0, integ_d1=0.0, dQdt_prev=None, integ_c2=0.0, Q_prev2=None, limit_state_in=None,
                 res_tol_floor=None):
        return eng.nr_solve(np.asarray(X_init), np.asarray(vsource_vals), np.asarray(isource_vals), np.asarray(Q_prev),
                            float(integ_c0), float(gmin), float(gshunt), float(integ_c1), float(integ_d1),
                            None if dQdt_prev is None else np.asarray(dQdt_prev), float(integ_c2),
                            None if Q_prev2 is None else np.asarray(Q_prev2),
                            None if limit_state_in is None else np.asarray(limit_state_in),
                            None if res_tol_floor is None else np.asarray(res_tol_floor), stream_ordered=True)

    nr_solve.engine = eng
    return nr_solve

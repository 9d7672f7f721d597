#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE's own Python modules.

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

JAX cannot be installed in the build container, so the reference's modules are
imported on top of ``oracle/refshim/jax`` - a NumPy-backed stand-in for the part of the JAX API that
the transient hot path uses (eager ``lax.while_loop``, functional ``.at[]`` updates, ``vmap`` as a
loop).  Everything that computes below is reference code, unmodified, read from /root/reference:

* ``vajax.analysis.integration`` / ``transient.adaptive`` / ``limiting`` / ``sources`` (unit vectors),
* ``openvaf_jax.OpenVAFToJAX.from_cache`` on this repo's MIR bundles (the reference's code generator
  produces the init / eval functions: model-level vectors),
* ``CircuitEngine`` (setup, node collapse, parameter split, stamp maps), ``make_mna_build_system_fn``
  (COO/dense and CSR-direct), ``_make_nr_solver_common`` and ``FullMNAStrategy.run`` with the
  while-loop body (engine-level vectors: build_system, nr_solve, whole transients with the
  accepted-step sequence).

What the reference cannot provide here is replaced by the thinnest possible stand-in, stated in
tests/golden/README.md: MIR comes from ``vajax_b200.models.fixtures`` (no OpenVAF compiler), the
sparse linear solve is ``scipy.sparse.linalg.splu`` behind the UMFPACK FFI call signature (no
SuiteSparse), the dense solve is ``numpy.linalg.solve`` (LAPACK, as in jaxlib).
The committed .npz files are what travels to the GPU box; this script does not run there.
"""
from __future__ import annotations

import io
import json
import logging
import os
import sys
import warnings
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VAJAX_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
os.environ.setdefault("VAJAX_NO_PROGRESS", "1")
os.environ["VAJAX_CACHE_DIR"] = "/tmp/vajax_golden_cache"

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402

assert jax.__version__.endswith("numpy-shim"), "make_golden.py must run on oracle/refshim/jax"

import openvaf_jax  # noqa: E402  (reference)
import vajax  # noqa: E402,F401  (reference)
from vajax.analysis import openvaf_models as ref_models  # noqa: E402
from vajax.analysis import solver_factories as ref_sf  # noqa: E402
from vajax.analysis.engine import CircuitEngine  # noqa: E402
from vajax.analysis.integration import IntegrationMethod, compute_coefficients  # noqa: E402
from vajax.analysis.limiting import fetlim, pnjlim  # noqa: E402
from vajax.analysis.transient import adaptive as ref_adaptive  # noqa: E402

from vajax_b200.benchmarks import circuits  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402

# openvaf_models.py drops its openvaf_jax import when the Rust crate openvaf_py is missing (:63-70);
# the code generator itself is pure Python, so hand it back.
ref_models.openvaf_jax = openvaf_jax
ref_models.HAS_OPENVAF = True
logging.getLogger("vajax").setLevel(logging.ERROR)
MODELS = ("resistor", "capacitor", "diode", "psp103", "opcov", "psp103w")


# ------------------------------------------------------------------------------------------------
# stand-in for the UMFPACK FFI module (vajax/sparse/umfpack_jax.py: solve(indptr, indices, data, b))
# ------------------------------------------------------------------------------------------------
def _install_sparse_solver_standin():
    import types

    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    def solve(indptr, indices, data, b):
        n = len(b)
        A = sp.csr_matrix((np.asarray(data), np.asarray(indices), np.asarray(indptr)), shape=(n, n))
        return jnp.asarray(spla.splu(A.tocsc()).solve(np.asarray(b)))

    mod = types.ModuleType("vajax.sparse.umfpack_jax")
    mod.solve = solve
    mod.is_available = lambda: True
    mod.clear_cache = lambda: None
    pkg = types.ModuleType("vajax.sparse")
    pkg.__path__ = [os.path.join(REF, "vajax", "sparse")]
    pkg.umfpack_jax = mod
    sys.modules["vajax.sparse"] = pkg
    sys.modules["vajax.sparse.umfpack_jax"] = mod
    ref_sf.is_umfpack_ffi_available = lambda: True


# ------------------------------------------------------------------------------------------------
# compiled_models from MIR bundles (mirror of openvaf_models.compile_openvaf_models:490-545,
# cache-hit branch: the translator comes from OpenVAFToJAX.from_cache)
# ------------------------------------------------------------------------------------------------
def compile_from_bundles():
    compiled = {}
    for mt in MODELS:
        b = get_bundle(mt)
        tr = openvaf_jax.OpenVAFToJAX.from_cache(b)
        dae = tr.get_dae_metadata()
        init_fn, init_meta = tr.translate_init_array()
        names = b["param_names"]
        e2i = {n.lower(): i for i, n in enumerate(names)}
        compiled[mt] = {
            "module": None, "translator": tr, "va_hash": "fixture", "dae_metadata": dae,
            "param_names": names, "param_kinds": b["param_kinds"], "nodes": b["nodes"],
            "collapsible_pairs": b["collapsible_pairs"], "num_collapsible": b["num_collapsible"],
            "init_fn": init_fn, "vmapped_init": jax.jit(jax.vmap(init_fn)),
            "init_param_names": list(init_meta["param_names"]), "init_param_kinds": list(init_meta["param_kinds"]),
            "cache_size": init_meta["cache_size"], "cache_mapping": init_meta["cache_mapping"],
            "init_param_defaults": init_meta.get("param_defaults", {}),
            "init_to_eval_indices": jnp.array([e2i.get(n.lower(), -1) for n in init_meta["param_names"]], dtype=jnp.int32),
            "uses_simparam_gmin": tr.uses_simparam_gmin, "uses_analysis": tr.uses_analysis,
            "analysis_type_map": tr.analysis_type_map,
        }
    return compiled


def make_engine(circ: dict, compiled: dict, limiting: bool) -> CircuitEngine:
    import copy

    eng = CircuitEngine(Path("/nonexistent/fixture.sim"))
    eng.devices = copy.deepcopy(circ["devices"])
    eng.num_nodes = circ["num_nodes"]
    eng.node_names = dict(circ["node_names"])
    used = {d["model"] for d in eng.devices if d.get("is_openvaf")}
    eng._compiled_models = {k: dict(v) for k, v in compiled.items() if k in used}
    eng._has_openvaf_devices = True
    eng.use_device_limiting = limiting
    ap = circ["analysis"]
    o = eng.options
    o.step, o.stop = float(ap["step"]), float(ap["stop"])
    if "maxstep" in ap:
        o.maxstep = float(ap["maxstep"])
    o.icmode = ap.get("icmode", "op")
    o.tran_method = IntegrationMethod.from_string(ap.get("tran_method", "trap"))
    for k in ("tran_lteratio", "nr_convtol", "reltol", "abstol"):
        if k in ap:
            setattr(o, k, float(ap[k]))
    eng._sync_options_to_analysis_params()
    eng._device_collapse_decisions = ref_models.compute_early_collapse_decisions(eng.devices, eng._compiled_models)
    eng._simulation_temperature = float(circ.get("temperature_k", 300.15))
    eng.options.temp = eng._simulation_temperature - 273.15
    return eng


sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import CIRCUITS  # noqa: E402  (one table for the generator and the tests)


def unit_vectors(out: dict):
    """Integration coefficients, predictor, LTE, limiting: reference functions on seeded inputs."""
    rng = np.random.default_rng(11)
    for m in ("be", "trap", "gear2"):
        c = compute_coefficients(IntegrationMethod.from_string(m), 1e-12)
        out[f"integ_{m}"] = np.array([c.c0, c.c1, c.d1, getattr(c, "c2", 0.0) or 0.0, c.error_coeff], dtype=np.float64)
    cases = []
    for hc in (1, 2, 3, 4):
        for _ in range(4):
            Vh = rng.normal(size=(4, 6))
            dth = rng.uniform(1e-12, 5e-12, 4)
            ndt = float(rng.uniform(1e-12, 5e-12))
            Vp, err = ref_adaptive.predict_voltage_jax(jnp.asarray(Vh), jnp.asarray(dth), jnp.asarray(hc), jnp.asarray(ndt), 2)
            cases.append(np.concatenate([[hc, ndt], dth, Vh.ravel(), np.asarray(Vp).ravel(), [float(err)]]))
    out["predict_cases"] = np.array(cases)
    cfg = ref_adaptive.AdaptiveConfig()
    lte = []
    for hc in (2, 3, 4):
        for ec in (-0.5, 1.0 / 24.0, -2.0 / 9.0):
            for _ in range(3):
                Vn = rng.normal(size=6)
                Vp = Vn + rng.normal(scale=10 ** rng.uniform(-6, -1), size=6)
                pe = float(rng.normal())
                dt = float(rng.uniform(1e-12, 1e-9))
                Vmax = np.abs(rng.normal(size=6)) * 2
                dt_new, norm = ref_adaptive.compute_lte_timestep_jax(
                    jnp.asarray(Vn), jnp.asarray(Vp), jnp.asarray(pe), jnp.asarray(dt), jnp.asarray(hc), cfg, ec,
                    V_max_historic=jnp.asarray(Vmax))
                lte.append(np.concatenate([[hc, ec, pe, dt], Vn, Vp, Vmax, [float(dt_new), float(norm)]]))
    out["lte_cases"] = np.array(lte)
    vn = rng.uniform(-3, 3, 400)
    vo = rng.uniform(-3, 3, 400)
    out["pnjlim_in"] = np.stack([vn, vo])
    out["pnjlim_out"] = np.asarray(pnjlim(jnp.asarray(vn), jnp.asarray(vo), 0.02585, 0.6145))
    out["fetlim_out"] = np.asarray(fetlim(jnp.asarray(vn), jnp.asarray(vo), 0.5))


def source_vectors(out: dict):
    from vajax.analysis.sources import build_source_fn

    devs = [
        dict(name="v_dc", model="vsource", nodes=[1, 0], params=dict(dc=1.5, type="dc")),
        dict(name="v_pulse", model="vsource", nodes=[1, 0], params=dict(type="pulse", val0=0.1, val1=1.2, rise=1e-9, fall=2e-9, width=5e-9, period=2e-8, delay=3e-9)),
        dict(name="v_sine", model="vsource", nodes=[1, 0], params=dict(type="sine", sinedc=0.5, ampl=2.0, freq=1e8, phase=0.3)),
        dict(name="v_pwl", model="vsource", nodes=[1, 0], params=dict(type="pwl", wave=[0, 0, 1e-9, 1, 5e-9, -1, 2e-8, 0.5])),
        dict(name="v_exp", model="vsource", nodes=[1, 0], params=dict(type="exp", val0=0, val1=1, delay=1e-9, td2=4e-9, tau1=1e-9, tau2=2e-9)),
        dict(name="i_pulse", model="isource", nodes=[0, 1], params=dict(type="pulse", val0=0, val1=1e-5, delay=1e-9, rise=1e-9, fall=1e-9, width=1e-9, period=1.0)),
    ]
    eng = CircuitEngine(Path("/nonexistent/fixture.sim"))
    fn = build_source_fn(devs, eng.parse_spice_number)
    ts = np.concatenate([np.linspace(0, 6e-8, 241), [3e-9, 4e-9, 9e-9, 1.1e-8]])
    names = [d["name"] for d in devs]
    vals = np.array([[float(fn(jnp.asarray(t))[n]) for n in names] for t in ts])
    out["source_times"] = ts
    out["source_values"] = vals
    out["source_defs"] = np.array(json.dumps(devs))


def model_vectors(compiled: dict, out: dict):
    """The reference code generator's init and eval functions on seeded instances."""
    rng = np.random.default_rng(5)
    for mt in MODELS:
        b = get_bundle(mt)
        c = compiled[mt]
        n = 6
        names = c["init_param_names"]
        defaults = {k.lower(): v for k, v in b["param_defaults"].items()}
        init_in = np.zeros((n, len(names)))
        for j, (nm, kd) in enumerate(zip(names, c["init_param_kinds"])):
            if kd == "temperature":
                init_in[:, j] = rng.uniform(250, 400, n)
            elif kd == "param_given":
                init_in[:, j] = 1.0
            else:
                d = defaults.get(nm.lower(), 0.0)
                init_in[:, j] = d * rng.uniform(0.7, 1.3, n) if d != 0 else rng.choice([0.0, 1.0, 5.0], n)
                if nm.lower() == "type":
                    init_in[:, j] = rng.choice([-1.0, 1.0], n)
        cache, collapse = c["vmapped_init"](jnp.asarray(init_in))
        out[f"model_{mt}_init_in"] = init_in
        out[f"model_{mt}_cache"] = np.asarray(cache)
        out[f"model_{mt}_collapse"] = np.asarray(collapse, dtype=np.float64)


def engine_vectors(name: str, limiting: bool, compiled: dict, out: dict, sparse: bool):
    circ = CIRCUITS[name]()
    tag = f"{name}_{'lim' if limiting else 'nolim'}_{'csr' if sparse else 'dense'}"
    eng = make_engine(circ, compiled, limiting)
    eng.prepare(use_sparse=sparse, backend="cpu")
    strat = eng._prepared_strategy
    setup = strat.ensure_setup()
    nr_solve = strat._ensure_full_mna_solver(setup)
    n_total, nu = setup.n_total, setup.n_unknowns
    sdd = setup.source_device_data
    n_vs = len(sdd.get("vsource", {}).get("names", []))
    out[f"{tag}_sizes"] = np.array([n_total, nu, n_vs, strat._total_limit_states if hasattr(strat, "_total_limit_states") else -1])
    bs = strat._cached_build_system_jit
    dev_arrays = strat._device_arrays_full_mna
    # ---- build_system on seeded inputs (same seeds as tests/test_gpu_parity.py)
    rng = np.random.default_rng(7)
    X = np.zeros(n_total + n_vs)
    X[1:] = rng.uniform(-0.3, 1.2, X.size - 1)
    Qp, dq, q2 = (rng.normal(0, 1e-13, nu) for _ in range(3))
    src = eng._build_source_fn()
    jse = strat._make_jit_source_eval(setup, src)
    vs, isv = jse(jnp.asarray(2.5e-10))
    nlim = int(getattr(strat, "_total_limit_states", 0) or 0)
    ls = rng.uniform(0, 0.7, nlim)
    out[f"{tag}_bs_X"] = X
    out[f"{tag}_bs_state"] = np.stack([Qp, dq, q2]) if nu else np.zeros((3, 0))
    out[f"{tag}_bs_src"] = np.concatenate([np.asarray(vs), np.asarray(isv)])
    out[f"{tag}_bs_lim"] = ls
    for ci, (c0, c1, d1, c2, it) in enumerate([(0.0, 0.0, 0.0, 0.0, 0), (2e12, -2e12, -1.0, 0.0, 1), (1.5e9, -2e9, 0.0, 0.5e9, 3)]):
        J, f, Q, Ivs, lso, mrc = bs(jnp.asarray(X), vs, isv, jnp.asarray(Qp), c0, dev_arrays, 1e-12, 1e-9, c1, d1,
                                    jnp.asarray(dq), c2, jnp.asarray(q2), jnp.asarray(ls), it)
        out[f"{tag}_bs{ci}_J"] = np.asarray(J)
        out[f"{tag}_bs{ci}_f"] = np.asarray(f)
        out[f"{tag}_bs{ci}_Q"] = np.asarray(Q)
        out[f"{tag}_bs{ci}_Ivs"] = np.asarray(Ivs)
        out[f"{tag}_bs{ci}_lim"] = np.asarray(lso)
        out[f"{tag}_bs{ci}_mrc"] = np.asarray(mrc)
    if sparse:
        out[f"{tag}_csr_indptr"] = np.asarray(strat._bcsr_indptr if hasattr(strat, "_bcsr_indptr") else [])
        out[f"{tag}_csr_indices"] = np.asarray(strat._bcsr_indices if hasattr(strat, "_bcsr_indices") else [])
    # ---- whole transient through FullMNAStrategy.run
    buf = io.StringIO()
    with redirect_stdout(buf):
        times, V_out, stats = strat.run(eng._prepared_t_stop, eng._prepared_dt, None)
    n = int(stats["n_steps"]) if "n_steps" in stats else int(stats.get("accepted_steps", 0))
    out[f"{tag}_tran_times"] = np.asarray(times)[:n]
    out[f"{tag}_tran_V"] = np.asarray(V_out)[:n]
    keep = {k: (float(v) if np.isscalar(v) or np.ndim(v) == 0 else None) for k, v in stats.items()
            if isinstance(v, (int, float, np.integer, np.floating)) or np.ndim(v) == 0 and not isinstance(v, (str, dict))}
    out[f"{tag}_tran_stats"] = np.array(json.dumps({k: v for k, v in keep.items() if v is not None}))
    print(f"  {tag}: n_total={n_total} steps={n} rejected={stats.get('rejected_steps')} nr={stats.get('total_nr_iterations')}")


def main():
    _install_sparse_solver_standin()
    compiled = compile_from_bundles()
    if not sys.argv[1:] or "units" in sys.argv[1:]:
        out: dict = {}
        unit_vectors(out)
        source_vectors(out)
        model_vectors(compiled, out)
        np.savez_compressed(os.path.join(HERE, "reference_units.npz"), **out)
        print("wrote reference_units.npz:", len(out), "arrays")
    only = set(sys.argv[1:])
    for name in CIRCUITS:
        if only and name not in only:
            continue
        eo: dict = {}
        for limiting in (False, True):
            for sparse in (False, True):
                engine_vectors(name, limiting, compiled, eo, sparse)
        np.savez_compressed(os.path.join(HERE, f"reference_{name}.npz"), **eo)
        print(f"wrote reference_{name}.npz:", len(eo), "arrays")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors for the DC continuation: the REFERENCE's ``vajax/analysis/homotopy.py``
(``gmin_stepping``, ``source_stepping``, ``run_homotopy_chain``) driving the REFERENCE's ``nr_solve``
(``FullMNAStrategy._ensure_full_mna_solver``), on the NumPy JAX shim like make_golden.py.

    python tests/golden/make_homotopy_golden.py      # needs /root/reference (build container only)

Each case records the sequence of solves the continuation made (source scale, gmin, gshunt, NR iterations,
converged) and its outcome (converged, method, totals, final vector).  NR failures are provoked the way a user
would meet them: a small ``tran_itl`` (the iteration cap ``nr_solve`` is built with, full_mna.py:610) and the
mid-rail start vector or a seeded random one; the continuation's thresholds come from ``HomotopyConfig.max_iterations``.
Output: tests/golden/reference_homotopy.npz.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up the shim + reference import paths)

from vajax.analysis import homotopy as ref_h  # noqa: E402  (reference)

jnp = mg.jnp

# (case name, circuit, tran_itl, HomotopyConfig overrides, which entry point[, seed of a random start vector])
CASES = [
    ("ring_chain_rand_itl3", "ring", 3, dict(max_iterations=3, gmin_max_steps=40, source_max_steps=40), "chain", 11),
    ("ring_gdev_rand_itl4", "ring", 4, dict(max_iterations=4, gmin_factor=1000.0), "gdev", 12),
    ("ring_gshunt_rand_itl4", "ring", 4, dict(max_iterations=4, gdev_start=1e-6, gmin_max=1e-2), "gshunt", 13),
    ("ring_src_rand_itl3", "ring", 3, dict(max_iterations=3, source_step=0.5, source_step_min=0.01), "src", 14),
    ("ring_chain_rand_itl2", "ring", 2, dict(max_iterations=2, gmin_max_steps=25, source_max_steps=25), "chain", 15),
    ("mul3_chain_rand_itl5", "mul3", 5, dict(max_iterations=5, gmin_max_steps=30, source_max_steps=30), "chain", 16),
    ("ring_chain_default", "ring", 20, {}, "chain"),
    ("ring_src_itl4", "ring", 4, dict(max_iterations=4, source_step=0.3), "src"),
    ("clipper_chain_itl3", "clipper", 3, dict(max_iterations=3), "chain"),
    ("mul3_gdev_itl6", "mul3", 6, dict(max_iterations=6, gdev_target=1e-9), "gdev"),
    ("mul3_src_itl6", "mul3", 6, dict(max_iterations=10, source_step=0.25, source_max_steps=30), "src"),
]


def run_case(name, circ_name, itl, over, entry, compiled, out, seed=None):
    circ = mg.CIRCUITS[circ_name]()
    eng = mg.make_engine(circ, compiled, False)
    eng.options.tran_itl = int(itl)
    eng._sync_options_to_analysis_params()
    eng.prepare(use_sparse=True, backend="cpu")
    strat = eng._prepared_strategy
    setup = strat.ensure_setup()
    nr_solve = strat._ensure_full_mna_solver(setup)
    dev_arrays = strat._device_arrays_full_mna
    n_total, nu = setup.n_total, setup.n_unknowns
    src = eng._build_source_fn()
    vs, isv = strat._make_jit_source_eval(setup, src)(jnp.asarray(0.0))
    n_vs = int(np.asarray(vs).size)
    X0 = np.zeros(n_total + n_vs)
    X0[:n_total] = np.asarray(strat._init_mid_rail(setup, n_total))
    if seed is not None:   # a start vector away from the symmetric mid-rail point, so that Newton needs its iterations
        X0[1:n_total] = np.random.default_rng(seed).uniform(0.0, 1.2, n_total - 1)
    trace = []

    def logged(X, v, i, Q, c0, da, gmin=1e-12, gshunt=0.0, *rest):
        res = nr_solve(X, v, i, Q, c0, da, gmin, gshunt, *rest)
        vs_abs = float(np.sum(np.abs(np.asarray(vs)))) or 1.0
        trace.append((float(np.sum(np.abs(np.asarray(v)))) / vs_abs, float(gmin), float(gshunt), int(res[1]), bool(res[2])))
        return res

    cfg = ref_h.HomotopyConfig(gmin=1e-12, **over)
    Qz = jnp.zeros(nu)
    if entry == "chain":
        r = ref_h.run_homotopy_chain(logged, jnp.asarray(X0), vs, isv, Qz, dev_arrays, cfg)
    elif entry in ("gdev", "gshunt"):
        r = ref_h.gmin_stepping(logged, jnp.asarray(X0), vs, isv, Qz, dev_arrays, 1.0, cfg, mode=entry)
    else:
        r = ref_h.source_stepping(logged, jnp.asarray(X0), vs, isv, Qz, dev_arrays, cfg)
    out[f"{name}_X0"] = X0
    out[f"{name}_V"] = np.asarray(r.V, dtype=np.float64)
    out[f"{name}_trace"] = np.array(trace, dtype=np.float64).reshape(-1, 5)
    out[f"{name}_result"] = np.array(json.dumps(dict(
        converged=bool(r.converged), method=r.method, iterations=int(r.iterations), homotopy_steps=int(r.homotopy_steps),
        final_gmin=float(r.final_gmin), final_source_scale=float(r.final_source_scale),
        circuit=circ_name, tran_itl=int(itl), config=over, entry=entry, seed=seed)))
    print(f"  {name}: solves={len(trace)} converged={r.converged} method={r.method} iters={r.iterations} "
          f"ok/fail={sum(1 for t in trace if t[4])}/{sum(1 for t in trace if not t[4])}")


def main():
    mg._install_sparse_solver_standin()
    compiled = mg.compile_from_bundles()
    out: dict = {}
    only = set(sys.argv[1:])
    for name, circ_name, itl, over, entry, *seed in CASES:
        if only and name not in only:
            continue
        run_case(name, circ_name, itl, over, entry, compiled, out, *seed)
    if not only:
        np.savez_compressed(os.path.join(HERE, "reference_homotopy.npz"), **out)
        print("wrote reference_homotopy.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Build container only (needs /root/reference): prepare the REFERENCE's CircuitEngine / FullMNAStrategy on the NumPy
JAX shim (same machinery as make_golden.py), hand its objects to vajax_b200.analysis.reference_adapter, and check
the arrays that would go into vjx_circuit_desc against (a) the reference's own arrays (the adapter's cross-check)
and (b) what this repo derives on its own from the flat circuit (oracle/engine.py, which the GPU engine is tested
against).  Host only - no GPU, no plugin compilation.  Prints one JSON line per circuit.

    python tests/golden/check_reference_adapter.py ring clipper mul3
"""
import json
import sys

import numpy as np

import make_golden as mg  # noqa: E402  (installs the shim, imports the reference)
from golden_util import CIRCUITS  # noqa: E402

from oracle.engine import OracleEngine  # noqa: E402
from vajax_b200.analysis.reference_adapter import engine_from_reference  # noqa: E402


def check(name: str, limiting: bool, compiled: dict) -> dict:
    circ = CIRCUITS[name]()
    ref = mg.make_engine(circ, compiled, limiting)
    ref.prepare(use_sparse=True, backend="cpu")
    strat = ref._prepared_strategy
    setup = strat.ensure_setup()
    strat._ensure_full_mna_solver(setup)
    eng = engine_from_reference(ref, setup, strat, host_only=True)     # raises on any mismatch with the reference
    d = eng.describe()
    o = OracleEngine(circ, {k: mg.get_bundle(k) for k in mg.MODELS}, use_device_limiting=limiting)
    o.setup()
    assert (d["n_total"], d["n_unknowns"], d["n_vs"], d["nse"]) == (o.n_total, o.n_unknowns, o.n_vs, o.nse)
    assert np.array_equal(d["csr_indptr"], o.csr_indptr) and np.array_equal(d["csr_indices"], o.csr_indices)
    assert sorted(d["noi_indices"].tolist()) == sorted(int(i) for i in o.noi_indices)
    n_arrays = 0
    for mt, m in d["models"].items():
        om = o.models[mt]
        assert m["shared_indices"] == list(om["shared_indices"]) and m["varying_indices"] == list(om["varying_indices"])
        assert m["shared_cache_indices"] == list(om["shared_cache_indices"])
        assert np.array_equal(m["shared_params"], om["shared_params"])
        nonv = [k for k, c in enumerate(m["varying_indices"]) if k not in set(int(x) for x in om["voltage_positions"])]
        assert np.array_equal(m["dev_params"], om["device_params"][:, nonv])
        assert np.array_equal(m["dev_cache"], om["cache"][:, m["varying_cache_indices"]])
        assert np.array_equal(m["shared_cache"], om["cache"][0, m["shared_cache_indices"]])
        for a, b in (("vnode1", "vnode1"), ("vnode2", "vnode2"), ("res_indices", "res_indices"),
                     ("jac_row_indices", "jac_row_indices"), ("jac_col_indices", "jac_col_indices")):
            assert np.array_equal(m[a], om[b]), (mt, a)
            n_arrays += 1
        assert m["num_limit_states"] == om["num_limit_states"]
    return {"circuit": name, "limiting": limiting, "n_total": int(d["n_total"]), "nse": int(d["nse"]), "models": sorted(d["models"]),
            "arrays_equal": n_arrays}


if __name__ == "__main__":
    mg._install_sparse_solver_standin()
    compiled = mg.compile_from_bundles()
    for nm in (sys.argv[1:] or ["ring", "clipper"]):
        for lim in (False, True):
            print(json.dumps(check(nm, lim, compiled)), flush=True)

#!/usr/bin/env python
"""GPU probe of the mul64-scale path: prepare, a few transient steps, per-phase times, LU stats and
a linear-solve residual check.  python scripts/mul64_probe.py [nbits] [steps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vajax_b200.analysis.engine import B200Engine  # noqa: E402
from vajax_b200.benchmarks.circuits import mul64_circuit  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402

nbits = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
model = sys.argv[3] if len(sys.argv) > 3 else "psp103w"
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103", "psp103w")}
t0 = time.perf_counter()
circ = mul64_circuit(nbits, mos_model=model)
print(f"netlist: {len(circ['devices'])} devices, {circ['num_nodes']} nodes, {time.perf_counter() - t0:.1f}s", flush=True)
t0 = time.perf_counter()
e = B200Engine(circ, bundles, use_device_graph=True).prepare()
print(f"prepare {time.perf_counter() - t0:.1f}s timings={ {k: round(v, 2) for k, v in e.timings.items()} }", flush=True)
print(f"n_aug={e.n_aug} nse={e.nse}", flush=True)
t0 = time.perf_counter()
e.tran_begin(max_steps=1 + K)
print(f"tran_begin {time.perf_counter() - t0:.1f}s", flush=True)
per = []
for s in range(1, 1 + K):
    t0 = time.perf_counter()
    a = e.tran_run(step_limit=s + 1)
    wall = time.perf_counter() - t0
    sc = e.tran_scalars()
    per.append(e.tran_run_ms()[0])
    print(f"step {s}: attempts={a} device_ms={per[-1]:.2f} wall_ms={wall * 1e3:.1f} t={sc['t']:.3e} dt={sc['dt']:.3e} "
          f"nr={sc['total_nr']} rej={sc['rejected']} err={sc['error_flags']}", flush=True)
    if a == 0 or sc["error_flags"]:
        break
print("lu:", json.dumps(e.lu_stats()), flush=True)
names = ["device_eval", "jacobian_stamp", "residual", "lu_factor_fwd", "(fused)", "lu_usolve", "nr_update"]
for i, nm in enumerate(names):
    ms, nl = e.time_phase(i, 3)
    print(f"phase {nm:15s} {ms:9.3f} ms ({nl} launches)", flush=True)
# linear-solve residual on the current Jacobian
import scipy.sparse as sp  # noqa: E402

X, Q, dQ = e.tran_state()
vs, isv = e.eval_sources(sc["t"])
nu = e.n_unknowns
c0 = 2.0 / max(sc["dt"], 1e-15)
out = e.build_system(X, vs, isv, Q, c0, 1e-12, 0.0, -c0, -1.0, dQ, 0.0, None, None, 1)
csr = out[0].copy()
rows = np.repeat(np.arange(e.n_aug), np.diff(e.csr_indptr))
noi = np.zeros(e.n_aug, bool)
noi[e.noi_indices - 1] = True
csr[noi[rows] | ((rows < nu) & noi[e.csr_indices])] = 0.0
csr[noi[rows] & (rows == e.csr_indices)] = 1.0
A = sp.csr_matrix((csr, e.csr_indices, e.csr_indptr), shape=(e.n_aug, e.n_aug))
rhs = np.random.default_rng(0).normal(size=e.n_aug)
x = e.lu_solve(csr, rhs)
r = A @ x - rhs
print(f"lu_solve residual max|Ax-b|={np.max(np.abs(r)):.3e}  scale={max(1.0, np.max(np.abs(A @ np.abs(x)))):.3e}", flush=True)
print(f"steady ms/step (steps 3..): {np.mean(per[2:]) if len(per) > 2 else float('nan'):.2f}")

#!/usr/bin/env python
"""Golden vector at BASELINE.json's largest size: the first accepted steps of the mul64 transient computed by
the CPU oracle (about 30 s per NR iteration, so this runs once, here, not in the test suite).

    python tests/golden/make_mul64_steps.py      ->  tests/golden/mul64_first_steps.npz

Stored: the accepted-step records (step, t_next, dt, nr_iters, accept), the output times, the waveform rows at
every 64th external node, and the max-norm / sum of every full row (a checksum of the columns not stored).
tests/test_gpu_fullsize.py::test_mul64_first_steps_match_the_oracle_golden compares the CUDA path with it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.engine import OracleEngine  # noqa: E402
from vajax_b200.benchmarks.circuits import mul64_circuit  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402

MODEL = sys.argv[1] if len(sys.argv) > 1 else "psp103"      # psp103 (light surrogate) or psp103w (PSP103-weight)
# rows: row 0 = initial condition, then the accepted steps.  3 = the interpreter's golden (two steps); 16 = the
# reference's 15-step timing window (README.md:80), generated with the oracle's accelerations for long runs
# (compiled host model functions + cached ordering: oracle/host_compiled.py, tests/test_host_compiled.py)
ROWS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ACCEL = ROWS > 3
OUT = ("mul64_first_steps.npz" if MODEL == "psp103" else f"mul64_first_steps_{MODEL}.npz") if not ACCEL \
    else f"mul64_steps{ROWS - 1}_{MODEL}.npz"
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103", "psp103w")}
o = OracleEngine(mul64_circuit(64, mos_model=MODEL), bundles)
rec = []
if ACCEL:
    from oracle import host_compiled  # noqa: E402

    o.reuse_ordering = True
    with host_compiled.installed():
        t, V, st = o.run_transient(max_steps=ROWS, record=rec)
else:
    t, V, st = o.run_transient(max_steps=ROWS, record=rec)
n = st["n_steps"]
cols = np.arange(0, V.shape[1], 64)
np.savez_compressed(
    os.path.join(ROOT, "tests", "golden", OUT),
    times=np.asarray(t[:n]), cols=cols, V_cols=np.asarray(V[:n][:, cols]),
    row_absmax=np.max(np.abs(V[:n]), axis=1), row_sum=np.sum(V[:n], axis=1),
    records=np.array([[r["step"], r["t_next"], r["dt"], r["nr_iters"], float(r["accept"])] for r in rec]),
    n_steps=n, total_nr=st["total_nr_iterations"], rejected=st["rejected_steps"], n_external=V.shape[1])
print("rows", n, "nr", st["total_nr_iterations"], "records", len(rec))

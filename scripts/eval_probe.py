#!/usr/bin/env python
"""Run only the device-evaluation phase of a benchmark circuit a few times (for `ncu -k vjx_eval_kernel`).
python scripts/eval_probe.py [nbits|64] [model] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vajax_b200.analysis.engine import B200Engine  # noqa: E402
from vajax_b200.benchmarks.circuits import mul64_circuit, multiplier_circuit  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402

nbits = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = sys.argv[2] if len(sys.argv) > 2 else "psp103w"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103", "psp103w")}
circ = mul64_circuit(nbits, mos_model=model) if nbits == 64 else multiplier_circuit(nbits, mos_model=model)
t0 = time.perf_counter()
e = B200Engine(circ, bundles, use_device_graph=False).prepare()
print(f"prepare {time.perf_counter() - t0:.1f}s n_aug={e.n_aug} nse={e.nse}", flush=True)
e.tran_begin(max_steps=4)
e.tran_run(step_limit=2)
for pid, nm in ((0, "device_eval"), (1, "jacobian_stamp"), (2, "residual")):
    ms, nl = e.time_phase(pid, reps)
    print(f"phase {nm}: {ms:.4f} ms ({nl} launches)", flush=True)
e.close()

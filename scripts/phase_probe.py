#!/usr/bin/env python
"""Per-phase times of one NR iteration (vjx_time_phase, CUDA events) + ms/step of a short device loop.
python scripts/phase_probe.py [nbits|64] [model] [steps]   (env knobs: VJX_DT_CLUSTER, VJX_EVAL_SLICES, ...)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vajax_b200.analysis.engine import B200Engine  # noqa: E402
from vajax_b200.benchmarks.circuits import mul64_circuit, multiplier_circuit  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402

nbits = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = sys.argv[2] if len(sys.argv) > 2 else "psp103w"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103", "psp103w")}
circ = mul64_circuit(nbits, mos_model=model) if nbits == 64 else multiplier_circuit(nbits, mos_model=model)
e = B200Engine(circ, bundles, use_device_graph=True).prepare()
e.run_transient(max_steps=8, read_outputs=False)
names = ["device_eval", "jacobian_stamp", "residual", "lu_factor_fwd", "(fused)", "lu_usolve", "nr_update"]
out = {}
for i, nm in enumerate(names):
    if i == 4:
        continue
    ms, nl = e.time_phase(i, 20 if nbits < 64 else 5)
    out[nm] = round(ms * 1e3, 2)
print("phases_us", json.dumps(out), flush=True)
e.tran_begin(max_steps=1 + steps)
e.tran_run(step_limit=1 + steps)
sc = e.tran_scalars()
print(f"ms/step {e.tran_run_ms()[0] / max(sc['step_idx'] - 1, 1):.4f} steps {sc['step_idx'] - 1} nr {sc['total_nr']} rej {sc['rejected']} err {sc['error_flags']}", flush=True)
print("lu", json.dumps(e.lu_stats()), flush=True)
e.close()

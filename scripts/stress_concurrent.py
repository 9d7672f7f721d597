#!/usr/bin/env python
"""Stress test of concurrent whole-circuit replicas on one GPU: the waveforms of run_corners(concurrent=k)
must equal the sequential sweep bit for bit, every time.  python scripts/stress_concurrent.py [reps] [k]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vajax_b200.analysis import corners as cz  # noqa: E402
from vajax_b200.analysis.engine import B200Engine  # noqa: E402
from vajax_b200.benchmarks import circuits  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
bundles = {n: get_bundle(n) for n in ("resistor", "capacitor", "diode", "psp103")}
keep = B200Engine(circuits.multiplier_circuit(16), bundles, use_device_graph=True).prepare()
keep.run_transient(max_steps=5)
circ = circuits.multiplier_circuit(4, stop=4e-10, step=2e-12)
cs = cz.sweep_corners(64)[:6]
a = cz.run_corners(circ, bundles, cs, use_device_graph=True)
bad = 0
for r in range(reps):
    b = cz.run_corners(circ, bundles, cs, use_device_graph=True, concurrent=k)
    for ra, rb in zip(a.results, b.results):
        same = ra.times.shape == rb.times.shape and np.array_equal(ra.times, rb.times) and np.array_equal(ra.voltages, rb.voltages)
        if not same or not rb.converged:
            bad += 1
            n = min(len(ra.times), len(rb.times))
            dv = np.abs(ra.voltages[:n] - rb.voltages[:n]).max() if n else float("nan")
            first = int(np.argmax(np.any(ra.voltages[:n] != rb.voltages[:n], axis=1))) if n else -1
            print(f"rep {r} corner {ra.corner.name}: steps {len(ra.times)} vs {len(rb.times)}, max|dV| {dv:.3e}, first differing row {first}, "
                  f"flags {rb.stats.get('error_flags')} nr {ra.stats.get('total_nr_iterations')} vs {rb.stats.get('total_nr_iterations')}", flush=True)
print(f"{reps} repetitions x {len(cs)} corners, concurrent={k}: {bad} mismatches")
keep.close()

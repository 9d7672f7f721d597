#!/usr/bin/env python
"""Run a few transient steps of a benchmark circuit with the host-polled loop (one kernel launch per
phase, the form ncu can list).  python scripts/run_steps.py [nbits] [steps] [graph]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vajax_b200.analysis.engine import B200Engine  # noqa: E402
from vajax_b200.benchmarks import circuits  # noqa: E402
from vajax_b200.models.fixtures import get_bundle  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103")}
circ = circuits.mul64_circuit(nb) if nb == 64 else circuits.multiplier_circuit(nb)
e = B200Engine(circ, bundles, use_device_graph=graph).prepare()
t, V, st = e.run_transient(max_steps=steps)
print({k: st[k] for k in ("n_steps", "total_nr_iterations", "rejected_steps") if k in st}, "launches", e.launch_count())
e.close()

import sys, time
sys.path.insert(0, '/root/repo')
from vajax_b200.analysis.engine import B200Engine
from vajax_b200.analysis import corners as cz
from vajax_b200.benchmarks import circuits
from vajax_b200.models.fixtures import get_bundle
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103", "psp103w")}
circ = circuits.multiplier_circuit(16, mos_model="psp103w")
cs = cz.sweep_corners(64)[:3]
import cProfile, pstats
for i, c in enumerate(cs):
    t0 = time.perf_counter()
    if i == 2:
        pr = cProfile.Profile(); pr.enable()
    e = B200Engine(cz.apply_corner(circ, c), bundles, use_device_graph=True, stream="own").prepare()
    if i == 2:
        pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
    t1 = time.perf_counter()
    e.run_transient(max_steps=3)
    t2 = time.perf_counter()
    print(i, c.name, f"prepare {t1-t0:.2f}s first 3 steps (graph build, LU analysis) {t2-t1:.2f}s", {k: round(v, 3) for k, v in e.timings.items()}, flush=True)
    e.close()

import sys, os
sys.path.insert(0, os.getcwd())
from vajax_b200.analysis.engine import B200Engine
from vajax_b200.benchmarks import circuits
from vajax_b200.models.fixtures import get_bundle
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103")}
e = B200Engine(circuits.multiplier_circuit(16), bundles, use_device_graph=True).prepare()
e.run_transient(max_steps=12)
for i,nm in [(0,"eval"),(1,"stamp"),(2,"residual"),(6,"update")]:
    print(nm, e.time_phase(i, 50))
e.tran_begin(max_steps=101); e.tran_run(step_limit=101); print("ms/step", e.tran_run_ms()[0]/100, e.tran_scalars()["total_nr"])
e.close()

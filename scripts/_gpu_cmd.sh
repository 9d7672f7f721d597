mkdir -p gpurun_out
for b in 128 64 32; do echo block=$b; VJX_EVAL_BLOCK=$b python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from vajax_b200.analysis.engine import B200Engine
from vajax_b200.benchmarks import circuits
from vajax_b200.models.fixtures import get_bundle
bundles = {k: get_bundle(k) for k in ("resistor", "capacitor", "diode", "psp103")}
e = B200Engine(circuits.multiplier_circuit(16), bundles, use_device_graph=True).prepare()
e.run_transient(max_steps=12)
print("eval ms", e.time_phase(0, 50))
e.tran_begin(max_steps=101); e.tran_run(step_limit=101); print("ms/step", e.tran_run_ms()[0]/100)
e.close()
PY
done

mkdir -p gpurun_out
T=s6j
./scripts/fp64_peak > gpurun_out/r1_${T}_fp64_peak.txt 2>&1; cat gpurun_out/r1_${T}_fp64_peak.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5
python bench.py --mul64-steps 0 --no-cpu-baseline > gpurun_out/r1_${T}_bench.json 2> gpurun_out/r1_${T}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r1_${T}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r1_${T}_bench.json')); print(d['value'], d['value_l2_resident'], d['sweep_concurrent'], d['e2e'])"

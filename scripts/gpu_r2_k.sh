set -x
mkdir -p gpurun_out
# FMA contraction in the model kernels: speed and parity (information only; the default stays -fmad=false)
VJX_MODEL_FMAD=1 timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py -m gpu -q -k "ringw or mul2w or mul3 or ring" > gpurun_out/r2_m_pytest_fmad.log 2>&1; echo "pytest fmad rc=$?" >> gpurun_out/r2_m_pytest_fmad.log
tail -4 gpurun_out/r2_m_pytest_fmad.log
V=gpurun_out/r2_m_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 900 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -E "phases_us|ms/step|Error|error" >> $V; }
NB=16
run VJX_MODEL_FMAD=1
NB=64
run VJX_MODEL_FMAD=1
cat $V
timeout 900 python bench.py --steps 30 --warmup 3 --corners 2 --no-cpu-baseline > gpurun_out/r2_m_bench_short.json 2> gpurun_out/r2_m_bench_short.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_m_bench_short.json').read())
print(d['value'], d['e2e'], d['mul64']['value'], d['mul64']['e2e'])
PY

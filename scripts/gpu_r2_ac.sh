set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu_structures.py -m gpu -x -q -k "block_rows" > gpurun_out/r2_ai_pytest.log 2>&1; tail -3 gpurun_out/r2_ai_pytest.log
V=gpurun_out/r2_ai_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 600 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -v "^lu " >> $V; }
NB=64
run VJX_HV_VERBOSE=1 VJX_BR_WORK=100000
run VJX_HV_VERBOSE=1 VJX_BR_WORK=150000
run VJX_HV_VERBOSE=1 VJX_BR_WORK=60000
VJX_BR_WORK=100000 timeout 600 python scripts/lu_timeline.py 64 > gpurun_out/r2_ai_timeline_mul64_br.txt 2>&1
cat $V

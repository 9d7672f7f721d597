set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu_structures.py -m gpu -x -q -k "block_rows or heavy_row" > gpurun_out/r2_af_pytest.log 2>&1; tail -5 gpurun_out/r2_af_pytest.log
V=gpurun_out/r2_af_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 600 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -v "^lu " >> $V; }
NB=64
run VJX_HV_VERBOSE=1 VJX_BR_WORK=20000
run VJX_HV_VERBOSE=1 VJX_BR_WORK=100000
NB=16
run VJX_HV_VERBOSE=1 VJX_BR_WORK=2000
run VJX_HV_VERBOSE=1 VJX_BR_WORK=6000
VJX_BR_WORK=20000 timeout 600 python scripts/lu_timeline.py 64 > gpurun_out/r2_af_timeline_mul64_br.txt 2>&1
cat $V

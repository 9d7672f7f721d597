set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu_structures.py -m gpu -q -k "agree or bordered" > gpurun_out/r2_p_pytest_lu.log 2>&1; echo "pytest lu rc=$?" >> gpurun_out/r2_p_pytest_lu.log
grep -E "AssertionError|passed|failed" gpurun_out/r2_p_pytest_lu.log | head
VJX_DT_CLUSTER=1 VJX_DT_TRACE=1 timeout 300 python scripts/lu_timeline.py 16 2>&1 | grep -E "dense" | head -3 | cut -c1-300 > gpurun_out/r2_p_probe.txt
VJX_DT_CLUSTER=2 VJX_DT_TRACE=1 timeout 300 python scripts/lu_timeline.py 16 2>&1 | grep -E "dense" | head -3 | cut -c1-300 >> gpurun_out/r2_p_probe.txt
cat gpurun_out/r2_p_probe.txt
V=gpurun_out/r2_p_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 900 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -E "phases_us|ms/step|Error|error" >> $V; }
NB=16
run VJX_DT_CLUSTER=0
run VJX_DT_CLUSTER=1
run VJX_DT_CLUSTER=2
run VJX_EVAL_MINBLOCKS=3
NB=64
run VJX_EVAL_MINBLOCKS=6
run VJX_EVAL_MINBLOCKS=8
run VJX_EVAL_MINBLOCKS=5 VJX_DT_CLUSTER=4
cat $V

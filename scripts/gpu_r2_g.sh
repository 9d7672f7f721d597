set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu_structures.py -m gpu -q -k "agree or bordered" > gpurun_out/r2_i_pytest_lu.log 2>&1; echo "pytest lu rc=$?" >> gpurun_out/r2_i_pytest_lu.log
grep -E "AssertionError|passed|failed" gpurun_out/r2_i_pytest_lu.log | head -20
V=gpurun_out/r2_i_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 900 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -E "phases_us|ms/step|Error|error" >> $V; }
NB=16
run VJX_DT_CLUSTER=1
run VJX_DT_CLUSTER=2
run VJX_DT_CLUSTER=4
run VJX_EVAL_NOINLINE_MATH=1
run VJX_EVAL_NOINLINE_MATH=1 VJX_EVAL_SLICES=1
run VJX_EVAL_NOINLINE_MATH=1 VJX_EVAL_SLICES=6
NB=64
run VJX_DT_CLUSTER=4
run VJX_EVAL_NOINLINE_MATH=1
run VJX_EVAL_NOINLINE_MATH=1 VJX_EVAL_MINBLOCKS=0
cat $V
VJX_DT_CLUSTER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lu_dense -s 6 -c 1 -o gpurun_out/r2_ncu_dense_v3_c1 python scripts/phase_probe.py 16 psp103w 4 > gpurun_out/r2_i_ncu_dense_c1.log 2>&1
VJX_DT_CLUSTER=1 VJX_DT_TRACE=1 timeout 300 python scripts/lu_timeline.py 16 2>&1 | grep -E "dense trailing|factor\+fwd|usolve" > gpurun_out/r2_i_timeline_c1.txt
VJX_DT_CLUSTER=2 VJX_DT_TRACE=1 timeout 300 python scripts/lu_timeline.py 16 2>&1 | grep -E "dense trailing|factor\+fwd|usolve" > gpurun_out/r2_i_timeline_c2.txt
cat gpurun_out/r2_i_timeline_c1.txt gpurun_out/r2_i_timeline_c2.txt

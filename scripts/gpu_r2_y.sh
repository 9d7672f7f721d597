set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu_structures.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_ac_pytest.log 2>&1; tail -3 gpurun_out/r2_ac_pytest.log
V=gpurun_out/r2_ac_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 600 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -v "^lu " >> $V; }
NB=16
run VJX_NO_FLAT_U=1
run VJX_X=1
run VJX_EVAL_SLICES=4 VJX_EVAL_MINBLOCKS=3
run VJX_EVAL_SLICES=5 VJX_EVAL_MINBLOCKS=3
run VJX_EVAL_SLICES=6 VJX_EVAL_MINBLOCKS=4
run VJX_EVAL_SLICES=8 VJX_EVAL_MINBLOCKS=5
run VJX_EVAL_SLICES=3 VJX_EVAL_MINBLOCKS=3
VJX_X=1 timeout 300 python scripts/lu_timeline.py 16 > gpurun_out/r2_ac_timeline.txt 2>&1
NB=64
run VJX_X=1
cat $V

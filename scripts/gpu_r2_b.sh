# round 2, pass B: full GPU test suite with the sliced eval kernels + stream-ordered nr_solve; eval-kernel variants
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_d_pytest.log
tail -5 gpurun_out/r2_d_pytest.log
V=gpurun_out/r2_d_eval_variants.txt
: > $V
run() { echo "== $* " >> $V; env "$@" timeout 600 python scripts/eval_probe.py $NB psp103w 20 2>&1 | grep -E "phase|prepare|Error|error" >> $V; }
NB=16
run VJX_EVAL_SLICES=1
run VJX_EVAL_SLICES=4
run VJX_EVAL_SLICES=6
run VJX_EVAL_SLICES=6 VJX_EVAL_BLOCK=64
run VJX_EVAL_SLICES=6 VJX_EVAL_BLOCK=32
run VJX_EVAL_SLICES=1 VJX_EVAL_BLOCK=32
NB=64
run VJX_EVAL_SLICES=1
run VJX_EVAL_SLICES=1 VJX_EVAL_MINBLOCKS=3
run VJX_EVAL_SLICES=1 VJX_EVAL_MINBLOCKS=4
run VJX_EVAL_SLICES=1 VJX_EVAL_BLOCK=64 VJX_EVAL_MINBLOCKS=6
run VJX_EVAL_SLICES=2
run VJX_EVAL_SLICES=6
cat $V

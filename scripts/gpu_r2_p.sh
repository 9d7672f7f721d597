set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants_are_bit" > gpurun_out/r2_s_pytest.log 2>&1; tail -3 gpurun_out/r2_s_pytest.log
V=gpurun_out/r2_s_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 900 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -v "^lu" >> $V; }
NB=16
run VJX_EVAL_LOCKSTEP=0
run VJX_EVAL_LOCKSTEP=150
run VJX_EVAL_LOCKSTEP=400
run VJX_EVAL_LOCKSTEP=1000
run VJX_EVAL_LOCKSTEP=400 VJX_EVAL_BLOCK=256
run VJX_EVAL_LOCKSTEP=400 VJX_EVAL_SLICES=1
run VJX_EVAL_LOCKSTEP=0 VJX_EVAL_SLICES=1
NB=64
run VJX_EVAL_LOCKSTEP=0
run VJX_EVAL_LOCKSTEP=150
run VJX_EVAL_LOCKSTEP=400
run VJX_EVAL_LOCKSTEP=400 VJX_EVAL_BLOCK=256 VJX_EVAL_MINBLOCKS=2
run VJX_EVAL_LOCKSTEP=400 VJX_EVAL_BLOCK=256 VJX_EVAL_MINBLOCKS=3
run VJX_EVAL_LOCKSTEP=0 VJX_EVAL_BLOCK=256 VJX_EVAL_MINBLOCKS=2
cat $V

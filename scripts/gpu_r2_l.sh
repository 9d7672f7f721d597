set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_o_pytest.log
tail -5 gpurun_out/r2_o_pytest.log
V=gpurun_out/r2_o_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 900 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -E "phases_us|ms/step|Error|error" >> $V; }
NB=16
run VJX_FULL_FINAL_EVAL=1
run VJX_X=1
NB=64
run VJX_FULL_FINAL_EVAL=1
run VJX_X=1
run VJX_EVAL_MINBLOCKS=5
cat $V

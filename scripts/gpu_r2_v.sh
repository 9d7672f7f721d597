set -x
mkdir -p gpurun_out
V=gpurun_out/r2_z_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 600 python scripts/phase_probe.py $NB psp103w 200 2>&1 | grep -v "^lu " >> $V; }
NB=16
run VJX_NRU_SPLIT=1 VJX_NO_EVAL_FORK=1
run VJX_NO_EVAL_FORK=1
run VJX_NRU_SPLIT=1
run VJX_X=1
run VJX_NRU_SPLIT=1 VJX_NO_EVAL_FORK=1
run VJX_X=1
cat $V
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_z_pytest.log
tail -5 gpurun_out/r2_z_pytest.log

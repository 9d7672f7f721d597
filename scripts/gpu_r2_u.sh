set -x
mkdir -p gpurun_out
V=gpurun_out/r2_y_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 600 python scripts/phase_probe.py $NB psp103w 60 >> $V 2>&1; }
NB=16
run VJX_HV_MIN=0
run VJX_HV_MIN=0 VJX_NRU_SPLIT=1 VJX_NO_EVAL_FORK=1
run VJX_HV_VERBOSE=1
run VJX_HV_VERBOSE=1 VJX_HV_MIN=96
VJX_HV_VERBOSE=1 timeout 300 python scripts/lu_timeline.py 16 > gpurun_out/r2_y_timeline_hv.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_dc_homotopy.py -m gpu -x -q > gpurun_out/r2_y_pytest_par.log 2>&1; tail -5 gpurun_out/r2_y_pytest_par.log
NB=64
run VJX_HV_VERBOSE=1 VJX_HV_MIN=256
cat $V

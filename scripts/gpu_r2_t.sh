set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lu_structures.py -m gpu -x -q > gpurun_out/r2_x_pytest_lu.log 2>&1; tail -5 gpurun_out/r2_x_pytest_lu.log
V=gpurun_out/r2_x_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 600 python scripts/phase_probe.py $NB psp103w 60 >> $V 2>&1; }
NB=16
run VJX_HV_MIN=0
run VJX_HV_VERBOSE=1
run VJX_HV_VERBOSE=1 VJX_HV_MIN=32
run VJX_HV_VERBOSE=1 VJX_HV_MIN=96
VJX_HV_VERBOSE=1 timeout 300 python scripts/lu_timeline.py 16 > gpurun_out/r2_x_timeline_hv.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r2_x_pytest_par.log 2>&1; tail -5 gpurun_out/r2_x_pytest_par.log
NB=64
run VJX_HV_VERBOSE=1
run VJX_HV_VERBOSE=1 VJX_HV_MIN=256
cat $V

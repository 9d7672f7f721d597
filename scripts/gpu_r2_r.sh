set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dc_homotopy.py tests/test_gpu_lu_structures.py -m gpu -q > gpurun_out/r2_v_pytest.log 2>&1; tail -15 gpurun_out/r2_v_pytest.log
V=gpurun_out/r2_v_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 900 python scripts/phase_probe.py $NB psp103w 60 2>&1 >> $V; }
NB=16
run VJX_ORDER=amd
run VJX_ORDER=mdh
run VJX_DT_ROWS=128
run VJX_DT_ROWS=96
run VJX_DT_ROWS=64
run VJX_DT_ROWS=48
NB=64
run VJX_ORDER=amd
run VJX_ORDER=mdh
NB=16
for o in amd mdh; do VJX_ORDER=$o timeout 600 python scripts/lu_timeline.py 16 > gpurun_out/r2_v_timeline_$o.txt 2>&1; done
cat $V

# round 2, pass D: blocked / cluster dense kernel (tests + timings), eval slice counts, ncu of the c6288 eval
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lu_structures.py -m gpu -x -q > gpurun_out/r2_f_pytest_lu.log 2>&1; echo "pytest lu rc=$?" >> gpurun_out/r2_f_pytest_lu.log
tail -5 gpurun_out/r2_f_pytest_lu.log
V=gpurun_out/r2_f_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 900 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -E "phases_us|ms/step|lu |Error|error" >> $V; }
NB=16
run VJX_DT_CLUSTER=0
run VJX_DT_CLUSTER=1
run VJX_DT_CLUSTER=2
run VJX_DT_CLUSTER=4
run VJX_EVAL_SLICES=2
run VJX_EVAL_SLICES=4
NB=64
run VJX_DT_CLUSTER=0
run VJX_DT_CLUSTER=2
run VJX_X=1
cat $V
timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_lu_structures.py > gpurun_out/r2_f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_f_pytest.log
tail -5 gpurun_out/r2_f_pytest.log
VJX_EVAL_SLICES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vjx_eval_kernel -s 4 -c 2 -o gpurun_out/r2_ncu_eval_c6288_k3 python scripts/eval_probe.py 16 psp103w 2 > gpurun_out/r2_f_ncu3.log 2>&1

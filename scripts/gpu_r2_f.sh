set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu_structures.py -m gpu -q > gpurun_out/r2_h_pytest_lu.log 2>&1; echo "pytest lu rc=$?" >> gpurun_out/r2_h_pytest_lu.log
grep -E "AssertionError|passed|failed" gpurun_out/r2_h_pytest_lu.log | head -20
V=gpurun_out/r2_h_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 900 python scripts/phase_probe.py $NB psp103w 60 2>&1 | grep -E "phases_us|ms/step|Error|error" >> $V; }
NB=16
run VJX_DT_CLUSTER=0
run VJX_DT_CLUSTER=1
run VJX_DT_CLUSTER=2
run VJX_DT_CLUSTER=4
NB=64
run VJX_DT_CLUSTER=0
run VJX_DT_CLUSTER=4
cat $V
for c in 1 2; do
VJX_DT_CLUSTER=$c timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lu_dense -s 6 -c 1 -o gpurun_out/r2_ncu_dense_v2_c$c python scripts/phase_probe.py 16 psp103w 4 > gpurun_out/r2_h_ncu_dense_c$c.log 2>&1
done
VJX_DT_CLUSTER=1 VJX_DT_TRACE=1 timeout 300 python scripts/lu_timeline.py 16 2>&1 | grep -E "dense trailing|factor\+fwd|usolve" > gpurun_out/r2_h_timeline_c1.txt
VJX_DT_CLUSTER=2 VJX_DT_TRACE=1 timeout 300 python scripts/lu_timeline.py 16 2>&1 | grep -E "dense trailing|factor\+fwd|usolve" > gpurun_out/r2_h_timeline_c2.txt
cat gpurun_out/r2_h_timeline_c1.txt gpurun_out/r2_h_timeline_c2.txt
timeout 900 python bench.py --corners 16 --steps 100 --warmup 3 --mul64-steps 0 --no-cpu-baseline > gpurun_out/r2_h_bench_16corners.json 2> gpurun_out/r2_h_bench_16corners.err; echo "bench16 rc=$?"
tail -c 400 gpurun_out/r2_h_bench_16corners.err

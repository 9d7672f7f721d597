set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu_structures.py -m gpu -q -k "agree" > gpurun_out/r2_g_pytest_lu.log 2>&1; echo "pytest lu rc=$?" >> gpurun_out/r2_g_pytest_lu.log
grep -E "AssertionError|passed|failed" gpurun_out/r2_g_pytest_lu.log | head -20
for c in 0 1 2; do
VJX_DT_CLUSTER=$c timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lu_dense -s 6 -c 1 -o gpurun_out/r2_ncu_dense_c$c python scripts/phase_probe.py 16 psp103w 4 > gpurun_out/r2_g_ncu_dense_c$c.log 2>&1
done
VJX_DT_CLUSTER=1 VJX_DT_TRACE=1 timeout 300 python scripts/lu_timeline.py 16 > gpurun_out/r2_g_timeline_c1.txt 2>&1
head -5 gpurun_out/r2_g_timeline_c1.txt

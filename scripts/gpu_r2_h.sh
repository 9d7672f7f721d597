set -x
mkdir -p gpurun_out
VJX_DT_CLUSTER=1 VJX_DT_TRACE=1 timeout 300 python scripts/lu_timeline.py 16 2>&1 | grep -E "dense" | head -12 > gpurun_out/r2_j_probe_c1.txt
VJX_DT_CLUSTER=2 VJX_DT_TRACE=1 timeout 300 python scripts/lu_timeline.py 16 2>&1 | grep -E "dense" | head -12 > gpurun_out/r2_j_probe_c2.txt
cat gpurun_out/r2_j_probe_c1.txt gpurun_out/r2_j_probe_c2.txt

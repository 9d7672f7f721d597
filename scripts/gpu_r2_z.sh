set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_ad_bench.json 2> gpurun_out/r2_ad_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload ring --no-cpu-baseline > gpurun_out/r2_ad_bench_ring.json 2> gpurun_out/r2_ad_bench_ring.err; echo "ring rc=$?"
timeout 900 python scripts/mul64_probe.py 64 6 > gpurun_out/r2_ad_mul64_probe.txt 2>&1; echo "probe rc=$?"
tail -c 300 gpurun_out/r2_ad_bench.json; tail -c 600 gpurun_out/r2_ad_bench_ring.json; tail -14 gpurun_out/r2_ad_mul64_probe.txt

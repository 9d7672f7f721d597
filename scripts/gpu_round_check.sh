set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1_s5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_s5_pytest.log
tail -5 gpurun_out/r1_s5_pytest.log
python bench.py > gpurun_out/r1_s5_bench.json 2> gpurun_out/r1_s5_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r1_s5_bench.json
timeout 900 python scripts/mul64_probe.py 64 6 > gpurun_out/r1_s5_mul64_probe.txt 2>&1; echo "probe rc=$?"
tail -30 gpurun_out/r1_s5_mul64_probe.txt

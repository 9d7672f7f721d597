set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_aa_pytest.log
tail -5 gpurun_out/r2_aa_pytest.log
timeout 900 python bench.py > gpurun_out/r2_aa_bench.json 2> gpurun_out/r2_aa_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_aa_bench.json

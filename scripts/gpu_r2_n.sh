set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_q_pytest.log
tail -4 gpurun_out/r2_q_pytest.log
timeout 1800 python bench.py > gpurun_out/r2_q_bench.json 2> gpurun_out/r2_q_bench.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r2_q_bench.json; tail -3 gpurun_out/r2_q_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_q_smoke.log 2>&1; tail -2 gpurun_out/r2_q_smoke.log

# Round-2 closing check: GPU suite, smoke, default bench line (new CPU leg), ring line.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final2_pytest.log
tail -4 gpurun_out/r2_final2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_final2_smoke.log
python bench.py > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_final2_bench.json
python bench.py --impl reference --steps 12 > gpurun_out/r2_final2_bench_ref.json 2> gpurun_out/r2_final2_bench_ref.err; echo "ref rc=$?"
tail -c 900 gpurun_out/r2_final2_bench_ref.json

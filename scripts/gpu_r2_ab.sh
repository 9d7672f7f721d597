set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_ag_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_ag_pytest.log
tail -4 gpurun_out/r2_ag_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_ag_smoke.log 2>&1; tail -2 gpurun_out/r2_ag_smoke.log

set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_l_pytest.log
tail -6 gpurun_out/r2_l_pytest.log
timeout 1500 python bench.py > gpurun_out/r2_l_bench.json 2> gpurun_out/r2_l_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_l_bench.json; tail -3 gpurun_out/r2_l_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_l_launches_c6288.csv python bench.py --steps 6 --warmup 3 --no-graph --corners 1 --mul64-steps 0 --no-cpu-baseline > gpurun_out/r2_l_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lu_rows|k_lu_usolve|k_lu_dense_tail|k_stamp_jacobian|k_residual" -s 40 -c 6 -o gpurun_out/r2_ncu_lu_c6288 python scripts/phase_probe.py 16 psp103w 4 > gpurun_out/r2_l_ncu_lu.log 2>&1

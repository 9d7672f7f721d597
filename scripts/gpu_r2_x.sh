set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_ab_pytest.log
tail -5 gpurun_out/r2_ab_pytest.log
# launch list of a short host-polled run (per-launch durations, cold and serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_ab_launches_c6288.csv python bench.py --steps 6 --warmup 3 --no-graph --corners 1 --mul64-steps 0 --no-cpu-baseline > gpurun_out/r2_ab_launches.log 2>&1
# full captures of the LU kernels and of the model kernels of one NR iteration
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lu_load|k_lu_rows|k_lu_usolve|k_lu_dense_tail" -s 40 -c 4 -o gpurun_out/r2_ncu_lu_c6288_final python scripts/phase_probe.py 16 psp103w 4 > gpurun_out/r2_ab_ncu_lu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vjx_eval_kernel -s 8 -c 4 -o gpurun_out/r2_ncu_eval_c6288_final python scripts/eval_probe.py 16 psp103w 2 > gpurun_out/r2_ab_ncu_eval.log 2>&1
ls -la gpurun_out/*.ncu-rep

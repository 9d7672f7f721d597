set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r1_s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_s4_pytest.log
tail -5 gpurun_out/r1_s4_pytest.log
python bench.py > gpurun_out/r1_s4_bench.json 2> gpurun_out/r1_s4_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r1_s4_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1_s4_launches.csv python bench.py --no-graph --corners 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_s4_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lu_factor -s 4 -c 2 -o gpurun_out/r1_s4_lu_factor python bench.py --no-graph --corners 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_s4_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vjx_eval_kernel -s 6 -c 2 -o gpurun_out/r1_s4_eval python bench.py --no-graph --corners 1 --steps 3 --warmup 3 --no-cpu-baseline >> gpurun_out/r1_s4_ncu_full.log 2>&1
ls -la gpurun_out

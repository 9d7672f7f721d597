mkdir -p gpurun_out
T=s6i
python bench.py > gpurun_out/r1_${T}_bench.json 2> gpurun_out/r1_${T}_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r1_${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_${T}_launches.csv python bench.py --no-graph --steps 3 --warmup 3 --corners 1 --mul64-steps 0 --no-cpu-baseline > gpurun_out/r1_${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lu_rows -s 6 -c 1 -f -o gpurun_out/r1_${T}_lu_rows python scripts/run_steps.py 16 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lu_usolve -s 6 -c 1 -f -o gpurun_out/r1_${T}_lu_usolve python scripts/run_steps.py 16 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lu_dense -s 6 -c 1 -f -o gpurun_out/r1_${T}_lu_dense python scripts/run_steps.py 16 4 > /dev/null 2>&1
ls -la gpurun_out | tail -8

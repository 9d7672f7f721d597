mkdir -p gpurun_out
T=s7o
python bench.py > gpurun_out/r1_${T}_bench.json 2> gpurun_out/r1_${T}_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r1_${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_${T}_launches.csv python bench.py --no-graph --steps 3 --warmup 3 --corners 1 --mul64-steps 0 --no-cpu-baseline > gpurun_out/r1_${T}_ncu_bench.log 2>&1
for k in k_lu_rows k_lu_usolve; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/r1_${T}_$k python scripts/run_steps.py 16 4 > /dev/null 2>&1
done
ncu --set full --clock-control none -k regex:k_lu_rows -s 4 -c 2 -f -o gpurun_out/r1_${T}_mul64_rows python scripts/run_steps.py 64 2 > /dev/null 2>&1
ls gpurun_out | grep $T

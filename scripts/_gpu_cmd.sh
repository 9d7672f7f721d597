mkdir -p gpurun_out
T=s6y
python bench.py > gpurun_out/r1_${T}_bench.json 2> gpurun_out/r1_${T}_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r1_${T}_bench.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r1_${T}_ref.json 2>/dev/null; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_${T}_launches.csv python bench.py --no-graph --steps 3 --warmup 3 --corners 1 --mul64-steps 0 --no-cpu-baseline > gpurun_out/r1_${T}_ncu_bench.log 2>&1
for k in k_lu_rows k_lu_usolve k_lu_dense k_lu_load; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/r1_${T}_$k python scripts/run_steps.py 16 4 > /dev/null 2>&1
done
ls gpurun_out | grep $T

mkdir -p gpurun_out
T=s7r
for k in k_residual k_stamp_jacobian vjx_eval_kernel k_nr_update; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/r1_${T}_$k python scripts/run_steps.py 16 4 > /dev/null 2>&1
done
ls gpurun_out | grep $T

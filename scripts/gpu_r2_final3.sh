# ncu --set full of the psp103w model kernel at mul64 size with the final launch configuration (5 CTAs / SM)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vjx_eval_kernel -s 4 -c 2 -o gpurun_out/r2_ncu_eval_mul64_final python scripts/eval_probe.py 64 psp103w 2 > gpurun_out/r2_final3_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r2_final3_ncu.log
python scripts/ncu_summary.py report gpurun_out/r2_ncu_eval_mul64_final.ncu-rep > gpurun_out/r2_ncu_eval_kernel_mul64_psp103w_final.txt 2>&1
cat gpurun_out/r2_ncu_eval_kernel_mul64_psp103w_final.txt | head -80
rm -f gpurun_out/r2_ncu_eval_mul64_final.ncu-rep

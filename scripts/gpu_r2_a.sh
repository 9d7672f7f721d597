# round 2, first GPU pass with the PSP103-weight model: new parity tests, bench line, eval-kernel ncu at mul64 size
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "ringw or psp103w" > gpurun_out/r2_c_pytest_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c_pytest_w.log
tail -5 gpurun_out/r2_c_pytest_w.log
timeout 1200 python bench.py --steps 100 --warmup 3 --cpu-steps 8 > gpurun_out/r2_c_bench.json 2> gpurun_out/r2_c_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_c_bench.json
tail -5 gpurun_out/r2_c_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vjx_eval_kernel -s 4 -c 1 -o gpurun_out/r2_ncu_eval_mul64 python scripts/eval_probe.py 64 psp103w 2 > gpurun_out/r2_c_ncu_eval.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r2_c_ncu_eval.log

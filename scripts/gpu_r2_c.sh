# round 2, pass C: tests (save list, batched tail), eval variants after the tail fix, ring + default bench lines
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_e_pytest.log
tail -5 gpurun_out/r2_e_pytest.log
V=gpurun_out/r2_e_eval_variants.txt
: > $V
run() { echo "== $* " >> $V; env "$@" timeout 600 python scripts/eval_probe.py $NB psp103w 20 2>&1 | grep -E "phase|prepare|Error|error" >> $V; }
NB=16
run VJX_EVAL_SLICES=1
run VJX_EVAL_SLICES=3
run VJX_EVAL_SLICES=6
NB=64
run VJX_EVAL_SLICES=1 VJX_EVAL_MINBLOCKS=0
run VJX_EVAL_SLICES=1 VJX_EVAL_MINBLOCKS=3
run VJX_EVAL_SLICES=1 VJX_EVAL_MINBLOCKS=4
cat $V
VJX_EVAL_SLICES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vjx_eval_kernel -s 4 -c 1 -o gpurun_out/r2_ncu_eval_c6288_k1 python scripts/eval_probe.py 16 psp103w 2 > gpurun_out/r2_e_ncu1.log 2>&1
VJX_EVAL_SLICES=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vjx_eval_kernel -s 4 -c 1 -o gpurun_out/r2_ncu_eval_c6288_k6 python scripts/eval_probe.py 16 psp103w 2 > gpurun_out/r2_e_ncu6.log 2>&1
timeout 900 python bench.py --workload ring --steps 300 --warmup 5 --cpu-steps 40 > gpurun_out/r2_e_bench_ring.json 2> gpurun_out/r2_e_bench_ring.err; echo "ring bench rc=$?"
tail -c 1200 gpurun_out/r2_e_bench_ring.json
timeout 1500 python bench.py --steps 100 --warmup 3 --cpu-steps 8 > gpurun_out/r2_e_bench.json 2> gpurun_out/r2_e_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_e_bench.json; tail -3 gpurun_out/r2_e_bench.err

set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants_are_bit" > gpurun_out/r2_r_pytest.log 2>&1; tail -3 gpurun_out/r2_r_pytest.log
V=gpurun_out/r2_r_sweep.txt
: > $V
run() { echo "== $* " >> $V; env "$@" timeout 900 python bench.py --steps 100 --warmup 3 --mul64-steps 0 --no-cpu-baseline --corners $NC 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['sweep_concurrent']['steps_per_s'], d['value'], d['value_l2_resident'])" >> $V; }
NC=8
run VJX_SWEEP_MINBLOCKS=5
run VJX_SWEEP_MINBLOCKS=3
run VJX_SWEEP_SLICES=2 VJX_SWEEP_MINBLOCKS=4
NC=16
run VJX_SWEEP_MINBLOCKS=4
NC=4
run VJX_SWEEP_MINBLOCKS=4
cat $V

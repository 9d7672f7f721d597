set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dc_homotopy.py -m gpu -q > gpurun_out/r2_t_pytest.log 2>&1; tail -15 gpurun_out/r2_t_pytest.log
V=gpurun_out/r2_t_sweep.txt
: > $V
run() { echo "== NC=$NC $* " >> $V; env "$@" timeout 900 python bench.py --steps 100 --warmup 3 --mul64-steps 0 --no-cpu-baseline --corners $NC 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['sweep_concurrent']['steps_per_s'], d['value'], d['value_l2_resident'])" >> $V; }
NC=8
run VJX_SWEEP_LU_DIV=1
run VJX_SWEEP_LU_DIV=2
run VJX_SWEEP_LU_DIV=4
run VJX_SWEEP_LU_DIV=8
NC=16
run VJX_SWEEP_LU_DIV=4
run VJX_SWEEP_LU_DIV=8
cat $V

set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_w_pytest.log
tail -5 gpurun_out/r2_w_pytest.log
V=gpurun_out/r2_w_variants.txt
: > $V
run() { echo "== $NB $* " >> $V; env "$@" timeout 900 python scripts/phase_probe.py $NB psp103w 60 2>&1 >> $V; }
NB=16
run VJX_ORDER=amd
run VJX_ORDER=mdh
run VJX_ORDER=mdh VJX_DT_ROWS=64
run VJX_ORDER=mdh VJX_DT_ROWS=48
run VJX_ORDER=mdh VJX_DT_ROWS=32
NB=64
run VJX_ORDER=amd
run VJX_ORDER=mdh
for o in amd mdh; do VJX_ORDER=$o timeout 600 python scripts/lu_timeline.py 16 > gpurun_out/r2_w_timeline_$o.txt 2>&1; done
timeout 900 python bench.py > gpurun_out/r2_w_bench.json 2> gpurun_out/r2_w_bench.err; echo "bench rc=$?"
cat $V
tail -c 600 gpurun_out/r2_w_bench.json

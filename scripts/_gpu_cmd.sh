mkdir -p gpurun_out
T=s7i
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
VJX_DT_TRACE=1 python scripts/lu_timeline.py 16 2>&1 | grep "==\|dense trail" | tail -n 3
python scripts/mul64_probe.py 64 6 2>&1 | grep "phase lu\|steady\|lu_solve residual"

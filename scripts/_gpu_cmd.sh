mkdir -p gpurun_out
T=s7e
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python scripts/lu_timeline.py 64 > gpurun_out/r1_${T}_timeline64.txt 2>&1; grep "==" gpurun_out/r1_${T}_timeline64.txt
python scripts/mul64_probe.py 64 6 2>&1 | grep "phase lu\|steady\|lu_solve residual"

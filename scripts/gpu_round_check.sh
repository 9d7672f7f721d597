set -x
mkdir -p gpurun_out
T=${TAG:-s6b}
python -m pytest tests -m gpu -x -q > gpurun_out/r1_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_${T}_pytest.log
tail -5 gpurun_out/r1_${T}_pytest.log
timeout 600 python scripts/lu_timeline.py 16 > gpurun_out/r1_${T}_timeline16.txt 2>&1; echo "tl16 rc=$?"
grep "==" gpurun_out/r1_${T}_timeline16.txt
timeout 900 python scripts/mul64_probe.py 64 6 > gpurun_out/r1_${T}_mul64_probe.txt 2>&1; echo "probe rc=$?"
tail -14 gpurun_out/r1_${T}_mul64_probe.txt
timeout 600 python scripts/lu_timeline.py 64 > gpurun_out/r1_${T}_timeline64.txt 2>&1; echo "tl64 rc=$?"
grep "==" gpurun_out/r1_${T}_timeline64.txt

# One gpurun call that re-establishes the state of the round: GPU tests, the default bench line, the mul64 probe.
#   gpurun --timeout 1800 -- 'TAG=x bash scripts/gpu_round_check.sh'
set -x
mkdir -p gpurun_out
T=${TAG:-check}
python -m pytest tests -m gpu -x -q > gpurun_out/r1_${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_${T}_pytest.log
tail -5 gpurun_out/r1_${T}_pytest.log
python bench.py > gpurun_out/r1_${T}_bench.json 2> gpurun_out/r1_${T}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r1_${T}_bench.json
timeout 900 python scripts/mul64_probe.py 64 6 > gpurun_out/r1_${T}_mul64_probe.txt 2>&1; echo "probe rc=$?"
tail -14 gpurun_out/r1_${T}_mul64_probe.txt

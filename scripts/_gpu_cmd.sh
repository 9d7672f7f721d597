mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/lu_timeline.py 16 2>&1 | grep "=="
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_lu -s 8 -c 8 --csv --log-file gpurun_out/r1_s6e_lu_launches.csv python scripts/run_steps.py 16 4 > gpurun_out/r1_s6e_ncu1.log 2>&1
grep -o '"k_lu[a-z_]*.*' gpurun_out/r1_s6e_lu_launches.csv | awk -F'","' '{print substr($1,1,20), $NF}'
python scripts/mul64_probe.py 64 6 2>&1 | grep "phase lu\|steady"

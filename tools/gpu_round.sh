#!/bin/bash
# One GPU-box visit: GPU tests, smoke, bench, reference arm; "ncu" adds the launch list and one full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/cpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-2500
if [ "$1" == "ncu" ]; then
  # launch list of ONE timed step at the headline size (skip the 3 warm-up evaluations)
  L=$(python - <<'PY'
import json
try:
    print(int(json.loads(open("gpurun_out/bench.log").readline())["gpu_launches"]) // 10)
except Exception:
    print(2200)
PY
)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L)) -c $L --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 400 -c 2 -o gpurun_out/prof_dgemm -f python bench.py --steps 1 --warmup 3 --n 8192 > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi

#!/bin/bash
# One GPU-box visit: GPU tests, smoke, bench, ncu launch list (+ optional full capture of the top kernel).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/cpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
if [ "$1" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --n 8192 > gpurun_out/ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 300 -c 3 -o gpurun_out/prof_dgemm -f python bench.py --steps 1 --warmup 3 --n 8192 > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi

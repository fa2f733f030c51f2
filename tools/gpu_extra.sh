#!/bin/bash
# secondary workloads + full ncu capture of the DMMA GEMM
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload fit > gpurun_out/bench_fit.log 2>&1; tail -1 gpurun_out/bench_fit.log | cut -c1-1200
timeout 900 python bench.py --workload fit --impl reference > gpurun_out/bench_fit_ref.log 2>&1; tail -1 gpurun_out/bench_fit_ref.log | cut -c1-1200
timeout 900 python bench.py --workload acq --steps 3 > gpurun_out/bench_acq.log 2>&1; tail -1 gpurun_out/bench_acq.log | cut -c1-1200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-600
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 400 -c 2 -o gpurun_out/prof_dgemm -f python bench.py --steps 1 --warmup 3 --n 8192 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300

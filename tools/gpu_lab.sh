#!/bin/bash
# One GPU-box visit for kernel development: factorisation lab, GPU tests, short bench variants.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 ./tools/chol_lab "$@" > gpurun_out/chol_lab.log 2>&1; echo "lab rc=$?" >> gpurun_out/chol_lab.log
cat gpurun_out/chol_lab.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-1800
GPP_GEMM_BM=64 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_bm64.log 2>&1
tail -1 gpurun_out/bench_bm64.log | cut -c1-1800

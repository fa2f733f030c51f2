#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-2300
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active --clock-control none --kernel-name-base mangled -k regex:dgemm_dmma_kernelILb0ELb0E -s 3 -c 1 --csv --log-file gpurun_out/lauum16k_traffic.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_lauum16k.log 2>&1
tail -6 gpurun_out/lauum16k_traffic.csv | cut -c150-400

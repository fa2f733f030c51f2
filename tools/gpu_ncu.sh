#!/bin/bash
# ncu artefacts: launch list of one timed step at the headline size + full captures of the dominant kernels
mkdir -p gpurun_out
L=435
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L)) -c $L --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/launches.csv | cut -c1-200
# LAUUM (k-strided operands) and the trailing SYRK (k-contiguous) at N=8192: one launch each, full set
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:dgemm_dmma_kernelILb0ELb0E -s 3 -c 1 -o gpurun_out/prof_lauum -f python bench.py --steps 1 --warmup 3 --n 8192 > gpurun_out/ncu_full_lauum.log 2>&1
tail -2 gpurun_out/ncu_full_lauum.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:grad_tile_kernel -s 3 -c 1 -o gpurun_out/prof_grad -f python bench.py --steps 1 --warmup 3 --n 8192 > gpurun_out/ncu_full_grad.log 2>&1
tail -2 gpurun_out/ncu_full_grad.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:cov_tile_kernel -s 3 -c 1 -o gpurun_out/prof_cov -f python bench.py --steps 1 --warmup 3 --n 8192 > gpurun_out/ncu_full_cov.log 2>&1
tail -2 gpurun_out/ncu_full_cov.log | cut -c1-200

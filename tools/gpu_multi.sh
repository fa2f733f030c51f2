#!/bin/bash
# multi-GPU visit: N = number of GPUs on the box
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_g$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_g$N.log
grep '^{' gpurun_out/bench_g$N.log | cut -c1-900; tail -2 gpurun_out/bench_g$N.log | cut -c1-300
timeout 900 $TR bench.py --gpus $N --workload fit > gpurun_out/bench_fit_g$N.log 2>&1
grep '^{' gpurun_out/bench_fit_g$N.log | cut -c1-700
timeout 900 $TR bench.py --gpus $N --workload acq --steps 3 > gpurun_out/bench_acq_g$N.log 2>&1
grep '^{' gpurun_out/bench_acq_g$N.log | cut -c1-700
timeout 600 $TR bench.py --gpus $N --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_g$N.log 2>&1
grep '^{' gpurun_out/bench_ref_g$N.log | cut -c1-300

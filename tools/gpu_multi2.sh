#!/bin/bash
# multi-GPU visit: default bench line (with extra.fit_c2 / extra.acq_c5) and the C4 fit with a maxiter cap under torchrun
# usage: gpu_multi2.sh N
N=$1
mkdir -p gpurun_out
( if [ "$N" == "1" ]; then timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline; else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5; fi ) > gpurun_out/bench_g$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_g$N.log
grep '^{' gpurun_out/bench_g$N.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('N=%d value %.3f evals/s e2e %.3f fit_c2 %.3f s (%s per rank) acq %.2f ms dev / %.2f ms e2e' % (d['n_gpus'], d['value'], d['e2e']['value'], d['extra']['fit_c2']['value'], d['extra']['fit_c2']['restarts_run_per_rank'], d['extra']['acq_c5']['ms_per_step'], d["extra"]["acq_c5"]["e2e"]["ms_per_step"]))
"
tail -2 gpurun_out/bench_g$N.log | cut -c1-300
( if [ "$N" == "1" ]; then timeout 900 python bench.py --workload fit --fit-config c4 --maxiter 3; else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --workload fit --fit-config c4 --maxiter 3 --gpus $N; fi ) > gpurun_out/fit_c4_g$N.log 2>&1; echo "rc=$?" >> gpurun_out/fit_c4_g$N.log
grep '^{' gpurun_out/fit_c4_g$N.log | cut -c1-900
tail -1 gpurun_out/fit_c4_g$N.log

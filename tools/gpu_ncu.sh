#!/bin/bash
# ncu launch list of ONE timed step at the headline size (after the 3 warm-up evaluations of bench.py)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
L=$(python -c "import json; print(int(json.loads(open('gpurun_out/bench.log').readline())['gpu_launches'])//10)")
echo "launches per evaluation: $L"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L)) -c $L --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
head -c 600 gpurun_out/bench.log; echo
python tools/launch_summary.py gpurun_out/launches.csv

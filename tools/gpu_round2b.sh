#!/bin/bash
# round 2, visit b: tests after the exp_nonpos / mirror-off changes, fit standalone vs fit inside the default bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_b.log
tail -4 gpurun_out/pytest_gpu_b.log
timeout 600 python bench.py --workload fit > gpurun_out/fit_standalone.log 2>&1; tail -1 gpurun_out/fit_standalone.log | cut -c1-700
timeout 600 python bench.py --workload fit > gpurun_out/fit_standalone2.log 2>&1; tail -1 gpurun_out/fit_standalone2.log | cut -c1-700
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_b.log
python - <<'PY'
import json
for l in open("gpurun_out/bench_b.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "ms", d["ms_per_step"], "stages", d["roofline"]["stages_ms"], "hbm", d["roofline"]["hbm_stage_gbs"])
        print("parity", d["parity"])
        print("fit_c2", d["extra"]["fit_c2"]["value"], d["extra"]["fit_c2"]["evals_per_s"], "acq", d["extra"]["acq_c5"]["ms_per_step"], d["extra"]["acq_c5"]["e2e"]["ms_per_step"])
PY
tail -3 gpurun_out/bench_b.log | cut -c1-300

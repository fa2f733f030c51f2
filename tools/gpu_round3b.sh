#!/bin/bash
# round 2, visit 3b: feature-major gradient panels, parallel transposed sweep -- parity, stage times, ncu, small-N probe, acquisition
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_headline_gpu.py tests/test_reference_pin_gpu.py -m gpu -q -x > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3b_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep"
timeout 600 $B > gpurun_out/r3b_bench.log 2>&1
grep '^{' gpurun_out/r3b_bench.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('ms/step %.2f' % d['ms_per_step'], d['roofline']['stages_ms'], d['roofline']['hbm_stage_gbs'], d['parity'])
"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
B1="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep"
timeout 900 $NCU -k 'regex:grad_tile_kernel' -s 3 -c 1 -o gpurun_out/r3b_prof_grad $B1 > gpurun_out/r3b_ncu_grad.log 2>&1; tail -1 gpurun_out/r3b_ncu_grad.log
python tools/ncu_summary.py gpurun_out/r3b_prof_grad.ncu-rep > gpurun_out/r3b_ncu_summary.json 2>gpurun_out/r3b_ncu_summary.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3b_ncu_summary.json"))
for k, v in d.items():
    for r in v:
        print(r["kernel"][:50], r.get("gpu__time_duration.sum"), "fp64", r.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
              "issue", r.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps", r.get("sm__warps_active.avg.pct_of_peak_sustained_active"), "regs", r.get("launch__registers_per_thread"), "bankconf", r.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"))
PY
GPP_GRAPH=0 timeout 300 python tools/small_eval_probe.py > gpurun_out/r3b_small_eval_probe.log 2>&1; tail -4 gpurun_out/r3b_small_eval_probe.log
LS=$(grep "eval 0" gpurun_out/r3b_small_eval_probe.log | sed 's/.*so far //')
GPP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*LS)) -c $LS --csv --log-file gpurun_out/r3b_launches_n500.csv python tools/small_eval_probe.py > gpurun_out/r3b_ncu_small.log 2>&1
python tools/launch_summary.py gpurun_out/r3b_launches_n500.csv
timeout 600 python tools/small_eval_probe.py > gpurun_out/r3b_small_eval_probe_graph.log 2>&1; tail -3 gpurun_out/r3b_small_eval_probe_graph.log
timeout 900 python bench.py --workload acq --steps 5 > gpurun_out/r3b_acq.log 2>&1; tail -1 gpurun_out/r3b_acq.log | cut -c1-900

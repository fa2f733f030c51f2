#!/bin/bash
# round 2 profiling visit: launch list of one timed step at N=16384 + full ncu captures of the shipped kernels
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep"
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep > gpurun_out/ncu_bench.log 2>&1
L=$(python -c "import json; print(int([json.loads(l) for l in open('gpurun_out/ncu_bench.log') if l.startswith('{')][0]['gpu_launches'])//5)")
echo "launches per evaluation: $L"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L)) -c $L --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv | head -20
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
# LAUUM (the only <false, false> GEMM), one trailing update (<true, true>), covariance, gradient, leaf: the launch of the timed step
timeout 900 $NCU -k 'regex:dgemm_dmma_kernel<false, false' -s 3 -c 1 -o gpurun_out/r02_prof_lauum $B > gpurun_out/ncu_full_lauum.log 2>&1; tail -1 gpurun_out/ncu_full_lauum.log
timeout 900 $NCU -k 'regex:cov_tile_kernel' -s 3 -c 1 -o gpurun_out/r02_prof_cov $B > gpurun_out/ncu_full_cov.log 2>&1; tail -1 gpurun_out/ncu_full_cov.log
timeout 900 $NCU -k 'regex:grad_tile_kernel' -s 3 -c 1 -o gpurun_out/r02_prof_grad $B > gpurun_out/ncu_full_grad.log 2>&1; tail -1 gpurun_out/ncu_full_grad.log
timeout 900 $NCU -k 'regex:leaf_potrf' -s 400 -c 1 -o gpurun_out/r02_prof_leaf $B > gpurun_out/ncu_full_leaf.log 2>&1; tail -1 gpurun_out/ncu_full_leaf.log
timeout 900 $NCU -k 'regex:dgemm_dmma_kernel<true, true' -s $((3*400+40)) -c 1 -o gpurun_out/r02_prof_syrk $B > gpurun_out/ncu_full_syrk.log 2>&1; tail -1 gpurun_out/ncu_full_syrk.log
python tools/ncu_summary.py gpurun_out/r02_prof_lauum.ncu-rep gpurun_out/r02_prof_cov.ncu-rep gpurun_out/r02_prof_grad.ncu-rep gpurun_out/r02_prof_leaf.ncu-rep gpurun_out/r02_prof_syrk.ncu-rep > gpurun_out/r02_ncu_full_summary.json 2>gpurun_out/ncu_summary.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_ncu_full_summary.json"))
for k, v in d.items():
    for r in v:
        print(k.split("/")[-1], r["kernel"][:60], r.get("gpu__time_duration.sum"), "dram rd", r.get("dram__bytes_read.sum"), "wr", r.get("dram__bytes_write.sum"),
              "issue", r.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps", r.get("sm__warps_active.avg.pct_of_peak_sustained_active"), "regs", r.get("launch__registers_per_thread"))
PY
# our DMMA GEMM on RANDOM operands vs cuBLAS (torch.matmul), 8192^3
python - <<'PY' > gpurun_out/r02_probe_dgemm.log 2>&1
import sys
sys.path.insert(0, "gp-plus_b200")
import torch
from gpplus_b200 import _engine as E
ms = min(E.probe_dgemm(8192, 8192, 8192, 10) for _ in range(3))
print("dgemm_dmma_kernel 8192^3 on random operands: %.3f ms = %.2f TFLOP/s" % (ms, 2 * 8192 ** 3 / ms / 1e9))
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
best = 1e9
for i in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
    if i >= 2: best = min(best, e0.elapsed_time(e1))
print("cuBLAS (torch.matmul f64) 8192^3: %.3f ms = %.2f TFLOP/s" % (best, 2 * 8192 ** 3 / best / 1e9))
PY
cat gpurun_out/r02_probe_dgemm.log
# the latency-bound small-N evaluation (n=500): kernel-by-kernel launch list with graph replay off
GPP_GRAPH=0 timeout 300 python tools/small_eval_probe.py > gpurun_out/small_eval_probe.log 2>&1; cat gpurun_out/small_eval_probe.log | tail -4
LS=$(grep "eval 0" gpurun_out/small_eval_probe.log | sed 's/.*so far //')
GPP_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*LS)) -c $LS --csv --log-file gpurun_out/r02_launches_n500.csv python tools/small_eval_probe.py > gpurun_out/ncu_small.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_n500.csv

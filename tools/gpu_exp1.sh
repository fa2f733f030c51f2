#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep"
show() { grep '^{' $1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); s=d['roofline']['stages_ms']
    print('$2', 'ms/step %.2f' % d['ms_per_step'], 'chol %.2f trtri %.2f lauum %.2f cov %.3f grad %.3f' % (s['cholesky'], s['trtri'], s['lauum'], s['covariance'], s['gradient']))
"; }
$B > gpurun_out/exp_base.log 2>&1; show gpurun_out/exp_base.log base
GPP_PANEL=16 $B > gpurun_out/exp_pb16.log 2>&1; show gpurun_out/exp_pb16.log PB16
GPP_PANEL=12 $B > gpurun_out/exp_pb12.log 2>&1; show gpurun_out/exp_pb12.log PB12
GPP_OVERLAP_INV=0 $B > gpurun_out/exp_noov.log 2>&1; show gpurun_out/exp_noov.log no-overlap
GPP_OVERLAP_INV=0 GPP_PANEL=16 $B > gpurun_out/exp_noov16.log 2>&1; show gpurun_out/exp_noov16.log no-overlap-PB16
# full ncu of LAUUM and one trailing update (fixed regex)
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
B1="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep"
timeout 900 $NCU -k 'regex:dgemm_dmma_kernel<\(bool\)0, \(bool\)0' -s 3 -c 1 -o gpurun_out/r02_prof_lauum $B1 > gpurun_out/ncu_full_lauum.log 2>&1; tail -1 gpurun_out/ncu_full_lauum.log
timeout 900 $NCU -k 'regex:dgemm_dmma_kernel<\(bool\)1, \(bool\)1' -s $((3*281+12)) -c 1 -o gpurun_out/r02_prof_syrk $B1 > gpurun_out/ncu_full_syrk.log 2>&1; tail -1 gpurun_out/ncu_full_syrk.log
python tools/ncu_summary.py gpurun_out/r02_prof_lauum.ncu-rep gpurun_out/r02_prof_syrk.ncu-rep > gpurun_out/r02_ncu_gemm_summary.json 2>gpurun_out/ncu_summary2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_ncu_gemm_summary.json"))
for k, v in d.items():
    for r in v:
        print(k.split("/")[-1], r["kernel"][:50], r["grid"], r.get("gpu__time_duration.sum"), "dram rd", r.get("dram__bytes_read.sum"), "wr", r.get("dram__bytes_write.sum"),
              "dmma", r.get("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"), "L2 hit", r.get("lts__t_sector_hit_rate.pct"))
PY

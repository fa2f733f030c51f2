#!/bin/bash
# round 2, INT8-sliced path: GPU test suite, driver-style bench line, launch list of one step, full ncu of oz_gemm_kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r4_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r4_pytest.log
timeout 900 python bench.py > gpurun_out/r4_bench.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/r4_bench.log > gpurun_out/r4_bench.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r4_bench.json").read())
r = d["roofline"]
print("value %.3f evals/s  ms/step %.2f  e2e %.3f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
print("roofline", r["kernel"][:40], "achieved %.1f peak %.1f frac %.3f" % (r["achieved"], r["peak"], r["frac"]), r.get("fp64_equivalent"))
print("stages", r["stages_ms"]); print("parity", d["parity"]); print("exact_dmma", d.get("exact_dmma"))
print("cpu", d["cpu_baseline"]); print("prior", d["prior_draws"]); print("clocks", d["clocks"])
print("fit_c2", {k: d["extra"]["fit_c2"][k] for k in ("value", "objective_evals", "first_fit_s")}); print("acq", d["extra"]["acq_c5"]["ms_per_step"], d["extra"]["acq_c5"]["e2e"]["ms_per_step"])
PY
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep --no-dmma-arm"
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep --no-dmma-arm > gpurun_out/r4_ncu_bench.log 2>&1
L=$(python -c "import json; print(int([json.loads(l) for l in open('gpurun_out/r4_ncu_bench.log') if l.startswith('{')][0]['gpu_launches'])//5)")
echo "launches per evaluation: $L"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L)) -c $L --csv --log-file gpurun_out/r4_launches.csv $B > gpurun_out/r4_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r4_launches.csv | head -24
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
# the K^-1 launch is the last oz_gemm_kernel of a step: count the oz launches of one step from the list
NOZ=$(grep -c "oz_gemm_kernel" gpurun_out/r4_launches.csv)
echo "oz_gemm launches per evaluation: $NOZ"
timeout 900 $NCU -k 'regex:oz_gemm_kernel' -s $((4*NOZ-1)) -c 1 -o gpurun_out/r4_prof_oz_lauum $B > gpurun_out/r4_ncu_full_oz.log 2>&1; tail -1 gpurun_out/r4_ncu_full_oz.log
timeout 900 $NCU -k 'regex:oz_gemm_kernel' -s $((3*NOZ+4)) -c 1 -o gpurun_out/r4_prof_oz_syrk $B > gpurun_out/r4_ncu_full_oz2.log 2>&1; tail -1 gpurun_out/r4_ncu_full_oz2.log
timeout 900 $NCU -k 'regex:oz_split_cols_kernel' -s 3 -c 1 -o gpurun_out/r4_prof_split $B > gpurun_out/r4_ncu_full_split.log 2>&1; tail -1 gpurun_out/r4_ncu_full_split.log
python tools/ncu_summary.py gpurun_out/r4_prof_oz_lauum.ncu-rep gpurun_out/r4_prof_oz_syrk.ncu-rep gpurun_out/r4_prof_split.ncu-rep > gpurun_out/r4_ncu_full_summary.json 2>gpurun_out/r4_ncu_summary.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r4_ncu_full_summary.json"))
for k, v in d.items():
    for r in v:
        print(k.split("/")[-1], r["kernel"][:40], r.get("gpu__time_duration.sum"), "grid", r.get("grid"), "dram rd", r.get("dram__bytes_read.sum"), "wr", r.get("dram__bytes_write.sum"),
              "lts", r.get("lts__t_bytes.sum"), "regs", r.get("launch__registers_per_thread"))
        for kk, vv in r.items():
            if "tensor" in kk or "lts__throughput" in kk: print("    ", kk, vv)
PY

#!/bin/bash
mkdir -p gpurun_out
for ex in ${@:-01_borehole_emulation 02_borehole_mixed_emulation 03_wing_multi_fidelity 04_mfbo_borehole 05_sine_1d}; do
  t0=$(date +%s)
  timeout 900 python examples/$ex.py > gpurun_out/example_$ex.log 2>&1
  rc=$?
  echo "== $ex rc=$rc $(( $(date +%s) - t0 )) s: $(tail -1 gpurun_out/example_$ex.log | cut -c1-120)"
  grep -E "RRMSE|Error|error" gpurun_out/example_$ex.log | head -3
done

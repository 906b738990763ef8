#!/usr/bin/env bash
# tuning sweep: rebuild with different compile-time knobs and time the forward kernel on the diag workload
mkdir -p gpurun_out
for defs in "$@"; do
  IRGS_NVCC_DEFS="$defs" python -m irgs_b200.build --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  echo "=== $defs" | tee -a gpurun_out/sweep.txt
  IMG=320 python scripts/diag_short.py 2>&1 | tail -n 2 | tee -a gpurun_out/sweep.txt
done
python -m irgs_b200.build --force > /dev/null 2>&1

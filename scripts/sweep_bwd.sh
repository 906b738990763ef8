#!/usr/bin/env bash
mkdir -p gpurun_out
for defs in "$@"; do
  IRGS_NVCC_DEFS="$defs" python -m irgs_b200.build --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  echo "=== $defs" | tee -a gpurun_out/sweep_bwd.txt
  python scripts/bwd_time.py 2>&1 | tail -n 3 | head -n 1 | tee -a gpurun_out/sweep_bwd.txt
done
python -m irgs_b200.build --force > /dev/null 2>&1

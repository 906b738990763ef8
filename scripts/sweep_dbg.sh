#!/usr/bin/env bash
for defs in "$@"; do
  IRGS_NVCC_DEFS="$defs" python -m irgs_b200.build --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  echo "=== $defs"
  python scripts/dbg_degenerate.py 2>&1 | grep "bad rays" | cut -c1-150
done
python -m irgs_b200.build --force > /dev/null 2>&1

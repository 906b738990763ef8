#!/usr/bin/env bash
# Same-box A/B of the forward kernel: the working tree's library against build/ab/lib_<name>.so, 2^24 rays, two rounds.
for i in 1 2; do
  IMG=256 python scripts/fused_time.py 2>&1 | tail -2 | head -1
  IRGS_LIB=build/ab/lib_$1.so IMG=256 python scripts/fused_time.py 2>&1 | tail -2 | head -1
done

#!/usr/bin/env bash
# forward time on 2^24 rays for the working tree's library and every build/ab/lib_*.so, two rounds
for i in 1 2; do
  echo -n "tree: "; IMG=256 python scripts/fused_time.py 2>&1 | tail -2 | head -1
  for f in build/ab/lib_*.so; do echo -n "$(basename $f): "; IRGS_LIB=$f IMG=256 python scripts/fused_time.py 2>&1 | tail -2 | head -1; done
done

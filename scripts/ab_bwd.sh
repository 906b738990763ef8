#!/usr/bin/env bash
# backward time on 2^22 rays for the working tree's library and every build/ab/lib_*.so, two rounds (CARVE = carve-out percent)
for i in 1 2; do
  echo -n "tree: "; MODES=0 IMG=400 python scripts/bwd_time.py 2>&1 | tail -1
  for f in build/ab/lib_*.so; do echo -n "$(basename $f): "; MODES=0 IRGS_LIB=$f IMG=400 python scripts/bwd_time.py 2>&1 | tail -1; done
done

#!/usr/bin/env bash
# backward time on 2^22 rays for the working tree's library and every build/ab/lib_*.so, two rounds
for i in 1 2; do
  echo -n "tree: "; IMG=400 python scripts/bwd_time.py 2>&1 | tail -3 | head -1
  for f in build/ab/lib_*.so; do echo -n "$(basename $f): "; IRGS_LIB=$f IMG=400 python scripts/bwd_time.py 2>&1 | tail -3 | head -1; done
done

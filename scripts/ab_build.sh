#!/usr/bin/env bash
# A/B builds of the native library for tuning experiments:
#   scripts/ab_build.sh <name> "<nvcc defines>" [file.cu ...]
# builds build/ab/lib_<name>.so from the current sources (use with IRGS_LIB=build/ab/lib_<name>.so).  Only the listed files are
# compiled with the defines (default: all); the others are linked from build/obj/base_<file>.o, compiled once per source state
# (delete build/obj to refresh them).
set -euo pipefail
cd "$(dirname "$0")/../irgs_b200/csrc"
name=$1; defs=$2; shift 2
all="capi lbvh trace trace_fwd shade surfel_params"
sel="${*:-$all}"
mkdir -p ../../build/ab ../../build/obj
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
objs=""
for f in $all; do
  if [[ " $sel " == *" $f "* || " $sel " == *" $f.cu "* ]]; then
    o=../../build/obj/${name}_$f.o
    $NV $defs -c $f.cu -o $o &
  else
    o=../../build/obj/base_$f.o
    if [[ ! -f $o || $f.cu -nt $o || internal.cuh -nt $o || trace_common.cuh -nt $o || shade_math.cuh -nt $o ]]; then $NV -c $f.cu -o $o & fi
  fi
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o ../../build/ab/lib_$name.so $objs
echo built build/ab/lib_$name.so

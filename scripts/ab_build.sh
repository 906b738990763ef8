#!/usr/bin/env bash
# A/B builds of the native library for tuning experiments: scripts/ab_build.sh <name> "<nvcc defines>" builds
# build/ab/lib_<name>.so from the current sources with the given -D flags (use with IRGS_LIB=build/ab/lib_<name>.so).
set -euo pipefail
cd "$(dirname "$0")/../irgs_b200/csrc"
mkdir -p ../../build/ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $2 \
    -o ../../build/ab/lib_$1.so capi.cu lbvh.cu trace.cu trace_fwd.cu shade.cu surfel_params.cu
echo built build/ab/lib_$1.so

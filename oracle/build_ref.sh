#!/usr/bin/env bash
# Build the UNMODIFIED reference tracer (/root/reference/submodules/surfel_tracer) into
# baseline/_ref/ (git-ignored; travels to the GPU box with gpurun).  Test/bench infrastructure only:
# nothing under irgs_b200/ imports it.  The reference sources are never copied into the repo history;
# the build happens in a scratch copy under /tmp because /root/reference is read-only.
#
# Two stages, exactly as the reference's own readme describes:
#   1. cmake + make  -> PTX of the three OptiX programs, embedded as build/optix_ptx.h
#   2. setup.py build_ext --inplace (TORCH_CUDA_ARCH_LIST=10.0: no GPU is visible here) -> surfel_tracer/_C*.so
set -euo pipefail
REF=${REF:-/root/reference/submodules/surfel_tracer}
REPO=$(cd "$(dirname "$0")/.." && pwd)
OUT=$REPO/baseline/_ref
SCRATCH=${SCRATCH:-/tmp/irgs_ref_build}
if [ ! -d "$REF" ]; then echo "reference not present at $REF (GPU box?) - keeping prebuilt $OUT"; exit 0; fi
rm -rf "$SCRATCH"; mkdir -p "$SCRATCH"
cp -r "$REF" "$SCRATCH/st"; chmod -R u+w "$SCRATCH/st"
mkdir -p "$SCRATCH/st/build"; cd "$SCRATCH/st/build"
cmake .. -DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc > "$SCRATCH/cmake.log" 2>&1
make -j"$(nproc)" > "$SCRATCH/make.log" 2>&1
cd "$SCRATCH/st"
TORCH_CUDA_ARCH_LIST=10.0 MAX_JOBS=$(nproc) python setup.py build_ext --inplace > "$SCRATCH/ext.log" 2>&1
mkdir -p "$OUT"; rm -rf "$OUT/surfel_tracer"
cp -r "$SCRATCH/st/surfel_tracer" "$OUT/surfel_tracer"
rm -rf "$OUT/surfel_tracer/__pycache__"
ls -la "$OUT/surfel_tracer"
echo "reference tracer built into $OUT"

#!/usr/bin/env bash
# backward experiments: (1) the flat kernel without its reductions, (2) the 1M-surfel scene
mkdir -p gpurun_out
IRGS_NVCC_DEFS="-DIRGS_DEBUG_SKIP_REDUCE" python -m irgs_b200.build --force > /dev/null 2>&1
echo "=== no reductions" | tee -a gpurun_out/exp_bwd.txt
python scripts/bwd_time.py 2>&1 | tail -n 3 | tee -a gpurun_out/exp_bwd.txt
python -m irgs_b200.build --force > /dev/null 2>&1
echo "=== normal" | tee -a gpurun_out/exp_bwd.txt
python scripts/bwd_time.py 2>&1 | tail -n 3 | tee -a gpurun_out/exp_bwd.txt
echo "=== 1M surfels" | tee -a gpurun_out/exp_bwd.txt
python bench.py --surfels 1000000 --img 400 --steps 2 --warmup 2 --no-e2e --no-fused --no-cpu-baseline 2>&1 | tail -n 1 | tee -a gpurun_out/exp_bwd.txt | cut -c1-1500

#!/usr/bin/env bash
# First GPU contact: probe the box, smoke, golden vectors from the reference, GPU tests, a short bench.
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv
  echo "nproc $(nproc)"; free -g | head -2
  ls -la /usr/lib/x86_64-linux-gnu/libnvoptix* /usr/lib/x86_64-linux-gnu/libnvidia-rtcore* 2>&1 | head
  ldconfig -p | grep -i -E "optix|rtcore" | head
} > gpurun_out/probe.log 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python oracle/gen_golden_ref.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden exit $?" >> gpurun_out/golden.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_first.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_first.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_first.log 2>&1; echo "ref exit $?" >> gpurun_out/bench_ref_first.log
tail -5 gpurun_out/*.log

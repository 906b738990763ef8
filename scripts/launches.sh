#!/usr/bin/env bash
mkdir -p gpurun_out
CMD="python bench.py --img ${IMG:-160} --steps 1 --warmup 1 --no-e2e --no-fused --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo done

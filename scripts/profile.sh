#!/usr/bin/env bash
# ncu evidence for the dominant kernels on a reduced C3 (same scene, fewer pixels): launch list + one full capture.
mkdir -p gpurun_out
CMD="python bench.py --img ${IMG:-160} --steps 1 --warmup 1 --no-e2e --no-fused --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'trace_forward|trace_backward_replay' -s 6 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/prof_plain.log gpurun_out/ncu_full.log

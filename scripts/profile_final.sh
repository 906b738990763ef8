#!/usr/bin/env bash
# Round evidence on the REAL C3 configuration: launch list of one timed step (kernel shares) + one full capture of the
# forward and the backward kernels on a 2^22-ray launch.  Outputs under gpurun_out/; summaries are copied to profiles/.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-fused --no-cpu-baseline --no-other"
$CMD > gpurun_out/final_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/final_plain.log; exit 1; }
IRGS_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ncu_launches.log 2>&1
IRGS_BENCH_PROFILE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'trace_forward|trace_backward_flat' -s 8 -c 2 -f -o gpurun_out/final_prof $CMD > gpurun_out/final_ncu_full.log 2>&1
tail -n 2 gpurun_out/final_plain.log | cut -c1-300

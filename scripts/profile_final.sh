#!/usr/bin/env bash
# Round evidence on the REAL C3 configuration: launch list (kernel shares) + one full capture of the forward and the
# backward-replay kernels on a 2^22-ray launch.  Outputs under gpurun_out/; copy the summaries into profiles/.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/final_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/final_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'trace_forward|trace_backward_replay' -s 8 -c 2 -f -o gpurun_out/final_prof $CMD > gpurun_out/final_ncu_full.log 2>&1
tail -n 2 gpurun_out/final_plain.log | cut -c1-300

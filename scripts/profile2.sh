#!/usr/bin/env bash
# one full ncu capture of the forward kernel on a 4M-ray chunk (diag workload)
mkdir -p gpurun_out
CMD="python scripts/prof_fwd.py"
$CMD > gpurun_out/prof2_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/prof2_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'trace_forward' -s 2 -c 1 -f -o gpurun_out/prof2 $CMD > gpurun_out/ncu_full2.log 2>&1
tail -n 3 gpurun_out/prof2_plain.log gpurun_out/ncu_full2.log

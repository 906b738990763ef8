#!/usr/bin/env bash
# Turns the artefacts of scripts/profile_final.sh (gpurun_out/final_*) into the committed summaries under profiles/.
set -e
R=gpurun_out/final_prof.ncu-rep
python scripts/ncu_regions.py $R trace_forward 0.4 > profiles/r${ROUND:-02}_fwd_regions_final.txt
python scripts/ncu_phases.py $R > profiles/r${ROUND:-02}_fwd_phases_final.txt
python scripts/ncu_regions.py $R trace_backward_flat 0.6 > profiles/r${ROUND:-02}_bwd_flat_regions.txt
python scripts/launch_summary.py gpurun_out/final_launches.csv > profiles/r${ROUND:-02}_step_launches.txt
cp gpurun_out/final_launches.csv profiles/r${ROUND:-02}_step_launches.csv
python scripts/ncu_summary.py $R ${RAYS_PER_LAUNCH:-16384000}

#!/bin/bash
# Whole missions with the round's final kernels: circle swaps with goal planning on the device (safety audit every step),
# the forest crossing through lsc_sim with the goal planner on the device and on the host.
OUT=${1:-gpurun_out/missions.txt}; : > $OUT
for N in 64 256; do timeout 600 python tools/mission_run.py $N 800 1 2>&1 | tail -1 >> $OUT; done
BT=tests/golden/worlds/simple_forest.bt
for PL in device host; do
  echo "lsc_sim synthetic_forest10 mode/goal=prior_based goal/planner=$PL:" >> $OUT
  lsc_planner_b200/host/lsc_sim mission=tests/golden/missions/synthetic_forest10.json world/file_name=$BT mode/goal=prior_based goal/planner=$PL \
      multisim/max_planner_iteration=400 multisim/save_result=false 2>&1 | grep -E "flight time|distance|planning time per agent|safety ratio|is_collided" >> $OUT
done
cat $OUT

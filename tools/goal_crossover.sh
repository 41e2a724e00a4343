#!/bin/bash
# Host grid planner (threaded) against k_goal_astar in the simulator, random swarms in simple_forest.bt.
# Usage: tools/goal_crossover.sh "<N> <N> ..." <out.txt>
OUT=${2:-gpurun_out/goal_crossover.txt}; : > $OUT
BT=tests/golden/worlds/simple_forest.bt
for N in $1; do
  python - $N <<'PY'
import json, sys, numpy as np
sys.path.insert(0, ".")
import lsc_planner_b200 as L
n = int(sys.argv[1])
tmp = L.ReplanEngine(2, L.Param(world_use_octomap=True)); tmp.set_octomap_file("tests/golden/worlds/simple_forest.bt"); dm = tmp.distmap(); tmp.close()
scn = L.scenarios.random_forest(n, dm["sqdist"], dm["off"], seed=0)
ms = {"quadrotors": {"crazyflie": {"max_vel": [1.0, 1.0, 1.0], "max_acc": [2.0, 2.0, 2.0], "radius": 0.15, "nominal_velocity": 1.0, "downwash": 2.0}},
      "world": [{"dimension": [-5.0, -5.0, 0.0, 5.0, 5.0, 2.5]}],
      "agents": [{"type": "crazyflie", "cid": k + 1, "start": [float(v) for v in scn.start[k]], "goal": [float(v) for v in scn.goal[k]]} for k in range(n)],
      "obstacles": []}
json.dump(ms, open(f"/tmp/rf{n}.json", "w"))
PY
  for PL in host device; do
    lsc_planner_b200/host/lsc_sim mission=/tmp/rf$N.json world/file_name=$BT mode/goal=prior_based goal/planner=$PL \
        multisim/max_planner_iteration=31 multisim/save_result=false > /tmp/sim_${N}_${PL}.log 2>&1
    echo "N=$N goal/planner=$PL: $(grep -E 'planning time per agent|goal planning time' /tmp/sim_${N}_${PL}.log | tr '\n' ' ')" >> $OUT
  done
done
cat $OUT

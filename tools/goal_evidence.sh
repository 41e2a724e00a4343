#!/bin/bash
# Goal planning on the device with an octomap: parity tests, bench lines beside the static-goal ones, lsc_sim with both planners.
T=${1:-goal}; O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_goal.py -x -q ) > $O/${T}_pytest_goal.log 2>&1; tail -3 $O/${T}_pytest_goal.log
for GM in prior_based static; do
  timeout 600 python bench.py --workload random_forest --agents 512 --goal-mode $GM --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_random_forest_512_$GM.json 2> $O/${T}_bench_rf512_$GM.err
  timeout 600 python bench.py --workload circle_forest --agents 1024 --goal-mode $GM --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_circle_forest_1024_$GM.json 2> $O/${T}_bench_cf1024_$GM.err
done
BT=tests/golden/worlds/simple_forest.bt
for PL in device host; do
  ( time lsc_planner_b200/host/lsc_sim mission=tests/golden/missions/synthetic_forest10.json world/file_name=$BT mode/goal=prior_based goal/planner=$PL \
      multisim/max_planner_iteration=400 result=$O/${T}_sim_$PL.csv summary=$O/${T}_sim_summary_$PL.csv ) > $O/${T}_sim_$PL.log 2>&1
  tail -12 $O/${T}_sim_$PL.log
done
python - "$T" <<'PY'
import json, sys
t = sys.argv[1]
for w in ("random_forest_512", "circle_forest_1024"):
    for gm in ("prior_based", "static"):
        try:
            d = json.loads(open(f"gpurun_out/{t}_bench_{w}_{gm}.json").read().strip().splitlines()[-1])
            print(w, gm, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), d["kernel_ms_per_step"], d.get("goal_planning"))
        except Exception as ex:
            print(w, gm, "failed", ex)
PY

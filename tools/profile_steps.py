#!/usr/bin/env python
"""Runs N device-resident closed-loop steps of a workload (for ncu captures): tools/profile_steps.py <steps> [workload agents]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import lsc_planner_b200 as L

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
workload = sys.argv[2] if len(sys.argv) > 2 else "circle_forest"
agents = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
scn, bt = bench.make_scenario(workload, agents)
e = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=scn.use_octomap), scn.agents)
if scn.use_octomap:
    e.set_octomap_file(bt)
e.set_states(scn.start); e.set_goals(scn.goal)
for _ in range(steps):
    e.replan_resident(1)
o = e.fetch()
print("steps", steps, "failed", int((o["qp_status"] != 0).sum()), "kept/agent", float(o["lsc_pairs_kept"].mean()))

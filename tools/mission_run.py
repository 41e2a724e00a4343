#!/usr/bin/env python
"""Closed-loop mission on the GPU with goal planning on the device (prior_based, no octomap): steps until every agent is
within goal_threshold of its goal, safety audit every step. Usage: tools/mission_run.py [agents] [max_steps] [goal_mode]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lsc_planner_b200 as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
max_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
goal_mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
scn = L.scenarios.circle_swap(n)
e = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max, goal_mode=goal_mode), scn.agents)
e.set_states(scn.start); e.set_goals(scn.goal)
worst = np.inf; fails = 0; done_at = None
for step in range(max_steps):
    e.replan_resident(1)
    out = e.fetch()
    r, c = e.safety_audit(0.1, 0.2)
    worst = min(worst, r.min()); fails += int((out["qp_status"] != 0).sum())
    d = np.linalg.norm(out["next_position"] - scn.goal, axis=1)
    if step % 50 == 0:
        print(f"step {step:4d} max dist to goal {d.max():.2f} mean {d.mean():.2f} min safety ratio so far {worst:.4f} qp failures so far {fails} retreats {int(out['goal_kind'].sum())}")
    if d.max() < 0.1:
        done_at = step + 1; break
print(f"RESULT agents {n} goal_mode {goal_mode} finished_at_step {done_at} flight_time_s {None if done_at is None else 0.2 * done_at} min_safety_ratio {worst:.5f} qp_failures {fails}")

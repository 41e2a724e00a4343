#!/usr/bin/env python
"""Wall-clock latency of the operator-level entry for ONE problem (a batch-of-one TrajOptimizer::solve through
lscgpu_qp_solve_batch: four copies in, one kernel chain, one copy out) — the cost a per-agent drop-in of the reference's
TrajOptimizer pays per call. Usage: tools/op_latency.py [agents]."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lsc_planner_b200 as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
scn = L.scenarios.circle_swap(n)
e = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max), scn.agents)
e.set_states(scn.start); e.set_goals(scn.goal)
e.replan_resident(14)
prev = e.fetch().copy()
pos, vel, acc = prev["next_position"], prev["next_velocity"], prev["next_acceleration"]
out = e.replan(pos, vel, acc, scn.goal).copy()
pred = e.initial_traj()
a = 0
nr, d = e.get_lsc(a)
others = [j for j in range(scn.n) if j != a]
state = np.concatenate([pos[a], vel[a], acc[a]]).astype(np.float64)[None]
args = (np.array([a]), state, scn.goal[a:a + 1].astype(np.float64), [0, len(others)], nr, pred[others], d)
for _ in range(50):
    r = e.qp_solve_batch(*args)
ts = []
for _ in range(500):
    t0 = time.perf_counter(); r = e.qp_solve_batch(*args); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e6
x = r["x"].reshape(1, 3, 5, 6).transpose(0, 2, 3, 1)
print(f"batch-of-one lscgpu_qp_solve_batch, {n - 1} obstacles ({(n - 1) * 27} LSC rows), through the ctypes binding: "
      f"p50 {np.percentile(ts, 50):.1f} us, p90 {np.percentile(ts, 90):.1f} us, min {ts.min():.1f} us; status {int(r['status'][0])}, "
      f"iterations {int(r['iterations'][0])}, max |x - step's trajectory| {np.abs(x[0] - out['traj'][a]).max():.1e}")

#!/usr/bin/env python
"""Diagnostic: teacher-forced forest10 run with goal planning on the device; prints where engine and oracle differ."""
import os, sys, pathlib, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import lsc_planner_b200 as L
import test_gpu_goal as T

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
bt = os.path.join(ROOT, "tests", "golden", "worlds", "simple_forest.bt")
tmp = pathlib.Path(tempfile.mkdtemp())
scn = T._forest_case("forest10", tmp, bt)
n = scn.n
omap = O.Map.from_bt(bt, scn.world_min, scn.world_max)
sw = T._oracle_swarm(scn, omap)
e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=True, goal_mode=1), scn.agents)
e.set_octomap_file(bt)
for step in range(steps):
    pos, vel, acc = sw.state()
    e.set_sfc(sw.boxes(), np.full(n, 1 if sw.seq == 0 else 0, np.int32))
    e.set_prev_traj(sw.traj(), sw.seq)
    sw.step()
    out = e.replan(pos, vel, acc, scn.goal)
    g_o, k_o = sw.goals()
    q = sw.qp()
    diffs = np.abs(out["traj"] - sw.traj()).reshape(n, -1).max(1)
    st = e.step_stats()
    for a in np.flatnonzero(diffs > 2e-6):
        print(f"step {step} agent {a} diff {diffs[a]:.3e} status {out['qp_status'][a]}/{q['status'][a]} flags {out['flags'][a]} "
              f"maxviol {q['maxviol'][a]:.3e} it {out['qp_iterations'][a]} active {out['qp_active'][a]} cost {out['qp_cost'][a]:.12g} "
              f"ts {out['terminal_segments'][a]} goal {out['current_goal'][a]} / {g_o[a]} kind {k_o[a]}")
        print("   oracle qp keys", {k: (v[a] if hasattr(v, '__len__') else v) for k, v in q.items()})
        bx = e.get_sfc()[0][a]; print("   sfc equal", np.array_equal(bx.view(np.uint32), sw.boxes()[a].view(np.uint32)))
    print(f"step {step} ms {st['ms_total']:.3f} astar {st['astar_expansions']} maxdiff {diffs.max():.2e}", flush=True)
    sw.advance()

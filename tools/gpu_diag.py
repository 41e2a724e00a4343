#!/usr/bin/env python
"""Per-step diagnostics of a closed-loop run on the GPU (iterations, failures, kernel times)."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import lsc_planner_b200 as L

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="circle_forest"); ap.add_argument("--agents", type=int, default=1024)
ap.add_argument("--steps", type=int, default=110); ap.add_argument("--every", type=int, default=10)
a = ap.parse_args()
scn, bt = bench.make_scenario(a.workload, a.agents)
if scn is None:
    tmp = L.ReplanEngine(2, L.Param(world_use_octomap=True)); tmp.set_octomap_file(bt); dm = tmp.distmap()
    scn = L.scenarios.random_forest(a.agents, dm["sqdist"], dm["off"], seed=0); tmp.close()
e = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=scn.use_octomap), scn.agents)
if scn.use_octomap: e.set_octomap_file(bt)
e.set_states(scn.start); e.set_goals(scn.goal); e.set_profiling(True)
for s in range(a.steps):
    e.replan_resident()
    if s % a.every == 0 or s == a.steps - 1:
        o = e.fetch(); st = e.step_stats()
        it = o["qp_iterations"]; kc = o["qp_kcycles"].astype(float) * 1024 / 1.965e3   # us
        slow = np.argsort(-kc)[:3]
        pk = o["qp_price_kcycles"].astype(float) * 1024 / 1.965e3
        print("      slowest agents (id, us, price us, iters, sweeps, status, active, kept):", [(int(a), f"{kc[a]:.0f}", f"{pk[a]:.0f}", int(it[a]), int(o["qp_sweeps"][a]), int(o["qp_status"][a]), int(o["qp_active"][a]), int(o["lsc_pairs_kept"][a])) for a in slow],
              f"| us p50 {np.percentile(kc,50):.0f} p90 {np.percentile(kc,90):.0f} p99 {np.percentile(kc,99):.0f}",
              "| sweeps p50 %.0f p99 %.0f max %.0f" % tuple(np.percentile(o["qp_sweeps"], [50, 99, 100])),
              f"| rows priced/iter {st['qp_rows_priced'] / max(st['qp_iterations'] + scn.n, 1):.0f}")
        dist = np.linalg.norm(o["next_position"] - scn.goal, axis=1)
        lk = o["lsc_kcycles"].astype(float) * 1024 / 1.965e3
        print(f"step {s:4d} ms tot {st['ms_total']:.3f} predict {st['ms_predict']:.3f} plan {st['ms_plan']:.3f} commit {st['ms_commit']:.3f} | corridor us p50 {np.percentile(lk,50):.0f} p99 {np.percentile(lk,99):.0f} max {lk.max():.0f} | iters mean {it.mean():.1f} "
              f"p50 {np.percentile(it,50):.0f} p99 {np.percentile(it,99):.0f} max {it.max()} | status {np.bincount(o['qp_status'], minlength=3)} "
              f"active max {o['qp_active'].max()} | kept/agent {st['lsc_pairs_kept']/scn.n:.0f} sweeps {st['qp_full_passes']/scn.n:.2f} "
              f"flags {np.bincount(o['flags'], minlength=4)} | sfc grown in block {o['sfc_in_block'].mean():.2f} | dist-to-goal mean {dist.mean():.2f}")

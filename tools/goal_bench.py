#!/usr/bin/env python
"""Goal planning on the host cores: the product's grid planner (lsc_planner_b200/host/libhostgoal.so)
beside the reference's own A* (oracle/_ref/libref_astar.so = <ref>/src/Astar-3D compiled as is) and the oracle's
restatement, on the occupancy grid of world/simple_forest.bt and on whole goal-planning problems. CPU only.
Usage: tools/goal_bench.py > profiles/<round>_goal_planning_cpu.json"""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS"); u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"); f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
H = C.CDLL(os.path.join(ROOT, "lsc_planner_b200", "host", "libhostgoal.so"))
H.host_astar.argtypes = [i32p, u8p, i32p, i32p, i32p, C.c_int, C.POINTER(C.c_longlong)]; H.host_astar.restype = C.c_int
H.host_goal_plan.argtypes = [C.c_int, C.c_int, f32p, f32p, f32p, f32p, f64p, f64p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_double, f32p, f32p] + [C.c_double] * 5 + [f32p, C.POINTER(C.c_longlong)]
H.host_goal_plan.restype = C.c_int
ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_astar.so")
R = C.CDLL(ref_so) if os.path.exists(ref_so) else None
if R:
    R.ref_astar_plan.argtypes = [i32p, u8p, i32p, i32p, i32p, C.c_int]; R.ref_astar_plan.restype = C.c_int

wmin, wmax = [-5, -5, 0], [5, 5, 2.5]
omap = O.Map.from_bt(os.path.join(ROOT, "tests", "golden", "worlds", "simple_forest.bt"), wmin, wmax)
# occupancy grid of GridBasedPlanner::updateGridMap for a crazyflie: 41 x 41 x 11 cells of 0.25 m, dist < 0.15 + 0.1
dim = np.array([41, 41, 11], np.int32)
grid = np.zeros(tuple(dim), np.uint8)
for i in range(41):
    for j in range(41):
        for k in range(11):
            p = np.float32([-5 + 0.25 * i, -5 + 0.25 * j, 0.25 * k])
            grid[i, j, k] = omap.distance(p) < 0.15 + np.float32(0.1)
rng = np.random.default_rng(0)
free = np.argwhere(grid[:, :, 4] == 0)
pairs = [(np.append(free[rng.integers(len(free))], 4).astype(np.int32), np.append(free[rng.integers(len(free))], 4).astype(np.int32))
         for _ in range(200)]
buf = np.zeros((4096, 3), np.int32); ex = C.c_longlong(0)
def timed(fn):
    t0 = time.perf_counter(); tot = 0
    for s, g in pairs: tot += fn(s, g)
    return time.perf_counter() - t0, tot
t_host, len_host = timed(lambda s, g: H.host_astar(dim, grid, s, g, buf, 4096, C.byref(ex)))
t_orc, len_orc = timed(lambda s, g: len(O.astar(grid, s, g)[0]))
out = {"grid": "simple_forest.bt, 41x41x11 cells of 0.25 m, %d occupied" % int(grid.sum()), "searches": len(pairs),
       "host_astar_ms_per_search": 1e3 * t_host / len(pairs), "oracle_astar_ms_per_search": 1e3 * t_orc / len(pairs),
       "path_cells_total": int(len_host), "same_total_as_oracle": bool(len_host == len_orc)}
if R:
    t_ref, len_ref = timed(lambda s, g: R.ref_astar_plan(dim, grid, s, g, buf, 4096))
    out["reference_astar_ms_per_search"] = 1e3 * t_ref / len(pairs); out["same_total_as_reference"] = bool(len_ref == len_host)
# whole goal-planning problems: 64 agents in the forest, one call per agent (grid build + priority + A* + line of sight)
n = 64
def free_pts(k):
    pts = []
    while len(pts) < k:
        p = (np.float32(wmin) + 0.6 + (np.float32(wmax) - np.float32(wmin) - 1.2) * rng.random(3)).astype(np.float32)
        if omap.distance(p) >= 0.4: pts.append(p)
    return np.array(pts, np.float32)
pos, desired = free_pts(n), free_pts(n)
prev = np.repeat(pos[:, None, :], 30, axis=1).astype(np.float32)
radius = np.full(n, 0.15); dw = np.full(n, 2.0)
sq = np.ascontiguousarray(np.minimum(omap.sqdist(), 255).astype(np.uint8))
size = np.asarray(omap.size, np.int32); off = np.asarray(omap.off, np.int32)
g = np.zeros(3, np.float32)
t0 = time.perf_counter()
for a in range(n):
    H.host_goal_plan(a, n, pos, desired, np.ascontiguousarray(prev.reshape(n, 90)), np.ascontiguousarray(prev[a, 29]), radius, dw,
                     sq.ctypes.data, size.ctypes.data, off.ctypes.data, 0.1, np.float32(wmin), np.float32(wmax), 0.25, 0.1, 0.1, 2.0, 0.4,
                     g, C.byref(ex))
t_hg = time.perf_counter() - t0
t0 = time.perf_counter()
for a in range(n):
    O.goal_plan(a, pos, desired, prev, prev[a, 29], radius, dw, omap, wmin, wmax)
t_og = time.perf_counter() - t0
out["goal_plan_64_agents"] = {"host_ms_per_agent_1_thread_uncached_grid": 1e3 * t_hg / n, "oracle_ms_per_agent_1_thread": 1e3 * t_og / n,
                              "reference_published_ms_per_agent": 1.95, "note": "reference figure: log/summary_LSC_16agents.csv "
                              "(goal_planning_time, author's workstation); lsc_sim additionally caches the static grid per radius "
                              "and spreads agents over the host threads"}
out["cores"] = os.cpu_count()
print(json.dumps(out, indent=1))

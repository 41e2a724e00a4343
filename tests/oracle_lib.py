"""ctypes binding of the CPU oracle (oracle/liblsc_oracle.so). Test infrastructure only.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(ORACLE_DIR, "liblsc_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("api.cpp", "geom.hpp", "edt.hpp", "qp.hpp", "swarm.hpp", "goal.hpp")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "liblsc_oracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_gjk.argtypes = [f64p, C.c_int, f64p]; L.orc_gjk.restype = C.c_int
        L.orc_lsc_pair.argtypes = [f32p, f32p, C.c_double, C.c_double, C.c_double, C.c_double, f32p, f64p, i32p]
        L.orc_map_from_bt.argtypes = [C.c_char_p, f32p, f32p, C.c_double]; L.orc_map_from_bt.restype = C.c_void_p
        L.orc_map_from_voxels.argtypes = [i32p, C.c_int, f32p, f32p, C.c_double]; L.orc_map_from_voxels.restype = C.c_void_p
        L.orc_map_info.argtypes = [C.c_void_p, i32p, i32p, i64p, i64p]
        L.orc_map_occupied.argtypes = [C.c_void_p, i32p]
        L.orc_map_sqdist.argtypes = [C.c_void_p, i32p]
        L.orc_map_distance.argtypes = [C.c_void_p, f32p]; L.orc_map_distance.restype = C.c_float
        L.orc_map_free.argtypes = [C.c_void_p]
        L.orc_sfc_expand.argtypes = [C.c_void_p, f32p, f32p, C.c_double, f32p, f32p, C.c_double, f32p, i64p]
        L.orc_sfc_expand.restype = C.c_int
        L.orc_tables_create.argtypes = [C.c_double] * 3; L.orc_tables_create.restype = C.c_void_p
        L.orc_tables_free.argtypes = [C.c_void_p]
        L.orc_tables_get.argtypes = [C.c_void_p] + [f64p] * 7
        L.orc_terminal_segments.argtypes = [f32p, f32p, C.c_double, C.c_double]; L.orc_terminal_segments.restype = C.c_int
        qp_args = [C.c_void_p, f64p, f64p, C.c_int, f64p, f64p, f64p, f64p, C.c_int, i32p, f64p, f64p]
        L.orc_qp_solve.argtypes = qp_args + [f64p, f64p]; L.orc_qp_solve.restype = C.c_int
        L.orc_qp_dense.argtypes = qp_args + [C.c_void_p, f64p, f64p, f64p, f64p, f64p, f64p, f64p, C.c_int]
        L.orc_qp_dense.restype = C.c_int
        L.orc_set_tier_threshold.argtypes = [C.c_double]
        L.orc_qp_solve_slack.argtypes = qp_args + [i32p, C.c_double, f64p, f64p, f64p]; L.orc_qp_solve_slack.restype = C.c_int
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        L.orc_swarm_set_slack_weight.argtypes = [C.c_void_p, C.c_double]
        L.orc_swarm_set_warm_start.argtypes = [C.c_void_p, C.c_int]
        L.orc_swarm_get_active.argtypes = [C.c_void_p, C.c_int, i32p]; L.orc_swarm_get_active.restype = C.c_int
        L.orc_swarm_get_warm_stats.argtypes = [C.c_void_p, i64p]
        L.orc_swarm_get_reset_ever.argtypes = [C.c_void_p, u8]
        L.orc_swarm_set_reset_ever.argtypes = [C.c_void_p, u8]
        L.orc_swarm_get_slack.argtypes = [C.c_void_p, f64p, i32p]
        L.orc_swarm_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                       f32p, f32p, f64p, f64p, f64p, f64p, f64p]
        L.orc_swarm_create.restype = C.c_void_p
        L.orc_swarm_free.argtypes = [C.c_void_p]
        L.orc_swarm_set_map.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_swarm_set_capture.argtypes = [C.c_void_p, C.c_int]
        L.orc_swarm_set_state.argtypes = [C.c_void_p, f32p, f32p, f32p]
        L.orc_swarm_set_goals.argtypes = [C.c_void_p, f32p]
        L.orc_swarm_set_traj.argtypes = [C.c_void_p, f32p, C.c_int]
        L.orc_swarm_set_boxes.argtypes = [C.c_void_p, f32p, i32p]
        L.orc_swarm_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_swarm_step_list.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int]
        L.orc_swarm_safety_audit.argtypes = [C.c_void_p, C.c_double, C.c_double, f64p, i32p]
        L.orc_swarm_set_goal_mode.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 5
        L.orc_swarm_set_desired_goals.argtypes = [C.c_void_p, f32p]
        L.orc_swarm_get_goals.argtypes = [C.c_void_p, f32p, i32p]
        L.orc_swarm_astar_expansions.argtypes = [C.c_void_p]; L.orc_swarm_astar_expansions.restype = C.c_longlong
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        L.orc_astar.argtypes = [i32p, u8p, i32p, i32p, i32p, C.c_int, C.POINTER(C.c_longlong)]; L.orc_astar.restype = C.c_int
        L.orc_goal_plan.argtypes = [C.c_int, C.c_int, f32p, f32p, f32p, f32p, f64p, f64p, C.c_void_p, C.c_double, f32p, f32p,
                                    C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, f32p, C.POINTER(C.c_longlong)]
        L.orc_goal_plan.restype = C.c_int
        L.orc_swarm_advance.argtypes = [C.c_void_p]
        L.orc_swarm_seq.argtypes = [C.c_void_p]; L.orc_swarm_seq.restype = C.c_int
        L.orc_swarm_get_traj.argtypes = [C.c_void_p, f32p]
        L.orc_swarm_get_pred.argtypes = [C.c_void_p, f32p]
        L.orc_swarm_get_state.argtypes = [C.c_void_p, f32p, f32p, f32p]
        L.orc_swarm_get_boxes.argtypes = [C.c_void_p, f32p]
        L.orc_swarm_get_qp.argtypes = [C.c_void_p, f64p, i32p, i32p, i32p, i32p, f64p, f64p]
        L.orc_swarm_get_counters.argtypes = [C.c_void_p, i64p]
        L.orc_swarm_reset_counters.argtypes = [C.c_void_p]
        L.orc_swarm_get_capture.argtypes = [C.c_void_p, f32p, f64p, i32p]
        _LIB = L
    return _LIB


# ------------------------------------------------------------------------------------------
def gjk(points: np.ndarray):
    pts = np.ascontiguousarray(points, np.float64)
    v = np.zeros(3)
    it = lib().orc_gjk(pts, len(pts), v)
    return v, it


def lsc_pair(own, obs, r_i=0.15, dw_i=2.0, r_j=0.15, dw_j=2.0):
    own = np.ascontiguousarray(own, np.float32).reshape(30, 3)
    obs = np.ascontiguousarray(obs, np.float32).reshape(30, 3)
    n = np.zeros((5, 3), np.float32); d = np.zeros((5, 6)); it = np.zeros(5, np.int32)
    lib().orc_lsc_pair(own, obs, r_i, dw_i, r_j, dw_j, n, d, it)
    return n, d, it


def astar(grid: np.ndarray, start, goal):
    """Oracle A* on an occupancy grid [i][j][k] (non-zero = occupied): (path cells [n][3], expansions)."""
    grid = np.ascontiguousarray(grid, np.uint8)
    dim = np.asarray(grid.shape, np.int32)
    path = np.zeros((int(dim.sum()) * 8 + 16, 3), np.int32); ex = C.c_longlong(0)
    n = lib().orc_astar(dim, grid, np.asarray(start, np.int32), np.asarray(goal, np.int32), path, len(path), C.byref(ex))
    assert n <= len(path)
    return path[:n].copy(), ex.value


def goal_plan(a, pos, desired, prev_traj, init_end, radius, downwash, omap, wmin, wmax, res=0.1, grid_resolution=0.25,
              grid_margin=0.1, goal_threshold=0.1, goal_radius=2.0, priority_dist_threshold=0.4):
    """Oracle goalPlanningWithPriority for agent a: (goal[3], kind, A* expansions)."""
    n = len(pos)
    out = np.zeros(3, np.float32); ex = C.c_longlong(0)
    kind = lib().orc_goal_plan(a, n, np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(desired, np.float32),
                               np.ascontiguousarray(prev_traj, np.float32).reshape(n, 90), np.ascontiguousarray(init_end, np.float32),
                               np.ascontiguousarray(np.broadcast_to(radius, (n,)), np.float64),
                               np.ascontiguousarray(np.broadcast_to(downwash, (n,)), np.float64),
                               omap.h if omap is not None else None, res, np.asarray(wmin, np.float32), np.asarray(wmax, np.float32),
                               grid_resolution, grid_margin, goal_threshold, goal_radius, priority_dist_threshold, out, C.byref(ex))
    return out, kind, ex.value


class Map:
    def __init__(self, handle, wmin, wmax, res):
        if not handle:
            raise RuntimeError("oracle map creation failed")
        self.h = handle; self.wmin = np.asarray(wmin, np.float32); self.wmax = np.asarray(wmax, np.float32); self.res = res
        size = np.zeros(3, np.int32); off = np.zeros(3, np.int32); nocc = np.zeros(1, np.int64); nn = np.zeros(1, np.int64)
        lib().orc_map_info(self.h, size, off, nocc, nn)
        self.size, self.off, self.n_occ, self.n_nodes = size, off, int(nocc[0]), int(nn[0])

    @classmethod
    def from_bt(cls, path, wmin, wmax, res=0.1):
        wmin = np.asarray(wmin, np.float32); wmax = np.asarray(wmax, np.float32)
        return cls(lib().orc_map_from_bt(path.encode(), wmin, wmax, res), wmin, wmax, res)

    @classmethod
    def from_voxels(cls, keys, wmin, wmax, res=0.1):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        wmin = np.asarray(wmin, np.float32); wmax = np.asarray(wmax, np.float32)
        return cls(lib().orc_map_from_voxels(keys, len(keys), wmin, wmax, res), wmin, wmax, res)

    def occupied(self):
        k = np.zeros((self.n_occ, 3), np.int32)
        lib().orc_map_occupied(self.h, k)
        return k

    def sqdist(self):
        out = np.zeros(tuple(self.size), np.int32)
        lib().orc_map_sqdist(self.h, out)
        return out

    def distance(self, p):
        return float(lib().orc_map_distance(self.h, np.asarray(p, np.float32)))

    def sfc_expand(self, point, goal, radius=0.15):
        box = np.zeros(6, np.float32); lk = np.zeros(1, np.int64)
        ok = lib().orc_sfc_expand(self.h, self.wmin, self.wmax, self.res, np.asarray(point, np.float32),
                                  np.asarray(goal, np.float32), radius, box, lk)
        return bool(ok), box, int(lk[0])

    def __del__(self):
        try:
            lib().orc_map_free(self.h)
        except Exception:
            pass


class Tables:
    def __init__(self, dt=0.2, w=0.01, wT=1.0):
        self.h = lib().orc_tables_create(dt, w, wT)
        self.dt, self.w, self.wT = dt, w, wT
        self.Qb = np.zeros((6, 6)); self.A17 = np.zeros((17, 30)); self.Xp = np.zeros((30, 3)); self.Z = np.zeros((30, 13))
        self.G = np.zeros((5, 30, 13)); self.Xs = np.zeros((5, 30, 3)); self.xg = np.zeros((5, 30))
        lib().orc_tables_get(self.h, self.Qb, self.A17, self.Xp, self.Z, self.G, self.Xs, self.xg)

    def _args(self, state, goal, ts, lb, ub, vmax, amax, rows):
        row_m = np.ascontiguousarray([r[0] for r in rows], np.int32) if rows else np.zeros(0, np.int32)
        row_a = np.ascontiguousarray([r[1] for r in rows], np.float64).reshape(-1, 3) if rows else np.zeros((0, 3))
        row_rhs = np.ascontiguousarray([r[2] for r in rows], np.float64).reshape(-1, 6) if rows else np.zeros((0, 6))
        return (self.h, np.ascontiguousarray(state, np.float64).reshape(9), np.ascontiguousarray(goal, np.float64),
                int(ts), np.ascontiguousarray(lb, np.float64), np.ascontiguousarray(ub, np.float64),
                np.ascontiguousarray(vmax, np.float64), np.ascontiguousarray(amax, np.float64), len(rows),
                row_m, row_a, row_rhs)

    def solve(self, state, goal, ts, lb, ub, vmax, amax, rows):
        """rows: list of (m, a3, rhs6). Returns dict(x, cost, status, iters, n_active, kkt, maxviol)."""
        x = np.zeros(90); info = np.zeros(5)
        st = lib().orc_qp_solve(*self._args(state, goal, ts, lb, ub, vmax, amax, rows), x, info)
        return dict(x=x, cost=info[0], status=st, iters=int(info[1]), n_active=int(info[2]), kkt=info[3], maxviol=info[4])

    def solve_slack(self, state, goal, ts, lb, ub, vmax, amax, rows, row_slack, slack_w=1.0):
        """The QP with slack variables: row_slack[r] = 1 for the (m, a3, rhs6) entries of obstacles in obs_slack_indices."""
        x = np.zeros(90); info = np.zeros(6); eps = np.zeros(max(len(rows), 1))
        st = lib().orc_qp_solve_slack(*self._args(state, goal, ts, lb, ub, vmax, amax, rows),
                                      np.ascontiguousarray(row_slack, np.int32), float(slack_w), x, eps, info)
        return dict(x=x, eps=eps[:len(rows)], cost=info[0], status=st, iters=int(info[1]), n_active=int(info[2]), kkt=info[3],
                    maxviol=info[4], slack_cost=info[5])

    def dense(self, state, goal, ts, lb, ub, vmax, amax, rows, boxes=None):
        max_in = 6 * len(rows) + 252 + 162 + 8
        P = np.zeros((90, 90)); q = np.zeros(90); c0 = np.zeros(1); Aeq = np.zeros((51, 90)); beq = np.zeros(51)
        Ain = np.zeros((max_in, 90)); bin_ = np.zeros(max_in)
        bx = None
        if boxes is not None:
            bx_arr = np.ascontiguousarray(boxes, np.float32).reshape(5, 6)
            bx = bx_arr.ctypes.data_as(C.c_void_p)
        n = lib().orc_qp_dense(*self._args(state, goal, ts, lb, ub, vmax, amax, rows), bx, P, q, c0, Aeq, beq, Ain, bin_, max_in)
        assert n >= 0
        return dict(P=P, q=q, c0=float(c0[0]), Aeq=Aeq, beq=beq, Ain=Ain[:n], bin=bin_[:n],
                    lb=np.asarray(lb, float), ub=np.asarray(ub, float))

    def __del__(self):
        try:
            lib().orc_tables_free(self.h)
        except Exception:
            pass


def terminal_segments(pos, goal, v_nom=1.0, dt=0.2):
    return lib().orc_terminal_segments(np.asarray(pos, np.float32), np.asarray(goal, np.float32), v_nom, dt)


class Swarm:
    """Oracle swarm stepper; mirrors lsc_planner_b200.Engine's stepping surface."""

    def __init__(self, n, world_min, world_max, radius=0.15, downwash=2.0, vmax=(1, 1, 1), amax=(2, 2, 2), v_nom=1.0,
                 dt=0.2, w=0.01, wT=1.0, res=0.1, reset_threshold=0.15, use_octomap=False, omap: Map | None = None):
        self.n = n
        bc = lambda v, k: np.ascontiguousarray(np.broadcast_to(np.asarray(v, np.float64), (n,) + ((k,) if k else ())), np.float64)
        self.h = lib().orc_swarm_create(n, dt, w, wT, res, reset_threshold, int(use_octomap),
                                        np.asarray(world_min, np.float32), np.asarray(world_max, np.float32),
                                        bc(radius, 0), bc(downwash, 0), bc(vmax, 3), bc(amax, 3), bc(v_nom, 0))
        self.map = omap
        if omap is not None:
            lib().orc_swarm_set_map(self.h, omap.h)

    def set_capture(self, on=True): lib().orc_swarm_set_capture(self.h, int(on))

    def set_slack_weight(self, w): lib().orc_swarm_set_slack_weight(self.h, float(w))

    def active_rows(self, a):
        """Canonical ids of the rows active at agent a's last successful solve (None when it failed / carried slack)."""
        ids = np.zeros(39, np.int32); n = lib().orc_swarm_get_active(self.h, a, ids)
        return None if n < 0 else ids[:n].copy()

    def set_warm_start(self, on=True): lib().orc_swarm_set_warm_start(self.h, int(on))

    def warm_stats(self):
        c = np.zeros(2, np.int64); lib().orc_swarm_get_warm_stats(self.h, c); return int(c[0]), int(c[1])

    def reset_ever(self):
        out = np.zeros(self.n, np.uint8); lib().orc_swarm_get_reset_ever(self.h, out); return out

    def set_reset_ever(self, flags): lib().orc_swarm_set_reset_ever(self.h, np.ascontiguousarray(flags, np.uint8))

    def slack(self):
        """(slack share of the QP cost, number of rows with eps < 0) per agent of the last step."""
        c = np.zeros(self.n); r = np.zeros(self.n, np.int32); lib().orc_swarm_get_slack(self.h, c, r); return c, r

    def set_state(self, pos, vel=None, acc=None):
        pos = np.ascontiguousarray(pos, np.float32).reshape(self.n, 3)
        vel = np.zeros_like(pos) if vel is None else np.ascontiguousarray(vel, np.float32).reshape(self.n, 3)
        acc = np.zeros_like(pos) if acc is None else np.ascontiguousarray(acc, np.float32).reshape(self.n, 3)
        lib().orc_swarm_set_state(self.h, pos, vel, acc)

    def set_goals(self, goal): lib().orc_swarm_set_goals(self.h, np.ascontiguousarray(goal, np.float32).reshape(self.n, 3))

    def set_traj(self, traj, seq): lib().orc_swarm_set_traj(self.h, np.ascontiguousarray(traj, np.float32).reshape(self.n, 90), seq)

    def set_boxes(self, boxes, init_sfc):
        lib().orc_swarm_set_boxes(self.h, np.ascontiguousarray(boxes, np.float32).reshape(self.n, 30),
                                  np.ascontiguousarray(init_sfc, np.int32))

    def set_goal_mode(self, mode, grid_resolution=0.25, grid_margin=0.1, goal_threshold=0.1, goal_radius=2.0,
                      priority_dist_threshold=0.4):
        """0: static (set_goals = current goals); 1: prior_based (set_desired_goals; goals planned every step)."""
        lib().orc_swarm_set_goal_mode(self.h, mode, grid_resolution, grid_margin, goal_threshold, goal_radius, priority_dist_threshold)

    def set_desired_goals(self, goal):
        lib().orc_swarm_set_desired_goals(self.h, np.ascontiguousarray(goal, np.float32).reshape(self.n, 3))

    def goals(self):
        g = np.zeros((self.n, 3), np.float32); k = np.zeros(self.n, np.int32)
        lib().orc_swarm_get_goals(self.h, g, k); return g, k

    def safety_audit(self, record_time_step=0.1, time_step=0.2):
        """(ratio[n], closest[n]) of the current trajectories over the recorded sub-times of one step."""
        r = np.zeros(self.n); c = np.zeros(self.n, np.int32)
        lib().orc_swarm_safety_audit(self.h, record_time_step, time_step, r, c); return r, c

    def astar_expansions(self): return int(lib().orc_swarm_astar_expansions(self.h))

    def step(self, a0=0, a1=None, threads=1): lib().orc_swarm_step(self.h, a0, self.n if a1 is None else a1, threads)

    def step_agents(self, ids, threads=1):
        """One synchronous step in which only the listed agents re-plan (all from the same snapshot)."""
        ids = np.ascontiguousarray(ids, np.int32)
        lib().orc_swarm_step_list(self.h, ids, len(ids), threads)

    def advance(self): lib().orc_swarm_advance(self.h)

    @property
    def seq(self): return lib().orc_swarm_seq(self.h)

    def traj(self):
        out = np.zeros((self.n, 5, 6, 3), np.float32); lib().orc_swarm_get_traj(self.h, out.reshape(self.n, 90)); return out

    def pred(self):
        out = np.zeros((self.n, 5, 6, 3), np.float32); lib().orc_swarm_get_pred(self.h, out.reshape(self.n, 90)); return out

    def state(self):
        p = np.zeros((self.n, 3), np.float32); v = np.zeros_like(p); a = np.zeros_like(p)
        lib().orc_swarm_get_state(self.h, p, v, a); return p, v, a

    def boxes(self):
        out = np.zeros((self.n, 5, 6), np.float32); lib().orc_swarm_get_boxes(self.h, out.reshape(self.n, 30)); return out

    def qp(self):
        n = self.n
        cost = np.zeros(n); st = np.zeros(n, np.int32); it = np.zeros(n, np.int32); na = np.zeros(n, np.int32)
        fl = np.zeros(n, np.int32); mv = np.zeros(n); kk = np.zeros(n)
        lib().orc_swarm_get_qp(self.h, cost, st, it, na, fl, mv, kk)
        return dict(cost=cost, status=st, iters=it, n_active=na, flags=fl, maxviol=mv, kkt=kk)

    def counters(self):
        c = np.zeros(4, np.int64); lib().orc_swarm_get_counters(self.h, c)
        return dict(gjk_iters=int(c[0]), qp_iters=int(c[1]), edt_lookups=int(c[2]), qp_rows=int(c[3]))

    def reset_counters(self): lib().orc_swarm_reset_counters(self.h)

    def capture(self):
        n = self.n
        nr = np.zeros((n, n, 5, 3), np.float32); d = np.zeros((n, n, 5, 6)); g = np.zeros((n, n, 5), np.int32)
        lib().orc_swarm_get_capture(self.h, nr.reshape(-1, 3), d.reshape(-1), g.reshape(-1)); return nr, d, g

    def __del__(self):
        try:
            lib().orc_swarm_free(self.h)
        except Exception:
            pass

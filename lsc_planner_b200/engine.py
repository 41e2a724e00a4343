"""Thin Python binding of the C-ABI engine (used by tests/, bench.py and __graft_entry__).

The product's host side is C++ (lsc_planner_b200/host/, mirroring the reference's TrajPlanner / TrajOptimizer /
MultiSyncSimulator); this module only marshals numpy arrays into include/lscgpu.h calls.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi as A


@dataclass
class Param:
    """The hot-path subset of the reference's Param (include/param.hpp; defaults launch/simulation.launch)."""
    dt: float = 0.2
    control_input_weight: float = 0.01
    terminal_weight: float = 1.0
    world_resolution: float = 0.1
    reset_threshold: float = 0.15
    world_use_octomap: bool = False
    world_min: Sequence[float] = (-5.0, -5.0, 0.0)
    world_max: Sequence[float] = (5.0, 5.0, 2.5)
    M: int = 5
    n: int = 5
    phi: int = 3
    dim: int = 3
    goal_mode: int = 0              # 0 static (goal = current goal), 1 prior_based on the GPU (goal = desired goal)
    goal_threshold: float = 0.1
    goal_radius: float = 2.0
    priority_dist_threshold: float = 0.4
    grid_resolution: float = 0.25   # grid/resolution, grid/margin: the planning grid of goal_mode 1 with an octomap
    grid_margin: float = 0.1

    def to_c(self) -> A.Params:
        p = A.Params()
        p.dt, p.control_input_weight, p.terminal_weight = self.dt, self.control_input_weight, self.terminal_weight
        p.world_resolution, p.reset_threshold = self.world_resolution, self.reset_threshold
        p.world_use_octomap = int(self.world_use_octomap)
        for k in range(3):
            p.world_min[k] = float(np.float32(self.world_min[k])); p.world_max[k] = float(np.float32(self.world_max[k]))
        p.M, p.n, p.phi, p.dim = self.M, self.n, self.phi, self.dim
        p.goal_mode, p.goal_threshold, p.goal_radius = self.goal_mode, self.goal_threshold, self.goal_radius
        p.priority_dist_threshold = self.priority_dist_threshold
        p.grid_resolution, p.grid_margin = self.grid_resolution, self.grid_margin
        return p


@dataclass
class AgentType:
    """Constant part of the reference's Agent (include/sp_const.hpp:153-165); defaults = `crazyflie` quadrotor."""
    radius: float = 0.15
    downwash: float = 2.0
    nominal_velocity: float = 1.0
    max_vel: Sequence[float] = (1.0, 1.0, 1.0)
    max_acc: Sequence[float] = (2.0, 2.0, 2.0)


def measure_fma_peaks(device: int = 0) -> dict:
    """FP32 / FP64 FMA peaks of the device in TFLOP/s (lscgpu_measure_fma_peaks)."""
    a, b = C.c_double(0), C.c_double(0)
    A.check(A.lib().lscgpu_measure_fma_peaks(device, C.byref(a), C.byref(b)))
    return {"fp32_tflops": a.value, "fp64_tflops": b.value}


def measure_latencies(device: int = 0) -> dict:
    """Dependent-issue latencies in cycles (lscgpu_measure_latencies)."""
    v = (C.c_double * 7)()
    A.check(A.lib().lscgpu_measure_latencies(device, v))
    return dict(zip(["dfma", "ffma", "lds", "shfl_dadd", "rsqrt_f64", "div_f64", "redux"], [float(x) for x in v]))


class ReplanEngine:
    """One GPU's replanning engine for a swarm of `n_agents` (include/lscgpu.h)."""

    def __init__(self, n_agents: int, param: Optional[Param] = None, agents: Optional[Sequence[AgentType]] = None,
                 device: int = 0):
        self.lib = A.lib()
        self.n = int(n_agents)
        self.param = param or Param()
        agents = list(agents) if agents is not None else [AgentType()] * self.n
        if len(agents) != self.n:
            raise ValueError("one AgentType per agent")
        arr = (A.AgentConst * self.n)()
        for i, a in enumerate(agents):
            arr[i].radius, arr[i].downwash, arr[i].nominal_velocity = a.radius, a.downwash, a.nominal_velocity
            for k in range(3):
                arr[i].max_vel[k] = a.max_vel[k]; arr[i].max_acc[k] = a.max_acc[k]
        self.agents = agents
        h = A.ptr()
        cp = self.param.to_c()
        A.check(self.lib.lscgpu_create(C.byref(cp), self.n, arr, device, C.byref(h)))
        self.h = h
        self.a0, self.a1 = 0, self.n
        self._in = np.zeros(self.n, A.AGENT_IN)
        self._out = np.zeros(self.n, A.AGENT_OUT)

    # ---- lifetime ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.lscgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- map -----------------------------------------------------------------------------------------------
    def set_octomap_file(self, path: str):
        A.check(self.lib.lscgpu_set_octomap_file(self.h, path.encode()))

    def set_octomap_voxels(self, keys):
        keys = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        A.check(self.lib.lscgpu_set_octomap_voxels(self.h, A.p(keys), len(keys)))

    def distmap(self):
        size = np.zeros(3, np.int32); off = np.zeros(3, np.int32); nocc = np.zeros(1, np.int64)
        A.check(self.lib.lscgpu_get_distmap_info(self.h, A.p(size), A.p(off), A.p(nocc)))
        sq = np.zeros(tuple(size), np.uint8)
        A.check(self.lib.lscgpu_get_distmap_sqdist(self.h, A.p(sq)))
        return dict(size=size, off=off, n_occupied=int(nocc[0]), sqdist=sq)

    # ---- sharding ------------------------------------------------------------------------------------------
    def set_shard(self, a0: int, a1: int):
        A.check(self.lib.lscgpu_set_shard(self.h, a0, a1)); self.a0, self.a1 = a0, a1

    def nccl_unique_id(self) -> bytes:
        buf = np.zeros(128, np.uint8)
        A.check(self.lib.lscgpu_nccl_unique_id(A.p(buf)))
        return buf.tobytes()

    def nccl_init(self, unique_id: bytes, rank: int, n_ranks: int):
        buf = np.frombuffer(unique_id, np.uint8).copy()
        A.check(self.lib.lscgpu_nccl_init(self.h, A.p(buf), rank, n_ranks))
        # the job's LPT order is dealt out round-robin: this rank plans every n_ranks-th entry, starting at `rank`
        self.a0, self.a1 = 0, self.n
        self.rank, self.n_ranks = rank, n_ranks

    def p2p_export(self) -> bytes:
        """CUDA IPC handle of this rank's exchange buffer (direct exchange over NVLink peer memory)."""
        h = np.zeros(64, np.uint8)
        A.check(self.lib.lscgpu_p2p_export(self.h, A.p(h)))
        return h.tobytes()

    def p2p_attach(self, handles: Sequence[bytes]):
        """Map every rank's exchange buffer (handles in rank order) and switch the step to the direct exchange."""
        buf = np.frombuffer(b"".join(handles), np.uint8).copy()
        assert len(buf) == 64 * len(handles)
        A.check(self.lib.lscgpu_p2p_attach(self.h, A.p(buf)))

    @property
    def n_planned(self) -> int:
        """Agents this engine plans per step."""
        if getattr(self, "n_ranks", 1) > 1:
            return max(0, (self.n - self.rank + self.n_ranks - 1) // self.n_ranks)
        return self.a1 - self.a0

    # ---- stepping ------------------------------------------------------------------------------------------
    def replan(self, pos, vel, acc, goal, out: Optional[np.ndarray] = None, inp: Optional[np.ndarray] = None):
        """One synchronous replanning step with host inputs (H2D + kernels + D2H). Returns the AGENT_OUT array."""
        inp = self._in if inp is None else inp
        inp["position"] = pos; inp["velocity"] = vel; inp["acceleration"] = acc; inp["goal"] = goal
        return self.replan_raw(inp, out)

    def replan_raw(self, inp: np.ndarray, out: Optional[np.ndarray] = None):
        out = self._out if out is None else out
        A.check(self.lib.lscgpu_replan_batch(self.h, A.p(inp), A.p(out)))
        return out

    def replan_ptr(self, in_ptr: int, out_ptr: int):
        """Same, on raw host addresses (e.g. pinned torch tensors)."""
        A.check(self.lib.lscgpu_replan_batch(self.h, A.ptr(in_ptr), A.ptr(out_ptr)))

    def advance_inputs_ptr(self, out_ptr: int, in_ptr: int):
        """in[a].position / velocity / acceleration = out[a].next_* on raw host addresses (the simulator's state hand-over)."""
        A.check(self.lib.lscgpu_advance_inputs(A.ptr(out_ptr), A.ptr(in_ptr), self.n))

    def set_goals(self, goal):
        g = np.ascontiguousarray(goal, np.float32).reshape(self.n, 3)
        A.check(self.lib.lscgpu_set_goals(self.h, A.p(g)))

    def set_states(self, pos, vel=None, acc=None):
        pos = np.ascontiguousarray(pos, np.float32).reshape(self.n, 3)
        vel = np.zeros_like(pos) if vel is None else np.ascontiguousarray(vel, np.float32).reshape(self.n, 3)
        acc = np.zeros_like(pos) if acc is None else np.ascontiguousarray(acc, np.float32).reshape(self.n, 3)
        A.check(self.lib.lscgpu_set_states(self.h, A.p(pos), A.p(vel), A.p(acc)))

    def replan_resident(self, steps: int = 1, sync: bool = True):
        """Enqueue `steps` device-resident closed-loop steps (inputs = the engine's own advanced states)."""
        for _ in range(steps):
            A.check(self.lib.lscgpu_replan_resident(self.h))
        if sync:
            self.synchronize()

    def synchronize(self):
        A.check(self.lib.lscgpu_synchronize(self.h))

    def fetch(self, out: Optional[np.ndarray] = None):
        out = self._out if out is None else out
        A.check(self.lib.lscgpu_fetch(self.h, A.p(out)))
        return out

    def safety_audit(self, record_time_step: float = 0.1, time_step: float = 0.2):
        """Minimum-distance audit of the step just planned (MultiSyncSimulator::savePlanningResult): per agent the
        smallest safety ratio over the recorded sub-times and the agent it is attained with."""
        r = np.zeros(self.n, np.float64); c = np.zeros(self.n, np.int32)
        A.check(self.lib.lscgpu_safety_audit(self.h, record_time_step, time_step, A.p(r), A.p(c)))
        return r, c

    def reset(self):
        A.check(self.lib.lscgpu_reset(self.h))

    def set_prev_traj(self, traj, planner_seq: int):
        t = np.ascontiguousarray(traj, np.float32).reshape(self.n, 90)
        A.check(self.lib.lscgpu_set_prev_traj(self.h, A.p(t), planner_seq))

    def set_sfc(self, boxes, init_flags):
        b = np.ascontiguousarray(boxes, np.float32).reshape(self.n, 30)
        f = np.ascontiguousarray(init_flags, np.int32).reshape(self.n)
        A.check(self.lib.lscgpu_set_sfc(self.h, A.p(b), A.p(f)))

    def get_sfc(self):
        b = np.zeros((self.n, 5, 6), np.float32); f = np.zeros(self.n, np.int32)
        A.check(self.lib.lscgpu_get_sfc(self.h, A.p(b), A.p(f)))
        return b, f

    @property
    def planner_seq(self) -> int:
        return self.lib.lscgpu_get_planner_seq(self.h)

    def set_slack_collision_weight(self, w: float):
        """opt/slack_collision_weight of the slack variables a state reset brings into the QPs."""
        A.check(self.lib.lscgpu_set_slack_collision_weight(self.h, float(w)))

    def reset_state(self) -> np.ndarray:
        """Sticky per-agent "was ever reset" bytes (who is in everybody's obs_slack_indices)."""
        out = np.zeros(self.n, np.uint8)
        A.check(self.lib.lscgpu_get_reset_state(self.h, A.p(out)))
        return out

    def set_reset_state(self, flags):
        f = np.ascontiguousarray(flags, np.uint8).reshape(self.n)
        A.check(self.lib.lscgpu_set_reset_state(self.h, A.p(f)))

    def get_lsc(self, agent: int):
        """(normals [N-1][5][3] float32, d [N-1][5][6] float64) of the last step, CollisionConstraints layout."""
        nr = np.zeros((max(self.n - 1, 0), 5, 3), np.float32); d = np.zeros((max(self.n - 1, 0), 5, 6), np.float64)
        A.check(self.lib.lscgpu_get_lsc(self.h, agent, A.p(nr), A.p(d)))
        return nr, d

    def set_capture_rows(self, on: bool = True):
        """Mirror the rows k_agent_plan builds into global memory, so that get_lsc_rows returns the production rows."""
        A.check(self.lib.lscgpu_set_capture_rows(self.h, int(on)))

    def get_lsc_rows(self, agent: int):
        """(normals, d, kept [N-1][5] bool): like get_lsc, with the kept (neighbour, segment) pairs decoded from the row
        store the planning kernel wrote in the last step; the culled pairs are recomputed."""
        nr = np.zeros((max(self.n - 1, 0), 5, 3), np.float32); d = np.zeros((max(self.n - 1, 0), 5, 6), np.float64)
        kept = np.zeros((max(self.n - 1, 0), 5), np.uint8)
        A.check(self.lib.lscgpu_get_lsc_ex(self.h, agent, A.p(nr), A.p(d), A.p(kept)))
        return nr, d, kept.astype(bool)

    def dump_qp_lp(self, agent: int, path: str):
        """The QP of `agent` in the last step as a CPLEX LP file (the reference's log/QPmodel.lp)."""
        A.check(self.lib.lscgpu_dump_qp_lp(self.h, agent, path.encode()))

    def set_lp_dump_dir(self, directory: Optional[str]):
        """Write the LP of every failed QP of every replan() into `directory` (None: off)."""
        A.check(self.lib.lscgpu_set_lp_dump_dir(self.h, directory.encode() if directory else None))

    def initial_traj(self):
        out = np.zeros((self.n, 5, 6, 3), np.float32)
        A.check(self.lib.lscgpu_get_initial_traj(self.h, A.p(out)))
        return out

    # ---- operator-level entries ----------------------------------------------------------------------------
    def qp_solve_batch(self, agent_index, state, goal, obs_offset, lsc_normal, lsc_point, lsc_d, sfc=None, obs_slack=None):
        """TrajOptimizer::solve for a batch. Returns dict(x [B][90], cost, status, iterations[, eps [obstacles][5]]).
        obs_slack: per obstacle, non-zero = member of obs_slack_indices (its rows get slack variables)."""
        ai = np.ascontiguousarray(agent_index, np.int32); nb = len(ai)
        st = np.ascontiguousarray(state, np.float64).reshape(nb, 9)
        gl = np.ascontiguousarray(goal, np.float64).reshape(nb, 3)
        off = np.ascontiguousarray(obs_offset, np.int32)
        assert len(off) == nb + 1
        tot = int(off[-1])
        nr = np.ascontiguousarray(lsc_normal, np.float32).reshape(tot, 5, 3)
        pt = np.ascontiguousarray(lsc_point, np.float32).reshape(tot, 5, 6, 3)
        dd = np.ascontiguousarray(lsc_d, np.float64).reshape(tot, 5, 6)
        bx = None if sfc is None else np.ascontiguousarray(sfc, np.float32).reshape(nb, 30)
        x = np.zeros((nb, 90)); cost = np.zeros(nb); status = np.zeros(nb, np.int32); iters = np.zeros(nb, np.int32)
        if obs_slack is not None:
            sl = np.ascontiguousarray(obs_slack, np.uint8).reshape(tot)
            eps = np.zeros((max(tot, 1), 5))
            A.check(self.lib.lscgpu_qp_solve_batch_slack(self.h, nb, A.p(ai), A.p(st), A.p(gl), None if bx is None else A.p(bx),
                                                         A.p(off), A.p(nr), A.p(pt), A.p(dd), A.p(sl), A.p(x), A.p(cost),
                                                         A.p(status), A.p(iters), A.p(eps)))
            return dict(x=x, cost=cost, status=status, iterations=iters, eps=eps[:tot])
        A.check(self.lib.lscgpu_qp_solve_batch(self.h, nb, A.p(ai), A.p(st), A.p(gl), None if bx is None else A.p(bx),
                                               A.p(off), A.p(nr), A.p(pt), A.p(dd), A.p(x), A.p(cost), A.p(status),
                                               A.p(iters)))
        return dict(x=x, cost=cost, status=status, iterations=iters)

    def gjk_batch(self, hulls):
        h = np.ascontiguousarray(hulls, np.float64).reshape(-1, 6, 3); n = len(h)
        v = np.zeros((n, 3)); it = np.zeros(n, np.int32)
        A.check(self.lib.lscgpu_gjk_batch(self.h, n, A.p(h), A.p(v), A.p(it)))
        return v, it

    def sfc_expand_batch(self, point, goal, radius=None):
        pt = np.ascontiguousarray(point, np.float32).reshape(-1, 3); n = len(pt)
        gl = np.ascontiguousarray(goal, np.float32).reshape(n, 3)
        r = np.full(n, self.agents[0].radius) if radius is None else np.ascontiguousarray(radius, np.float64).reshape(n)
        box = np.zeros((n, 6), np.float32); ok = np.zeros(n, np.int32)
        A.check(self.lib.lscgpu_sfc_expand_batch(self.h, n, A.p(pt), A.p(gl), A.p(r), A.p(box), A.p(ok)))
        return box, ok

    # ---- instrumentation -----------------------------------------------------------------------------------
    def set_profiling(self, on: bool):
        A.check(self.lib.lscgpu_set_profiling(self.h, int(on)))

    def step_stats(self) -> dict:
        s = A.StepStats()
        A.check(self.lib.lscgpu_get_step_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in A.StepStats._fields_}

    @property
    def stream(self) -> int:
        return int(self.lib.lscgpu_stream(self.h) or 0)

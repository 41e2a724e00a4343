"""Deterministic swarm scenarios of BASELINE.json / SURVEY.md §8(d) and the mission-JSON reader.

Mission schema: the reference's missions/*.json (missions/readme.txt:1-30, src/mission.cpp:20-319):
  quadrotors{type: {max_vel, max_acc, radius, nominal_velocity, downwash}}, world[0].dimension = [xmin,ymin,zmin,xmax,ymax,zmax],
  agents[{type, cid?, start, goal}], obstacles[].
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass
from typing import List

import numpy as np

from .engine import AgentType


@dataclass
class Scenario:
    name: str
    start: np.ndarray          # [N][3] float32
    goal: np.ndarray           # [N][3] float32
    world_min: tuple
    world_max: tuple
    agents: List[AgentType]
    use_octomap: bool = False

    @property
    def n(self) -> int:
        return len(self.start)


def load_mission(path: str) -> Scenario:
    ms = json.load(open(path))
    quad = ms["quadrotors"]
    dim = ms["world"][0]["dimension"]
    starts, goals, agents = [], [], []
    for ag in ms["agents"]:
        q = quad[ag["type"]]
        agents.append(AgentType(radius=q["radius"], downwash=q["downwash"], nominal_velocity=q["nominal_velocity"],
                                max_vel=tuple(q["max_vel"]), max_acc=tuple(q["max_acc"])))
        starts.append(ag["start"]); goals.append(ag["goal"])
    return Scenario(path.rsplit("/", 1)[-1], np.asarray(starts, np.float32), np.asarray(goals, np.float32),
                    tuple(dim[:3]), tuple(dim[3:]), agents)


def circle_swap(n: int, z: float = 1.0, spacing: float = 0.8, r0: float = 4.0, ring_gap: float = 1.0,
                forest: bool = False) -> Scenario:
    """concentric rings R_k = 4 + k, floor(2 pi R_k / 0.8) agents per ring, goal = antipode."""
    starts = []
    k = 0
    while len(starts) < n:
        R = r0 + ring_gap * k
        cnt = min(int(math.floor(2 * math.pi * R / spacing)), n - len(starts))
        for t in range(cnt):
            th = 2 * math.pi * t / cnt
            starts.append((R * math.cos(th), R * math.sin(th), z))
        k += 1
    start = np.asarray(starts, np.float64)
    goal = start * np.array([-1.0, -1.0, 1.0])
    rmax = r0 + ring_gap * (k - 1)
    ext = float(math.ceil(rmax + 2.0))
    return Scenario(f"circle_swap_{n}", start.astype(np.float32), goal.astype(np.float32), (-ext, -ext, 0.0),
                    (ext, ext, 2.5), [AgentType()] * n, use_octomap=forest)


def random_forest(n: int, sqdist: np.ndarray, off, res: float = 0.1, seed: int = 0, world_min=(-5, -5, 0),
                  world_max=(5, 5, 2.5), clearance_m: float = 0.4, pair_dist: float = 0.6, downwash: float = 2.0) -> Scenario:
    """rejection-sampled starts and goals in the world shrunk by 0.5 m with EDT >= 0.4 m
    and pairwise downwash-scaled distance >= 0.6 m inside each set. `sqdist` is the engine's squared cell distance grid.
    0.6 m spacing saturates the 10 x 10 x 2.5 m world near 200 agents (random sequential packing), so for larger swarms
    the spacing is reduced stepwise (0.5, 0.4, 0.35 m; the collision distance is 0.3 m) until all agents fit."""
    vol = float(np.prod((np.asarray(world_max, float) - np.asarray(world_min, float) - 1.0) * [1, 1, 1.0 / downwash]))
    for pd in [d for d in (pair_dist, 0.5, 0.4, 0.35) if d <= pair_dist]:
        if n > 0.3 * vol / (math.pi / 6 * pd ** 3) and pd > 0.35:
            continue                       # beyond the random-sequential-packing capacity at this spacing
        try:
            return _random_forest(n, sqdist, off, res, seed, world_min, world_max, clearance_m, pd, downwash)
        except RuntimeError:
            continue
    raise RuntimeError("cannot place agents")


def _random_forest(n, sqdist, off, res, seed, world_min, world_max, clearance_m, pair_dist, downwash) -> Scenario:
    rng = np.random.default_rng(seed)
    lo = np.asarray(world_min, float) + 0.5; hi = np.asarray(world_max, float) - 0.5
    need_sq = (clearance_m / res) ** 2

    def sample_set():
        pts = []
        tries = 0
        while len(pts) < n:
            tries += 1
            if tries > 400 * n + 20000:
                raise RuntimeError("cannot place agents")
            p = rng.uniform(lo, hi)
            c = np.floor(p / res).astype(int) - np.asarray(off)
            if (c < 0).any() or (c >= np.asarray(sqdist.shape)).any() or sqdist[tuple(c)] < need_sq:
                continue
            if pts:
                d = (np.asarray(pts) - p) * np.array([1, 1, 1.0 / downwash])
                if (np.einsum("ij,ij->i", d, d) < pair_dist ** 2).any():
                    continue
            pts.append(p)
        return np.asarray(pts)

    start = sample_set(); goal = sample_set()
    return Scenario(f"random_forest_{n}", start.astype(np.float32), goal.astype(np.float32), tuple(world_min),
                    tuple(world_max), [AgentType()] * n, use_octomap=True)

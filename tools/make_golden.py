#!/usr/bin/env python
"""Regenerates tests/golden/* from the read-only reference tree (run in the build container only;
/root/reference does not exist on the GPU box, so the tests read the committed fixtures).

Fixtures written (all DATA, no reference source code):
  golden/qpmodel_lp.npz          dense matrices parsed from <ref>/log/QPmodel.lp (the reference's only
                                 recorded QP: an INFEASIBLE failure dump) plus the
                                 planner inputs recovered from it (state, goal, ts, LSC rows, limits)
  golden/gjk_ref_vectors.npz     6-point hulls and the outputs of the REFERENCE's own openGJK
                                 (oracle/_ref/libref_gjk.so, compiled from <ref>/src/openGJK)
  golden/missions/*.json         the two mission files BASELINE.json names (re-dumped JSON)
  golden/worlds/simple_forest.bt the octomap BASELINE.json names (binary data, byte copy)
  golden/simple_forest_voxels.npz occupied finest voxels of that map as decoded by the oracle's reader
  golden/astar_ref_vectors.npz   occupancy grids, start / goal cells and the cell paths returned by the REFERENCE's own
                                 A* (oracle/_ref/libref_astar.so, compiled from <ref>/src/Astar-3D), including grids
                                 built to provoke the (F, g) ties that the reference resolves by hash-map iteration order
"""
import ctypes as C
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O   # noqa: E402
import qp_pyref as R     # noqa: E402

REF = os.environ.get("LSC_REFERENCE", "/root/reference")
G = os.path.join(ROOT, "tests", "golden")


def astar_cases(seed=2024, count=160):
    """Deterministic test grids: random clutter, forests of vertical columns, and mirror-symmetric layouts (start and
    goal in the same grid column with an obstacle between them) where equal-length detours tie."""
    rng = np.random.default_rng(seed)
    cases = []
    for t in range(count):
        dim = (int(rng.integers(3, 42)), int(rng.integers(3, 42)), int(rng.integers(1, 12)))
        dens = float(rng.choice([0.0, 0.05, 0.15, 0.3]))
        grid = (rng.random(dim) < dens).astype(np.uint8)
        if t % 3 == 0:
            grid[:] = 0
            for _ in range(int(dens * 40)):
                a, b = int(rng.integers(0, dim[0])), int(rng.integers(0, dim[1])); grid[a:a + 2, b:b + 2, :] = 1
        s = [int(rng.integers(0, d)) for d in dim]; g = [int(rng.integers(0, d)) for d in dim]
        if t % 4 == 0:
            g[1] = s[1]
        if t % 5 == 0:
            g[2] = s[2]
        if t % 8 == 0 and dim[0] > 8 and dim[1] > 8:          # symmetric wall across the straight line
            grid[:] = 0
            s = [1, dim[1] // 2, dim[2] // 2]; g = [dim[0] - 2, dim[1] // 2, dim[2] // 2]
            w = int(rng.integers(1, max(2, dim[1] // 2 - 1)))
            grid[dim[0] // 2, dim[1] // 2 - w:dim[1] // 2 + w + 1, :] = 1
        grid[tuple(s)] = 0
        cases.append((np.ascontiguousarray(grid), np.array(s, np.int32), np.array(g, np.int32)))
    return cases


def astar_fixture():
    so = os.path.join(ROOT, "oracle", "_ref", "libref_astar.so")
    R_ = C.CDLL(so)
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS"); u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    R_.ref_astar_plan.argtypes = [i32p, u8p, i32p, i32p, i32p, C.c_int]; R_.ref_astar_plan.restype = C.c_int
    out = {}
    n_found = 0
    for k, (grid, s, g) in enumerate(astar_cases()):
        path = np.zeros((4096, 3), np.int32)
        n = R_.ref_astar_plan(np.asarray(grid.shape, np.int32), grid, s, g, path, len(path))
        out[f"grid{k}"] = np.packbits(grid); out[f"dim{k}"] = np.asarray(grid.shape, np.int32)
        out[f"start{k}"] = s; out[f"goal{k}"] = g; out[f"path{k}"] = path[:n].astype(np.int16)
        n_found += n > 0
    out["count"] = np.array(len(astar_cases()))
    np.savez_compressed(os.path.join(G, "astar_ref_vectors.npz"), **out)
    print("astar_ref_vectors.npz:", len(astar_cases()), "cases,", n_found, "with a path")


def lp_fixture():
    lp = R.parse_lp(open(os.path.join(REF, "log", "QPmodel.lp")).read())
    state = np.zeros((3, 3)); state[0] = lp["beq"][[0, 15, 30]]
    goal = -lp["q"][[29, 59, 89]] / 2
    ts = int(round(sum(lp["P"][k * 30 + m * 6 + 5, k * 30 + m * 6 + 5] > 22500.5 for k in (0,) for m in range(5))))
    n_lsc = len(lp["bin"]) - 252
    rows_m, rows_a, rows_rhs = [], [], []
    r = 0
    for oi in range(n_lsc // 27):
        for m in range(5):
            a = None; rhs = np.zeros(6)
            for i in range(6):
                if m == 0 and i < 3:
                    continue
                row = lp["Ain"][r]
                aa = np.array([row[k * 30 + m * 6 + i] for k in range(3)])
                assert np.count_nonzero(row) == np.count_nonzero(aa)
                a = aa if a is None else a
                assert np.array_equal(a, aa)
                rhs[i] = lp["bin"][r]; r += 1
            rows_m.append(m); rows_a.append(a); rows_rhs.append(rhs)
    dyn = lp["bin"][n_lsc:]
    vmax = np.array([-dyn[0], -dyn[84], -dyn[168]])              # first velocity row of each axis
    amax = np.array([-dyn[46 * 0 + 6], -dyn[84 + 6], -dyn[168 + 6]])  # m=0: 6 velocity rows, then acceleration
    np.savez_compressed(os.path.join(G, "qpmodel_lp.npz"), P=lp["P"], q=lp["q"], c0=lp["c0"], Aeq=lp["Aeq"],
                        beq=lp["beq"], Ain=lp["Ain"], bin=lp["bin"], lb=lp["lb"], ub=lp["ub"], state=state, goal=goal,
                        ts=ts, rows_m=np.array(rows_m, np.int32), rows_a=np.array(rows_a), rows_rhs=np.array(rows_rhs),
                        vmax=vmax, amax=amax)
    print("qpmodel_lp.npz: ts", ts, "lsc rows", r, "vmax", vmax, "amax", amax)


def gjk_fixture():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_gjk.so"))
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    lib.ref_gjk_point_hull.restype = C.c_double
    lib.ref_gjk_point_hull.argtypes = [f64p, C.c_int, f64p, f64p]
    rng = np.random.default_rng(20260101)
    hulls, vs, ds = [], [], []
    for t in range(4000):
        kind = t % 8
        P = rng.normal(size=(6, 3))
        if kind == 0: P += rng.normal(size=3) * 3
        elif kind == 1: P = P[0] + np.outer(np.linspace(0, 1, 6), P[1])       # collinear
        elif kind == 2: P = np.repeat(P[:1], 6, 0)                              # one point
        elif kind == 3: P += np.array([2.0, 0, 0])
        elif kind == 4: P[:, 2] = 0.5                                           # coplanar
        elif kind == 5: P *= 0.05                                               # origin inside / tiny
        elif kind == 6: P = P * 0.2 + rng.normal(size=3) * 6                    # far, small hull
        elif kind == 7: P[3:] = P[:3]                                           # duplicated vertices
        P = np.ascontiguousarray(P.astype(np.float32).astype(np.float64))      # float32-representable
        v = np.zeros(3)
        d = lib.ref_gjk_point_hull(P, 6, np.zeros(3), v)
        hulls.append(P); vs.append(v); ds.append(d)
    np.savez_compressed(os.path.join(G, "gjk_ref_vectors.npz"), hulls=np.array(hulls), v=np.array(vs), dist=np.array(ds))
    print("gjk_ref_vectors.npz:", len(hulls), "hulls")


def data_fixtures():
    for name in ("multi_simple3.json", "multi_circle20.json"):
        ms = json.load(open(os.path.join(REF, "missions", name)))
        with open(os.path.join(G, "missions", name), "w") as f:
            json.dump(ms, f, indent=1)
    shutil.copyfile(os.path.join(REF, "world", "simple_forest.bt"), os.path.join(G, "worlds", "simple_forest.bt"))
    m = O.Map.from_bt(os.path.join(G, "worlds", "simple_forest.bt"), [-5, -5, 0], [5, 5, 2.5])
    np.savez_compressed(os.path.join(G, "simple_forest_voxels.npz"), keys=m.occupied(), n_nodes=m.n_nodes)
    print("simple_forest: nodes", m.n_nodes, "voxels", m.n_occ)


if __name__ == "__main__":
    os.makedirs(os.path.join(G, "missions"), exist_ok=True)
    os.makedirs(os.path.join(G, "worlds"), exist_ok=True)
    if "--astar-only" in sys.argv:
        astar_fixture(); sys.exit(0)
    lp_fixture(); gjk_fixture(); data_fixtures(); astar_fixture()

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
    lp_fixture(); gjk_fixture(); data_fixtures()

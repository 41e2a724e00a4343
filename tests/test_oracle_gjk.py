"""Pins the oracle's GJK / LSC restatement (oracle/geom.hpp).

Golden vectors: tests/golden/gjk_ref_vectors.npz = outputs of the REFERENCE's own openGJK
(src/openGJK/openGJK.cpp compiled into oracle/_ref by oracle/Makefile; generator tools/make_golden.py).
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O


def test_gjk_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "gjk_ref_vectors.npz"))
    worst = 0.0
    for P, v_ref, d_ref in zip(g["hulls"], g["v"], g["dist"]):
        v, it = O.gjk(P)
        assert 1 <= it <= 25
        worst = max(worst, float(np.abs(v - v_ref).max()))
        assert abs(np.linalg.norm(v) - d_ref) <= 1e-9
    assert worst <= 1e-9, worst     # tolerance: 1e-9 m on the closest point (reference eps_rel = 1e-10)


def test_gjk_matches_reference_live():
    so = os.path.join(O.ORACLE_DIR, "_ref", "libref_gjk.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    lib = C.CDLL(so)
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    lib.ref_gjk_point_hull.restype = C.c_double
    lib.ref_gjk_point_hull.argtypes = [f64p, C.c_int, f64p, f64p]
    rng = np.random.default_rng(7)
    for t in range(3000):
        P = rng.normal(size=(6, 3)) * rng.choice([0.1, 1.0, 3.0]) + rng.normal(size=3) * rng.choice([0.0, 1.0, 5.0])
        P = np.ascontiguousarray(P.astype(np.float32).astype(np.float64))
        v, _ = O.gjk(P)
        vr = np.zeros(3)
        lib.ref_gjk_point_hull(P, 6, np.zeros(3), vr)
        assert np.abs(v - vr).max() <= 1e-9


def test_gjk_properties():
    rng = np.random.default_rng(3)
    for t in range(500):
        P = rng.normal(size=(6, 3)) + np.array([3.0, 0, 0])
        v, _ = O.gjk(P)
        # v is the min-norm point of the hull: every vertex satisfies p.v >= |v|^2 (separating plane)
        assert (P @ v >= v @ v * (1 - 1e-9) - 1e-12).all()
    v, _ = O.gjk(np.array([[1, 1, 1], [-1, 1, -1], [1, -1, -1], [-1, -1, 1], [0, 0, 2], [0, 2, 0]], float))
    assert np.abs(v).max() == 0.0                      # origin inside the hull -> zero vector


def test_lsc_pair_semantics():
    """generateLSC (src/traj_planner.cpp:1310-1407): unit normal in downwash-scaled space, margins
    d_i = 0.5 (r_i + r_j + (c_i - o_i).n), z un-scaled afterwards."""
    rng = np.random.default_rng(5)
    own = rng.normal(size=(30, 3)).astype(np.float32)
    obs = (rng.normal(size=(30, 3)) + [4, 0, 0]).astype(np.float32)
    n, d, it = O.lsc_pair(own, obs)
    for m in range(5):
        nt = n[m].astype(np.float64).copy(); nt[2] *= 2.0          # back to scaled space
        assert abs(np.linalg.norm(nt) - 1) < 1e-6
        rel = (own[m * 6:(m + 1) * 6] - obs[m * 6:(m + 1) * 6]).astype(np.float64); rel[:, 2] /= 2.0
        np.testing.assert_allclose(d[m], 0.5 * (0.3 + rel @ nt), atol=2e-6)
        # the LSC evaluated at the own control points is feasible iff hull distance >= r_i + r_j
        assert ((rel @ nt) >= np.linalg.norm(O.gjk(rel.astype(np.float32).astype(np.float64))[0]) - 1e-5).all()
    # coincident hulls: zero normal (octomath normalize() leaves the zero vector), d = 0.5 (r_i + r_j)
    n, d, it = O.lsc_pair(own, own)
    assert np.abs(n).max() == 0 and np.allclose(d, 0.15)

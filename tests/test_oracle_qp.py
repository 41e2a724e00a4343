"""Pins the oracle's QP restatement (oracle/qp.hpp) against
 (a) the reference's only recorded QP, log/QPmodel.lp (golden/qpmodel_lp.npz; coefficient-exact
     assembly incl. row order, and the INFEASIBLE verdict), and
 (b) the independent null-space + NNLS solve of tests/qp_pyref.py.
"""
import os

import numpy as np

import oracle_lib as O
import qp_pyref as R


def _lp(golden_dir):
    return np.load(os.path.join(golden_dir, "qpmodel_lp.npz"))


def _rows(g, skip_obstacle=None):
    rows = []
    for r in range(len(g["rows_m"])):
        if skip_obstacle is not None and r // 5 == skip_obstacle:
            continue
        rows.append((int(g["rows_m"][r]), g["rows_a"][r], g["rows_rhs"][r]))
    return rows


def test_assembly_matches_reference_lp_dump(golden_dir):
    g = _lp(golden_dir)
    T = O.Tables(0.2, 0.01, 1.0)
    D = T.dense(g["state"], g["goal"], int(g["ts"]), g["lb"], g["ub"], g["vmax"], g["amax"], _rows(g))
    # LP text carries 15 significant digits -> 1e-12 relative on coefficients up to 2.25e4
    assert np.abs(D["P"] - g["P"]).max() <= 1e-10
    assert np.abs(D["q"] - g["q"]).max() <= 1e-14 and abs(D["c0"] - float(g["c0"])) <= 1e-13
    assert D["Aeq"].shape == g["Aeq"].shape and np.abs(D["Aeq"] - g["Aeq"]).max() <= 1e-12
    assert np.abs(D["beq"] - g["beq"]).max() == 0
    assert D["Ain"].shape == g["Ain"].shape            # 243 LSC + 252 dynamic rows, reference row order
    assert np.abs(D["Ain"] - g["Ain"]).max() <= 1e-12 and np.abs(D["bin"] - g["bin"]).max() == 0
    assert np.array_equal(np.isfinite(D["lb"]), np.isfinite(g["lb"]))
    fin = np.isfinite(g["lb"])
    assert np.array_equal(D["lb"][fin], g["lb"][fin]) and np.array_equal(D["ub"][fin], g["ub"][fin])


def test_reference_lp_dump_is_infeasible(golden_dir):
    g = _lp(golden_dir)
    T = O.Tables()
    res = T.solve(g["state"], g["goal"], int(g["ts"]), g["lb"], g["ub"], g["vmax"], g["amax"], _rows(g))
    assert res["status"] == 1                           # QP_INFEASIBLE, as CPLEX reported for this dump
    lp = {k: g[k] for k in ("P", "q", "c0", "Aeq", "beq", "Ain", "bin", "lb", "ub")}
    assert R.solve_ldp(lp)[2] == "infeasible"


def test_lp_dump_without_conflicting_neighbour(golden_dir):
    g = _lp(golden_dir)
    T = O.Tables()
    rows = _rows(g, skip_obstacle=3)
    res = T.solve(g["state"], g["goal"], int(g["ts"]), g["lb"], g["ub"], g["vmax"], g["amax"], rows)
    D = T.dense(g["state"], g["goal"], int(g["ts"]), g["lb"], g["ub"], g["vmax"], g["amax"], rows)
    x, obj, st = R.solve_ldp(D)
    assert st == "ok" and res["status"] == 0
    assert abs(res["cost"] - obj) <= 1e-9 * max(1, abs(obj)) and np.abs(res["x"] - x).max() <= 1e-8


def test_tables():
    T = O.Tables(0.2, 0.01, 1.0)
    np.testing.assert_allclose(T.Qb, R.q_base(0.2), rtol=0, atol=1e-6)
    np.testing.assert_array_equal(T.A17, R.aeq_axis(0.2))
    assert np.abs(T.A17 @ T.Z).max() < 1e-9
    e = np.zeros((17, 3)); e[:3] = np.eye(3)
    assert np.abs(T.A17 @ T.Xp - e).max() < 1e-9
    for ts in range(1, 6):
        P = np.kron(np.eye(5), 0.01 * T.Qb)
        for m in range(5 - ts, 5):
            P[6 * m + 5, 6 * m + 5] += 1.0
        G = T.G[ts - 1]
        np.testing.assert_allclose(G.T @ P @ G, np.eye(13), atol=1e-9)      # whitened basis
        assert np.abs(T.A17 @ G).max() < 1e-8
        # x0 = Xs s + xg g minimises the equality-constrained cost: gradient orthogonal to null space
        s = np.array([0.3, -0.2, 0.1]); gl = 1.7
        x0 = T.Xs[ts - 1] @ s + T.xg[ts - 1] * gl
        eT = np.zeros(30); eT[[6 * m + 5 for m in range(5 - ts, 5)]] = 1
        grad = 2 * P @ x0 - 2 * gl * eT
        assert np.abs(T.Z.T @ grad).max() < 1e-7
        assert np.abs(T.A17 @ x0 - e @ s).max() < 1e-9


def test_solver_matches_independent_nnls_on_swarm_steps():
    """Closed-loop 12-agent circle swap; every 4th step 3 agents are re-solved by the independent path."""
    N = 12
    ang = 2 * np.pi * np.arange(N) / N
    start = np.stack([3 * np.cos(ang), 3 * np.sin(ang), np.ones(N)], 1).astype(np.float32)
    goal = (-start * [1, 1, -1]).astype(np.float32)
    wmin, wmax = [-5, -5, 0], [5, 5, 2.5]
    sw = O.Swarm(N, wmin, wmax); sw.set_state(start); sw.set_goals(goal); sw.set_capture(True)
    T = O.Tables()
    lb = np.full(90, -np.inf); ub = np.full(90, np.inf)
    for k in range(3):
        for m in range(5):
            for i in range(6):
                if not (m == 0 and i < 3):
                    lb[k * 30 + m * 6 + i] = wmin[k]; ub[k * 30 + m * 6 + i] = wmax[k]
    for step in range(24):
        pos, vel, acc = sw.state()
        sw.step()
        q = sw.qp()
        assert (q["status"] == 0).all() and q["maxviol"].max() <= 1e-6 + 1e-12 and q["kkt"].max() <= 1e-9
        if step % 4 == 1:
            pred = sw.pred(); nr, d, _ = sw.capture(); traj = sw.traj()
            for a in (0, 5, 7):
                rows = []
                for j in range(N):
                    if j == a:
                        continue
                    for m in range(5):
                        aa = nr[a, j, m].astype(np.float64)
                        rows.append((m, aa, d[a, j, m] + pred[j, m].astype(np.float64) @ aa))
                st = np.stack([pos[a], vel[a], acc[a]]).astype(np.float64)
                ts = O.terminal_segments(pos[a], goal[a])
                D = T.dense(st, goal[a].astype(np.float64), ts, lb, ub, [1, 1, 1], [2, 2, 2], rows)
                x, obj, status = R.solve_ldp(D)
                assert status == "ok"
                xo = traj[a].transpose(2, 0, 1).reshape(90)
                # the independent path solves the exact problem; the oracle ignores rows violated by <= 1e-6 (CPLEX's
                # feasibility tolerance), so the two may differ by the effect of such rows
                assert np.abs(xo - x).max() <= 2e-5
                assert abs(q["cost"][a] - obj) <= 1e-5 * max(1.0, abs(obj)) and q["cost"][a] <= obj + 1e-9
                v_eq, v_in = R.kkt_violation(D, T.solve(st, goal[a].astype(np.float64), ts, lb, ub, [1, 1, 1], [2, 2, 2], rows)["x"])
                assert v_eq <= 1e-7 and v_in <= 1e-6
        sw.advance()

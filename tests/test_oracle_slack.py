"""Pins the oracle's disturbance branch (SURVEY.md §8 f4 + a16):
 * qp_solve_slack (oracle/qp.hpp; src/traj_optimizer.cpp:317-326,383-390,455-457) against the independent
   null-space + NNLS solve of tests/qp_pyref.py on the extended problem (x, eps), and against qp_solve when no row
   carries a slack variable (must be identical);
 * the reset handling of the swarm stepper (oracle/swarm.hpp; src/traj_planner.cpp:866-878,1047-1061,
   src/multi_sync_simulator.cpp:229-246): prediction / initial trajectory collapsed to the observed position, sticky
   obs_slack_indices, flag_initialize_sfc re-armed.
CPLEX is not available, so these are the pins the slack branch has.
"""
import numpy as np

import oracle_lib as O
import qp_pyref as R

WMIN, WMAX = [-5, -5, 0], [5, 5, 2.5]


def _bounds():
    lb = np.full(90, -np.inf); ub = np.full(90, np.inf)
    for k in range(3):
        for m in range(5):
            for i in range(6):
                if not (m == 0 and i < 3):
                    lb[k * 30 + m * 6 + i] = WMIN[k]; ub[k * 30 + m * 6 + i] = WMAX[k]
    return lb, ub


def _circle(n, radius=3.0):
    ang = 2 * np.pi * np.arange(n) / n
    start = np.stack([radius * np.cos(ang), radius * np.sin(ang), np.ones(n)], 1).astype(np.float32)
    return start, (-start * [1, 1, -1]).astype(np.float32)


def disturb(sw, agent, offset):
    """What MultiSyncSimulator::update does when the observed pose is off by more than reset_threshold
    (src/multi_sync_simulator.cpp:229-246): the state becomes the observed position at rest."""
    pos, vel, acc = sw.state()
    pos[agent] += np.asarray(offset, np.float32); vel[agent] = 0; acc[agent] = 0
    sw.set_state(pos, vel, acc)


def _rows_of(sw, a, pred, nr, d):
    rows = []
    for j in range(sw.n):
        if j == a:
            continue
        for m in range(5):
            aa = nr[a, j, m].astype(np.float64)
            rows.append((m, aa, d[a, j, m] + pred[j, m].astype(np.float64) @ aa))
    return rows


def test_slack_solver_without_slack_rows_is_the_plain_solver():
    N = 10
    start, goal = _circle(N)
    sw = O.Swarm(N, WMIN, WMAX); sw.set_state(start); sw.set_goals(goal); sw.set_capture(True)
    T = O.Tables(); lb, ub = _bounds()
    for step in range(14):
        pos, vel, acc = sw.state()
        sw.step()
        if step >= 8:
            pred = sw.pred(); nr, d, _ = sw.capture()
            for a in (0, 3):
                rows = _rows_of(sw, a, pred, nr, d)
                st = np.stack([pos[a], vel[a], acc[a]]).astype(np.float64)
                ts = O.terminal_segments(pos[a], goal[a])
                r0 = T.solve(st, goal[a].astype(np.float64), ts, lb, ub, [1, 1, 1], [2, 2, 2], rows)
                r1 = T.solve_slack(st, goal[a].astype(np.float64), ts, lb, ub, [1, 1, 1], [2, 2, 2], rows, np.zeros(len(rows), np.int32))
                assert r0["status"] == r1["status"] == 0 and r0["iters"] == r1["iters"]
                assert np.array_equal(r0["x"], r1["x"]) and r0["cost"] == r1["cost"] and r1["slack_cost"] == 0
        sw.advance()


def test_slack_solver_matches_independent_nnls():
    """Agents thrown into their neighbours' collision region: without slack the QPs are infeasible, with slack they are
    solved and agree with the independent extended solve (x, eps, objective)."""
    N = 8
    start, goal = _circle(N, 1.2)
    T = O.Tables(); lb, ub = _bounds()
    checked = 0
    for w in (1.0, 100000.0):                       # src/param.cpp:75 default and launch/simulation.launch:69
        sw = O.Swarm(N, WMIN, WMAX); sw.set_state(start); sw.set_goals(goal); sw.set_capture(True); sw.set_slack_weight(w)
        for step in range(12):
            if step == 5:
                pos, _, _ = sw.state()
                disturb(sw, 2, (pos[3] - pos[2]) * 0.7)        # lands 0.3 of the way from agent 3: inside r_i + r_j if close
            pos, vel, acc = sw.state()
            sw.step()
            q = sw.qp()
            assert (q["status"] == 0).all(), (step, q["status"])
            assert q["maxviol"].max() <= 1e-6 + 1e-12 and q["kkt"].max() <= 1e-8
            if step >= 5:
                ever = sw.reset_ever()
                assert ever[2] == 1 and ever.sum() == 1
                pred = sw.pred(); nr, d, _ = sw.capture(); traj = sw.traj()
                for a in (2, 3, 6):
                    rows = _rows_of(sw, a, pred, nr, d)
                    # obs_slack_indices: every obstacle for the reset agent, the reset agent for everybody else
                    sl = np.array([1 if (a == 2 or j == 2) else 0 for j in range(N) if j != a for _ in range(5)], np.int32)
                    st = np.stack([pos[a], vel[a], acc[a]]).astype(np.float64)
                    ts = O.terminal_segments(pos[a], goal[a])
                    res = T.solve_slack(st, goal[a].astype(np.float64), ts, lb, ub, [1, 1, 1], [2, 2, 2], rows, sl, w)
                    assert res["status"] == 0
                    xo = traj[a].transpose(2, 0, 1).reshape(90)
                    assert np.abs(xo - res["x"].astype(np.float32)).max() == 0          # the stepper ran this very solve
                    D = T.dense(st, goal[a].astype(np.float64), ts, lb, ub, [1, 1, 1], [2, 2, 2], rows)
                    E, idx = R.with_slack(D, rows, sl, w)
                    xe, obj, status = R.solve_ldp(E)
                    assert status == "ok"
                    assert np.abs(res["x"] - xe[:90]).max() <= 2e-5
                    assert np.abs(res["eps"][idx] - xe[90:]).max() <= 2e-5
                    assert abs(res["cost"] - obj) <= 1e-5 * max(1.0, abs(obj))
                    assert (res["eps"] <= 1e-12).all()
                    checked += 1
            sw.advance()
        sc, srows = sw.slack()
    assert checked >= 30


def test_slack_is_used_and_costs():
    """One step after a violent disturbance the disturbed agent needs its slack variables: eps < 0, cost share > 0, and
    the same QP without them is infeasible."""
    N = 8
    start, goal = _circle(N, 1.2)
    sw = O.Swarm(N, WMIN, WMAX); sw.set_state(start); sw.set_goals(goal); sw.set_capture(True); sw.set_slack_weight(100000.0)
    T = O.Tables(); lb, ub = _bounds()
    for step in range(7):
        if step == 5:
            pos, _, _ = sw.state()
            disturb(sw, 2, (pos[3] - pos[2]) * 0.8)
        pos, vel, acc = sw.state()
        sw.step()
        if step == 5:
            sc, srows = sw.slack()
            assert srows[2] > 0 and sc[2] > 0
            q = sw.qp()
            assert q["flags"][2] & 1 and (q["flags"][np.arange(N) != 2] & 1).sum() == 0
            pred = sw.pred(); nr, d, _ = sw.capture()
            assert np.array_equal(pred[2].reshape(30, 3), np.broadcast_to(pos[2], (30, 3)))     # collapsed to the position
            rows = _rows_of(sw, 2, pred, nr, d)
            st = np.stack([pos[2], vel[2], acc[2]]).astype(np.float64)
            hard = T.solve(st, goal[2].astype(np.float64), O.terminal_segments(pos[2], goal[2]), lb, ub, [1, 1, 1], [2, 2, 2], rows)
            assert hard["status"] == 1
        sw.advance()


def test_reset_rearms_the_corridor(golden_dir):
    """flag_initialize_sfc after a reset: all five boxes are regrown from the observed position (traj_planner.cpp:1059,
    1454-1462)."""
    import os
    bt = os.path.join(golden_dir, "worlds", "simple_forest.bt")
    if not os.path.exists(bt):
        import pytest
        pytest.skip("forest map fixture not present")
    wmin, wmax = [-6, -6, 0], [6, 6, 2.5]
    omap = O.Map.from_bt(bt, wmin, wmax)
    N = 6
    start, goal = _circle(N, 5.0)
    sw = O.Swarm(N, wmin, wmax, use_octomap=True, omap=omap); sw.set_state(start); sw.set_goals(goal)
    for step in range(8):
        if step == 4:
            disturb(sw, 1, (0.0, 0.0, 0.4))
        sw.step()
        if step == 4:
            b = sw.boxes()
            assert (b[1] == b[1][0]).all()                # five identical boxes again
            pos, _, _ = sw.state()
            assert (b[1][0][:3] <= pos[1]).all() and (pos[1] <= b[1][0][3:]).all()
            ok, box, _ = omap.sfc_expand(pos[1], goal[1])
            assert ok and np.array_equal(box, b[1][0])
        sw.advance()

"""GPU parity of the disturbance branch (SURVEY.md §8 f4 + a16) through the C-ABI, against the oracle:
state resets (src/traj_planner.cpp:866-878,1047-1061) and the slack-variable QP (src/traj_optimizer.cpp:317-326,
383-390,455-457). Teacher-forced: before every step the engine is loaded with the oracle's planner state; both see the
same disturbed states, reported the way MultiSyncSimulator::update does (observed position, zero velocity and
acceleration; src/multi_sync_simulator.cpp:229-246).

Tolerances as tests/test_gpu_parity.py: status identical; trajectories <= 2e-6 m and relative cost <= 1e-6 unless a row
sits inside the 1e-6 feasibility band (then 2e-5 m / 1e-5).
"""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

SLACK_NEEDED, SFC_BLOCKED, SLACK_MODE, SLACK_USED, SLACK_OVERFLOW = 1, 2, 4, 8, 16


def _ring(n, radius, wmin=(-5, -5, 0), wmax=(5, 5, 2.5)):
    import lsc_planner_b200 as L
    ang = 2 * np.pi * np.arange(n) / n
    start = np.stack([radius * np.cos(ang), radius * np.sin(ang), np.ones(n)], 1).astype(np.float32)
    goal = (start * [-1, -1, 1]).astype(np.float32)
    return L.scenarios.Scenario(f"ring_{n}", start, goal, tuple(wmin), tuple(wmax), [L.AgentType()] * n)


def _run(scn, steps, disturbances, slack_w, omap=None, bt=None, capture=True):
    """disturbances: {step: [(agent, fn(pos) -> new position)]}. Returns a summary dict."""
    import lsc_planner_b200 as L
    n = scn.n
    sw = O.Swarm(n, scn.world_min, scn.world_max, use_octomap=omap is not None, omap=omap)
    sw.set_state(scn.start); sw.set_goals(scn.goal); sw.set_slack_weight(slack_w)
    e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=omap is not None),
                       scn.agents)
    if bt:
        e.set_octomap_file(bt)
    e.set_slack_collision_weight(slack_w)
    summary = dict(used=0, mode=0, resets=0, worst=0.0, infeasible=0, in_band=0, agent_steps=0, slack_cost=0.0)
    ever = np.zeros(n, bool)
    for step in range(steps):
        pos, vel, acc = sw.state()
        for (agent, fn) in disturbances.get(step, []):
            pos[agent] = np.asarray(fn(pos), np.float32); vel[agent] = 0; acc[agent] = 0
            ever[agent] = True
        sw.set_state(pos, vel, acc)
        if omap is not None:
            e.set_sfc(sw.boxes(), np.full(n, 1 if step == 0 else 0, np.int32))
        e.set_prev_traj(sw.traj(), sw.seq)
        assert np.array_equal(e.reset_state(), sw.reset_ever())
        sw.step()
        out = e.replan(pos, vel, acc, scn.goal)
        q = sw.qp()
        sc, srows = sw.slack()
        assert np.array_equal(e.initial_traj().view(np.uint32), sw.pred().view(np.uint32)), step
        assert np.array_equal(e.reset_state(), sw.reset_ever()) and np.array_equal(sw.reset_ever().astype(bool), ever)
        if omap is not None:
            bx, _ = e.get_sfc()
            assert np.array_equal(bx.view(np.uint32), sw.boxes().view(np.uint32)), step
        assert np.array_equal(out["qp_status"], q["status"]), (step, out["qp_status"], q["status"])
        assert np.array_equal(out["flags"] & 3, q["flags"]), step
        assert ((out["flags"] & SLACK_MODE) != 0).all() == bool(ever.any()) and not (out["flags"] & SLACK_OVERFLOW).any()
        diffs = np.abs(out["traj"] - sw.traj()).reshape(n, -1).max(1)
        in_band = (q["maxviol"] > 1e-9) | ((out["flags"] & 64) != 0)     # rows inside the feasibility band, at either solution
        assert diffs[~in_band].max(initial=0) <= 2e-6, (step, diffs.max())
        assert diffs.max() <= 2e-5, (step, diffs.max())
        ok = q["status"] == 0
        rel = np.abs(out["qp_cost"] - q["cost"]) / np.maximum(1.0, np.abs(q["cost"]))
        assert rel[ok & ~in_band].max(initial=0) <= 1e-6 and rel[ok].max(initial=0) <= 1e-5, (step, rel.max())
        # slack variables in use where the oracle's are (beyond noise), nowhere else
        used = (out["flags"] & SLACK_USED) != 0
        assert used[ok & (sc > 1e-9)].all() and not used[ok & (srows == 0) & ~in_band].any(), step
        summary["used"] += int(used.sum()); summary["mode"] += int(((out["flags"] & SLACK_MODE) != 0).sum())
        summary["resets"] += int((out["flags"] & SLACK_NEEDED != 0).sum())
        summary["worst"] = max(summary["worst"], float(diffs.max()))
        summary["infeasible"] += int((~ok).sum()); summary["in_band"] += int(in_band.sum()); summary["agent_steps"] += n
        summary["slack_cost"] = max(summary["slack_cost"], float(sc.max()))
        sw.advance()
    return summary


@pytest.mark.parametrize("slack_w", [1.0, 100000.0])
def test_reset_into_a_neighbour(slack_w):
    """Agent 2 of a tight ring is thrown most of the way onto agent 3: without slack variables its QP (and its
    neighbours') is infeasible; with them everything is solved and agrees with the oracle."""
    scn = _ring(8, 1.2)
    s = _run(scn, 30, {5: [(2, lambda p: p[2] + (p[3] - p[2]) * 0.8)]}, slack_w)
    assert s["resets"] == 1 and s["used"] > 0 and s["mode"] == 25 * 8 and s["slack_cost"] > 0
    assert s["infeasible"] == 0


def test_several_resets_crowd():
    """24 agents crossing; three agents are displaced at different steps (one of them twice), among them displacements
    that do not create any conflict: the slack variables exist but stay zero."""
    scn = _ring(24, 2.0)
    dist = {4: [(1, lambda p: p[1] + np.float32([0.0, 0.0, 0.5]))],
            9: [(7, lambda p: p[7] + (p[8] - p[7]) * 0.6)],
            15: [(1, lambda p: p[1] + np.float32([0.3, -0.2, 0.0])), (13, lambda p: p[13] + (p[12] - p[13]) * 0.7)]}
    s = _run(scn, 45, dist, 100000.0)
    assert s["resets"] == 4 and s["used"] > 0 and s["worst"] <= 2e-5


def test_reset_in_the_forest(golden_dir):
    """With an octomap: the reset re-arms the corridor (all five boxes regrown from the observed position); boxes stay
    bit-identical to the oracle's."""
    import lsc_planner_b200 as L
    bt = os.path.join(golden_dir, "worlds", "simple_forest.bt")
    wmin, wmax = [-5, -5, 0], [5, 5, 2.5]
    om = O.Map.from_bt(bt, wmin, wmax)
    e = L.ReplanEngine(2, L.Param(world_use_octomap=True, world_min=wmin, world_max=wmax), None)
    e.set_octomap_file(bt)
    dm = e.distmap()
    scn = L.scenarios.random_forest(16, dm["sqdist"], dm["off"], seed=3)
    dist = {6: [(4, lambda p: p[4] + np.float32([0.0, 0.0, 0.35]))], 12: [(9, lambda p: p[9] + np.float32([0.25, 0.0, 0.0]))]}
    s = _run(scn, 25, dist, 100000.0, omap=om, bt=bt)
    assert s["resets"] == 2 and s["mode"] > 0


def test_closed_loop_after_reset_resident():
    """A reset uploaded with lscgpu_set_states, then device-resident steps (the CUDA-graphed path): the slack kernel keeps
    planning, nobody is reset again, all QPs are solved."""
    import lsc_planner_b200 as L
    scn = _ring(12, 1.8)
    e = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max), scn.agents)
    e.set_slack_collision_weight(100000.0)
    e.set_states(scn.start); e.set_goals(scn.goal)
    e.replan_resident(6)
    out = e.fetch()
    assert not (out["flags"] & SLACK_MODE).any()
    pos = out["next_position"].copy(); vel = out["next_velocity"].copy(); acc = out["next_acceleration"].copy()
    pos[5] += (pos[6] - pos[5]) * 0.5; vel[5] = 0; acc[5] = 0
    e.set_states(pos, vel, acc)
    e.replan_resident(1)
    out = e.fetch()
    assert (out["flags"] & SLACK_NEEDED).nonzero()[0].tolist() == [5] and (out["flags"] & SLACK_MODE).all()
    e.replan_resident(20)
    out = e.fetch()
    assert (out["flags"] & SLACK_MODE).all() and not (out["flags"] & SLACK_NEEDED).any()
    assert (out["qp_status"] == 0).all()
    assert e.reset_state().nonzero()[0].tolist() == [5]
    e.reset()
    assert not e.reset_state().any()


def test_reference_lp_dump_with_slack(golden_dir):
    """The reference's only recorded QP (log/QPmodel.lp) is infeasible — CPLEX said so, and so do the oracle and the engine.
    Put its conflicting neighbour (obstacle 3) into obs_slack_indices, as the reference does after a disturbance: the QP
    is solved, the operator-level entry (TrajOptimizer::solve with CollisionConstraints::getSlackIndices) agrees with the
    oracle and with the independent NNLS solve of the extended problem on x, eps and the objective."""
    import lsc_planner_b200 as L
    import qp_pyref as R
    from test_gpu_parity import _lp_problem
    g = np.load(os.path.join(golden_dir, "qpmodel_lp.npz"))
    agents = [L.AgentType(max_vel=tuple(g["vmax"]), max_acc=tuple(g["amax"]))]
    lb, ub = g["lb"], g["ub"]
    wmin = [lb[3], lb[33], lb[63]]; wmax = [ub[3], ub[33], ub[63]]
    e = L.ReplanEngine(1, L.Param(world_min=wmin, world_max=wmax), agents)
    T = O.Tables()
    normal, point, d, rows = _lp_problem(g, None)
    n_obs = len(normal)
    st = np.zeros(9); st[:3] = g["state"][0]
    for w in (1.0, 100000.0):
        e.set_slack_collision_weight(w)
        for members in ([3], list(range(n_obs))):
            obs_slack = np.zeros(n_obs, np.uint8); obs_slack[members] = 1
            r = e.qp_solve_batch([0], st, g["goal"], [0, n_obs], normal, point, d, obs_slack=obs_slack)
            row_slack = np.repeat(obs_slack, 5).astype(np.int32)              # rows are (obstacle, segment) entries
            ro = T.solve_slack(g["state"], g["goal"], int(g["ts"]), lb, ub, g["vmax"], g["amax"], rows, row_slack, w)
            assert r["status"][0] == 0 and ro["status"] == 0
            assert np.abs(r["x"][0] - ro["x"]).max() <= 1e-6
            assert np.abs(r["eps"].reshape(-1) - ro["eps"]).max() <= 1e-6 and r["eps"].min() < 0 and r["eps"].max() <= 0
            assert abs(r["cost"][0] - ro["cost"]) <= 1e-8 * max(1.0, abs(ro["cost"]))
            D = T.dense(g["state"], g["goal"], int(g["ts"]), lb, ub, g["vmax"], g["amax"], rows)
            E, idx = R.with_slack(D, rows, row_slack, w)
            xe, obj, stt = R.solve_ldp(E)
            assert stt == "ok"
            assert np.abs(r["x"][0] - xe[:90]).max() <= 1e-6 and abs(r["cost"][0] - obj) <= 1e-7 * max(1.0, abs(obj))
            assert np.abs(r["eps"].reshape(-1)[idx] - xe[90:]).max() <= 1e-6
    # no member: the plain entry, infeasible as recorded
    r = e.qp_solve_batch([0], st, g["goal"], [0, n_obs], normal, point, d, obs_slack=np.zeros(n_obs, np.uint8))
    assert r["status"][0] == 1 and not r["eps"].any()


def test_lp_dump_of_a_failed_qp(tmp_path):
    """The engine's LP dump (the reference's log/QPmodel.lp, written on a failed solve): parsed back, it is the problem the
    oracle assembles for the same agent — coefficient for coefficient up to the 15 printed digits — and the independent
    NNLS solve calls it infeasible too; the dump of a solved agent reproduces the engine's trajectory."""
    import lsc_planner_b200 as L
    import qp_pyref as R
    scn = _ring(8, 1.2)
    n = scn.n
    sw = O.Swarm(n, scn.world_min, scn.world_max, reset_threshold=1e9); sw.set_state(scn.start); sw.set_goals(scn.goal); sw.set_capture(True)
    e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max, reset_threshold=1e9), scn.agents)
    e.set_lp_dump_dir(str(tmp_path))
    T = O.Tables()
    failed_seen = 0
    for step in range(8):
        pos, vel, acc = sw.state()
        if step == 5:            # thrown onto a neighbour, but below the (huge) reset threshold: no slack variables, the QP fails
            pos[2] = pos[2] + (pos[3] - pos[2]) * 0.8; vel[2] = 0; acc[2] = 0
            sw.set_state(pos, vel, acc)
        e.set_prev_traj(sw.traj(), sw.seq)
        sw.step()
        out = e.replan(pos, vel, acc, scn.goal)
        if step != 5:
            sw.advance()
            continue
        files = sorted(os.listdir(tmp_path))
        bad = np.flatnonzero(out["qp_status"] != 0)
        assert len(bad) > 0 and files == [f"QPmodel_agent{a}_seq{e.planner_seq}.lp" for a in bad]
        nr, d, _ = sw.capture(); pred = sw.pred()
        for a in list(bad[:2]) + [int(np.flatnonzero(out["qp_status"] == 0)[0])]:
            path = str(tmp_path / f"dump_{a}.lp")
            e.dump_qp_lp(int(a), path)
            lp = R.parse_lp(open(path).read())
            rows = []
            for j in range(n):
                if j == a:
                    continue
                for m in range(5):
                    aa = nr[a, j, m].astype(np.float64)
                    rows.append((m, aa, d[a, j, m] + pred[j, m].astype(np.float64) @ aa))
            st = np.stack([pos[a], vel[a], acc[a]]).astype(np.float64)
            ts = O.terminal_segments(pos[a], scn.goal[a])
            lb = np.full(90, -np.inf); ub = np.full(90, np.inf)
            for k in range(3):
                for m in range(5):
                    for i in range(6):
                        if not (m == 0 and i < 3):
                            lb[k * 30 + m * 6 + i] = scn.world_min[k]; ub[k * 30 + m * 6 + i] = scn.world_max[k]
            D = T.dense(st, scn.goal[a].astype(np.float64), ts, lb, ub, [1, 1, 1], [2, 2, 2], rows)
            assert lp["Ain"].shape == D["Ain"].shape and lp["Aeq"].shape == D["Aeq"].shape
            assert np.abs(lp["P"] - D["P"]).max() <= 1e-9 * np.abs(D["P"]).max()
            assert np.abs(lp["q"] - D["q"]).max() <= 1e-12 and abs(lp["c0"] - D["c0"]) <= 1e-12
            assert np.abs(lp["Aeq"] - D["Aeq"]).max() <= 1e-9 and np.abs(lp["beq"] - D["beq"]).max() <= 1e-12
            assert np.abs(lp["Ain"] - D["Ain"]).max() <= 1e-6 * max(1.0, np.abs(D["Ain"]).max())
            assert np.abs(lp["bin"] - D["bin"]).max() <= 1e-6
            fin = np.isfinite(D["lb"])
            assert np.array_equal(np.isfinite(lp["lb"]), fin) and np.abs(lp["lb"][fin] - D["lb"][fin]).max() <= 1e-6
            x, obj, status = R.solve_ldp(lp)
            if out["qp_status"][a] != 0:
                assert status == "infeasible"
                failed_seen += 1
            else:
                assert status == "ok"
                xe = out["traj"][a].transpose(2, 0, 1).reshape(90)
                assert np.abs(xe - x).max() <= 2e-5 and abs(out["qp_cost"][a] - obj) <= 1e-5 * max(1.0, abs(obj))
        break
    assert failed_seen >= 1

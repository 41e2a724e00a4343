"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle and the committed golden fixtures.

Tolerances:
  GJK closest point      <= 1e-9 m vs the reference's own openGJK outputs (golden/gjk_ref_vectors.npz)
  distance field, SFC    bit-exact (integer squared distances; boxes compared as float32 bit patterns)
  LSC normals / margins  <= 1e-6 (float32 normals; in practice bit-identical)
  QP                     status identical; trajectory <= 2e-6 m (two float32 ulps at 8 m: the trajectory is handed on as
                         float32) and relative objective gap <= 1e-6 whenever no row of
                         the agent's QP sits inside the 1e-6 feasibility band (CPLEX EpRHS, oracle/qp.hpp): there
                         the minimiser is unique. Rows violated by less than the band are ignored by both solvers, and
                         which of them end up active can depend on pricing order; those agent-steps (a few %) must
                         agree within 2e-5 m / 1e-5 relative cost, and every solution must satisfy every row to 1e-6.
"""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

WMIN, WMAX = [-5, -5, 0], [5, 5, 2.5]


def _engine(n, **kw):
    import lsc_planner_b200 as L
    agents = kw.pop("agents", None)
    return L.ReplanEngine(n, L.Param(**kw), agents)


@pytest.fixture(scope="module")
def forest_path(golden_dir):
    return os.path.join(golden_dir, "worlds", "simple_forest.bt")


# ---------------------------------------------------------------------------------------------------------
def test_gjk_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "gjk_ref_vectors.npz"))
    e = _engine(2)
    v, it = e.gjk_batch(g["hulls"])
    assert ((it >= 1) & (it <= 25)).all()
    assert np.abs(v - g["v"]).max() <= 1e-9
    assert np.abs(np.linalg.norm(v, axis=1) - g["dist"]).max() <= 1e-9
    # and against the oracle on fresh random hulls (float32-representable, as the planner feeds them)
    rng = np.random.default_rng(11)
    H = (rng.normal(size=(2000, 6, 3)) * rng.choice([0.05, 1.0, 4.0], size=(2000, 1, 1)) +
         rng.normal(size=(2000, 1, 3)) * rng.choice([0.0, 2.0, 8.0], size=(2000, 1, 1))).astype(np.float32).astype(np.float64)
    v, _ = e.gjk_batch(H)
    vo = np.array([O.gjk(h)[0] for h in H])
    assert np.abs(v - vo).max() <= 1e-9


def test_distance_field_bit_exact(forest_path):
    e = _engine(2, world_use_octomap=True, world_min=WMIN, world_max=WMAX)
    e.set_octomap_file(forest_path)
    dm = e.distmap()
    om = O.Map.from_bt(forest_path, WMIN, WMAX)
    assert list(dm["size"]) == list(om.size) == [101, 101, 26] and list(dm["off"]) == list(om.off)
    assert dm["n_occupied"] == om.n_occ == 4384
    assert np.array_equal(dm["sqdist"].astype(np.int32), om.sqdist())
    # voxel-list upload gives the same field
    e2 = _engine(2, world_use_octomap=True, world_min=WMIN, world_max=WMAX)
    e2.set_octomap_voxels(om.occupied())
    assert np.array_equal(e2.distmap()["sqdist"], dm["sqdist"])


def test_sfc_expand_bit_exact(forest_path):
    e = _engine(2, world_use_octomap=True, world_min=WMIN, world_max=WMAX)
    e.set_octomap_file(forest_path)
    om = O.Map.from_bt(forest_path, WMIN, WMAX)
    rng = np.random.default_rng(3)
    n = 300
    pts = rng.uniform([-4.9, -4.9, 0.05], [4.9, 4.9, 2.45], size=(n, 3)).astype(np.float32)
    pts[:60] = np.round(pts[:60] * 10) / 10            # lattice seeds (degenerate initial boxes)
    pts[60:90, 2] = np.round(pts[60:90, 2] * 10) / 10
    gls = rng.uniform([-4.9, -4.9, 0.05], [4.9, 4.9, 2.45], size=(n, 3)).astype(np.float32)
    box, ok = e.sfc_expand_batch(pts, gls)
    n_ok = 0
    for i in range(n):
        oko, bo, _ = om.sfc_expand(pts[i], gls[i])
        assert bool(ok[i]) == oko, i
        if oko:
            n_ok += 1
            assert np.array_equal(box[i].view(np.uint32), bo.view(np.uint32)), (i, box[i], bo)
    assert n_ok > 150
    # empty map: the box fills the world
    e3 = _engine(2, world_use_octomap=True, world_min=WMIN, world_max=WMAX)
    e3.set_octomap_voxels(np.zeros((0, 3), np.int32))
    box, ok = e3.sfc_expand_batch([[1.03, -2.0, 1.0]], [[3, 3, 1]])
    assert ok[0] == 1 and np.allclose(box[0], [-5, -5, 0, 5, 5, 2.5], atol=1e-5)


# ---------------------------------------------------------------------------------------------------------
def _lp_problem(g, skip=None):
    keep = [r for r in range(len(g["rows_m"])) if skip is None or r // 5 != skip]
    n_obs = len(keep) // 5
    normal = np.zeros((n_obs, 5, 3), np.float32); point = np.zeros((n_obs, 5, 6, 3), np.float32); d = np.zeros((n_obs, 5, 6))
    rows = []
    for t, r in enumerate(keep):
        o, m = t // 5, int(g["rows_m"][r])
        assert m == t % 5
        normal[o, m] = g["rows_a"][r]                   # float32-representable (widened floats in the LP dump)
        d[o, m] = g["rows_rhs"][r]                      # obstacle point = origin -> rhs = d
        rows.append((m, g["rows_a"][r], g["rows_rhs"][r]))
    return normal, point, d, rows


def test_qp_on_reference_lp_dump(golden_dir):
    """log/QPmodel.lp (the reference's only recorded QP): INFEASIBLE like CPLEX said; solvable without the
    conflicting neighbour, where the engine must match the oracle and the independent NNLS solve."""
    import lsc_planner_b200 as L
    import qp_pyref as R
    g = np.load(os.path.join(golden_dir, "qpmodel_lp.npz"))
    agents = [L.AgentType(max_vel=tuple(g["vmax"]), max_acc=tuple(g["amax"]))]
    lb, ub = g["lb"], g["ub"]
    wmin = [lb[3], lb[33], lb[63]]; wmax = [ub[3], ub[33], ub[63]]
    e = _engine(1, agents=agents, world_min=wmin, world_max=wmax)
    T = O.Tables()
    for skip, expect in ((None, 1), (3, 0)):
        normal, point, d, rows = _lp_problem(g, skip)
        # the engine derives ts from |goal - pos| / v_nom: the dump's ts must be reproduced
        st = np.zeros(9); st[:3] = g["state"][0]
        r = e.qp_solve_batch([0], st, g["goal"], [0, len(normal)], normal, point, d)
        assert r["status"][0] == expect
        ro = T.solve(g["state"], g["goal"], int(g["ts"]), lb, ub, g["vmax"], g["amax"], rows)
        assert ro["status"] == expect
        if expect == 0:
            assert np.abs(r["x"][0] - ro["x"]).max() <= 1e-7
            assert abs(r["cost"][0] - ro["cost"]) <= 1e-9 * max(1.0, abs(ro["cost"]))
            D = T.dense(g["state"], g["goal"], int(g["ts"]), lb, ub, g["vmax"], g["amax"], rows)
            x, obj, stt = R.solve_ldp(D)
            assert stt == "ok" and np.abs(r["x"][0] - x).max() <= 1e-7 and abs(r["cost"][0] - obj) <= 1e-8 * max(1, abs(obj))
            v_eq, v_in = R.kkt_violation(D, r["x"][0])
            assert v_eq <= 1e-7 and v_in <= 1e-7


# ---------------------------------------------------------------------------------------------------------
def _teacher_forced(scn, steps, omap=None, bt=None, check_lsc_agents=(0,), traj_tol=2e-6):
    """Oracle runs closed loop; before every step the engine is loaded with the oracle's planner state, so both
    plan from identical inputs. Returns per-step max |traj diff|."""
    import lsc_planner_b200 as L
    n = scn.n
    sw = O.Swarm(n, scn.world_min, scn.world_max, use_octomap=omap is not None, omap=omap,
                 radius=[a.radius for a in scn.agents], downwash=[a.downwash for a in scn.agents],
                 vmax=[a.max_vel for a in scn.agents], amax=[a.max_acc for a in scn.agents],
                 v_nom=[a.nominal_velocity for a in scn.agents])
    sw.set_state(scn.start); sw.set_goals(scn.goal); sw.set_capture(True)
    e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=omap is not None),
                       scn.agents)
    if bt:
        e.set_octomap_file(bt)
    e.set_capture_rows(True)          # the LSC assertions below read the rows the planning kernel itself built
    worst = []
    n_culled = n_kept = 0
    for step in range(steps):
        pos, vel, acc = sw.state()
        traj_prev = sw.traj(); seq = sw.seq
        if omap is not None:
            e.set_sfc(sw.boxes(), _init_flags(sw, step))
        e.set_prev_traj(traj_prev, seq)
        sw.step()
        out = e.replan(pos, vel, acc, scn.goal)
        q = sw.qp()
        assert e.planner_seq == sw.seq
        # initial trajectories / predictions: same float arithmetic -> bit-identical
        assert np.array_equal(e.initial_traj().view(np.uint32), sw.pred().view(np.uint32))
        nr_o, d_o, _ = sw.capture()
        pred_o = sw.pred().astype(np.float64)
        t_o_now = sw.traj().astype(np.float64)
        for a in check_lsc_agents:
            nr, d, kept = e.get_lsc_rows(a)       # kept pairs: decoded from the production row store; culled: recomputed
            others = [j for j in range(n) if j != a]
            assert np.abs(nr - nr_o[a, others]).max() <= 1e-6
            assert np.abs(d - d_o[a, others]).max() <= 1e-6
            assert kept.sum() == out["lsc_pairs_kept"][a]
            n_kept += int(kept.sum()); n_culled += int((~kept).sum())
            # exact culling: a dropped pair's rows are strictly satisfied at the oracle's (un-culled) solution
            if q["status"][a] == 0 and (~kept).any():
                rel = t_o_now[a][None] - pred_o[others]                                   # [n-1][5][6][3]
                lhs = np.einsum("omik,omk->omi", rel, nr_o[a, others].astype(np.float64)) - d_o[a, others]
                lhs[:, 0, :3] = np.inf                                                    # rows skipped for the initial state
                assert lhs[~kept].min() > 0.0, (step, a, lhs[~kept].min())
        if omap is not None:
            bx, _ = e.get_sfc()
            assert np.array_equal(bx.view(np.uint32), sw.boxes().view(np.uint32)), step
        assert np.array_equal(out["qp_status"], q["status"]), (step, out["qp_status"], q["status"])
        assert (out["report"] == 5).all()
        t_o = sw.traj()
        diffs = np.abs(out["traj"] - t_o).reshape(n, -1).max(1)
        worst.append(diffs)
        in_band = (q["maxviol"] > 1e-9) | ((out["flags"] & 64) != 0)     # rows inside the feasibility band, at either solution
        assert diffs[~in_band].max(initial=0) <= traj_tol, (step, diffs.max())
        assert diffs.max() <= 2e-5, (step, diffs.max())
        ok = q["status"] == 0
        rel = np.abs(out["qp_cost"] - q["cost"]) / np.maximum(1.0, np.abs(q["cost"]))
        assert rel[ok & ~in_band].max(initial=0) <= 1e-6 and rel[ok].max(initial=0) <= 1e-5
        assert np.array_equal(out["flags"] & 3, q["flags"])          # bits 0-1: the reference's conditions; the rest: engine diagnostics
        sw.advance()
        p2, v2, a2 = sw.state()
        assert np.abs(out["next_position"] - p2).max() <= 2e-5
        assert np.abs(out["next_velocity"] - v2).max() <= 1e-3 and np.abs(out["next_acceleration"] - a2).max() <= 5e-2
    worst = np.concatenate(worst)
    assert (worst > traj_tol).mean() <= 0.05
    assert n_kept > 0
    return worst


def _init_flags(sw, step):
    # oracle: flag_initialize_sfc is true only before the first step
    return np.full(sw.n, 1 if step == 0 else 0, np.int32)


def test_swarm_circle20_teacher_forced(golden_dir):
    import lsc_planner_b200 as L
    scn = L.scenarios.load_mission(os.path.join(golden_dir, "missions", "multi_circle20.json"))
    _teacher_forced(scn, 110, check_lsc_agents=(0, 7, 19))     # through the contact phase at the centre


def test_swarm_circle64_teacher_forced():
    import lsc_planner_b200 as L
    _teacher_forced(L.scenarios.circle_swap(64), 90, check_lsc_agents=(0, 40))


def test_swarm_simple3_teacher_forced(golden_dir):
    import lsc_planner_b200 as L
    scn = L.scenarios.load_mission(os.path.join(golden_dir, "missions", "multi_simple3.json"))
    _teacher_forced(scn, 30, check_lsc_agents=(0, 1, 2))


def test_swarm_forest_teacher_forced(forest_path):
    import lsc_planner_b200 as L
    om = O.Map.from_bt(forest_path, WMIN, WMAX)
    e = _engine(2, world_use_octomap=True, world_min=WMIN, world_max=WMAX)
    e.set_octomap_file(forest_path)
    dm = e.distmap()
    scn = L.scenarios.random_forest(24, dm["sqdist"], dm["off"], seed=0)
    _teacher_forced(scn, 25, omap=om, bt=forest_path, check_lsc_agents=(0, 5))


def test_swarm_closed_loop_circle20(golden_dir):
    """Engine alone, closed loop on its own outputs (device-resident stepping), vs the oracle closed loop."""
    import lsc_planner_b200 as L
    scn = L.scenarios.load_mission(os.path.join(golden_dir, "missions", "multi_circle20.json"))
    n = scn.n
    sw = O.Swarm(n, scn.world_min, scn.world_max)
    sw.set_state(scn.start); sw.set_goals(scn.goal)
    e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max), scn.agents)
    e.set_states(scn.start); e.set_goals(scn.goal)
    min_dist = np.inf
    for step in range(60):
        sw.step(); sw.advance()
        e.replan_resident()
        out = e.fetch()
        assert (out["qp_status"] == 0).all()
        p = out["next_position"].astype(np.float64)
        d = (p[:, None] - p[None]) * [1, 1, 0.5]
        dist = np.sqrt((d ** 2).sum(-1)) + np.eye(n) * 99
        min_dist = min(min_dist, dist.min())
        # same trajectories up to accumulated float noise
        assert np.abs(out["traj"] - sw.traj()).max() <= 1e-3, step
    assert min_dist >= 0.3 - 1e-3            # no collision: downwash-scaled distance >= r_i + r_j


def test_swarm_property_n256():
    """Config 3 size: every agent's solution satisfies every row of ITS OWN QP (checked in numpy from the engine's
    captured constraints), the dynamic limits and the equality constraints; spot agents are re-solved by the oracle."""
    import lsc_planner_b200 as L
    scn = L.scenarios.circle_swap(256)
    n = scn.n
    e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max), scn.agents)
    e.set_states(scn.start); e.set_goals(scn.goal)
    sw = O.Swarm(n, scn.world_min, scn.world_max)
    for step in range(12):
        e.replan_resident()
        out = e.fetch().copy()
        assert (out["qp_status"] == 0).all()
        pred = e.initial_traj()
        x = out["traj"].astype(np.float64)                       # [n][5][6][3]
        # continuity + dynamic limits
        vel = (x[:, :, 1:] - x[:, :, :-1]) * 25.0
        acc = (x[:, :, 2:] - 2 * x[:, :, 1:-1] + x[:, :, :-2]) * 500.0
        assert np.abs(vel[:, 1:]).max() <= 1.0 + 1e-4 and np.abs(acc[:, 1:]).max() <= 2.0 + 2e-2
        assert np.abs(x[:, 1:, 0] - x[:, :-1, 5]).max() <= 1e-6
        assert np.abs(x[:, 4, 5] - x[:, 4, 4]).max() <= 1e-6 and np.abs(x[:, 4, 5] - x[:, 4, 3]).max() <= 1e-6
        for a in (0, 100, 255):
            nr, d = e.get_lsc(a)
            others = [j for j in range(n) if j != a]
            rel = x[a][None] - pred[others].astype(np.float64)    # [n-1][5][6][3]
            lhs = np.einsum("omik,omk->omi", rel, nr.astype(np.float64)) - d
            lhs[:, 0, :3] = 0                                     # rows skipped for the initial state
            assert lhs.min() >= -3e-6, (step, a, lhs.min())      # 1e-6 band + float32 rounding of the trajectory
    # teacher-forced oracle check at this size: one more step, agent 3 re-planned by the oracle from the same inputs
    sw.set_state(out["next_position"], out["next_velocity"], out["next_acceleration"]); sw.set_goals(scn.goal)
    sw.set_traj(out["traj"], e.planner_seq)
    sw.step(3, 4)
    o2 = e.replan(out["next_position"], out["next_velocity"], out["next_acceleration"], scn.goal)
    assert e.planner_seq == sw.seq
    assert sw.qp()["status"][3] == 0
    assert np.abs(o2["traj"][3] - sw.traj()[3]).max() <= 1e-6


def _full_size_check(workload, agents, steps, n_sample):
    """BASELINE configs 4 / 5 at full size: closed loop on the device for `steps` steps, then (i) size-independent
    properties of EVERY agent's solution (continuity, terminal stop, dynamic limits, SFC boxes, LSC rows of spot
    agents) and (ii) the next step of a sample of agents — the most crowded ones and a spread over the swarm —
    re-planned by the oracle from the very same inputs."""
    import bench
    import lsc_planner_b200 as L
    scn, bt = bench.make_scenario(workload, agents)
    if scn is None:
        tmp = L.ReplanEngine(2, L.Param(world_use_octomap=True)); tmp.set_octomap_file(bt); dm = tmp.distmap()
        scn = L.scenarios.random_forest(agents, dm["sqdist"], dm["off"], seed=0); tmp.close()
    n = scn.n
    e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=True), scn.agents)
    e.set_octomap_file(bt)
    e.set_states(scn.start); e.set_goals(scn.goal)
    e.replan_resident(steps)
    out = e.fetch().copy()
    boxes, _ = e.get_sfc()
    ok = out["qp_status"] == 0
    assert ok.mean() > 0.85                                       # static goals: some crowded agents are infeasible (as in the oracle)
    x = out["traj"].astype(np.float64)[ok]
    vel = (x[:, :, 1:] - x[:, :, :-1]) * 25.0
    acc = (x[:, :, 2:] - 2 * x[:, :, 1:-1] + x[:, :, :-2]) * 500.0
    assert np.abs(vel[:, 1:]).max() <= 1.0 + 1e-4 and np.abs(acc[:, 1:]).max() <= 2.0 + 2e-2
    assert np.abs(x[:, 1:, 0] - x[:, :-1, 5]).max() <= 1e-6
    assert np.abs(x[:, 4, 5] - x[:, 4, 4]).max() <= 1e-6 and np.abs(x[:, 4, 5] - x[:, 4, 3]).max() <= 1e-6
    b = boxes.astype(np.float64)[ok]                              # [n][5][6] = min xyz, max xyz of every segment's box
    inside_lo = x - b[:, :, None, :3]; inside_hi = b[:, :, None, 3:] - x
    inside_lo[:, 0, :3] = 0; inside_hi[:, 0, :3] = 0              # the first three points are fixed by the state
    assert inside_lo.min() >= -4e-6 and inside_hi.min() >= -4e-6       # 1e-6 band + float32 rounding at |x| ~ 17 m
    # (ii) oracle re-plan of sampled agents, one step further, from identical inputs
    crowded = np.argsort(-out["lsc_pairs_kept"])[:n_sample // 2]
    spread = np.linspace(0, n - 1, n_sample - len(crowded)).astype(int)
    sample = sorted(set(crowded.tolist()) | set(spread.tolist()))
    pos, vel_, acc_ = out["next_position"].copy(), out["next_velocity"].copy(), out["next_acceleration"].copy()
    seq = e.planner_seq
    sw = bench.oracle_swarm(scn, bt)
    o2 = e.replan(pos, vel_, acc_, scn.goal).copy()
    boxes2, _ = e.get_sfc()
    worst = 0.0
    for a in sample:
        sw.set_state(pos, vel_, acc_); sw.set_goals(scn.goal)
        sw.set_traj(out["traj"], seq); sw.set_boxes(boxes, np.zeros(n, np.int32))
        sw.step(a, a + 1)
        q = sw.qp()
        assert q["status"][a] == o2["qp_status"][a], (a, q["status"][a], o2["qp_status"][a])
        assert np.array_equal(sw.boxes()[a].view(np.uint32), boxes2[a].view(np.uint32)), a
        if q["status"][a] == 0:
            d = float(np.abs(o2["traj"][a] - sw.traj()[a]).max())
            worst = max(worst, d)
            assert d <= (2e-5 if q["maxviol"][a] > 1e-9 else 4e-6), (a, d, q["maxviol"][a])   # 2 float32 ulps at 17 m
            assert abs(o2["qp_cost"][a] - q["cost"][a]) <= 1e-5 * max(1.0, abs(q["cost"][a]))
    e.close()
    return worst


def test_full_size_random_forest_512():
    _full_size_check("random_forest", 512, 30, 16)


def test_full_size_circle_forest_1024():
    _full_size_check("circle_forest", 1024, 60, 16)


# ---------------------------------------------------------------------------------------------------------
# BASELINE configs 3 / 4 / 5 at full size, EVERY agent compared with the oracle (threaded), teacher-forced, in two
# mission phases: from rest, and after `skip` closed-loop steps on the device (agents in contact).
# ---------------------------------------------------------------------------------------------------------
def _report(entry):
    import json
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(entry) + "\n")


def _teacher_forced_all_agents(workload, agents, skip, steps):
    import bench
    import lsc_planner_b200 as L
    scn, bt = bench.make_scenario(workload, agents)
    if scn is None:
        tmp = L.ReplanEngine(2, L.Param(world_use_octomap=True)); tmp.set_octomap_file(bt); dm = tmp.distmap()
        scn = L.scenarios.random_forest(agents, dm["sqdist"], dm["off"], seed=0); tmp.close()
    n = scn.n
    use_map = bool(scn.use_octomap)
    e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=use_map), scn.agents)
    if use_map:
        e.set_octomap_file(bt)
    sw = bench.oracle_swarm(scn, bt)
    threads = os.cpu_count() or 1
    ulp = float(np.spacing(np.float32(np.abs(np.concatenate([scn.world_min, scn.world_max])).max())))
    tol_exact = 2.0 * ulp                       # trajectories are handed on as float32: two ulps at the world's extent
    if skip:
        # reach the phase on the device, then load the oracle with the engine's planner state
        e.set_states(scn.start); e.set_goals(scn.goal)
        e.replan_resident(skip)
        o = e.fetch().copy()
        sw.set_state(o["next_position"], o["next_velocity"], o["next_acceleration"])
        sw.set_traj(o["traj"], e.planner_seq)
        if use_map:
            sw.set_boxes(e.get_sfc()[0], np.zeros(n, np.int32))
    stats = dict(agent_steps=0, in_band=0, above_exact_tol=0, failed=0, worst_exact=0.0, worst_band=0.0, worst_cost_gap=0.0)
    for step in range(steps):
        pos, vel, acc = sw.state()
        first = sw.seq == 0
        if use_map:
            e.set_sfc(sw.boxes(), np.full(n, 1 if first else 0, np.int32))
        e.set_prev_traj(sw.traj(), sw.seq)
        sw.step(0, n, threads)
        out = e.replan(pos, vel, acc, scn.goal)
        q = sw.qp()
        assert e.planner_seq == sw.seq
        assert np.array_equal(e.initial_traj().view(np.uint32), sw.pred().view(np.uint32))
        if use_map:
            assert np.array_equal(e.get_sfc()[0].view(np.uint32), sw.boxes().view(np.uint32)), step
        assert np.array_equal(out["qp_status"], q["status"]), (step, np.flatnonzero(out["qp_status"] != q["status"]))
        assert np.array_equal(out["flags"] & 3, q["flags"])          # bits 0-1: the reference's conditions; the rest: engine diagnostics
        diffs = np.abs(out["traj"] - sw.traj()).reshape(n, -1).max(1)
        in_band = (q["maxviol"] > 1e-9) | ((out["flags"] & 64) != 0)     # rows inside the feasibility band, at either solution
        ok = q["status"] == 0
        assert diffs[~in_band].max(initial=0) <= tol_exact, (step, diffs[~in_band].max(), tol_exact)
        assert diffs.max() <= 2e-5, (step, diffs.max())
        rel = np.abs(out["qp_cost"] - q["cost"]) / np.maximum(1.0, np.abs(q["cost"]))
        assert rel[ok & ~in_band].max(initial=0) <= 1e-6 and rel[ok].max(initial=0) <= 1e-5
        stats["agent_steps"] += n; stats["in_band"] += int(in_band.sum()); stats["failed"] += int((~ok).sum())
        stats["above_exact_tol"] += int((diffs > tol_exact).sum())
        stats["worst_exact"] = max(stats["worst_exact"], float(diffs[~in_band].max(initial=0)))
        stats["worst_band"] = max(stats["worst_band"], float(diffs[in_band].max(initial=0)))
        stats["worst_cost_gap"] = max(stats["worst_cost_gap"], float(rel[ok].max(initial=0)))
        sw.advance()
    stats.update(workload=f"{workload}_{n}", skip=skip, steps=steps, tol_exact=tol_exact,
                 in_band_share=stats["in_band"] / stats["agent_steps"], above_tol_share=stats["above_exact_tol"] / stats["agent_steps"])
    _report(stats)
    assert stats["above_tol_share"] <= 0.02, stats
    e.close()
    return stats


@pytest.mark.parametrize("workload,agents,skip", [("circle", 256, 0), ("circle", 256, 50),
                                                  ("random_forest", 512, 0), ("random_forest", 512, 25),
                                                  ("circle_forest", 1024, 0), ("circle_forest", 1024, 60)])
def test_full_size_every_agent_teacher_forced(workload, agents, skip):
    _teacher_forced_all_agents(workload, agents, skip, 10)

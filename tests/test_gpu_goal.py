"""GPU side of goal planning: k_goal_plan (prior_based without an octomap, on the device) and the
lsc_sim closed loop with the host grid planner in a forest, both against the oracle's restatement of
goalPlanningWithPriority. Goals are float32 points compared bit for bit."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "lsc_planner_b200", "host")
MISSIONS = os.path.join(ROOT, "tests", "golden", "missions")


def _oracle_swarm(scn, omap=None):
    sw = O.Swarm(scn.n, scn.world_min, scn.world_max, use_octomap=omap is not None, omap=omap,
                 radius=[a.radius for a in scn.agents], downwash=[a.downwash for a in scn.agents],
                 vmax=[a.max_vel for a in scn.agents], amax=[a.max_acc for a in scn.agents],
                 v_nom=[a.nominal_velocity for a in scn.agents])
    sw.set_state(scn.start); sw.set_goal_mode(1); sw.set_desired_goals(scn.goal)
    return sw


@pytest.mark.parametrize("name,steps", [("multi_circle20.json", 60), ("crowd", 40)])
def test_device_goal_planning_matches_oracle(name, steps):
    """Teacher-forced: before every step the engine is loaded with the oracle's planner state; both then choose the
    goals and plan. Goals and kinds identical, trajectories within the QP tolerance."""
    import lsc_planner_b200 as L
    if name == "crowd":
        scn = L.scenarios.circle_swap(48)
        # squeeze the swarm so that agents come within priority_dist_threshold of each other (retreat branch)
        scn.start[:] = (scn.start * np.float32([0.35, 0.35, 1.0])).astype(np.float32)
        scn.goal[:] = (scn.goal * np.float32([0.35, 0.35, 1.0])).astype(np.float32)
    else:
        scn = L.scenarios.load_mission(os.path.join(MISSIONS, name))
    n = scn.n
    sw = _oracle_swarm(scn)
    e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max, goal_mode=1), scn.agents)
    kinds = np.zeros(2, int); clipped = 0
    for step in range(steps):
        pos, vel, acc = sw.state()
        e.set_prev_traj(sw.traj(), sw.seq)
        sw.step()
        out = e.replan(pos, vel, acc, scn.goal)
        g_o, k_o = sw.goals()
        assert np.array_equal(out["current_goal"].view(np.uint32), g_o.view(np.uint32)), (step, np.abs(out["current_goal"] - g_o).max())
        assert np.array_equal(out["goal_kind"], k_o), step
        q = sw.qp()
        assert np.array_equal(out["qp_status"], q["status"]), step
        diffs = np.abs(out["traj"] - sw.traj()).reshape(n, -1).max(1)
        in_band = (q["maxviol"] > 1e-9) | ((out["flags"] & 64) != 0)     # rows inside the feasibility band, at either solution
        assert diffs[~in_band].max(initial=0) <= 2e-6 and diffs.max() <= 2e-5, (step, diffs.max())
        kinds += np.bincount(k_o, minlength=2)
        clipped += int((np.linalg.norm(g_o - scn.goal, axis=1) > 1e-3).sum())
        sw.advance()
    assert clipped > 0                      # goals further than goal_radius from the end of the initial trajectory
    if name == "crowd":
        assert kinds[1] > 0                 # the retreat branch was exercised
    e.close()


def _forest_case(case, tmp_path, bt):
    import lsc_planner_b200 as L
    if case == "forest10":
        mission = str(tmp_path / "forest10.json"); _forest_mission(mission)
        return L.scenarios.load_mission(mission)
    tmp = L.ReplanEngine(2, L.Param(world_use_octomap=True)); tmp.set_octomap_file(bt); dm = tmp.distmap(); tmp.close()
    return L.scenarios.random_forest(int(case[len("random"):]), dm["sqdist"], dm["off"], seed=3)


@pytest.mark.parametrize("case,steps", [("forest10", 60), ("random64", 30), ("random512", 6)])
def test_device_goal_planning_with_octomap_matches_oracle(case, steps, tmp_path, golden_dir):
    """goal_mode 1 WITH an octomap: k_goal_astar (priority rule, occupancy grid with the higher-priority agents stamped in,
    A* with the reference's hash-order tie-breaking, line-of-sight goal by ray casting, clip) against the oracle's
    goalPlanningWithPriority, teacher-forced. Goals are float32 points compared bit for bit; the A* expansion counts of
    the whole run are equal too (same searches, not just same goals). random512 is BASELINE config 4's swarm (the dense
    one: agents walled in by higher-priority neighbours re-plan without them)."""
    import lsc_planner_b200 as L
    bt = os.path.join(golden_dir, "worlds", "simple_forest.bt")
    scn = _forest_case(case, tmp_path, bt)
    n = scn.n
    omap = O.Map.from_bt(bt, scn.world_min, scn.world_max)
    sw = _oracle_swarm(scn, omap)
    e = L.ReplanEngine(n, L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=True, goal_mode=1), scn.agents)
    e.set_octomap_file(bt)
    expanded = 0; kinds = np.zeros(2, int); moved = 0
    for step in range(steps):
        pos, vel, acc = sw.state()
        e.set_sfc(sw.boxes(), np.full(n, 1 if sw.seq == 0 else 0, np.int32))
        e.set_prev_traj(sw.traj(), sw.seq)
        sw.step(0, n, os.cpu_count() or 1)
        out = e.replan(pos, vel, acc, scn.goal)
        expanded += e.step_stats()["astar_expansions"]
        g_o, k_o = sw.goals()
        bad = np.flatnonzero((out["current_goal"].view(np.uint32) != g_o.view(np.uint32)).any(axis=1))
        assert len(bad) == 0, (step, bad[:8], out["current_goal"][bad[:4]], g_o[bad[:4]])
        assert np.array_equal(out["goal_kind"], k_o), step
        q = sw.qp()
        assert np.array_equal(out["qp_status"], q["status"]), step
        diffs = np.abs(out["traj"] - sw.traj()).reshape(n, -1).max(1)
        in_band = (q["maxviol"] > 1e-9) | ((out["flags"] & 64) != 0)
        assert diffs[~in_band].max(initial=0) <= 2e-6 and diffs.max() <= 2e-5, (step, diffs.max())
        kinds += np.bincount(k_o, minlength=2)
        moved += int((np.linalg.norm(g_o - scn.goal, axis=1) > 1e-3).sum())
        sw.advance()
    assert moved > 0                               # goals that are not simply the desired goal
    assert expanded == sw.astar_expansions() and expanded > 1000, (expanded, sw.astar_expansions())
    e.close()


def test_device_goal_with_octomap_closed_loop_resident(golden_dir, tmp_path):
    """Device-resident closed loop (goal planning, corridors and QP without a host round trip) equals the host-driven one."""
    import lsc_planner_b200 as L
    bt = os.path.join(golden_dir, "worlds", "simple_forest.bt")
    scn = _forest_case("forest10", tmp_path, bt)
    prm = L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=True, goal_mode=1)
    e1 = L.ReplanEngine(scn.n, prm, scn.agents); e2 = L.ReplanEngine(scn.n, prm, scn.agents)
    e1.set_octomap_file(bt); e2.set_octomap_file(bt)
    e1.set_states(scn.start); e1.set_goals(scn.goal)
    pos = scn.start.copy(); vel = np.zeros_like(pos); acc = np.zeros_like(pos)
    for _ in range(40):
        e1.replan_resident(1)
        out = e2.replan(pos, vel, acc, scn.goal)
        pos, vel, acc = out["next_position"].copy(), out["next_velocity"].copy(), out["next_acceleration"].copy()
    o1 = e1.fetch()
    assert np.array_equal(o1["traj"].view(np.uint32), out["traj"].view(np.uint32))
    assert np.array_equal(o1["current_goal"].view(np.uint32), out["current_goal"].view(np.uint32))
    e1.close(); e2.close()


def test_device_goal_closed_loop_resident():
    """Device-resident closed loop (engine advances its own states) equals the host-driven one with goal_mode 1."""
    import lsc_planner_b200 as L
    scn = L.scenarios.load_mission(os.path.join(MISSIONS, "multi_circle20.json"))
    prm = L.Param(world_min=scn.world_min, world_max=scn.world_max, goal_mode=1)
    e1 = L.ReplanEngine(scn.n, prm, scn.agents); e2 = L.ReplanEngine(scn.n, prm, scn.agents)
    e1.set_states(scn.start); e1.set_goals(scn.goal)
    pos = scn.start.copy(); vel = np.zeros_like(pos); acc = np.zeros_like(pos)
    for _ in range(30):
        e1.replan_resident(1)
        out = e2.replan(pos, vel, acc, scn.goal)
        pos, vel, acc = out["next_position"].copy(), out["next_velocity"].copy(), out["next_acceleration"].copy()
    o1 = e1.fetch()
    assert np.array_equal(o1["traj"].view(np.uint32), out["traj"].view(np.uint32))
    assert np.array_equal(o1["current_goal"].view(np.uint32), out["current_goal"].view(np.uint32))
    e1.close(); e2.close()


def _forest_mission(path, n=10):
    """n crazyflies crossing simple_forest.bt (occupied x in [-2.9,2.6], y in [-2.5,3.2]) from opposite sides."""
    agents = []
    for k in range(n):
        y = -3.6 + 7.2 * k / (n - 1)
        side = -1 if k % 2 == 0 else 1
        agents.append({"type": "crazyflie", "cid": k + 1, "start": [4.2 * side, y, 1.0], "goal": [-4.2 * side, -y, 1.0]})
    ms = {"quadrotors": {"crazyflie": {"max_vel": [1.0, 1.0, 1.0], "max_acc": [2.0, 2.0, 2.0], "radius": 0.15,
                                       "nominal_velocity": 1.0, "downwash": 2.0}},
          "world": [{"dimension": [-5.0, -5.0, 0.0, 5.0, 5.0, 2.5]}], "agents": agents, "obstacles": []}
    with open(path, "w") as f:
        json.dump(ms, f)


@pytest.mark.parametrize("planner", ["device", "host"])
def test_simulator_prior_based_forest_matches_oracle(planner, tmp_path, golden_dir):
    """lsc_sim with mode/goal=prior_based and an octomap: goals from k_goal_astar inside the step (goal/planner=device, the
    default) or from the host grid planner (A* + line of sight, threaded; goal/planner=host), corridors and QP from the
    kernels — against the oracle's closed loop with goal_mode 1."""
    import lsc_planner_b200 as L
    from lsc_planner_b200 import build as B
    B.build()
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    mission = str(tmp_path / "forest10.json"); _forest_mission(mission)
    bt = os.path.join(golden_dir, "worlds", "simple_forest.bt")
    res, summ = str(tmp_path / "result.csv"), str(tmp_path / "summary.csv")
    steps = 40
    r = subprocess.run([os.path.join(HOST, "lsc_sim"), "mission=" + mission, "world/file_name=" + bt, "mode/goal=prior_based",
                        "goal/planner=" + planner, "multisim/record_time_step=0.2", f"multisim/max_planner_iteration={steps + 1}", "result=" + res,
                        "summary=" + summ], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    scn = L.scenarios.load_mission(mission)
    rows = np.genfromtxt(res, delimiter=",", skip_header=1)
    rec = rows.reshape(len(rows), scn.n, 15)
    assert len(rec) >= steps
    omap = O.Map.from_bt(bt, scn.world_min, scn.world_max)
    sw = _oracle_swarm(scn, omap)
    n_astar = 0
    for k in range(steps):
        pos, _, _ = sw.state()
        assert np.abs(rec[k, :, 2:5] - pos).max() <= 2e-3, k
        sw.step(); sw.advance()
        n_astar = sw.astar_expansions()
    assert n_astar > 1000                     # the grid planner really ran
    assert "goal planning" in r.stdout


@pytest.mark.parametrize("record", [0.1, 0.05, 0.07])
def test_safety_audit_matches_oracle(record):
    """lscgpu_safety_audit (the O(N^2) minimum-distance audit of savePlanningResult, on the device) against the oracle on
    the same trajectories. Dyadic sub-times (t/dt = 0.5, 0.25, ...) make the Bernstein powers exact: identical ratios;
    otherwise pow() may differ in the last place and positions by one float32 ulp."""
    import lsc_planner_b200 as L
    scn = L.scenarios.circle_swap(64)
    sw = O.Swarm(scn.n, scn.world_min, scn.world_max)
    sw.set_state(scn.start); sw.set_goals(scn.goal)
    e = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max), scn.agents)
    for step in range(25):
        pos, vel, acc = sw.state()
        e.set_prev_traj(sw.traj(), sw.seq)
        sw.step()
        out = e.replan(pos, vel, acc, scn.goal)
        if step % 6 == 0:
            # audit the ORACLE's trajectories on both sides (the engine's own differ by the QP tolerance)
            e.set_prev_traj(sw.traj(), sw.seq)
            r_g, c_g = e.safety_audit(record, 0.2)
            r_o, c_o = sw.safety_audit(record, 0.2)
            if record in (0.1, 0.05):
                assert np.array_equal(r_g, r_o) and np.array_equal(c_g, c_o), step
            else:
                assert np.abs(r_g - r_o).max() <= 1e-6 and (c_g == c_o).mean() > 0.95
            assert r_o.min() > 0.99
        sw.advance()
    e.close()


def test_mission_completes_with_device_goal_planning():
    """Whole missions, device resident: with prior_based goal planning on the GPU the 64-agent circle swap reaches every
    goal without a collision (safety ratio >= 1 at every recorded sub-time, audited on the device) and without a single
    QP failure; with static goals the same swarm deadlocks at the centre (the reason the reference plans goals)."""
    import lsc_planner_b200 as L
    scn = L.scenarios.circle_swap(64)

    def fly(goal_mode, max_steps):
        e = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max, goal_mode=goal_mode), scn.agents)
        e.set_states(scn.start); e.set_goals(scn.goal)
        worst, fails, done = np.inf, 0, None
        for step in range(max_steps):
            e.replan_resident(1)
            out = e.fetch()
            r, _ = e.safety_audit(0.1, 0.2)
            worst = min(worst, float(r.min())); fails += int((out["qp_status"] != 0).sum())
            if np.linalg.norm(out["next_position"] - scn.goal, axis=1).max() < 0.1:
                done = step + 1
                break
        e.close()
        return done, worst, fails

    done, worst, fails = fly(1, 400)
    assert done is not None and done <= 300, done          # 231 steps = 46.2 s of flight when this test was written
    assert worst >= 1.0 - 1e-4 and fails == 0
    done0, worst0, _ = fly(0, 300)
    assert done0 is None and worst0 >= 1.0 - 1e-4          # static goals: safe, but stuck

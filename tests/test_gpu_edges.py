"""Edge cases of the GPU path against the oracle: degenerate swarm sizes, heterogeneous agents, coincident hulls,
blocked corridor seeds, sharded engines, reset."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
WMIN, WMAX = [-5, -5, 0], [5, 5, 2.5]


def _pair(n, start, goal, agents=None, **kw):
    import lsc_planner_b200 as L
    agents = agents or [L.AgentType()] * n
    e = L.ReplanEngine(n, L.Param(world_min=WMIN, world_max=WMAX, **kw), agents)
    sw = O.Swarm(n, WMIN, WMAX, radius=[a.radius for a in agents], downwash=[a.downwash for a in agents],
                 vmax=[a.max_vel for a in agents], amax=[a.max_acc for a in agents],
                 v_nom=[a.nominal_velocity for a in agents])
    sw.set_state(np.asarray(start, np.float32)); sw.set_goals(np.asarray(goal, np.float32))
    return e, sw


def _lockstep(e, sw, goal, steps, tol=2e-6):
    for step in range(steps):
        pos, vel, acc = sw.state()
        e.set_prev_traj(sw.traj(), sw.seq)
        sw.step()
        out = e.replan(pos, vel, acc, np.asarray(goal, np.float32))
        q = sw.qp()
        assert np.array_equal(out["qp_status"], q["status"]), (step, out["qp_status"], q["status"])
        d = np.abs(out["traj"] - sw.traj()).max()
        in_band = q["maxviol"].max() > 1e-9 or ((out["flags"] & 64) != 0).any()      # a row inside the feasibility band somewhere
        assert d <= (2e-5 if in_band else tol), (step, d)
        sw.advance()
    return out


def test_single_agent_no_neighbours():
    e, sw = _pair(1, [[-3, 0, 1]], [[3, 0, 1]])
    out = _lockstep(e, sw, [[3, 0, 1]], 45)
    assert np.linalg.norm(out["next_position"][0] - [3, 0, 1]) < 0.05       # arrives (6 m at 1 m/s = 30 steps)
    nr, d = e.get_lsc(0)
    assert nr.shape == (0, 5, 3)


def test_two_agents_head_on():
    start = [[-2, 0, 1], [2, 0, 1]]; goal = [[2, 0, 1], [-2, 0, 1]]
    e, sw = _pair(2, start, goal)
    _lockstep(e, sw, goal, 40)


def test_heterogeneous_agents():
    """Different radii / downwash / limits per agent: per-pair downwash ratio and collision distance (src/traj_planner.cpp:1339-1345)."""
    import lsc_planner_b200 as L
    agents = [L.AgentType(radius=0.15, downwash=2.0), L.AgentType(radius=0.25, downwash=1.0, max_vel=(0.5, 0.5, 0.5)),
              L.AgentType(radius=0.1, downwash=3.0, max_acc=(1, 1, 1), nominal_velocity=0.7), L.AgentType()]
    start = [[-2, -2, 1], [2, 2, 1.2], [2, -2, 0.8], [-2, 2, 1]]
    goal = [[2, 2, 1], [-2, -2, 1], [-2, 2, 1.5], [2, -2, 1]]
    e, sw = _pair(4, start, goal, agents)
    _lockstep(e, sw, goal, 50)


def test_coincident_agents_are_infeasible_and_keep_previous_trajectory():
    """Two agents on the same spot: GJK returns the zero vector, the LSC normal is zero and the row 0 >= (r_i+r_j)/2
    is infeasible; the optimizer keeps its previous trajectory and the planner still reports SUCCESS."""
    start = [[0, 0, 1], [0, 0, 1], [3, 3, 1]]; goal = [[2, 0, 1], [-2, 0, 1], [-3, -3, 1]]
    e, sw = _pair(3, start, goal)
    pos, vel, acc = sw.state()
    sw.step()
    out = e.replan(pos, vel, acc, np.asarray(goal, np.float32))
    assert list(sw.qp()["status"]) == [1, 1, 0] and list(out["qp_status"]) == [1, 1, 0]
    assert (out["report"] == 5).all()
    assert np.all(out["traj"][:2] == 0)                         # traj_curr starts zero-initialised (src/traj_planner.cpp:36-39)
    assert np.abs(out["traj"][2] - sw.traj()[2]).max() <= 2e-6


def test_blocked_corridor_seed_is_flagged(golden_dir):
    import lsc_planner_b200 as L
    bt = os.path.join(golden_dir, "worlds", "simple_forest.bt")
    om = O.Map.from_bt(bt, WMIN, WMAX)
    vox = (om.occupied()[200] + 0.5) * 0.1                       # centre of an occupied voxel
    start = np.array([vox, [4, 4, 1]], np.float32); goal = np.array([[4, -4, 1], [-4, -4, 1]], np.float32)
    e = L.ReplanEngine(2, L.Param(world_min=WMIN, world_max=WMAX, world_use_octomap=True))
    e.set_octomap_file(bt)
    sw = O.Swarm(2, WMIN, WMAX, use_octomap=True, omap=om)
    sw.set_state(start); sw.set_goals(goal)
    pos, vel, acc = sw.state()
    sw.step()
    out = e.replan(pos, vel, acc, goal)
    assert list(out["flags"]) == list(sw.qp()["flags"]) and out["flags"][0] & 2 and not out["flags"][1] & 2
    assert np.array_equal(e.get_sfc()[0].view(np.uint32), sw.boxes().view(np.uint32))
    assert np.array_equal(out["qp_status"], sw.qp()["status"])


def test_octomap_required_before_stepping():
    import lsc_planner_b200 as L
    from lsc_planner_b200 import _capi as A
    e = L.ReplanEngine(2, L.Param(world_min=WMIN, world_max=WMAX, world_use_octomap=True))
    with pytest.raises(A.EngineError) as ei:
        e.replan(np.zeros((2, 3), np.float32), 0, 0, np.ones((2, 3), np.float32))
    assert ei.value.code == -5
    with pytest.raises(A.EngineError):
        e.set_octomap_file("/nonexistent/map.bt")


def test_shard_plans_only_its_block_and_reset():
    """A sharded engine without a communicator plans agents [a0,a1) only (the others keep zero records); reset
    restores planner_seq = 0 and the first-step behaviour."""
    import lsc_planner_b200 as L
    scn = L.scenarios.circle_swap(24)
    full = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max), scn.agents)
    part = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max), scn.agents)
    part.set_shard(5, 17)
    z = np.zeros_like(scn.start)
    o_full = full.replan(scn.start, z, z, scn.goal).copy()
    o_part = part.replan(scn.start, z, z, scn.goal).copy()
    assert np.array_equal(o_part["traj"][5:17], o_full["traj"][5:17])
    assert np.all(o_part["traj"][:5] == 0) and np.all(o_part["traj"][17:] == 0)
    # second step of the hand-sharded engine: the caller gathers and refreshes the replicas (include/lscgpu.h); the
    # engine itself only commits the agents it planned and must leave the other replicas alone
    part.set_prev_traj(o_full["traj"], 1)
    o_part2 = part.replan(o_full["next_position"], o_full["next_velocity"], o_full["next_acceleration"], scn.goal).copy()
    o_full2 = full.replan(o_full["next_position"], o_full["next_velocity"], o_full["next_acceleration"], scn.goal).copy()
    assert np.array_equal(o_part2["traj"][5:17], o_full2["traj"][5:17])
    assert np.array_equal(o_part2["qp_status"][5:17], o_full2["qp_status"][5:17])
    full.reset()
    o_full = full.replan(scn.start, z, z, scn.goal).copy()
    o2 = full.replan(o_full["next_position"], o_full["next_velocity"], o_full["next_acceleration"], scn.goal).copy()
    assert full.planner_seq == 2
    full.reset()
    assert full.planner_seq == 0
    o3 = full.replan(scn.start, z, z, scn.goal)
    assert np.array_equal(o3["traj"], o_full["traj"]) and not np.array_equal(o2["traj"], o_full["traj"])


def test_operator_batch_matches_swarm_step():
    """lscgpu_qp_solve_batch fed with the constraints the swarm step built (getLSC layout) reproduces the step."""
    import lsc_planner_b200 as L
    scn = L.scenarios.circle_swap(12)
    e = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max), scn.agents)
    e.set_states(scn.start); e.set_goals(scn.goal)
    e.replan_resident(14)
    prev = e.fetch().copy()
    pos, vel, acc = prev["next_position"], prev["next_velocity"], prev["next_acceleration"]
    out = e.replan(pos, vel, acc, scn.goal).copy()
    pred = e.initial_traj()
    normals, points, ds, states, offs = [], [], [], [], [0]
    for a in range(scn.n):
        nr, d = e.get_lsc(a)
        others = [j for j in range(scn.n) if j != a]
        normals.append(nr); ds.append(d); points.append(pred[others])
        states.append(np.concatenate([pos[a], vel[a], acc[a]]).astype(np.float64))
        offs.append(offs[-1] + len(others))
    r = e.qp_solve_batch(np.arange(scn.n), np.array(states), scn.goal.astype(np.float64), offs,
                         np.concatenate(normals), np.concatenate(points), np.concatenate(ds))
    assert np.array_equal(r["status"], out["qp_status"])
    x = r["x"].reshape(scn.n, 3, 5, 6).transpose(0, 2, 3, 1)
    assert np.abs(x - out["traj"]).max() <= 2e-6
    assert np.abs(r["cost"] - out["qp_cost"]).max() <= 1e-6 * max(1.0, np.abs(out["qp_cost"]).max())


def test_results_do_not_depend_on_the_block_size(monkeypatch):
    """k_agent_plan runs with 128, 256 or 512 threads per agent depending on the share of the swarm an engine plans; the
    row order is canonical, so the trajectories must be bit-identical whichever configuration is forced."""
    import lsc_planner_b200 as L
    scn = L.scenarios.circle_swap(96)
    outs = {}
    for threads in (128, 256, 512):
        monkeypatch.setenv("LSCGPU_PLAN_THREADS", str(threads))
        e = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max), scn.agents)
        e.set_states(scn.start); e.set_goals(scn.goal)
        e.replan_resident(70)
        outs[threads] = e.fetch()
        e.close()
    for threads in (256, 512):
        assert np.array_equal(outs[128]["traj"].view(np.uint32), outs[threads]["traj"].view(np.uint32)), threads
        assert np.array_equal(outs[128]["qp_status"], outs[threads]["qp_status"])
        assert np.array_equal(outs[128]["qp_iterations"], outs[threads]["qp_iterations"])
    assert (outs[128]["lsc_pairs_kept"] > 0).any()


def test_pinned_result_buffer_is_written_by_the_planning_blocks(golden_dir):
    """lscgpu_replan_batch with `out` in pinned host memory: the planning blocks store their records straight into it
    (no copy after the step). Every field equals what the copy path returns from a twin engine fed the same inputs — the
    cycle counters aside — over a closed loop that alternates with device-resident steps (which must leave the buffer
    alone), with and without an octomap."""
    import torch
    import lsc_planner_b200 as L
    from lsc_planner_b200 import _capi as A
    for use_map in (False, True):
        scn = L.scenarios.circle_swap(96)
        prm = L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=use_map)
        e1 = L.ReplanEngine(scn.n, prm, scn.agents); e2 = L.ReplanEngine(scn.n, prm, scn.agents)
        if use_map:
            bt = os.path.join(golden_dir, "worlds", "simple_forest.bt")
            e1.set_octomap_file(bt); e2.set_octomap_file(bt)
        n = scn.n
        pin_in = torch.zeros(n * A.AGENT_IN.itemsize, dtype=torch.uint8).pin_memory()
        pin_out = torch.zeros(n * A.AGENT_OUT.itemsize, dtype=torch.uint8).pin_memory()
        h_in = pin_in.numpy().view(A.AGENT_IN); h_out = pin_out.numpy().view(A.AGENT_OUT)
        h_in["position"] = scn.start; h_in["goal"] = scn.goal
        pos = scn.start.copy(); vel = np.zeros_like(pos); acc = np.zeros_like(pos)
        timing = ("qp_kcycles", "qp_price_kcycles", "lsc_kcycles", "sfc_in_block")
        for step in range(12):
            e1.replan_ptr(pin_in.data_ptr(), pin_out.data_ptr())
            ref = e2.replan(pos, vel, acc, scn.goal)
            for name in A.AGENT_OUT.names:
                if name not in timing:
                    assert np.array_equal(h_out[name], ref[name]), (use_map, step, name)
            if step == 5:
                # a resident step on both engines: the pinned buffer keeps the previous step's records
                before = h_out.copy()
                e1.replan_resident(1); e2.replan_resident(1)
                assert np.array_equal(h_out.view(np.uint8), before.view(np.uint8))
                ref = e2.fetch(); got = e1.fetch()
                assert np.array_equal(got["traj"], ref["traj"])
                h_in["position"] = got["next_position"]; h_in["velocity"] = got["next_velocity"]; h_in["acceleration"] = got["next_acceleration"]
            else:
                e1.advance_inputs_ptr(pin_out.data_ptr(), pin_in.data_ptr())
            pos, vel, acc = h_in["position"].copy(), h_in["velocity"].copy(), h_in["acceleration"].copy()
        e1.close(); e2.close()

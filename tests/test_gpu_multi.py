"""Multi-GPU path (one process per GPU, NCCL all-gather inside liblscgpu.so). Needs >= 2 GPUs; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["LSC_ROOT"])
import numpy as np, torch, torch.distributed as dist
import lsc_planner_b200 as L
from lsc_planner_b200 import sharding
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scn = L.scenarios.circle_swap(150)
prm = L.Param(world_min=scn.world_min, world_max=scn.world_max)
sharded = L.ReplanEngine(scn.n, prm, scn.agents, device=local)
sharding.connect(sharded, rank, world)
assert sharded.n_planned == len(sharding.deal(np.arange(scn.n), world, rank))
single = L.ReplanEngine(scn.n, prm, scn.agents, device=local)
pos = scn.start.copy(); vel = np.zeros_like(pos); acc = np.zeros_like(pos)
for step in range(25):
    o1 = sharded.replan(pos, vel, acc, scn.goal).copy()
    o2 = single.replan(pos, vel, acc, scn.goal).copy()
    assert np.array_equal(o1["qp_status"], o2["qp_status"]), step
    assert np.abs(o1["traj"] - o2["traj"]).max() <= 1e-6, (step, np.abs(o1["traj"] - o2["traj"]).max())
    assert np.abs(o1["qp_cost"] - o2["qp_cost"]).max() <= 1e-6 * max(1.0, np.abs(o2["qp_cost"]).max())
    pos, vel, acc = o2["next_position"].copy(), o2["next_velocity"].copy(), o2["next_acceleration"].copy()
# device-resident closed loop on the sharded engine: every replica must hold the same swarm state
sharded.reset(); sharded.set_states(scn.start); sharded.set_goals(scn.goal)
sharded.replan_resident(20)
mine = torch.from_numpy(sharded.fetch()["traj"].copy()).cuda()
ref = mine.clone(); dist.broadcast(ref, src=0)
assert torch.equal(mine, ref)
dist.barrier()
if rank == 0:
    print("MULTI_OK", world)
'''


def test_two_gpu_all_gather(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, LSC_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_OK 2" in r.stdout


WORKER_GOAL = r'''
import os, sys
sys.path.insert(0, os.environ["LSC_ROOT"])
import numpy as np, torch, torch.distributed as dist
import lsc_planner_b200 as L
from lsc_planner_b200 import sharding
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bt = os.path.join(os.environ["LSC_ROOT"], "tests", "golden", "worlds", "simple_forest.bt")
tmp = L.ReplanEngine(2, L.Param(world_use_octomap=True), device=local); tmp.set_octomap_file(bt); dm = tmp.distmap(); tmp.close()
scn = L.scenarios.random_forest(96, dm["sqdist"], dm["off"], seed=5)
prm = L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=True, goal_mode=1)
sharded = L.ReplanEngine(scn.n, prm, scn.agents, device=local); sharded.set_octomap_file(bt)
sharding.connect(sharded, rank, world)
single = L.ReplanEngine(scn.n, prm, scn.agents, device=local); single.set_octomap_file(bt)
pos = scn.start.copy(); vel = np.zeros_like(pos); acc = np.zeros_like(pos)
moved = 0
for step in range(15):
    o1 = sharded.replan(pos, vel, acc, scn.goal).copy()
    o2 = single.replan(pos, vel, acc, scn.goal).copy()
    # every rank plans the goals of its own agents only; the records of the others arrive through the exchange
    assert np.array_equal(o1["current_goal"].view(np.uint32), o2["current_goal"].view(np.uint32)), step
    assert np.array_equal(o1["goal_kind"], o2["goal_kind"]), step
    assert np.array_equal(o1["qp_status"], o2["qp_status"]), step
    assert np.array_equal(o1["sfc_box"].view(np.uint32), o2["sfc_box"].view(np.uint32)), step
    assert np.abs(o1["traj"] - o2["traj"]).max() <= 1e-6, (step, np.abs(o1["traj"] - o2["traj"]).max())
    moved += int((np.linalg.norm(o2["current_goal"] - scn.goal, axis=1) > 1e-3).sum())
    pos, vel, acc = o2["next_position"].copy(), o2["next_velocity"].copy(), o2["next_acceleration"].copy()
assert moved > 0
dist.barrier()
if rank == 0:
    print("MULTI_GOAL_OK", world)
'''


def test_two_gpu_goal_planning_with_octomap(tmp_path):
    """goal_mode 1 with an octomap on a dealt job: every rank runs k_goal_astar (and the corridor toward the planned goal)
    for the agents it plans; goals, kinds, SFC boxes and statuses equal a single engine's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker_goal.py"
    script.write_text(WORKER_GOAL)
    env = dict(os.environ, LSC_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29523", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_GOAL_OK 2" in r.stdout

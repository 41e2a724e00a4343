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

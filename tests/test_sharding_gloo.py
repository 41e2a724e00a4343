"""N>1 host logic on CPU: two gloo ranks each plan their block of the swarm with the oracle, exchange the blocks the
way the engine's all-gather does, and must reproduce the single-process step bit for bit."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["LSC_ROOT"]); sys.path.insert(0, os.path.join(os.environ["LSC_ROOT"], "tests"))
import numpy as np, torch.distributed as dist
import oracle_lib as O
from lsc_planner_b200 import scenarios, sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
scn = scenarios.circle_swap(21)                      # 21 agents over 2 ranks: blocks of 11 and 10
n = scn.n
a0, a1 = sharding.partition(n, world, rank)
uid = sharding.broadcast_bytes(b"unique-id-%d" % 7 if rank == 0 else None)
assert uid == b"unique-id-7"
sw = O.Swarm(n, scn.world_min, scn.world_max); sw.set_state(scn.start); sw.set_goals(scn.goal)
ref = O.Swarm(n, scn.world_min, scn.world_max); ref.set_state(scn.start); ref.set_goals(scn.goal)
for step in range(8):
    sw.step(a0, a1)                                    # this rank's block only
    full = sharding.all_gather_blocks(sw.traj()[a0:a1], n, world)
    sw.set_traj(full, sw.seq)                          # replica of every agent's new trajectory
    sw.advance()
    ref.step(); ref.advance()
    assert np.array_equal(full.view(np.uint32), ref.traj().view(np.uint32)), (rank, step)
    assert all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(sw.state(), ref.state()))
dist.barrier()
if rank == 0:
    print("SHARD_OK", a0, a1)
'''


def test_partition_rule():
    from lsc_planner_b200 import sharding
    for n, w in ((1024, 8), (21, 2), (5, 8), (256, 1), (1000, 3)):
        blocks = [sharding.partition(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(b0[1] == b1[0] for b0, b1 in zip(blocks, blocks[1:]))
        assert max(b - a for a, b in blocks) == sharding.block_size(n, w)


def test_two_rank_gloo_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, LSC_ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARD_OK 0 11" in r.stdout

"""N>1 host logic on CPU: two gloo ranks each plan the agents the engine's dealing rule gives them (LPT order of the
previous step's cost, dealt round-robin) with the oracle, exchange the records the way the engine's all-gather + commit do,
and must reproduce the single-process step bit for bit."""
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
uid = sharding.broadcast_bytes(b"unique-id-%d" % 7 if rank == 0 else None)
assert uid == b"unique-id-7"
sw = O.Swarm(n, scn.world_min, scn.world_max); sw.set_state(scn.start); sw.set_goals(scn.goal)
ref = O.Swarm(n, scn.world_min, scn.world_max); ref.set_state(scn.start); ref.set_goals(scn.goal)
cost = np.zeros(n, np.int64)                           # replicated: every rank sees every record of the previous step
planned = 0
for step in range(8):
    mine = sharding.deal(sharding.lpt_order(cost), world, rank)
    assert len(mine) in (n // world, n // world + 1)
    sw.step_agents(mine)                               # this rank's share only
    planned += len(mine)
    rec = np.concatenate([sw.traj().reshape(n, 90)[mine], sw.qp()["iters"][mine, None].astype(np.float32)], axis=1)
    full = sharding.all_gather_dealt(rec, mine, n, world)
    sw.set_traj(full[:, :90], sw.seq)                  # replica of every agent's new trajectory
    cost = full[:, 90].astype(np.int64)                # ... and of the cost the next step's order is built from
    sw.advance()
    ref.step(); ref.advance()
    assert np.array_equal(full[:, :90].copy().view(np.uint32), ref.traj().reshape(n, 90).view(np.uint32)), (rank, step)
    assert np.array_equal(cost, ref.qp()["iters"]), (rank, step)
    assert all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(sw.state(), ref.state()))
tot = [None] * world
dist.all_gather_object(tot, planned)
assert sum(tot) == 8 * n
dist.barrier()
if rank == 0:
    print("SHARD_OK", planned)
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
    assert "SHARD_OK" in r.stdout


def test_dealing_rule():
    from lsc_planner_b200 import sharding
    rng = np.random.default_rng(5)
    for n, w in ((1024, 8), (21, 2), (5, 8), (1000, 3)):
        cost = rng.integers(0, 50, n)
        order = sharding.lpt_order(cost)
        assert sorted(order.tolist()) == list(range(n))
        assert all(cost[order[i]] > cost[order[i + 1]] or (cost[order[i]] == cost[order[i + 1]] and order[i] < order[i + 1])
                   for i in range(n - 1))
        shares = [sharding.deal(order, w, r) for r in range(w)]
        assert sorted(np.concatenate(shares).tolist()) == list(range(n))
        assert max(len(s) for s in shares) == sharding.block_size(n, w)
        # balanced: the dealt shares' total costs differ by at most one agent's cost
        if n >= 8 * w:
            tot = [cost[s].sum() for s in shares]
            assert max(tot) - min(tot) <= cost.max()
        ids = np.full(sharding.block_size(n, w) * w, -1)
        recs = np.zeros((len(ids), 2))
        for r, sh in enumerate(shares):
            ids[r * sharding.block_size(n, w): r * sharding.block_size(n, w) + len(sh)] = sh
            recs[r * sharding.block_size(n, w): r * sharding.block_size(n, w) + len(sh), 0] = sh
        out = sharding.scatter_records(recs, ids, n)
        assert np.array_equal(out[:, 0], np.arange(n))

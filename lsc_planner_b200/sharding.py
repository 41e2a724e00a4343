"""Agent sharding across ranks (one process per GPU) and the rendezvous of the engine's own NCCL communicator.

Within one synchronous replanning step every agent depends only on the PREVIOUS step's trajectories of all agents
(Jacobi snapshot, reference src/multi_sync_simulator.cpp:190-318), so any rank can plan any agent from its replica of
the swarm state. The step time of a rank is the sum of its agents' plans and those differ by 10x between a crowded and
a free agent, so the agents are DEALT OUT: every rank derives the same longest-processing-time-first order of all agents
from the previous step's records (`lpt_order`, the rule of k_qp_order) and rank r plans entries r, r + G, r + 2G, ...
(`deal`). Records land in a rank-major gather buffer of ceil(N / G) slots per rank, carry their agent id, and ONE
all-gather (496 B per record), issued by liblscgpu.so on its own stream with its own communicator (lscgpu_nccl_init),
completes every replica; `scatter_records` is the host mirror of what k_commit does with the gathered buffer.
torch.distributed is only the out-of-band channel that carries the NCCL unique id. `partition` (contiguous blocks) is
the rule for callers that shard by hand with lscgpu_set_shard.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def block_size(n_agents: int, world: int) -> int:
    return (n_agents + world - 1) // world


def lpt_order(cost: np.ndarray) -> np.ndarray:
    """Agents by decreasing cost of their previous plan, ties by id (k_qp_order, csrc/kernels_plan.cu)."""
    cost = np.asarray(cost)
    return np.lexsort((np.arange(len(cost)), -cost.astype(np.int64))).astype(np.int32)


def deal(order: np.ndarray, world: int, rank: int) -> np.ndarray:
    """The agents rank `rank` plans, in the order its blocks are issued (lscgpu_nccl_init's rule)."""
    return np.asarray(order)[rank::world]


def scatter_records(gathered: np.ndarray, agent_ids: np.ndarray, n_agents: int) -> np.ndarray:
    """k_commit's rule: slot s of the gather buffer belongs to agent agent_ids[s]; negative ids are empty slots."""
    out = np.zeros((n_agents,) + gathered.shape[1:], gathered.dtype)
    ok = agent_ids >= 0
    out[agent_ids[ok]] = gathered[ok]
    return out


def partition(n_agents: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous blocks (for lscgpu_set_shard callers)."""
    b = block_size(n_agents, world)
    a0 = min(n_agents, rank * b)
    return a0, min(n_agents, a0 + b)


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """Broadcast a small byte string over the default torch.distributed group (any backend)."""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def connect(engine, rank: int, world: int, p2p: bool | None = None):
    """Create the engine's NCCL communicator: rank 0 makes the unique id, everybody joins. Then (p2p, default on unless
    LSCGPU_P2P=0) switch the exchange to direct stores over NVLink peer memory: the ranks swap the CUDA IPC handles of their
    exchange buffers over torch.distributed and map each other's."""
    import os
    import torch.distributed as dist
    uid = broadcast_bytes(engine.nccl_unique_id() if rank == 0 else None)
    engine.nccl_init(uid, rank, world)
    if p2p is None:
        p2p = os.environ.get("LSCGPU_P2P", "1") != "0"
    if p2p and world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, engine.p2p_export())
        engine.p2p_attach(handles)
        dist.barrier()
    return engine


def all_gather_dealt(local: np.ndarray, my_agents: np.ndarray, n_agents: int, world: int) -> np.ndarray:
    """Host-side mirror of the device exchange (CPU tests with the gloo backend): `local[i]` is the record of agent
    my_agents[i]; every rank contributes a block of ceil(N / world) slots (unused ones carry id -1), the blocks are
    all-gathered rank-major and scattered by agent id. Returns all n_agents records in agent order."""
    import torch
    import torch.distributed as dist
    b = block_size(n_agents, world)
    rec = local.reshape(local.shape[0], -1)
    pad = np.zeros((b, rec.shape[1]), rec.dtype)
    pad[: rec.shape[0]] = rec
    ids = np.full(b, -1, np.int64)
    ids[: len(my_agents)] = my_agents
    out = [torch.empty_like(torch.from_numpy(pad)) for _ in range(world)]
    out_ids = [torch.empty_like(torch.from_numpy(ids)) for _ in range(world)]
    dist.all_gather(out, torch.from_numpy(pad))
    dist.all_gather(out_ids, torch.from_numpy(ids))
    full = scatter_records(np.concatenate([t.numpy() for t in out]), np.concatenate([t.numpy() for t in out_ids]), n_agents)
    return full.reshape((n_agents,) + local.shape[1:])

"""Agent sharding across ranks (one process per GPU) and the rendezvous of the engine's own NCCL communicator.

Within one synchronous replanning step every agent depends only on the PREVIOUS step's trajectories of all agents
(Jacobi snapshot, reference src/multi_sync_simulator.cpp:190-318), so agents are block-partitioned: rank r plans
agents [r*B, min(N, (r+1)*B)), B = ceil(N / world). Every rank keeps a replica of all trajectories; the step ends with
ONE all-gather of the per-agent result records (464 B each), issued by liblscgpu.so on its own stream with its own
communicator (lscgpu_nccl_init). torch.distributed is only the out-of-band channel that carries the NCCL unique id.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def block_size(n_agents: int, world: int) -> int:
    return (n_agents + world - 1) // world


def partition(n_agents: int, world: int, rank: int) -> Tuple[int, int]:
    """Same rule as lscgpu_nccl_init (csrc/engine.cu)."""
    b = block_size(n_agents, world)
    a0 = min(n_agents, rank * b)
    return a0, min(n_agents, a0 + b)


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """Broadcast a small byte string over the default torch.distributed group (any backend)."""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def connect(engine, rank: int, world: int):
    """Create the engine's NCCL communicator: rank 0 makes the unique id, everybody joins."""
    uid = broadcast_bytes(engine.nccl_unique_id() if rank == 0 else None)
    engine.nccl_init(uid, rank, world)
    return engine


def all_gather_blocks(local: np.ndarray, n_agents: int, world: int) -> np.ndarray:
    """Host-side mirror of the in-place device all-gather (used by the CPU tests with the gloo backend):
    `local` holds this rank's block of per-agent records; returns all n_agents records in agent order."""
    import torch
    import torch.distributed as dist
    b = block_size(n_agents, world)
    rec = local.reshape(local.shape[0], -1)
    pad = np.zeros((b, rec.shape[1]), rec.dtype)
    pad[: rec.shape[0]] = rec
    out = [torch.empty_like(torch.from_numpy(pad)) for _ in range(world)]
    dist.all_gather(out, torch.from_numpy(pad))
    full = np.concatenate([t.numpy() for t in out])[:n_agents]
    return full.reshape((n_agents,) + local.shape[1:])

#!/usr/bin/env python
"""Multi-GPU step-time diagnostics (torchrun): resident steps with / without the L2 flush, per-kernel breakdown and the
spread of the ranks' plan-kernel times."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
import lsc_planner_b200 as L
from lsc_planner_b200 import sharding

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scn, bt = bench.make_scenario("circle_forest", 1024)
eng = L.ReplanEngine(scn.n, L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=True), scn.agents, device=local)
eng.set_octomap_file(bt)
if world > 1:
    sharding.connect(eng, rank, world)
stream = torch.cuda.ExternalStream(eng.stream, device=local)
buf = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=f"cuda:{local}")
steps = int(os.environ.get("DIAG_STEPS", 30))
for flush in (True,):
    for prof in (False, True):
        eng.reset(); eng.set_states(scn.start); eng.set_goals(scn.goal)
        eng.replan_resident(5); eng.set_profiling(prof)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        for _ in range(steps):
            if flush:
                with torch.cuda.stream(stream): buf.zero_()
            eng.replan_resident(1, sync=False)
        eng.synchronize()
        st = eng.step_stats()
        eng.set_profiling(False)
        print(f"rank {rank} flush {int(flush)} prof {int(prof)}: ms/step {st['ms_steps'] / steps:.4f} predict {st['ms_predict'] / steps:.4f} "
              f"plan {st['ms_plan'] / steps:.4f} sfc {st['ms_sfc'] / steps:.4f} exch {st['ms_exchange'] / steps:.4f} commit {st['ms_commit'] / steps:.4f}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()

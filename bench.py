#!/usr/bin/env python
"""agent-replans/sec of the batched replanning hot path (LSC + SFC + trajectory QP) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--agents A]

A "step" is one synchronous replanning step of the whole swarm (every agent replans once), closed loop: the next
step's states are the new trajectories evaluated at t = dt. Default workload = BASELINE.json configs[4]'s swarm
(synthetic circle-swap, 1024 agents, simple_forest.bt at the centre), which fits one GPU; with --gpus N the same
swarm is block-partitioned over N ranks (strong scaling) and every step ends with one NCCL all-gather.
Prints ONE JSON line (rank 0). See DESIGN.md §6 for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FOREST = os.path.join(ROOT, "tests", "golden", "worlds", "simple_forest.bt")
L2_BYTES = 126 * 2 ** 20


GOAL_MODES = {"static": 0, "prior_based": 1}


def make_scenario(workload: str, agents: int):
    from lsc_planner_b200 import scenarios as S
    if workload == "circle_forest":
        return S.circle_swap(agents, forest=True), FOREST
    if workload == "circle":
        return S.circle_swap(agents, forest=False), None
    if workload == "random_forest":
        # needs the distance field: built by a throw-away engine (GPU) or by the oracle (reference arm)
        return None, FOREST
    raise SystemExit(f"unknown workload {workload}")


def peaks():
    """HBM copy bandwidth of this pool's B200s: MEASURED_PEAKS.json (driver-written) when present, else the profiling
    recipe's fallback (6.65 TB/s, /opt/skills/guides/B200_PROFILING.md). The file's key for the bandwidth is looked up
    leniently: any (possibly nested) key mentioning hbm / copy / bandwidth whose value looks like GB/s."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if isinstance(d.get("hbm_gbs"), (int, float)):
                return {"hbm_gbs": float(d["hbm_gbs"])}, "measured"
            found = []

            def walk(o, path=""):
                if isinstance(o, dict):
                    for k, v in o.items():
                        walk(v, path + "/" + str(k).lower())
                elif isinstance(o, (int, float)) and any(t in path for t in ("hbm", "copy", "bandwidth", "bw")):
                    v = float(o)
                    if 1.0 <= v <= 20.0:
                        v *= 1e3            # TB/s
                    if 2000.0 <= v <= 9000.0:
                        found.append((("sustain" in path), v))
            walk(d)
            if found:
                return {"hbm_gbs": sorted(found)[0][1]}, "measured"      # burst figure preferred (kernel timed alone)
        except (ValueError, OSError):
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def fma_peaks():
    """FP32 / FP64 FMA peaks measured on this pool's B200 by tools/fma_peak.cu (profiles/fma_peaks.json)."""
    p = os.path.join(ROOT, "profiles", "fma_peaks.json")
    return json.load(open(p)) if os.path.exists(p) else None


def alg_flops_per_replan(n_agents: int, octomap: bool, gjk_iters: float, qp_iters: float):
    """F_lsc = (N-1) M (90 + 46 I_gjk) [FP64 GJK + float margins], F_qp = (K+1)(2 nnz(A) + 1080 + 2*39^2)
    + 2/3 39^3; I_gjk (per hull) and K come from the ORACLE's counters on the same inputs."""
    hulls = (n_agents - 1) * 5
    f_lsc = hulls * 90 + 46.0 * gjk_iters          # gjk_iters = total GJK iterations of the agent's hulls
    nnz = 81 * (n_agents - 1) + 618 + (162 if octomap else 0) + 174
    f_qp = (qp_iters + 1) * (2 * nnz + 1080 + 2 * 39 ** 2) + (2.0 / 3.0) * 39 ** 3
    return f_lsc, f_qp


def alg_bytes_per_replan(n_agents: int, l_sfc: float) -> float:
    """neighbours' previous trajectories + radius/downwash, own trajectory, state, goal, SFC window
    in/out, EDT lookups (4 B each, oracle count), output trajectory, status + cost."""
    return (n_agents - 1) * 368 + 360 + 36 + 12 + 48 + 4.0 * l_sfc + 360 + 16


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's CPU path; CPLEX/ROS/octomap are not installable here)
# ------------------------------------------------------------------------------------------------------------
def oracle_swarm(scn, bt, goal_mode=0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    omap = O.Map.from_bt(bt, scn.world_min, scn.world_max) if (bt and scn.use_octomap) else None
    sw = O.Swarm(scn.n, scn.world_min, scn.world_max, use_octomap=omap is not None, omap=omap,
                 radius=[a.radius for a in scn.agents], downwash=[a.downwash for a in scn.agents],
                 vmax=[a.max_vel for a in scn.agents], amax=[a.max_acc for a in scn.agents],
                 v_nom=[a.nominal_velocity for a in scn.agents])
    sw.set_state(scn.start); sw.set_goals(scn.goal)
    if goal_mode:
        sw.set_goal_mode(1); sw.set_desired_goals(scn.goal)
    return sw


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scn, bt = make_scenario(args.workload, args.agents)
    threads = os.cpu_count() or 1
    if scn is None:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        from lsc_planner_b200 import scenarios as S
        m = O.Map.from_bt(bt, [-5, -5, 0], [5, 5, 2.5])
        scn = S.random_forest(args.agents, m.sqdist(), m.off, seed=0)
    sw = oracle_swarm(scn, bt, GOAL_MODES[args.goal_mode])
    n = scn.n
    budget = 150.0
    t_start = time.perf_counter()
    # warm-up steps plan the whole swarm (closed loop); the first one calibrates the per-agent cost
    sample = n
    per_agent = None
    for w in range(args.warmup):
        t0 = time.perf_counter(); sw.step(0, sample, threads); sw.advance()
        dt_ = time.perf_counter() - t0
        per_agent = dt_ / sample
        remaining = budget - (time.perf_counter() - t_start)
        steps_left = args.warmup - w - 1 + args.steps
        sample = int(min(n, max(threads, remaining / max(steps_left, 1) / per_agent)))
    if per_agent is None:
        t0 = time.perf_counter(); sw.step(0, min(n, threads), threads); sw.advance()
        per_agent = (time.perf_counter() - t0) / min(n, threads)
        sample = int(min(n, max(threads, budget / args.steps / per_agent)))
    sw.reset_counters()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sw.step(0, sample, threads); sw.advance()
    el = time.perf_counter() - t0
    value = sample * args.steps / el
    c = sw.counters()
    what = (f"agents [0,{sample}) of {n} planned per step (each against all {n - 1} neighbours), "
            f"{args.steps} closed-loop steps, {threads} threads")
    line = {
        "impl": "reference", "metric": "agent-replans/sec", "value": value, "unit": "agent-replans/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}_{n}", "agents": n, "octomap": bool(scn.use_octomap)},
        "cpu_baseline": {"value": value, "unit": "agent-replans/s", "cores": threads, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": "agent-replans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "oracle_counters_per_replan": {k: v / max(sample * args.steps, 1) for k, v in c.items()},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import lsc_planner_b200 as L
    from lsc_planner_b200 import _capi as A

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scn, bt = make_scenario(args.workload, args.agents)
    if scn is None:   # random_forest: sample starts/goals against the engine's distance field
        tmp = L.ReplanEngine(2, L.Param(world_use_octomap=True), device=local)
        tmp.set_octomap_file(bt)
        dm = tmp.distmap()
        scn = L.scenarios.random_forest(args.agents, dm["sqdist"], dm["off"], seed=0)
        tmp.close()
    n = scn.n
    if GOAL_MODES[args.goal_mode] and scn.use_octomap:
        raise SystemExit("--goal-mode prior_based on the GPU needs a workload without octomap (--workload circle): with an "
                         "octomap the goals come from the host grid planner (lsc_sim)")
    prm = L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=scn.use_octomap,
                  goal_mode=GOAL_MODES[args.goal_mode])
    eng = L.ReplanEngine(n, prm, scn.agents, device=local)
    if scn.use_octomap:
        eng.set_octomap_file(bt)
    if world > 1:
        from lsc_planner_b200 import sharding
        sharding.connect(eng, rank, world)
    n_local = eng.a1 - eng.a0
    stream = torch.cuda.ExternalStream(eng.stream, device=local)

    row_store_bytes = n_local * 5 * (n - 1) * 64
    # every step is timed on its own (begin/end events on the engine stream) with a 256 MB write flushing the 126 MB
    # L2 in between: the data one step touches (kept rows, distance field tables) is smaller than L2
    flush = True
    flush_buf = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=f"cuda:{local}") if flush else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_steps(k):
        for _ in range(k):
            if flush:
                with torch.cuda.stream(stream):
                    flush_buf.zero_()
            eng.replan_resident(1, sync=False)

    # ---- per-kernel breakdown: the same W + K closed-loop steps with the kernels serialised and CUDA events between
    # them (profiling mode); the timed pass below overlaps kernels of different agent groups, where per-kernel
    # durations are not separable
    eng.set_states(scn.start); eng.set_goals(scn.goal)
    eng.set_profiling(True)
    resident_steps(args.warmup)
    eng.synchronize()
    resident_steps(args.steps)
    eng.synchronize()
    st = eng.step_stats()
    kernel_ms = st["ms_predict"] + st["ms_sfc"] + st["ms_lsc"] + st["ms_qp"] + st["ms_exchange"] + st["ms_commit"]
    ms_serial = st["ms_total"]

    # ---- device-resident throughput (`value`) -----------------------------------------------------------------
    eng.reset()
    eng.set_states(scn.start); eng.set_goals(scn.goal)
    eng.set_profiling(False)
    resident_steps(args.warmup)
    eng.synchronize()
    barrier()
    clocks = ClockSampler(local)
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    resident_steps(args.steps)
    ev1.record(stream)
    eng.synchronize()
    barrier()
    st_timed = eng.step_stats()
    # with an L2 flush between steps only the steps themselves count; otherwise the whole bracket
    ms_region = st_timed["ms_steps"] if flush else ev0.elapsed_time(ev1)
    t = torch.tensor([ms_region], dtype=torch.float64, device=f"cuda:{local}")
    launches = torch.tensor([st_timed["kernel_launches"]], dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    out_mid = eng.fetch().copy()
    boxes_mid = eng.get_sfc()[0] if scn.use_octomap else None
    assert (out_mid["report"] == 5).all()
    qp_fail = int((out_mid["qp_status"] != 0).sum())
    value = n * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C-ABI with pinned HOST buffers (`e2e`) ----------------------------------------
    eng.reset()
    eng.set_profiling(False)
    pin_in = torch.zeros(n * A.AGENT_IN.itemsize, dtype=torch.uint8).pin_memory()
    pin_out = torch.zeros(n * A.AGENT_OUT.itemsize, dtype=torch.uint8).pin_memory()
    h_in = pin_in.numpy().view(A.AGENT_IN); h_out = pin_out.numpy().view(A.AGENT_OUT)
    h_in["position"] = scn.start; h_in["velocity"] = 0; h_in["acceleration"] = 0; h_in["goal"] = scn.goal

    def host_step():
        eng.replan_ptr(pin_in.data_ptr(), pin_out.data_ptr())      # H2D + kernels (+ all-gather) + D2H, synchronous
        h_in["position"] = h_out["next_position"]; h_in["velocity"] = h_out["next_velocity"]
        h_in["acceleration"] = h_out["next_acceleration"]

    for _ in range(args.warmup):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n * args.steps / float(e2e_s.item())
    clk = clocks.stop()

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    pk, pk_kind = peaks()
    per_kernel = {"k_lsc_build": st["ms_lsc"], "k_qp_solve": st["ms_qp"], "k_sfc_expand": st["ms_sfc"],
                  "k_predict": st["ms_predict"], "k_commit": st["ms_commit"], "nccl_all_gather": st["ms_exchange"]}
    dom = max(("k_lsc_build", "k_qp_solve", "k_sfc_expand"), key=lambda k: per_kernel[k])
    dom_ms = per_kernel[dom] / max(st["steps"], 1)

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, single-GPU runs only) -------------------
    cpu = None
    l_sfc = 0.0
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sw = oracle_swarm(scn, bt, GOAL_MODES[args.goal_mode])
        # same step as the end of the timed region: the oracle is loaded with the engine's planner state
        sw.set_state(out_mid["next_position"], out_mid["next_velocity"], out_mid["next_acceleration"])
        def restore():
            sw.set_traj(out_mid["traj"], args.warmup + args.steps)
            if boxes_mid is not None:
                sw.set_boxes(boxes_mid, np.zeros(n, np.int32))
        # calibrate on one batch of `threads` agents, then re-plan a sample sized for ~10 s (repeating the same step
        # when the whole swarm takes less than that)
        s_cal = min(n, threads)
        restore()
        t0 = time.perf_counter(); sw.step(0, s_cal, threads); cal = max(time.perf_counter() - t0, 1e-4)
        s = int(min(n, max(s_cal, s_cal * int(10.0 / cal))))
        restore(); sw.reset_counters()
        reps = 0; el = 0.0
        while el < 8.0 and reps < 200:
            restore()
            t0 = time.perf_counter(); sw.step(0, s, threads); el += time.perf_counter() - t0
            reps += 1
        s_total = s * reps
        c = sw.counters()
        l_sfc = c["edt_lookups"] / s_total
        # the reference's own execution model: one thread, agents one after the other, on a smaller sample
        s1 = int(min(n, max(4, 2.0 / max(cal, 1e-6))))          # `cal` = one agent per thread, i.e. one agent's time
        restore()
        t0 = time.perf_counter(); sw.step(0, s1, 1); one_thread = s1 / max(time.perf_counter() - t0, 1e-9)
        cpu = {"value": s_total / el, "unit": "agent-replans/s", "cores": threads, "kind": "port",
               "sample": f"oracle (CPU port of the reference path, oracle/) re-plans agents [0,{s}) of the {n}-agent swarm "
                         f"from the engine's state after the timed region, {reps}x, {threads} threads, {el:.1f} s",
               "one_thread": {"value": one_thread, "unit": "agent-replans/s", "sample": f"agents [0,{s1}) once, 1 thread"},
               "per_replan": {"gjk_iterations": c["gjk_iters"] / s_total, "qp_iterations": c["qp_iters"] / s_total,
                              "edt_lookups": l_sfc, "qp_rows": c["qp_rows"] / s_total}}
    # Algorithmic bytes split by the kernel that consumes / produces each term:
    #   k_lsc_build : neighbours' previous trajectories + radius/downwash, own trajectory, the kept rows it writes
    #   k_sfc_expand: the reference's EDT lookups (4 B each, oracle count), SFC window in/out, goal
    #   k_qp_solve  : the kept rows (64 B record + 8 B gate + 4 B pair index), state, goal, SFC window, result record
    kept_per_replan = st["lsc_pairs_kept"] / max(st["steps"] * n_local, 1)
    kernel_bytes = {"k_lsc_build": (n - 1) * 368 + 360 + 76.0 * kept_per_replan,
                    "k_sfc_expand": 4.0 * l_sfc + 48 + 12 + 36,
                    "k_qp_solve": 76.0 * kept_per_replan + 36 + 12 + 48 + 360 + 16}
    b_alg = kernel_bytes[dom] * n_local
    achieved = b_alg / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{args.workload}_{n}", {}).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json ({pk_kind})",
                "ms_per_launch": dom_ms, "algorithmic_bytes_per_launch": b_alg,
                "kept_pairs_per_replan": kept_per_replan,
                "note": "latency bound, not bandwidth bound (SURVEY.md §8d): the kernel's data is L2-resident and its time "
                        "is the slowest agent's chain of dependent FP64 steps; the fraction is low by construction; "
                        "see DESIGN.md §5 and fma_roofline"}
    # whole path, SURVEY.md §8(d) B_alg per agent-replan (every neighbour read and every EDT lookup of the reference's
    # algorithm counted as if from HBM) over the whole step's device time
    b_path = alg_bytes_per_replan(n, l_sfc) * n_local
    step_ms = ms_total / args.steps
    path_roofline = {"algorithmic_bytes_per_step": b_path, "ms_per_step": step_ms,
                     "achieved": b_path / (step_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": b_path / (step_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                     "note": "upper bound on necessary traffic: counts the reference's brute-force EDT sampling "
                             "(4 B x L_sfc) that k_sfc_expand replaces by summed-volume-table reads"}
    fma = None
    fp = fma_peaks()
    if cpu is not None and fp is not None:
        f_lsc, f_qp = alg_flops_per_replan(n, scn.use_octomap, cpu["per_replan"]["gjk_iterations"], cpu["per_replan"]["qp_iterations"])
        rate = value
        fma = {"algorithmic_flops_per_replan": {"lsc": f_lsc, "qp": f_qp},
               "achieved_tflops": (f_lsc + f_qp) * rate * 1e-12,
               "fp64_peak_tflops": fp["fp64_tflops"], "fp32_peak_tflops": fp["fp32_tflops"],
               "frac_of_fp64_peak": (f_lsc + f_qp) * rate * 1e-12 / fp["fp64_tflops"],
               "note": "all hot-path arithmetic runs in FP64 (GJK, QP); counters from the oracle on the same step"}
    if rank == 0:
        line = {
            "metric": "agent-replans/sec", "value": value, "unit": "agent-replans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}_{n}", "agents": n, "agents_per_gpu": n_local,
                       "octomap": bool(scn.use_octomap), "parallelism": f"agents block-partitioned x{world}",
                       "l2": "L2 flushed (256 MB write) between steps; every step timed on its own with CUDA events on the "
                             "engine stream; e2e is not flushed (its inputs arrive from host memory every step)",
                       "goals": ("prior_based goal planning on the GPU every step (k_goal_plan, SURVEY.md §8f #1)" if GOAL_MODES[args.goal_mode]
                                 else "fixed to the mission goals (goal planning is outside the path, SURVEY.md §8f)")},
            "e2e": {"value": e2e_value, "unit": "agent-replans/s", "h2d_bytes_per_step": n * A.AGENT_IN.itemsize,
                    "d2h_bytes_per_step": n * A.AGENT_OUT.itemsize},
            "gpu_launches": int(launches.item()),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "roofline": roofline, "path_roofline": path_roofline,
            "fma_roofline": fma,
            "cpu_baseline": cpu,
            "kernel_ms_per_step": {k: v / max(st["steps"], 1) for k, v in per_kernel.items()},
            "qp": {"iterations_per_replan": st["qp_iterations"] / max(n_local * st["steps"], 1),
                   "rows_priced_per_replan": st["qp_rows_priced"] / max(n_local * st["steps"], 1),
                   "full_sweeps_per_replan": st["qp_full_passes"] / max(n_local * st["steps"], 1),
                   "gjk_iterations_per_hull": st["gjk_iterations"] / max(st["lsc_pairs"], 1),
                   "failed_last_step": qp_fail},
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="circle_forest", choices=["circle_forest", "circle", "random_forest"])
    ap.add_argument("--agents", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--goal-mode", default="static", choices=list(GOAL_MODES))
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

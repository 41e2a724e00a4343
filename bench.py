#!/usr/bin/env python
"""agent-replans/sec of the batched replanning hot path (LSC + SFC + trajectory QP) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--agents A]

A "step" is one synchronous replanning step of the whole swarm (every agent replans once), closed loop: the next
step's states are the new trajectories evaluated at t = dt. Default workload = BASELINE.json configs[4]'s swarm
(synthetic circle-swap, 1024 agents, simple_forest.bt at the centre), which fits one GPU; with --gpus N the same
swarm is dealt out over N ranks in longest-plan-first order (strong scaling) and every step ends with one NCCL
all-gather; rank 0 then replays the run on a single engine and asserts identical trajectories.
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


def alg_flops_per_replan(n_agents: int, octomap: bool, gjk_iters: float, qp_iters: float):
    """F_lsc = (N-1) M (90 + 46 I_gjk) [FP64 GJK + float margins], F_qp = (K+1)(2 nnz(A) + 1080 + 2*39^2)
    + 2/3 39^3; I_gjk (per hull) and K come from the ORACLE's counters on the same inputs."""
    hulls = (n_agents - 1) * 5
    f_lsc = hulls * 90 + 46.0 * gjk_iters          # gjk_iters = total GJK iterations of the agent's hulls
    nnz = 81 * (n_agents - 1) + 618 + (162 if octomap else 0) + 174
    f_qp = (qp_iters + 1) * (2 * nnz + 1080 + 2 * 39 ** 2) + (2.0 / 3.0) * 39 ** 3
    return f_lsc, f_qp


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
        # nvidia-smi takes ~100 ms to come up and holds driver locks meanwhile: wait for its first sample, so that the
        # start-up does not land inside a timed region that is only a few milliseconds long
        t0 = time.perf_counter()
        while self.proc is not None and time.perf_counter() - t0 < 3.0:
            try:
                if os.path.getsize(self.path) > 0:
                    break
            except OSError:
                pass
            time.sleep(0.02)

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
    for _ in range(args.preroll):           # untimed: bring the swarm to the same mission phase as the GPU arm
        sw.step(0, n, threads); sw.advance()
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
        "run": {"preroll_steps": args.preroll},
        "cpu_baseline": {"value": value, "unit": "agent-replans/s", "cores": threads, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": "agent-replans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "oracle_counters_per_replan": {k: v / max(sample * args.steps, 1) for k, v in c.items()},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
def _traj_digest(traj: np.ndarray) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(traj).tobytes()).hexdigest()[:16]


def run_ours(args):
    import torch
    import torch.distributed as dist

    import lsc_planner_b200 as L
    from lsc_planner_b200 import _capi as A
    from lsc_planner_b200 import engine as E

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
    prm = L.Param(world_min=scn.world_min, world_max=scn.world_max, world_use_octomap=scn.use_octomap,
                  goal_mode=GOAL_MODES[args.goal_mode])

    def new_engine():
        e = L.ReplanEngine(n, prm, scn.agents, device=local)
        if scn.use_octomap:
            e.set_octomap_file(bt)
        return e

    eng = new_engine()
    if world > 1:
        from lsc_planner_b200 import sharding
        sharding.connect(eng, rank, world)
    n_local = eng.n_planned
    stream = torch.cuda.ExternalStream(eng.stream, device=local)

    # every step is timed on its own (begin/end events on the engine stream) with a 256 MB write flushing the 126 MB
    # L2 in between: the data one step touches (swarm state, distance-field tables) is smaller than L2
    flush_buf = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_steps(k, flush=True):
        for _ in range(k):
            if flush:
                with torch.cuda.stream(stream):
                    flush_buf.zero_()
            eng.replan_resident(1, sync=False)

    def restart():
        eng.reset(); eng.set_states(scn.start); eng.set_goals(scn.goal)
        resident_steps(args.preroll, flush=False)       # untimed closed-loop steps: choose the mission phase
        eng.synchronize()

    # ---- FMA peaks of THIS device, in this run (not in MEASURED_PEAKS.json) ----------------------
    fp = E.measure_fma_peaks(local)
    lat = E.measure_latencies(local)

    # ---- per-kernel breakdown: the same closed-loop steps with events between the kernels (profiling mode) --------
    restart()
    eng.set_profiling(True)
    resident_steps(args.warmup)
    eng.synchronize()
    resident_steps(args.steps)
    eng.synchronize()
    st = eng.step_stats()
    prof_out = eng.fetch().copy()

    # ---- device-resident throughput (`value`) -----------------------------------------------------------------
    eng.set_profiling(False)
    clocks = ClockSampler(local)          # samples every 100 ms through the timed pass and the e2e pass
    restart()
    resident_steps(args.warmup)
    eng.synchronize()
    barrier()
    resident_steps(args.steps)
    eng.synchronize()
    barrier()
    st_timed = eng.step_stats()
    ms_region = st_timed["ms_steps"]             # sum of the per-step (begin, end) event pairs: the L2 flushes are outside
    t = torch.tensor([ms_region], dtype=torch.float64, device=f"cuda:{local}")
    launches = torch.tensor([st_timed["kernel_launches"]], dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    out_mid = eng.fetch().copy()
    boxes_mid = eng.get_sfc()[0] if scn.use_octomap else None
    assert (out_mid["report"] == 5).all()
    assert sorted(out_mid["agent_id"].tolist()) == list(range(n)), "every agent must have been planned exactly once"
    qp_fail = int((out_mid["qp_status"] != 0).sum())
    value = n * args.steps / (ms_total * 1e-3)
    total_steps = args.preroll + args.warmup + args.steps

    # ---- multi-GPU correctness: every replica holds the same swarm, and it is the swarm ONE engine computes --------
    sharded_check = None
    if world > 1:
        mine = torch.from_numpy(out_mid["traj"].copy()).cuda()
        ref = mine.clone(); dist.broadcast(ref, src=0)
        same = torch.tensor([int(torch.equal(mine, ref))], device=f"cuda:{local}")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        single_digest = None
        if rank == 0:
            single = new_engine()
            single.set_states(scn.start); single.set_goals(scn.goal)
            single.replan_resident(total_steps)
            so = single.fetch()
            single_digest = _traj_digest(so["traj"])
            status_equal = bool(np.array_equal(so["qp_status"], out_mid["qp_status"]))
            single.close()
            sharded_check = {"replicas_identical": bool(same.item()), "single_engine_digest": single_digest,
                             "sharded_digest": _traj_digest(out_mid["traj"]), "qp_status_equal": status_equal,
                             "steps_replayed": total_steps}
            assert sharded_check["replicas_identical"], "replicas diverged"
            assert sharded_check["single_engine_digest"] == sharded_check["sharded_digest"] and status_equal, \
                f"sharded run differs from the single-engine replay: {sharded_check}"

    # ---- end to end through the C-ABI with pinned HOST buffers (`e2e`) ----------------------------------------
    restart()
    st0 = eng.fetch().copy() if args.preroll else None
    pin_in = torch.zeros(n * A.AGENT_IN.itemsize, dtype=torch.uint8).pin_memory()
    pin_out = torch.zeros(n * A.AGENT_OUT.itemsize, dtype=torch.uint8).pin_memory()
    h_in = pin_in.numpy().view(A.AGENT_IN); h_out = pin_out.numpy().view(A.AGENT_OUT)
    if st0 is None:
        h_in["position"] = scn.start; h_in["velocity"] = 0; h_in["acceleration"] = 0
    else:
        h_in["position"] = st0["next_position"]; h_in["velocity"] = st0["next_velocity"]; h_in["acceleration"] = st0["next_acceleration"]
    h_in["goal"] = scn.goal

    def host_step():
        eng.replan_ptr(pin_in.data_ptr(), pin_out.data_ptr())      # H2D + kernels (+ all-gather) + D2H, synchronous
        eng.advance_inputs_ptr(pin_out.data_ptr(), pin_in.data_ptr())   # the host closes the loop: next state = plan at t = dt

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
    e2e_device_ms = float(eng.step_stats()["ms_steps"])      # device time of the last end-to-end step (no L2 flush before it)
    clk = clocks.stop()
    e2e_matches_resident = bool(np.array_equal(h_out["traj"], out_mid["traj"]))

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    pk, pk_kind = peaks()
    n_steps_prof = max(st["steps"], 1)
    per_kernel = {"k_agent_plan": st["ms_plan"] / n_steps_prof, "k_predict": st["ms_predict"] / n_steps_prof,
                  "k_commit": st["ms_commit"] / n_steps_prof, "nccl_all_gather": st["ms_exchange"] / n_steps_prof}
    dom = "k_agent_plan"
    dom_ms = per_kernel[dom]
    # split of the fused kernel per agent, from the cycle counters of the result records (last profiled step)
    cyc = {"corridors_lsc_sfc": float(prof_out["lsc_kcycles"].mean()),
           "qp": float((prof_out["qp_kcycles"] - prof_out["lsc_kcycles"]).mean()),
           "qp_pricing": float(prof_out["qp_price_kcycles"].mean())}
    sm_khz = A.lib().lscgpu_sm_clock_khz(eng.h)
    per_agent_us = {k: v * 1024.0 / (sm_khz * 1e-3) for k, v in cyc.items()}

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, single-GPU runs only) -------------------
    cpu = None
    l_sfc = 0.0
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sw = oracle_swarm(scn, bt, GOAL_MODES[args.goal_mode])
        # same step as the end of the timed region: the oracle is loaded with the engine's planner state
        sw.set_state(out_mid["next_position"], out_mid["next_velocity"], out_mid["next_acceleration"])
        def restore():
            sw.set_traj(out_mid["traj"], total_steps)
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
               "sample": f"oracle (CPU port of the reference path, oracle/, -O3 -march=x86-64-v3) re-plans agents [0,{s}) of the "
                         f"{n}-agent swarm from the engine's state after the timed region, {reps}x, {threads} threads, {el:.1f} s",
               "one_thread": {"value": one_thread, "unit": "agent-replans/s", "sample": f"agents [0,{s1}) once, 1 thread"},
               "thread_scaling": (s_total / el) / max(one_thread, 1e-9),
               "per_replan": {"gjk_iterations": c["gjk_iters"] / s_total, "qp_iterations": c["qp_iters"] / s_total,
                              "edt_lookups": l_sfc, "qp_rows": c["qp_rows"] / s_total}}
    # Algorithmic bytes of one launch of the dominant kernel = SURVEY.md §8(d)'s per-agent-replan figure without its
    # `4 L_sfc` term (the reference's brute-force EDT sampling, which this kernel replaces by a few hundred reads of the
    # L2-resident summed-volume table): every neighbour's prediction + radius/downwash, own trajectory, state, goal, SFC
    # window in/out, the result record. x agents planned per launch / the kernel's mean duration in the profiled pass.
    kept_per_replan = st["lsc_pairs_kept"] / max(n_steps_prof * n_local, 1)
    bytes_per_replan = (n - 1) * 368 + 360 + 36 + 12 + 48 + A.AGENT_OUT.itemsize
    b_alg = bytes_per_replan * n_local
    achieved = b_alg / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    traffic_note = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        ent = json.load(open(tp)).get(f"{args.workload}_{n}", {}).get(dom)
        if isinstance(ent, dict):
            traffic, traffic_note = ent.get("bytes_per_launch"), ent.get("captured")
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "traffic_captured": traffic_note,
                "peak_source": f"MEASURED_PEAKS.json ({pk_kind})",
                "ms_per_launch": dom_ms, "algorithmic_bytes_per_launch": b_alg,
                "kept_pairs_per_replan": kept_per_replan,
                "note": "the swarm state is L2-resident (SURVEY.md §8d): the kernel is bound by FP64 issue (GJK) and by the "
                        "latency of each agent's active-set chain, see fma_roofline and DESIGN.md §5"}
    step_ms = ms_total / args.steps
    b_path = bytes_per_replan * n_local
    path_roofline = {"algorithmic_bytes_per_step": b_path, "ms_per_step": step_ms,
                     "achieved": b_path / (step_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": b_path / (step_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                     "note": "SURVEY.md §8(d) B_alg without the 4 L_sfc term, over the whole step's device time; every neighbour "
                             "read counted as if from HBM (upper bound on necessary traffic)"}
    fma = None
    if cpu is not None:
        f_lsc, f_qp = alg_flops_per_replan(n, scn.use_octomap, cpu["per_replan"]["gjk_iterations"], cpu["per_replan"]["qp_iterations"])
        fma = {"algorithmic_flops_per_replan": {"lsc": f_lsc, "qp": f_qp},
               "achieved_tflops": (f_lsc + f_qp) * value * 1e-12,
               "fp64_peak_tflops": fp["fp64_tflops"], "fp32_peak_tflops": fp["fp32_tflops"],
               "peak_source": "lscgpu_measure_fma_peaks, measured in this run on this device",
               "frac_of_fp64_peak": (f_lsc + f_qp) * value * 1e-12 / max(fp["fp64_tflops"], 1e-9),
               "note": "SURVEY.md §8d F_alg: the ORACLE's counters on the step after the timed region (un-culled row set, "
                       "full pricing passes) — work the kernel partly avoids, so this overstates its FMA utilisation"}
    if rank == 0:
        line = {
            "metric": "agent-replans/sec", "value": value, "unit": "agent-replans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}_{n}", "agents": n, "octomap": bool(scn.use_octomap)},
            "run": {"agents_per_gpu": n_local, "preroll_steps": args.preroll,
                    "parallelism": f"LPT order of the swarm dealt round-robin over {world} rank(s); one NCCL all-gather of "
                                   f"{A.AGENT_OUT.itemsize} B records per step" if world > 1 else "one engine plans every agent",
                    "l2": "L2 flushed (256 MB write) between steps; every step timed on its own with CUDA events on the "
                          "engine stream; e2e is not flushed (its inputs arrive from host memory every step)",
                    "goals": (("prior_based goal planning on the GPU every step, inside the timed region (SURVEY.md §8f #1): "
                               + ("k_goal_astar = priority rule + occupancy grid + A* + line-of-sight goal" if scn.use_octomap
                                  else "k_goal_plan = priority rule + clip (no octomap)")) if GOAL_MODES[args.goal_mode]
                              else "fixed to the mission goals (goal planning is outside the path, SURVEY.md §8f)")},
            "e2e": {"value": e2e_value, "unit": "agent-replans/s", "h2d_bytes_per_step": n * A.AGENT_IN.itemsize,
                    "d2h_bytes_per_step": n * A.AGENT_OUT.itemsize, "same_trajectories_as_resident_pass": e2e_matches_resident,
                    "ms_per_step": 1e3 * float(e2e_s.item()) / args.steps, "device_ms_last_step": e2e_device_ms,
                    "results": "written by the planning blocks straight into the pinned result array (no copy after the step)"
                               if world == 1 else "one device-to-host copy after the step"},
            "gpu_launches": int(launches.item()),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "roofline": roofline, "path_roofline": path_roofline,
            "fma_roofline": fma,
            "cpu_baseline": cpu,
            "kernel_ms_per_step": per_kernel,
            "k_agent_plan_per_agent_us": per_agent_us,
            "latency_cycles": lat,
            "goal_planning": {"astar_expansions_per_replan": st["astar_expansions"] / max(n_local * n_steps_prof, 1)},
            "qp": {"iterations_per_replan": st["qp_iterations"] / max(n_local * n_steps_prof, 1),
                   "rows_priced_per_replan": st["qp_rows_priced"] / max(n_local * n_steps_prof, 1),
                   "pricing_passes_per_replan": st["qp_full_passes"] / max(n_local * n_steps_prof, 1),
                   "gjk_iterations_per_hull": st["gjk_iterations"] / max(st["lsc_pairs"], 1),
                   "warm_start": {"tried_fraction": st["qp_warm_tried"] / max(n_local * n_steps_prof, 1),
                                  "accepted_of_tried": st["qp_warm_accepted"] / max(st["qp_warm_tried"], 1),
                                  "rows_per_accepted": st["qp_warm_rows"] / max(st["qp_warm_accepted"], 1)},
                   "failed_last_step": qp_fail, "qp_failed_fraction": qp_fail / n},
            "sharded_check": sharded_check,
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
    ap.add_argument("--preroll", type=int, default=0,
                    help="untimed closed-loop steps before the warm-up (e.g. 80: the contact phase of the circle swap)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""ctypes declarations of include/lscgpu.h (liblscgpu.so). No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LSCGPU_LIB: an instrumented build of the same library (tools/gpu_diag.py with -DLSCGPU_QP_SECTION_TIMERS); never a fallback
LIB_PATH = os.environ.get("LSCGPU_LIB") or os.path.join(HERE, "liblscgpu.so")

OK = 0
QP_OK, QP_INFEASIBLE, QP_MAXITER = 0, 1, 2
REPORT_SUCCESS = 5
FLAG_SLACK_NEEDED, FLAG_SFC_SEED_BLOCKED = 1, 2


class Params(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("control_input_weight", C.c_double), ("terminal_weight", C.c_double),
        ("world_resolution", C.c_double), ("reset_threshold", C.c_double), ("world_use_octomap", C.c_int),
        ("world_min", C.c_float * 3), ("world_max", C.c_float * 3),
        ("M", C.c_int), ("n", C.c_int), ("phi", C.c_int), ("dim", C.c_int),
        ("goal_mode", C.c_int), ("goal_threshold", C.c_double), ("goal_radius", C.c_double),
        ("priority_dist_threshold", C.c_double), ("grid_resolution", C.c_double), ("grid_margin", C.c_double),
    ]


class AgentConst(C.Structure):
    _fields_ = [("radius", C.c_double), ("downwash", C.c_double), ("nominal_velocity", C.c_double),
                ("max_vel", C.c_double * 3), ("max_acc", C.c_double * 3)]


class StepStats(C.Structure):
    _fields_ = [("steps", C.c_int32), ("ms_total", C.c_float), ("ms_predict", C.c_float), ("ms_plan", C.c_float),
                ("ms_sfc", C.c_float), ("ms_reserved1_", C.c_float), ("ms_exchange", C.c_float), ("ms_commit", C.c_float),
                ("kernel_launches", C.c_int32), ("lsc_pairs", C.c_int64), ("lsc_pairs_kept", C.c_int64), ("gjk_iterations", C.c_int64),
                ("qp_rows_priced", C.c_int64), ("qp_iterations", C.c_int64), ("qp_full_passes", C.c_int64),
                ("ms_steps", C.c_float), ("reserved_", C.c_float),
                ("qp_warm_tried", C.c_int64), ("qp_warm_accepted", C.c_int64), ("qp_warm_rows", C.c_int64),
                ("astar_expansions", C.c_int64)]


# numpy views of lscgpu_agent_in / lscgpu_agent_out (C layout, natural alignment)
AGENT_IN = np.dtype([("position", np.float32, 3), ("velocity", np.float32, 3), ("acceleration", np.float32, 3),
                     ("goal", np.float32, 3)], align=True)
AGENT_OUT = np.dtype([("traj", np.float32, (5, 6, 3)), ("next_position", np.float32, 3),
                      ("next_velocity", np.float32, 3), ("next_acceleration", np.float32, 3),
                      ("agent_id", np.int32), ("qp_cost", np.float64), ("report", np.int32), ("qp_status", np.int32),
                      ("qp_iterations", np.int32), ("qp_active", np.int32), ("flags", np.int32),
                      ("terminal_segments", np.int32), ("qp_sweeps", np.int32), ("qp_kcycles", np.int32),
                      ("qp_price_kcycles", np.int32), ("lsc_pairs_kept", np.int32),
                      ("current_goal", np.float32, 3), ("goal_kind", np.int32), ("sfc_box", np.float32, 6),
                      ("lsc_kcycles", np.int32), ("sfc_in_block", np.int32)], align=True)
assert AGENT_IN.itemsize == 48 and AGENT_OUT.itemsize == 496, (AGENT_IN.itemsize, AGENT_OUT.itemsize)

_lib = None
ptr = C.c_void_p


def symbols():
    """Every entry point include/lscgpu.h declares."""
    return ["lscgpu_last_error", "lscgpu_version", "lscgpu_create", "lscgpu_destroy", "lscgpu_set_octomap_file",
            "lscgpu_set_octomap_voxels", "lscgpu_get_distmap_info", "lscgpu_get_distmap_sqdist", "lscgpu_set_shard",
            "lscgpu_nccl_unique_id", "lscgpu_nccl_init", "lscgpu_p2p_export", "lscgpu_p2p_attach", "lscgpu_replan_batch",
            "lscgpu_advance_inputs",            "lscgpu_safety_audit", "lscgpu_set_goals",
            "lscgpu_set_states", "lscgpu_replan_resident", "lscgpu_synchronize", "lscgpu_fetch", "lscgpu_reset", "lscgpu_set_prev_traj",
            "lscgpu_set_sfc", "lscgpu_get_sfc", "lscgpu_get_planner_seq", "lscgpu_get_lsc", "lscgpu_get_lsc_ex",
            "lscgpu_set_slack_collision_weight", "lscgpu_get_reset_state", "lscgpu_set_reset_state",
            "lscgpu_set_capture_rows", "lscgpu_dump_qp_lp", "lscgpu_set_lp_dump_dir",
            "lscgpu_get_initial_traj", "lscgpu_qp_solve_batch", "lscgpu_qp_solve_batch_slack", "lscgpu_gjk_batch",
            "lscgpu_sfc_expand_batch",
            "lscgpu_get_step_stats", "lscgpu_set_profiling", "lscgpu_sm_clock_khz", "lscgpu_stream",
            "lscgpu_measure_fma_peaks", "lscgpu_measure_latencies"]


def lib():
    """Loads liblscgpu.so. Raises if it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m lsc_planner_b200.build` "
                           "(the engine has no CPU or pure-Python path)")
    L = C.CDLL(LIB_PATH)
    L.lscgpu_last_error.restype = C.c_char_p
    L.lscgpu_version.restype = C.c_int
    L.lscgpu_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(AgentConst), C.c_int, C.POINTER(ptr)]
    L.lscgpu_destroy.argtypes = [ptr]; L.lscgpu_destroy.restype = None
    L.lscgpu_set_octomap_file.argtypes = [ptr, C.c_char_p]
    L.lscgpu_set_octomap_voxels.argtypes = [ptr, ptr, C.c_int]
    L.lscgpu_get_distmap_info.argtypes = [ptr, ptr, ptr, ptr]
    L.lscgpu_get_distmap_sqdist.argtypes = [ptr, ptr]
    L.lscgpu_set_shard.argtypes = [ptr, C.c_int, C.c_int]
    L.lscgpu_nccl_unique_id.argtypes = [ptr]
    L.lscgpu_nccl_init.argtypes = [ptr, ptr, C.c_int, C.c_int]
    L.lscgpu_p2p_export.argtypes = [ptr, ptr]
    L.lscgpu_p2p_attach.argtypes = [ptr, ptr]
    L.lscgpu_replan_batch.argtypes = [ptr, ptr, ptr]
    L.lscgpu_advance_inputs.argtypes = [ptr, ptr, C.c_int]
    L.lscgpu_safety_audit.argtypes = [ptr, C.c_double, C.c_double, ptr, ptr]
    L.lscgpu_set_goals.argtypes = [ptr, ptr]
    L.lscgpu_set_states.argtypes = [ptr, ptr, ptr, ptr]
    L.lscgpu_replan_resident.argtypes = [ptr]
    L.lscgpu_synchronize.argtypes = [ptr]
    L.lscgpu_fetch.argtypes = [ptr, ptr]
    L.lscgpu_reset.argtypes = [ptr]
    L.lscgpu_set_prev_traj.argtypes = [ptr, ptr, C.c_int]
    L.lscgpu_set_sfc.argtypes = [ptr, ptr, ptr]
    L.lscgpu_get_sfc.argtypes = [ptr, ptr, ptr]
    L.lscgpu_get_planner_seq.argtypes = [ptr]
    L.lscgpu_set_slack_collision_weight.argtypes = [ptr, C.c_double]
    L.lscgpu_get_reset_state.argtypes = [ptr, ptr]
    L.lscgpu_set_reset_state.argtypes = [ptr, ptr]
    L.lscgpu_get_lsc.argtypes = [ptr, C.c_int, ptr, ptr]
    L.lscgpu_get_lsc_ex.argtypes = [ptr, C.c_int, ptr, ptr, ptr]
    L.lscgpu_set_capture_rows.argtypes = [ptr, C.c_int]
    L.lscgpu_dump_qp_lp.argtypes = [ptr, C.c_int, C.c_char_p]
    L.lscgpu_set_lp_dump_dir.argtypes = [ptr, C.c_char_p]
    L.lscgpu_get_initial_traj.argtypes = [ptr, ptr]
    L.lscgpu_qp_solve_batch.argtypes = [ptr, C.c_int] + [ptr] * 12
    L.lscgpu_qp_solve_batch_slack.argtypes = [ptr, C.c_int] + [ptr] * 14
    L.lscgpu_gjk_batch.argtypes = [ptr, C.c_int, ptr, ptr, ptr]
    L.lscgpu_sfc_expand_batch.argtypes = [ptr, C.c_int, ptr, ptr, ptr, ptr, ptr]
    L.lscgpu_get_step_stats.argtypes = [ptr, C.POINTER(StepStats)]
    L.lscgpu_set_profiling.argtypes = [ptr, C.c_int]
    L.lscgpu_sm_clock_khz.argtypes = [ptr]
    L.lscgpu_measure_fma_peaks.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.lscgpu_measure_latencies.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.lscgpu_stream.argtypes = [ptr]; L.lscgpu_stream.restype = ptr
    _lib = L
    return L


class EngineError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"lscgpu error {code}: {what}")
        self.code = code


def check(rc: int):
    if rc != OK:
        raise EngineError(rc, lib().lscgpu_last_error().decode(errors="replace"))


def p(a: np.ndarray):
    """Pointer to a C-contiguous numpy array."""
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ptr)

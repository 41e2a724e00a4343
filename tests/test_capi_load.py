"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol include/lscgpu.h declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lsc_planner_b200
from lsc_planner_b200 import _capi as A
from lsc_planner_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build()
    return A.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "lscgpu.h")).read()
    declared = set(re.findall(r"\b(lscgpu_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(A.symbols()), declared ^ set(A.symbols())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lscgpu_version() >= 100


def test_struct_layouts_match_header():
    assert A.AGENT_IN.itemsize == 48
    assert A.AGENT_OUT.itemsize == 496 and A.AGENT_OUT.fields["qp_cost"][1] == 400
    assert C.sizeof(A.Params) == 128 and C.sizeof(A.AgentConst) == 72


def test_no_cpu_fallback(lib):
    """Without a CUDA device lscgpu_create must fail loudly (LSCGPU_ERR_CUDA), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(A.EngineError) as ei:
        lsc_planner_b200.ReplanEngine(4)
    assert ei.value.code == -2


def test_argument_validation(lib):
    p = lsc_planner_b200.Param(M=4).to_c()
    arr = (A.AgentConst * 1)()
    h = A.ptr()
    assert lib.lscgpu_create(C.byref(p), 1, arr, 0, C.byref(h)) == -1       # LSCGPU_ERR_ARG: only M=5,n=5,phi=3
    assert b"M=5" in lib.lscgpu_last_error()
    assert lib.lscgpu_create(None, 1, arr, 0, C.byref(h)) == -1


def test_scenarios_are_deterministic():
    s1 = lsc_planner_b200.scenarios.circle_swap(256)
    s2 = lsc_planner_b200.scenarios.circle_swap(256)
    assert s1.n == 256 and np.array_equal(s1.start, s2.start)
    assert np.allclose(s1.goal[:, :2], -s1.start[:, :2]) and np.allclose(s1.goal[:, 2], 1.0)
    # ring k has floor(2 pi (4 + k) / 0.8) agents
    r = np.hypot(s1.start[:, 0], s1.start[:, 1])
    assert (np.abs(r - 4.0) < 1e-5).sum() == 31 and (np.abs(r - 5.0) < 1e-5).sum() == 39
    d = np.linalg.norm(s1.start[:, None] - s1.start[None], axis=-1) + np.eye(256) * 9
    assert d.min() > 0.6
    m = lsc_planner_b200.scenarios.load_mission(os.path.join(ROOT, "tests", "golden", "missions", "multi_simple3.json"))
    assert m.n == 3 and m.agents[0].radius == 0.15


def test_advance_inputs_is_the_state_hand_over(lib):
    """lscgpu_advance_inputs (host only, no device needed): in[a].position / velocity / acceleration = out[a].next_*, goals
    untouched (MultiSyncSimulator::update, src/multi_sync_simulator.cpp:203)."""
    rng = np.random.default_rng(0)
    n = 7
    out = np.zeros(n, A.AGENT_OUT); inp = np.zeros(n, A.AGENT_IN)
    for f in ("next_position", "next_velocity", "next_acceleration"):
        out[f] = rng.standard_normal((n, 3)).astype(np.float32)
    inp["goal"] = rng.standard_normal((n, 3)).astype(np.float32)
    inp["position"] = 9.0
    goal = inp["goal"].copy()
    assert lib.lscgpu_advance_inputs(A.p(out), A.p(inp), n) == 0
    assert np.array_equal(inp["position"], out["next_position"]) and np.array_equal(inp["velocity"], out["next_velocity"])
    assert np.array_equal(inp["acceleration"], out["next_acceleration"]) and np.array_equal(inp["goal"], goal)
    assert lib.lscgpu_advance_inputs(None, A.p(inp), n) == -1          # LSCGPU_ERR_ARG

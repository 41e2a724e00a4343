"""lsc_planner_b200 — B200-native replanning engine for the LSC swarm planner's inner loop.

csrc/    CUDA kernels (sm_100a) + the C-ABI (include/lscgpu.h) -> liblscgpu.so
host/    C++ host side mirroring the reference's TrajOptimizer / TrajPlanner / MultiSyncSimulator
engine.py  ctypes binding used by tests and bench
"""
from .engine import AgentType, Param, ReplanEngine  # noqa: F401
from . import scenarios  # noqa: F401

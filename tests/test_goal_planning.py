"""Goal planning, the step before the hot path: the oracle's restatement and the product's host-side
planner (lsc_planner_b200/host/grid_based_planner.hpp) against the REFERENCE's own A* (golden vectors written by
tools/make_golden.py from oracle/_ref/libref_astar.so = <ref>/src/Astar-3D compiled as is; compared live as well when
that library is present), and against each other on whole goal-planning problems. Bit-exact: integer cell paths and
float32 goal points."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "lsc_planner_b200", "host")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def host():
    subprocess.run(["make", "-C", HOST, "libhostgoal.so"], check=True, capture_output=True)
    H = C.CDLL(os.path.join(HOST, "libhostgoal.so"))
    H.host_astar.argtypes = [i32p, u8p, i32p, i32p, i32p, C.c_int, C.POINTER(C.c_longlong)]; H.host_astar.restype = C.c_int
    H.host_goal_plan.argtypes = [C.c_int, C.c_int, f32p, f32p, f32p, f32p, f64p, f64p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_double, f32p, f32p] + [C.c_double] * 5 + [f32p, C.POINTER(C.c_longlong)]
    H.host_goal_plan.restype = C.c_int
    H.devcore_astar.argtypes = H.host_astar.argtypes + [C.c_int]; H.devcore_astar.restype = C.c_int
    return H


def host_astar(H, grid, start, goal):
    grid = np.ascontiguousarray(grid, np.uint8)
    path = np.zeros((4096, 3), np.int32); ex = C.c_longlong(0)
    n = H.host_astar(np.asarray(grid.shape, np.int32), grid, np.asarray(start, np.int32), np.asarray(goal, np.int32), path,
                     len(path), C.byref(ex))
    return path[:n].copy(), ex.value


def devcore_astar(H, grid, start, goal, bits=32):
    """The search core of the GPU kernel (csrc/astar_core.cuh) compiled for the host, with 16- or 32-bit indices."""
    grid = np.ascontiguousarray(grid, np.uint8)
    path = np.zeros((4096, 3), np.int32); ex = C.c_longlong(0)
    n = H.devcore_astar(np.asarray(grid.shape, np.int32), grid, np.asarray(start, np.int32), np.asarray(goal, np.int32), path,
                        len(path), C.byref(ex), bits)
    assert n >= 0
    return path[:n].copy(), ex.value


def golden_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "astar_ref_vectors.npz"))
    for k in range(int(g["count"])):
        dim = tuple(int(v) for v in g[f"dim{k}"])
        grid = np.unpackbits(g[f"grid{k}"])[:dim[0] * dim[1] * dim[2]].reshape(dim)
        yield grid, g[f"start{k}"], g[f"goal{k}"], g[f"path{k}"].astype(np.int32)


def test_oracle_astar_matches_reference_golden(golden_dir):
    n = found = 0
    for grid, s, g, ref_path in golden_cases(golden_dir):
        path, _ = O.astar(grid, s, g)
        assert path.shape == ref_path.shape and (path == ref_path).all(), (grid.shape, s, g)
        n += 1; found += len(ref_path) > 0
    assert n >= 100 and found >= 60


def test_host_astar_matches_reference_golden(golden_dir, host):
    for grid, s, g, ref_path in golden_cases(golden_dir):
        path, _ = host_astar(host, grid, s, g)
        assert path.shape == ref_path.shape and (path == ref_path).all(), (grid.shape, s, g)


def test_device_astar_core_matches_reference_golden_and_host(golden_dir, host):
    """k_goal_astar's search (flat scratch arrays, packed cell bytes, hash-order model without containers) path for path
    against the reference's goldens, and against the host planner incl. the expansion count on random grids (walls force
    long detours, large rows force several rehashes of the row containers)."""
    for grid, s, g, ref_path in golden_cases(golden_dir):
        for bits in (16, 32):
            path, _ = devcore_astar(host, grid, s, g, bits)
            assert path.shape == ref_path.shape and (path == ref_path).all(), (grid.shape, s, g, bits)
    rng = np.random.default_rng(7)
    for t in range(150):
        dim = (int(rng.integers(3, 60)), int(rng.integers(3, 60)), int(rng.integers(1, 13)))
        grid = np.ascontiguousarray((rng.random(dim) < rng.choice([0.0, 0.1, 0.25, 0.35])).astype(np.uint8))
        if t % 4 == 0 and dim[0] > 8:
            grid[dim[0] // 2, :, :] = 1; grid[dim[0] // 2, int(rng.integers(0, dim[1])), :] = 0     # a wall with one gap
        s = np.array([rng.integers(0, d) for d in dim], np.int32); g = np.array([rng.integers(0, d) for d in dim], np.int32)
        grid[tuple(s)] = 0
        ph, eh = host_astar(host, grid, s, g)
        for bits in (16, 32):
            pd, ed = devcore_astar(host, grid, s, g, bits)
            assert ph.shape == pd.shape and (ph == pd).all() and eh == ed, (dim, s, g, bits)


def test_astar_live_against_reference_library(host):
    so = os.path.join(ROOT, "oracle", "_ref", "libref_astar.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_astar.so not built (reference tree absent)")
    R = C.CDLL(so)
    R.ref_astar_plan.argtypes = [i32p, u8p, i32p, i32p, i32p, C.c_int]; R.ref_astar_plan.restype = C.c_int
    rng = np.random.default_rng(99)
    for t in range(120):
        dim = (int(rng.integers(4, 50)), int(rng.integers(4, 50)), int(rng.integers(1, 12)))
        grid = np.ascontiguousarray((rng.random(dim) < rng.choice([0.0, 0.1, 0.25])).astype(np.uint8))
        s = np.array([rng.integers(0, d) for d in dim], np.int32); g = np.array([rng.integers(0, d) for d in dim], np.int32)
        if t % 3 == 0:
            g[1] = s[1]; g[2] = s[2]
        grid[tuple(s)] = 0
        ref = np.zeros((4096, 3), np.int32)
        n = R.ref_astar_plan(np.asarray(dim, np.int32), grid, s, g, ref, len(ref))
        po, _ = O.astar(grid, s, g); ph, _ = host_astar(host, grid, s, g)
        assert len(po) == n == len(ph) and (po == ref[:n]).all() and (ph == ref[:n]).all(), (dim, s, g)


def test_hash_order_model_matches_unordered_map(host):
    """The product's A* reproduces the reference's tie-breaking through an explicit model of libstdc++'s unordered_map
    node order (host/grid_based_planner.hpp); here that model against the real container, operation by operation."""
    host.host_hash_order_check.argtypes = [C.c_uint, C.c_int, C.c_int]; host.host_hash_order_check.restype = C.c_int
    for seed, key_range, ops in [(1, 8, 2000), (2, 40, 5000), (3, 451, 20000), (4, 1771, 40000), (5, 5000, 60000)]:
        assert host.host_hash_order_check(seed, key_range, ops) == 0, (seed, key_range)


def test_bucket_sequence_is_the_containers(host):
    """The bucket counts the engine records from libstdc++'s rehash policy (goal_bucket_sequence in csrc/engine.cu, same
    code as devcore_bucket_sequence) are the ones a real std::unordered_map goes through while it grows one element at a
    time — the device A* sizes and re-threads its row containers by them."""
    host.devcore_bucket_sequence.argtypes = [C.c_int, i32p, C.c_int]; host.devcore_bucket_sequence.restype = C.c_int
    host.host_unordered_map_bucket_counts.argtypes = [C.c_int, i32p, C.c_int]; host.host_unordered_map_bucket_counts.restype = C.c_int
    seq = np.zeros(16, np.int32); real = np.zeros(32, np.int32)
    n_seq = host.devcore_bucket_sequence(1771, seq, 16)              # the longest row of the 40 m world: 161 x 11 cells
    n_real = host.host_unordered_map_bucket_counts(int(seq[n_seq - 1]), real, 32)
    assert seq[0] == 1 and seq[n_seq - 1] >= 1771
    assert n_real >= n_seq and (real[:n_seq] == seq[:n_seq]).all(), (seq[:n_seq], real[:n_real])


def test_astar_properties(host):
    """6-connected unit steps, free cells only, shortest length in an empty grid, goal test ignores the altitude
    (src/Astar-3D/isearch.cpp:74), no path when the goal column is walled in."""
    grid = np.zeros((12, 9, 5), np.uint8)
    path, _ = O.astar(grid, [1, 2, 0], [10, 7, 4])
    assert (path[0] == [1, 2, 0]).all() and (path[-1][:2] == [10, 7]).all()
    assert 9 + 5 <= len(path) - 1 <= 9 + 5 + 4       # the goal test looks at the column only; the heuristic still climbs
    assert (np.abs(np.diff(path, axis=0)).sum(axis=1) == 1).all()
    grid[5, :, :] = 1
    assert len(O.astar(grid, [1, 2, 0], [10, 7, 4])[0]) == 0 and len(host_astar(host, grid, [1, 2, 0], [10, 7, 4])[0]) == 0
    grid[5, 4, 2] = 0
    path, _ = O.astar(grid, [1, 2, 0], [10, 7, 4])
    assert [5, 4, 2] in path.tolist() and all(grid[tuple(c)] == 0 for c in path)


def _swarm_case(rng, n, wmin, wmax, omap):
    lo = np.asarray(wmin, np.float32) + 0.6; hi = np.asarray(wmax, np.float32) - 0.6
    def free_points(k):
        pts = []
        while len(pts) < k:
            p = (lo + (hi - lo) * rng.random(3)).astype(np.float32)
            if omap is None or omap.distance(p) >= 0.4:
                pts.append(p)
        return np.array(pts, np.float32)
    pos = free_points(n); desired = free_points(n)
    prev = np.zeros((n, 30, 3), np.float32)
    for j in range(n):           # previous trajectory: a straight segment from pos toward a random direction
        d = rng.normal(size=3).astype(np.float32) * 0.3
        prev[j] = pos[j] + np.linspace(0, 1, 30, dtype=np.float32)[:, None] * d
    return pos, desired, prev


@pytest.mark.parametrize("with_map", [False, True])
def test_host_goal_planning_matches_oracle(golden_dir, host, with_map):
    wmin, wmax = [-5, -5, 0], [5, 5, 2.5]
    omap = O.Map.from_bt(os.path.join(golden_dir, "worlds", "simple_forest.bt"), wmin, wmax) if with_map else None
    sq = np.ascontiguousarray(np.minimum(omap.sqdist(), 255).astype(np.uint8)) if with_map else None
    rng = np.random.default_rng(5 + with_map)
    kinds = np.zeros(2, int); moved = 0
    for case in range(6):
        n = 12
        pos, desired, prev = _swarm_case(rng, n, wmin, wmax, omap)
        if case % 2 == 0:            # two agents closer than priority_dist_threshold: retreat branch
            pos[1] = pos[0] + np.float32([0.3, 0.05, 0.0]); desired[1] = pos[1] + np.float32([0.5, 0, 0])
            prev[0] = 0; prev[1] = 0     # no direction information: the same-direction exemption cannot apply
        if case == 3:
            desired[2] = pos[2]      # an agent already at its goal: everybody else has priority over it
        radius = np.full(n, 0.15); dw = np.full(n, 2.0)
        for a in range(n):
            init_end = prev[a, 29]
            go, ko, ex = O.goal_plan(a, pos, desired, prev, init_end, radius, dw, omap, wmin, wmax)
            gh = np.zeros(3, np.float32); exh = C.c_longlong(0)
            kh = host.host_goal_plan(a, n, pos, desired, np.ascontiguousarray(prev.reshape(n, 90)), np.ascontiguousarray(init_end),
                                     radius, dw, sq.ctypes.data if with_map else None,
                                     np.asarray(omap.size, np.int32).ctypes.data if with_map else None,
                                     np.asarray(omap.off, np.int32).ctypes.data if with_map else None, 0.1,
                                     np.asarray(wmin, np.float32), np.asarray(wmax, np.float32), 0.25, 0.1, 0.1, 2.0, 0.4, gh, C.byref(exh))
            assert kh == ko and (gh.view(np.uint32) == go.view(np.uint32)).all(), (case, a, gh, go)
            kinds[ko] += 1
            moved += not np.array_equal(go, desired[a])
            assert np.linalg.norm(go - init_end) <= 2.0 + 1e-5 or ko == 1
            if not with_map and ko == 0:     # no octomap: the line-of-sight goal is the desired goal, clipped to goal_radius
                d = desired[a] - init_end
                exp = desired[a] if np.linalg.norm(d) <= 2.0 else None
                if exp is not None:
                    assert np.array_equal(go, exp)
    assert kinds[1] >= 3 and kinds[0] >= 30 and moved >= 20

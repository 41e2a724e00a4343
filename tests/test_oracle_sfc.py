"""Oracle .bt reader, distance map and corridor expansion (oracle/edt.hpp)."""
import os

import numpy as np

import oracle_lib as O

WMIN, WMAX = [-5, -5, 0], [5, 5, 2.5]


def _forest(golden_dir):
    return O.Map.from_bt(os.path.join(golden_dir, "worlds", "simple_forest.bt"), WMIN, WMAX)


def test_bt_decode(golden_dir):
    m = _forest(golden_dir)
    g = np.load(os.path.join(golden_dir, "simple_forest_voxels.npz"))
    assert m.n_nodes == int(g["n_nodes"]) == 15949 and m.n_occ == 4384
    assert np.array_equal(m.occupied(), g["keys"])
    assert list(m.size) == [101, 101, 26] and list(m.off) == [-50, -50, 0]  # inclusive key range (App. C.2)


def test_edt_exact_and_clamped(golden_dir):
    m = _forest(golden_dir)
    sq = m.sqdist()
    assert sq.max() == 121 and (sq == 0).sum() == 4384
    pts = m.occupied() - m.off
    rng = np.random.default_rng(0)
    for _ in range(400):
        c = rng.integers(0, m.size)
        assert sq[tuple(c)] == min(121, int(((pts - c) ** 2).sum(1).min()))
    assert m.distance([20, 0, 1]) == -1.0                                   # outside the map
    assert abs(m.distance([4.9, 4.9, 2.4]) - 1.1) < 1e-6                    # clamp = 11 cells


def _blocked(sq):
    return sq <= 3      # dist < 0.15 + 0.05 - 1e-5  <=>  integer squared cell distance <= 3


def test_sfc_box_invariants(golden_dir):
    m = _forest(golden_dir)
    sq = m.sqdist()
    rng = np.random.default_rng(1)
    n_ok = 0
    for _ in range(60):
        p = rng.uniform([-4.5, -4.5, 0.3], [4.5, 4.5, 2.2]).astype(np.float32)
        gl = rng.uniform([-4.5, -4.5, 0.3], [4.5, 4.5, 2.2]).astype(np.float32)
        ok, box, lookups = m.sfc_expand(p, gl)
        if not ok:
            continue
        n_ok += 1
        assert lookups > 0
        # contains the seed point (within the 0.01 snapping of expandBoxFromPoint)
        assert (p >= box[:3] - 0.0101).all() and (p <= box[3:] + 0.0101).all()
        # inside the world, on the lattice
        assert (box[:3] >= np.array(WMIN) - 1e-6).all() and (box[3:] <= np.array(WMAX) + 1e-6).all()
        assert np.abs(box / 0.1 - np.round(box / 0.1)).max() < 1e-4
        # every voxel strictly inside the box is free of inflated obstacles
        lo = np.round(box[:3] / 0.1).astype(int) - m.off
        hi = np.round(box[3:] / 0.1).astype(int) - m.off
        inner = _blocked(sq)[lo[0] + 1:max(hi[0], lo[0] + 1), lo[1] + 1:max(hi[1], lo[1] + 1), lo[2] + 1:max(hi[2], lo[2] + 1)]
        assert not inner.any()
    assert n_ok > 20


def test_sfc_empty_map_fills_world():
    m = O.Map.from_voxels(np.zeros((0, 3), np.int32), WMIN, WMAX)
    ok, box, _ = m.sfc_expand([1.0, -2.0, 1.0], [3, 3, 1])
    assert ok
    np.testing.assert_allclose(box, [-5, -5, 0, 5, 5, 2.5], atol=1e-5)
    ok, box, _ = m.sfc_expand([1.03, -2.0, 1.0], [3, 3, 1])                  # off-lattice seed
    np.testing.assert_allclose(box, [-5, -5, 0, 5, 5, 2.5], atol=1e-5)


def test_sfc_seed_in_obstacle(golden_dir):
    m = _forest(golden_dir)
    k = m.occupied()[100]
    p = (k + 0.5) * 0.1
    ok, _, _ = m.sfc_expand(p.astype(np.float32), [4, 4, 1])
    assert not ok       # reference throws std::invalid_argument (corridor_constructor.hpp:35-38)

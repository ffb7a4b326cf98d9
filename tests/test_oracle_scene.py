"""The numpy restatement of the whole-scene tiler / vote (oracle/scene_oracle.py) against outputs of the
unmodified reference class and add_vote (tests/golden/scene.npz, made by tests/golden/make_golden_scene.py).
The reference pads and shuffles every window with np.random, so the comparison is per window: same windows in
the same order, same member set, same number of blocks, padding drawn from the window's own points, and every
entry's 9 channels bit-equal to the oracle's formula."""
import numpy as np

import parity
from oracle import scene_oracle as so


def _golden():
    g = parity.load("scene.npz")
    return g, int(g["block_points"])


def test_windows_members_and_features_match_reference():
    g, bp = _golden()
    pts = g["points"].astype(np.float64)
    wins, grid, cmin, cmax = so.tile_windows(pts)
    blk = 0
    for w, s_x, s_y, idx in wins:
        nb = int(np.ceil(idx.size / bp))
        ref_idx = g["index"][blk:blk + nb].reshape(-1)
        assert set(ref_idx.tolist()) == set(idx.tolist()), f"window {w}: member set"
        feats = so.entry_features(pts, ref_idx, s_x, s_y, cmin, cmax)
        assert np.array_equal(feats.view(np.uint32), g["data"][blk:blk + nb].reshape(-1, 9).view(np.uint32)), f"window {w}: features"
        blk += nb
    assert blk == g["index"].shape[0], "number of blocks"


def test_oracle_tile_scene_has_the_reference_structure():
    g, bp = _golden()
    data, pidx, wid, grid = so.tile_scene(g["points"], block_points=bp)
    assert data.shape == g["data"].shape and pidx.shape == g["index"].shape
    # cyclic padding: every member of a window appears floor or ceil (entries / members) times
    for w in np.unique(wid):
        e = pidx[wid == w].reshape(-1)
        _, cnt = np.unique(e, return_counts=True)
        assert cnt.max() - cnt.min() <= 1


def test_vote_matches_reference():
    g, _ = _golden()
    pool = so.add_vote(np.zeros(g["pool"].shape), g["index"], g["pred"])
    assert np.array_equal(pool.astype(np.int32), g["pool"])
    assert np.array_equal(so.vote_argmax(pool).astype(np.uint8), g["labels"])


def test_product_window_bounds_equal_oracle_windows():
    """scene.window_bounds (host side of the CUDA tiler, float64) reproduces the reference's window starts / ends
    exactly, including the clamped last windows and scenes barely larger than one block."""
    from pointcloud_bridge_b200 import scene
    rng = np.random.default_rng(0)
    for ext in (6.3, 1.0004, 1.5, 0.7, 12.0 + 1e-3):
        pts = np.zeros((50, 6))
        pts[:, 0] = rng.uniform(0, ext, 50)
        pts[:, 1] = rng.uniform(0, 3.2, 50)
        pts[0, 0], pts[1, 0] = 0.0, ext
        wins, grid, cmin, cmax = so.tile_windows(pts.astype(np.float32))
        cmin, cmax = np.amin(pts.astype(np.float32).astype(np.float64), 0), np.amax(pts.astype(np.float32).astype(np.float64), 0)
        sx, ex, lox, hix = scene.window_bounds(cmin[0], cmax[0], 1.0, 0.5, 0.001)
        sy, ey, loy, hiy = scene.window_bounds(cmin[1], cmax[1], 1.0, 0.5, 0.001)
        assert (len(sx), len(sy)) == grid
        for w, s_x, s_y, idx in wins:
            assert sx[w % grid[0]] == s_x and sy[w // grid[0]] == s_y

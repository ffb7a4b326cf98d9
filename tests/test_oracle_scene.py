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

"""GPU whole-scene tiler and vote scatter-back (pointcloud_bridge_b200/scene.py, csrc/scene.cu) against the CPU
oracle and the reference's golden outputs: windows, member sets, block structure and per-entry features
bit-exact; vote counts and labels identical."""
import numpy as np
import pytest
import torch

import parity
from oracle import scene_oracle as so
from pointcloud_bridge_b200 import scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_tiler_matches_oracle_and_reference_golden():
    g = parity.load("scene.npz")
    bp = int(g["block_points"])
    pts = g["points"]
    tiles = scene.SceneTiler(block_points=bp).tile(torch.from_numpy(pts).to(DEV))
    o_data, o_idx, o_wid, o_grid = so.tile_scene(pts, block_points=bp)
    assert tiles.grid == o_grid
    assert tiles.data.shape == g["data"].shape                        # the reference yields as many blocks
    assert np.array_equal(tiles.window_of_block.cpu().numpy(), o_wid)
    pidx = tiles.point_idx.cpu().numpy()
    data = tiles.data.cpu().numpy()
    wins, _, cmin, cmax = so.tile_windows(pts.astype(np.float64))
    for w, s_x, s_y, members in wins:
        sel = o_wid == w
        e = pidx[sel].reshape(-1)
        assert np.array_equal(np.unique(e), members), f"window {w}: member set"
        _, cnt = np.unique(e, return_counts=True)
        assert cnt.max() - cnt.min() <= 1, "padding draws without replacement while it is shorter than the window"
        feats = so.entry_features(pts.astype(np.float64), e, s_x, s_y, cmin, cmax)
        assert np.array_equal(feats.view(np.uint32), data[sel].reshape(-1, 9).view(np.uint32)), f"window {w}: features"
    assert tiles.model_input().shape == (pidx.shape[0], 9, bp)
    # the block composition itself (hashed member order, padding, affine interleave) is the oracle's, bit for bit
    assert np.array_equal(pidx, o_idx)
    assert np.array_equal(data.view(np.uint32), o_data.view(np.uint32))


def test_tiler_is_deterministic_mixes_blocks_and_redraws_per_vote():
    """Advisor finding of round 1: blocks were contiguous chunks of the atomics' fill order.  Now (a) two runs give
    identical blocks, (b) every block of a multi-block window is spread over the whole window (the reference
    shuffles: each block is a uniform subsample), (c) duplicates are spread over the blocks, (d) another vote gives
    another composition of the same member sets."""
    rng = np.random.default_rng(5)
    n = 200_000
    xyz = np.stack([rng.uniform(0, 3.0, n), rng.uniform(0, 2.0, n), rng.uniform(0, 1.0, n)], 1).astype(np.float32)
    xyz = xyz[np.argsort(xyz[:, 0], kind="stable")]                  # scan order = x order: the worst case for chunking
    pts = torch.from_numpy(np.concatenate([xyz, rng.uniform(0, 1, (n, 3)).astype(np.float32)], 1)).to(DEV)
    tiler = scene.SceneTiler(block_points=4096, seed=3)
    a, b = tiler.tile(pts), tiler.tile(pts)
    assert torch.equal(a.point_idx, b.point_idx) and torch.equal(a.data, b.data)
    o_data, o_idx, o_wid, _ = so.tile_scene(pts.cpu().numpy(), block_points=4096, seed=3)
    assert np.array_equal(a.point_idx.cpu().numpy(), o_idx)
    pidx = a.point_idx.cpu().numpy()
    wid = a.window_of_block.cpu().numpy()
    w = np.bincount(wid).argmax()                                     # the window with most blocks
    blocks = pidx[wid == w]
    assert blocks.shape[0] >= 8
    x_all = xyz[np.unique(blocks), 0]
    spread = [np.ptp(xyz[blk, 0]) / np.ptp(x_all) for blk in blocks]
    assert min(spread) > 0.95, spread                                # no block is a spatial slice of the window
    means = [xyz[blk, 0].mean() for blk in blocks]
    assert np.ptp(means) < 0.05 * np.ptp(x_all)
    members, counts = np.unique(blocks, return_counts=True)
    dup = members[counts > 1]
    if dup.size >= blocks.shape[0] * 8:                              # duplicates (padding) land in every block
        per_block = [np.isin(blk, dup).sum() for blk in blocks]
        assert min(per_block) > 0.3 * np.mean(per_block), per_block
    c = tiler.tile(pts, vote=1)
    assert not torch.equal(a.point_idx, c.point_idx)
    pc = c.point_idx.cpu().numpy()
    for ww in np.unique(wid)[::3]:
        assert np.array_equal(np.unique(pidx[wid == ww]), np.unique(pc[wid == ww]))


def test_vote_matches_reference_golden():
    g = parity.load("scene.npz")
    P = g["points"].shape[0]
    pool = scene.new_vote_pool(P, 3, DEV)
    half = g["index"].shape[0] // 2                                    # two batches, as the evaluation loop votes
    for sl in (slice(0, half), slice(half, None)):
        scene.add_vote(pool, torch.from_numpy(g["index"][sl]).to(DEV), torch.from_numpy(g["pred"][sl]).to(DEV))
    assert np.array_equal(pool.cpu().numpy(), g["pool"])
    assert np.array_equal(scene.vote_argmax(pool).cpu().numpy(), g["labels"])


@pytest.mark.parametrize("n,stride,block", [(300_000, 0.5, 1.0), (50_000, 1.0, 1.0), (80_000, 0.7, 1.5)])
def test_tiler_properties_on_larger_scenes(n, stride, block):
    rng = np.random.default_rng(n)
    xyz = np.stack([rng.uniform(0, 20.0, n), rng.uniform(0, 6.0, n), rng.uniform(0, 3.0, n)], 1).astype(np.float32)
    pts = np.concatenate([xyz, rng.uniform(0, 255, (n, 3)).astype(np.float32)], 1)
    t = scene.SceneTiler(block_points=4096, stride=stride, block_size=block).tile(torch.from_numpy(pts).to(DEV))
    o_data, o_idx, o_wid, _ = so.tile_scene(pts, block_points=4096, block_size=block, stride=stride)
    pidx = t.point_idx.cpu().numpy()
    assert np.array_equal(t.window_of_block.cpu().numpy(), o_wid)
    assert np.unique(pidx).size == n                                   # every point is evaluated at least once
    for w in np.unique(o_wid)[::7]:
        assert np.array_equal(np.unique(pidx[o_wid == w]), np.unique(o_idx[o_wid == w]))
    # identity vote: predicting (point index mod 5) for every entry gives that label back
    pool = scene.new_vote_pool(n, 5, DEV)
    scene.add_vote(pool, t.point_idx, (t.point_idx % 5).to(torch.uint8))
    assert torch.equal(scene.vote_argmax(pool).long().cpu(), torch.arange(n) % 5)
    assert int(pool.sum()) == pidx.size


def test_segment_scene_end_to_end():
    """Tile -> block inference (PointNet++ SSG, eval) -> votes -> labels: every point gets a label, and the
    result equals voting the same per-block predictions with the CPU oracle."""
    from pointcloud_bridge_b200 import engine
    from pointcloud_bridge_b200.partsize import pointnet2_sem_seg as ssg
    rng = np.random.default_rng(0)
    n = 30_000
    xyz = np.stack([rng.uniform(0, 4.0, n), rng.uniform(0, 2.0, n), rng.uniform(0, 1.0, n)], 1).astype(np.float32)
    pts = torch.from_numpy(np.concatenate([xyz, rng.uniform(0, 1, (n, 3)).astype(np.float32)], 1)).to(DEV)
    torch.manual_seed(0)
    net = ssg.get_model(5).to(DEV).eval()
    infer = engine.BlockInference(net, batch_blocks=4, amp=False, graph=False)
    torch.manual_seed(1)
    labels = engine.segment_scene(net, pts, 5, block_points=2048, batch_blocks=4, amp=False, infer=infer)
    assert labels.shape == (n,) and labels.dtype == torch.uint8 and int(labels.max()) < 5
    tiles = scene.SceneTiler(block_points=2048).tile(pts)
    torch.manual_seed(1)
    pred = infer.run(tiles.model_input())
    pool = so.add_vote(np.zeros((n, 5)), tiles.point_idx.cpu().numpy(), pred.cpu().numpy())
    assert np.array_equal(so.vote_argmax(pool).astype(np.uint8), labels.cpu().numpy())

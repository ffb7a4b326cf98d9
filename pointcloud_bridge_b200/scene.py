"""Whole-scene tiling and vote scatter-back on the GPU -- the steps either
side of the block-sharded inference path (BASELINE config 5).

Mirrors `ScannetDatasetWholeScene.__getitem__` (Highway_bridge/utils/BridgeDataLoader.py:214-277; same class in
Partsize-identical/data_prep/BridgeDataLoader.py:168-231) and the vote loop of
Partsize-identical/test_sem_seg.py:58-65, 132-162:

    tiler = SceneTiler(block_points=4096, stride=0.5, block_size=1.0, padding=0.001)
    tiles = tiler.tile(points_xyzrgb)                 # [P,6] CUDA tensor -> data [nb,4096,9], point_idx [nb,4096]
    pool = new_vote_pool(P, num_classes, device)
    add_vote(pool, tiles.point_idx, pred_labels)      # pred_labels [nb,4096] uint8
    labels = vote_argmax(pool)                        # [P] uint8, np.argmax semantics

Window membership is bit-identical to the reference (bounds computed in float64 like numpy does, compared
in double on the device).  The reference pads every window to a multiple of `block_points` with random
re-draws of its own points and shuffles members + padding (np.random), so each block is a uniform random
subsample of its window.  Here the same structure comes from a counter-based hash of (seed, vote, point,
window): members are sorted by hash, the padding re-uses the first members of that order, and an affine
permutation interleaves members and padding over the blocks -- deterministic for a given seed, different for
every vote, independent of the order in which the atomics filled the windows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import ops

__all__ = ["SceneTiler", "SceneTiles", "new_vote_pool", "add_vote", "vote_argmax", "window_bounds"]


def window_bounds(coord_min: float, coord_max: float, block_size: float, stride: float, padding: float):
    """Start / end of every window along one axis exactly as BridgeDataLoader.py:218-230 computes them
    (float64): s = min + i*stride; e = min(s + block, max); s = e - block.  Returns (s, e, lo, hi) arrays
    where [lo, hi] = [s - padding, e + padding] is the closed membership interval."""
    grid = int(np.ceil(float(coord_max - coord_min - block_size) / stride) + 1)
    if grid <= 0:
        return (np.zeros(0),) * 4
    s = np.empty(grid, np.float64)
    e = np.empty(grid, np.float64)
    for i in range(grid):
        s_i = coord_min + i * stride
        e_i = min(s_i + block_size, coord_max)
        s[i] = e_i - block_size
        e[i] = e_i
    return s, e, s - padding, e + padding


@dataclass
class SceneTiles:
    data: torch.Tensor            # [nb, block_points, 9] fp32: x - cx, y - cy, z, r, g, b, x/ex, y/ey, z/ez
    point_idx: torch.Tensor       # [nb, block_points] int64 index of every entry in the scene
    window_of_block: torch.Tensor  # [nb] int64 window id (iy * grid_x + ix) -- blocks follow the reference's window order
    window_counts: torch.Tensor   # [grid_y * grid_x] int32 points per window (before padding)
    grid: tuple                   # (grid_x, grid_y)
    num_blocks: int = 0           # blocks of the whole scene
    block_offset: int = 0         # first block held in `data` / `point_idx` (tile(..., block_range=...))

    def model_input(self) -> torch.Tensor:
        """[nb, 9, block_points] view, the layout the PointNet++ networks take."""
        return self.data.permute(0, 2, 1)


class SceneTiler:
    def __init__(self, block_points: int = 4096, stride: float = 0.5, block_size: float = 1.0, padding: float = 0.001,
                 seed: int = 0):
        self.block_points, self.stride, self.block_size, self.padding = int(block_points), float(stride), float(block_size), float(padding)
        self.seed = int(seed)

    def vote_seed(self, vote: int) -> int:
        """32-bit seed of the block composition of one vote (a new pseudo-random subsample per vote, as the
        reference's np.random draws give on every pass over the scene)."""
        return (self.seed * 0x9E3779B1 + int(vote) * 0x632BE5AB + 0x7F4A7C15) & 0xFFFFFFFF

    @torch.no_grad()
    def tile(self, points: torch.Tensor, vote: int = 0, block_range=None) -> SceneTiles:
        """points [P, >=6] fp32 CUDA tensor (x, y, z, r, g, b, ...); `vote`: which pass over the scene.
        block_range: callable (num_blocks) -> range of the blocks to materialise (a rank's shard of the block list:
        windows are counted and ordered for the whole scene, only the shard's [.., block_points, 9] rows are built)."""
        if not points.is_cuda or points.dtype != torch.float32 or points.dim() != 2 or points.shape[1] < 6:
            raise ValueError("SceneTiler.tile expects a [P, >=6] fp32 CUDA tensor")
        points = points.contiguous()
        dev = points.device
        P, pstride = points.shape
        # np.amin / np.amax over the (fp32-valued) coordinates, then float64 arithmetic as in the reference
        cmin = points[:, :3].amin(dim=0).double().cpu().numpy()
        cmax = points[:, :3].amax(dim=0).double().cpu().numpy()
        sx, ex, lox, hix = window_bounds(cmin[0], cmax[0], self.block_size, self.stride, self.padding)
        sy, ey, loy, hiy = window_bounds(cmin[1], cmax[1], self.block_size, self.stride, self.padding)
        gx, gy = len(sx), len(sy)
        if gx == 0 or gy == 0:
            raise ValueError("scene is smaller than one block minus one stride: no windows (the reference yields nothing)")
        reach = int(math.ceil(self.block_size / self.stride)) + 1
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        lox_d, hix_d, loy_d, hiy_d = d(lox), d(hix), d(loy), d(hiy)
        nwin = gx * gy
        counts = torch.zeros(nwin, dtype=torch.int32, device=dev)
        common = (points.data_ptr(), P, pstride, gx, gy, lox_d.data_ptr(), hix_d.data_ptr(), loy_d.data_ptr(), hiy_d.data_ptr(),
                  float(cmin[0]), float(cmin[1]), self.stride, reach)
        ops._call("pcb_scene_window_count_f32", dev, *common, counts.data_ptr(), alg_bytes=P * 8)
        offsets = torch.cumsum(counts.long(), 0) - counts.long()
        total = int(counts.sum().item())                      # host sync: sizes of the outputs depend on it
        seed = self.vote_seed(vote)
        keys = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
        cursor = torch.zeros(nwin, dtype=torch.int32, device=dev)
        ops._call("pcb_scene_window_fill_f32", dev, *common, offsets.data_ptr(), cursor.data_ptr(), seed, keys.data_ptr(),
                  alg_bytes=P * 8 + total * 8)
        # one radix sort over all windows: (window, hash, point) ascending = every window in its pseudo-random order
        members = (torch.sort(keys)[0] & 0x7FFFFFFF).to(torch.int32)
        del keys
        # per-block metadata (small, nwin-sized tensors): blocks of a window are consecutive, windows in (iy, ix) order
        nblk_w = (counts.long() + self.block_points - 1) // self.block_points
        nb = int(nblk_w.sum().item())
        if nb == 0:
            raise ValueError("no points fall into any window")
        win = torch.repeat_interleave(torch.arange(nwin, device=dev), nblk_w)
        first_blk_of_win = torch.cumsum(nblk_w, 0) - nblk_w
        blk_in_win = torch.arange(nb, device=dev) - first_blk_of_win[win]
        blk_off = offsets[win].contiguous()
        blk_cnt = counts[win].contiguous()
        blk_first = (blk_in_win * self.block_points).contiguous()
        cx = d(sx + self.block_size / 2.0)[win % gx]
        cy = d(sy + self.block_size / 2.0)[win // gx]
        blk_center = torch.stack([cx, cy], dim=1).contiguous()
        ext = cmax - cmin
        rng = range(nb) if block_range is None else block_range(nb)
        lo, n_loc = rng.start, len(rng)
        data = torch.empty(n_loc, self.block_points, 9, dtype=torch.float32, device=dev)
        pidx = torch.empty(n_loc, self.block_points, dtype=torch.long, device=dev)
        if n_loc:
            ops._call("pcb_scene_blocks_f32", dev, points.data_ptr(), pstride, members.data_ptr(),
                      blk_off[lo:].data_ptr(), blk_cnt[lo:].data_ptr(), blk_first[lo:].data_ptr(),
                      blk_center[lo:].data_ptr(), n_loc, self.block_points, float(ext[0]), float(ext[1]), float(ext[2]), seed,
                      data.data_ptr(), pidx.data_ptr(), alg_bytes=n_loc * self.block_points * (24 + 36 + 8 + 4))
        return SceneTiles(data, pidx, win[lo:lo + n_loc], counts, (gx, gy), nb, lo)


def new_vote_pool(num_points: int, num_classes: int, device) -> torch.Tensor:
    """vote_label_pool of test_sem_seg.py:130, as int32 counts."""
    return torch.zeros(num_points, num_classes, dtype=torch.int32, device=device)


@torch.no_grad()
def add_vote(pool: torch.Tensor, point_idx: torch.Tensor, pred_label: torch.Tensor) -> torch.Tensor:
    """pool[point_idx[b,n], pred_label[b,n]] += 1 for every entry (test_sem_seg.py:58-65; the reference's
    per-point weight is the label weight of the ground truth, always non-zero for the whole-scene set)."""
    if pool.dtype != torch.int32 or not pool.is_contiguous():
        raise ValueError("vote pool must be a contiguous int32 tensor [P, classes]")
    point_idx = ops._i64(point_idx, "point_idx")
    pred = pred_label.to(torch.uint8).contiguous()
    total = point_idx.numel()
    ops._call("pcb_scene_vote", pool.device, point_idx.data_ptr(), pred.data_ptr(), total, pool.shape[0], pool.shape[1],
              pool.data_ptr(), alg_bytes=total * 13)
    return pool


@torch.no_grad()
def vote_argmax(pool: torch.Tensor) -> torch.Tensor:
    """np.argmax(vote_label_pool, 1) (test_sem_seg.py:162): first maximum; [P] uint8."""
    labels = torch.empty(pool.shape[0], dtype=torch.uint8, device=pool.device)
    ops._call("pcb_scene_vote_argmax", pool.device, pool.data_ptr(), pool.shape[0], pool.shape[1], labels.data_ptr(),
              alg_bytes=pool.numel() * 4 + pool.shape[0])
    return labels

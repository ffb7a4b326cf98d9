"""Seeded synthetic "bridge-like" point blocks.

The reference trains on 4096-point blocks cut from bridge scans and normalised into the
unit ball (Highway_bridge/utils/BriPCDMulti_new.py:70-81, 254-273).  The data is not public,
so the benchmarks and parity tests use this generator: a mixture of thin slabs and boxes
(deck, girders, piers, parapets) plus uniform clutter, centred and scaled like
`normalize_points`.  numpy's PCG64 makes the stream identical on every host.
"""
from __future__ import annotations

import numpy as np

__all__ = ["bridge_block", "bridge_batch", "uniform_batch", "duplicated_batch", "bridge_scene",
           "sem_seg_input", "poly_features"]


def _box(rng, n, lo, hi):
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    return lo + (hi - lo) * rng.random((n, 3))


def _bridge_parts(rng, n, length=30.0, width=10.0):
    """Points (float64, metres) and labels of one span: 0 noise, 1 abutment/pier, 2 girder,
    3 deck, 4 parapet -- the reference's five classes."""
    n_girders = int(rng.integers(2, 5))
    frac = np.array([0.10, 0.14, 0.22, 0.42, 0.12])
    counts = np.floor(frac * n).astype(int)
    counts[3] += n - counts.sum()
    pts, lab = [], []
    hl, hw = length / 2, width / 2
    # deck: wide horizontal slab, thickness noise sigma = 5 mm
    deck = _box(rng, counts[3], [-hl, -hw, 0.0], [hl, hw, 0.0])
    deck[:, 2] = rng.normal(0.0, 0.005, counts[3])
    pts.append(deck); lab.append(np.full(counts[3], 3))
    # girders: long boxes under the deck (surfaces only: sample a box then snap one axis)
    per = np.full(n_girders, counts[2] // n_girders)
    per[0] += counts[2] - per.sum()
    ys = np.linspace(-hw * 0.7, hw * 0.7, n_girders)
    for y, m in zip(ys, per):
        g = _box(rng, m, [-hl, y - 0.3, -1.4], [hl, y + 0.3, -0.05])
        side = rng.integers(0, 3, m)
        g[side == 0, 1] = y - 0.3
        g[side == 1, 1] = y + 0.3
        g[side == 2, 2] = -1.4
        pts.append(g); lab.append(np.full(m, 2))
    # two piers / abutments: vertical boxes
    per = [counts[1] // 2, counts[1] - counts[1] // 2]
    for x, m in zip((-hl * 0.8, hl * 0.8), per):
        p = _box(rng, m, [x - 0.8, -hw * 0.8, -7.0], [x + 0.8, hw * 0.8, -1.4])
        side = rng.integers(0, 2, m)
        p[side == 0, 0] = x - 0.8
        p[side == 1, 0] = x + 0.8
        pts.append(p); lab.append(np.full(m, 1))
    # parapets: thin vertical strips along both deck edges
    per = [counts[4] // 2, counts[4] - counts[4] // 2]
    for y, m in zip((-hw, hw), per):
        q = _box(rng, m, [-hl, y - 0.02, 0.0], [hl, y + 0.02, 1.1])
        pts.append(q); lab.append(np.full(m, 4))
    # clutter
    pts.append(_box(rng, counts[0], [-hl, -hw * 1.3, -7.5], [hl, hw * 1.3, 3.0]))
    lab.append(np.full(counts[0], 0))
    return np.concatenate(pts), np.concatenate(lab)


def _normalize(p):
    """Centre on the centroid and scale so the farthest point sits on the unit sphere."""
    p = p - p.mean(axis=0, keepdims=True)
    return p / np.sqrt((p ** 2).sum(axis=1)).max()


def bridge_block(seed: int, n: int = 4096):
    """One block: xyz [n,3] fp32 in the unit ball, rgb [n,3] fp32 in [0,1), labels [n] int64."""
    rng = np.random.default_rng(int(seed))
    p, lab = _bridge_parts(rng, n)
    perm = rng.permutation(n)
    p, lab = p[perm], lab[perm]
    xyz = _normalize(p).astype(np.float32)
    rgb = rng.random((n, 3)).astype(np.float32)
    return xyz, rgb, lab.astype(np.int64)


def bridge_batch(seed: int, batch: int, n: int = 4096):
    """Batch of independent blocks: xyz [B,n,3], rgb [B,n,3], labels [B,n]."""
    blocks = [bridge_block(seed * 100003 + b, n) for b in range(batch)]
    return tuple(np.stack(x) for x in zip(*blocks))


def uniform_batch(seed: int, batch: int, n: int = 4096):
    """Uniform cube control cloud: no exact distance ties in practice."""
    rng = np.random.default_rng(int(seed) + 7919)
    return rng.random((batch, n, 3)).astype(np.float32)


def duplicated_batch(seed: int, batch: int, n: int = 4096, dup_frac: float = 0.15):
    """Bridge blocks where `dup_frac` of the points are exact copies of other points --
    what the reference's loaders produce when they pad a short block by resampling with
    replacement (BriPCDMulti_new.py:224-228).  Stresses tie-breaking."""
    xyz, rgb, lab = bridge_batch(seed, batch, n)
    rng = np.random.default_rng(int(seed) + 104729)
    m = int(n * dup_frac)
    for b in range(batch):
        dst = rng.choice(n, m, replace=False)
        src = rng.integers(0, n, m)
        xyz[b, dst] = xyz[b, src]
        rgb[b, dst] = rgb[b, src]
        lab[b, dst] = lab[b, src]
    return xyz, rgb, lab


def sem_seg_input(xyz, rgb):
    """The 9-channel PointNet++ sem-seg input [B,9,N] = [xyz | rgb | xyz / extent], as
    LWBridgeDataset assembles it (Partsize-identical/data_prep/BridgeDataLoader.py:98-113)."""
    lo = xyz.min(axis=1, keepdims=True)
    ext = (xyz - lo).max(axis=1, keepdims=True)
    ext = np.where(ext > 0, ext, 1.0)
    x = np.concatenate([xyz, rgb, (xyz - lo) / ext], axis=-1).astype(np.float32)
    return np.ascontiguousarray(np.transpose(x, (0, 2, 1)))


def bridge_scene(seed: int, n_points: int, block: int = 4096):
    """A long multi-span scene cut into `ceil(n_points / block)` blocks by position along the
    span, each block normalised on its own.  Returns a generator of (xyz, rgb) blocks so that
    50 M-point scenes never exist in host memory at once."""
    nblocks = -(-n_points // block)
    for i in range(nblocks):
        xyz, rgb, _ = bridge_block(seed * 1000003 + i, block)
        yield xyz, rgb


def poly_features(xyz, dim: int, seed: int = 0):
    """Deterministic [B,dim,N] fp32 features: fixed random quadratic forms of the coordinates,
    evaluated with fp32 multiplies and adds only (bit-identical on every host; no libm)."""
    rng = np.random.default_rng(int(seed) + 15485863)
    w = rng.standard_normal((dim, 6)).astype(np.float32)
    x, y, z = (np.ascontiguousarray(xyz[..., i], dtype=np.float32) for i in range(3))
    xy, yz, zx = x * y, y * z, z * x
    out = np.empty((xyz.shape[0], dim, xyz.shape[1]), np.float32)
    for c in range(dim):
        t = w[c, 0] * x
        t = t + w[c, 1] * y
        t = t + w[c, 2] * z
        t = t + w[c, 3] * xy
        t = t + w[c, 4] * yz
        t = t + w[c, 5] * zx
        out[:, c, :] = t
    return out

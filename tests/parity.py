"""Shared parity helpers: golden loading and tie-aware index comparison.

The reference's sort/topk leave the order of exactly-equal distances unspecified (SURVEY.md
Appendix A); the build orders by (distance, index).  `assert_topk_equivalent` therefore demands
bit-equal *distance sequences* everywhere and equal indices wherever the distance is untied.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PN2_LEVELS = [(1024, [(0.05, 16), (0.1, 32)]), (256, [(0.1, 16), (0.2, 32)]),
              (64, [(0.2, 16), (0.4, 32)]), (16, [(0.4, 16), (0.8, 32)])]
BRI_LEVELS = [(1024, [(0.1, 16), (0.2, 32)]), (512, [(0.2, 16), (0.4, 32)]),
              (128, [(0.4, 16), (0.8, 32)])]


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def bits_checksum(d):
    return np.ascontiguousarray(d, np.float32).view(np.uint32).astype(np.uint64).sum(axis=-1)


def gather_rows(full, idx):
    """full [B,N,M] distances, idx [B,N,k] -> [B,N,k]."""
    return np.take_along_axis(full, idx.astype(np.int64), axis=2)


def _ulps(a, b):
    """|a - b| in units of the last place of the larger magnitude (non-negative fp32 inputs)."""
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def assert_topk_equivalent(idx_ours, idx_ref, dist_of, what="", ulp_tol=0):
    """idx_* [B,N,k]; dist_of(idx) -> fp32 distances of those (row, idx) pairs under the oracle's
    arithmetic.  Ours must be ordered by (distance, index); the reference's row must carry the
    same distance sequence (bit for bit when ulp_tol == 0); indices must agree wherever a
    distance is separated from its neighbours by more than ulp_tol ulps.

    ulp_tol = 1 is used only for the cdist flavour, whose reference square root (MKL VML) is not
    correctly rounded, so that candidates within 1 ulp are effectively tied in the reference.
    Returns the fraction of differing index entries (all of them inside tie groups)."""
    idx_ours = np.asarray(idx_ours, np.int64)
    idx_ref = np.asarray(idx_ref, np.int64)
    assert idx_ours.shape == idx_ref.shape, what
    d_o = np.ascontiguousarray(dist_of(idx_ours), np.float32)
    d_r = np.ascontiguousarray(dist_of(idx_ref), np.float32)
    pos = (d_o >= 0).all() and (d_r >= 0).all()
    if ulp_tol == 0:
        assert np.array_equal(d_o.view(np.uint32), d_r.view(np.uint32)), f"{what}: distance sequences differ"
    else:
        assert pos and (_ulps(d_o, d_r) <= ulp_tol).all(), f"{what}: distance sequences differ by > {ulp_tol} ulp"
    # ours ordered by (distance, index), no repeats
    dd = np.diff(d_o, axis=-1)
    di = np.diff(idx_ours, axis=-1)
    assert (dd >= 0).all(), f"{what}: distances not ascending"
    assert ((dd > 0) | (di > 0)).all(), f"{what}: ties not in ascending index order"
    srt = np.sort(idx_ours, axis=-1)
    assert (np.diff(srt, axis=-1) > 0).all(), f"{what}: repeated index in a row"
    # untied positions must match exactly
    if ulp_tol == 0:
        close = dd == 0
    else:
        close = _ulps(d_o[..., 1:], d_o[..., :-1]) <= ulp_tol
    tied = np.zeros(d_o.shape, bool)
    tied[..., 1:] |= close
    tied[..., :-1] |= close
    tied[..., -1] = True          # the k-th may tie with the unseen (k+1)-th
    mism = (idx_ours != idx_ref) & ~tied
    assert not mism.any(), f"{what}: {int(mism.sum())} untied indices differ"
    return float((idx_ours != idx_ref).mean())


def seeded_fill_(model, seed=0):
    """Deterministic, construction-order-independent weights: every parameter / buffer is filled
    from a numpy generator keyed by its *name*, so the reference network and the B200 network
    (same state_dict keys) get bit-identical values on any host."""
    import zlib

    import torch
    with torch.no_grad():
        for name, t in list(model.named_parameters()) + list(model.named_buffers()):
            rng = np.random.default_rng((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
            if name.endswith("num_batches_tracked") or name.endswith("freqs") or "base_weights" in name:
                continue
            if name.endswith("running_var"):
                v = rng.uniform(0.5, 1.5, t.shape)
            elif name.endswith("running_mean"):
                v = rng.standard_normal(t.shape) * 0.1
            elif t.dim() > 1:
                fan_in = int(np.prod(t.shape[1:]))
                v = rng.standard_normal(t.shape) / np.sqrt(fan_in)
            elif name.endswith("weight"):                     # norm-layer scale
                v = rng.uniform(0.5, 1.5, t.shape)
            else:                                             # biases
                v = rng.standard_normal(t.shape) * 0.1
            t.copy_(torch.from_numpy(np.asarray(v, np.float32)))
    return model

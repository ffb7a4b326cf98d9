"""TEST INFRASTRUCTURE ONLY (never imported by pointcloud_bridge_b200/): numpy restatement of the input rows of
BridgeStructureEncoding -- Highway_bridge/models/attention_modules.py of the reference:
  :552-574  compute_absolute_position_encoding   -> abs_encoding()
  :590-597  neighbours - centre                  -> rel_pos()
  :622-687  get_structure_features               -> structure_features()
  :603-613  expand + cat                          -> rows()
Pinned against outputs of the unmodified reference class (tests/golden/make_golden_structure.py ->
tests/golden/structure.npz) by tests/test_oracle_structure.py; the CUDA kernel (csrc/structure.cu) is checked
against both on the GPU (tests/test_gpu_structure.py)."""
import numpy as np


def abs_encoding(xyz, freqs, grid_size=1.0):
    """:552-574: sin / cos of the grid-snapped coordinates at every frequency -> [B,N,6F]."""
    xyz = np.asarray(xyz, np.float32)
    grid = np.floor(xyz / np.float32(grid_size)) * np.float32(grid_size)
    enc = []
    for f in np.asarray(freqs, np.float32):
        enc += [np.sin(grid * f), np.cos(grid * f)]
    return np.concatenate(enc, axis=-1).astype(np.float32)


def rel_pos(xyz, idx):
    """:590-597: xyz[b, idx[b,n,j]] - xyz[b,n] -> [B,N,k,3]."""
    xyz = np.asarray(xyz, np.float32)
    b = np.arange(xyz.shape[0])[:, None, None]
    return xyz[b, idx] - xyz[:, :, None, :]


def structure_features(rel):
    """:622-687: the 13 statistics of the neighbour offsets rel [B,N,k,3] (fp32 throughout, as the reference)."""
    rel = np.asarray(rel, np.float32)
    B, N, k, _ = rel.shape
    flat = rel.reshape(B * N, k, 3)
    cov = np.matmul(flat.transpose(0, 2, 1), flat) / np.float32(k - 1)                       # :628-629
    ev = np.linalg.eigvalsh(cov.astype(np.float32)).astype(np.float32).reshape(B, N, 3)       # :632, ascending
    den = ev[..., 0] + np.float32(1e-8)
    shape = np.stack([(ev[..., 0] - ev[..., 1]) / den, (ev[..., 1] - ev[..., 2]) / den, ev[..., 2] / den], -1)
    centre = rel.mean(axis=2, keepdims=True)                                                  # :646
    dist = np.linalg.norm(rel - centre, axis=-1)
    local = np.stack([dist.max(-1), dist.mean(-1), dist.std(-1, ddof=1)], -1)                 # :649-653
    unit = rel / (np.linalg.norm(rel, axis=-1, keepdims=True) + np.float32(1e-8))             # :656
    sim = np.matmul(unit.reshape(B * N, k, 3), unit.reshape(B * N, k, 3).transpose(0, 2, 1))  # :657-661
    direction = sim.reshape(B, N, k * k).mean(-1, keepdims=True)
    z = rel[..., 2]
    zst = np.stack([z.std(-1, ddof=1), z.max(-1) - z.min(-1)], -1)                            # :664-667
    mean_rel = rel.mean(axis=2)                                                               # :670
    spread = np.linalg.norm(rel.std(axis=2, ddof=1), axis=-1, keepdims=True)                  # :679
    return np.concatenate([shape, local, direction, zst, mean_rel, spread], -1).astype(np.float32)


def rows(xyz, idx, freqs, grid_size=1.0):
    """:603-613: every neighbour row = [abs encoding | rel_pos | structure features] -> [B*N*k, 6F+16]."""
    rel = rel_pos(xyz, idx)
    B, N, k, _ = rel.shape
    a = abs_encoding(xyz, freqs, grid_size)
    s = structure_features(rel)
    r = np.concatenate([np.broadcast_to(a[:, :, None, :], (B, N, k, a.shape[-1])), rel,
                        np.broadcast_to(s[:, :, None, :], (B, N, k, 13))], -1)
    return r.reshape(B * N * k, -1).astype(np.float32)


def well_conditioned(rel, ratio=1e-3):
    """Mask of neighbourhoods whose smallest covariance eigenvalue is not negligible: the shape features divide by it,
    so an fp32 solver's absolute error of ~1e-7 * largest eigenvalue moves them arbitrarily otherwise."""
    rel = np.asarray(rel, np.float64)
    B, N, k, _ = rel.shape
    cov = np.einsum("bnki,bnkj->bnij", rel, rel) / (k - 1)
    ev = np.linalg.eigvalsh(cov)
    return ev[..., 0] > ratio * ev[..., 2]

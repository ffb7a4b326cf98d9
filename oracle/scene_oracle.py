"""CPU restatement (numpy) of the reference's whole-scene tiling and vote scatter-back -- TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product path never does.

  tile_windows / tile_scene : ScannetDatasetWholeScene.__getitem__, Highway_bridge/utils/BridgeDataLoader.py:214-277
                              (same class: Partsize-identical/data_prep/BridgeDataLoader.py:168-231)
  add_vote / vote_argmax    : Partsize-identical/test_sem_seg.py:58-65, 162

The reference draws the padding of a window with np.random.choice and shuffles the window (:240-242): those
two steps are random, so the oracle exposes the deterministic part (window order, exact membership, block
count, per-entry features) and builds the blocks with the product's counter-based rule (include/pcbridge.h:
members ordered by (hash16(seed, point, window), point), padding = the first members of that order again,
members and padding interleaved by the affine permutation s = (1000003 q + seed % 999983) mod tot).  Pinned
against outputs of the unmodified reference class in tests/golden/scene.npz (tests/golden/make_golden_scene.py).
"""
import numpy as np


def scene_hash(seed, i, w):
    """lowbias32 mix of (seed, point index, window id), uint32 arithmetic (csrc/scene.cu:scene_hash)."""
    m = np.uint64(0xFFFFFFFF)
    h = (np.uint64(seed) ^ ((np.asarray(i, np.uint64) * np.uint64(0x9E3779B1)) & m) ^ ((np.uint64(w) * np.uint64(0x85EBCA77)) & m)) & m
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x7FEB352D)) & m
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x846CA68B)) & m
    h ^= h >> np.uint64(16)
    return h


def vote_seed(seed, vote):
    return (int(seed) * 0x9E3779B1 + int(vote) * 0x632BE5AB + 0x7F4A7C15) & 0xFFFFFFFF


def tile_windows(points, block_size=1.0, stride=0.5, padding=0.001):
    """-> list of (window id, s_x, s_y, ascending member indices) for the non-empty windows, in the reference's
    (index_y, index_x) order (BridgeDataLoader.py:218-236)."""
    points = np.asarray(points, np.float64)
    coord_min, coord_max = np.amin(points, axis=0)[:3], np.amax(points, axis=0)[:3]
    grid_x = int(np.ceil(float(coord_max[0] - coord_min[0] - block_size) / stride) + 1)
    grid_y = int(np.ceil(float(coord_max[1] - coord_min[1] - block_size) / stride) + 1)
    out = []
    for index_y in range(0, grid_y):
        for index_x in range(0, grid_x):
            s_x = coord_min[0] + index_x * stride
            e_x = min(s_x + block_size, coord_max[0])
            s_x = e_x - block_size
            s_y = coord_min[1] + index_y * stride
            e_y = min(s_y + block_size, coord_max[1])
            s_y = e_y - block_size
            idx = np.where((points[:, 0] >= s_x - padding) & (points[:, 0] <= e_x + padding) &
                           (points[:, 1] >= s_y - padding) & (points[:, 1] <= e_y + padding))[0]
            if idx.size == 0:
                continue
            out.append((index_y * grid_x + index_x, s_x, s_y, idx))
    return out, (grid_x, grid_y), coord_min, coord_max


def entry_features(points, idx, s_x, s_y, coord_min, coord_max, block_size=1.0):
    """The 9 channels of the entries `idx` of one window (BridgeDataLoader.py:243-259), float64 like numpy
    computes them, rounded to fp32 as `torch.Tensor(batch_data)` does (test_sem_seg.py:150)."""
    points = np.asarray(points, np.float64)
    d = points[idx, :6].copy()
    norm = np.zeros((idx.size, 3))
    norm[:, 0] = d[:, 0] / (coord_max[0] - coord_min[0])
    norm[:, 1] = d[:, 1] / (coord_max[1] - coord_min[1])
    norm[:, 2] = d[:, 2] / (coord_max[2] - coord_min[2])
    d[:, 0] = d[:, 0] - (s_x + block_size / 2.0)
    d[:, 1] = d[:, 1] - (s_y + block_size / 2.0)
    return np.concatenate((d, norm), axis=1).astype(np.float32)


def tile_scene(points, block_points=4096, block_size=1.0, stride=0.5, padding=0.001, seed=0, vote=0):
    """Blocks in the product's pseudo-random composition: data [nb, block_points, 9] fp32,
    point_idx [nb, block_points] int64, window id per block."""
    wins, grid, cmin, cmax = tile_windows(points, block_size, stride, padding)
    sd = vote_seed(seed, vote)
    data, pidx, wid = [], [], []
    for w, s_x, s_y, idx in wins:
        nb = int(np.ceil(idx.size / block_points))
        n, tot = idx.size, nb * block_points
        key = ((scene_hash(sd, idx, w) >> np.uint64(16)) << np.uint64(31)) | idx.astype(np.uint64)
        order = idx[np.argsort(key, kind="stable")]              # members by (hash16, point index)
        q = np.arange(tot, dtype=np.uint64)
        s = ((q * np.uint64(1000003) + np.uint64(sd % 999983)) % np.uint64(tot)).astype(np.int64)
        rep = order[np.where(s < n, s, (s - n) % n)]
        data.append(entry_features(points, rep, s_x, s_y, cmin, cmax, block_size).reshape(nb, block_points, 9))
        pidx.append(rep.reshape(nb, block_points))
        wid += [w] * nb
    return np.concatenate(data), np.concatenate(pidx).astype(np.int64), np.asarray(wid, np.int64), grid


def add_vote(vote_label_pool, point_idx, pred_label, weight=None):
    """test_sem_seg.py:58-65."""
    B, N = pred_label.shape
    for b in range(B):
        for n in range(N):
            if weight is None or weight[b, n]:
                vote_label_pool[int(point_idx[b, n]), int(pred_label[b, n])] += 1
    return vote_label_pool


def vote_argmax(vote_label_pool):
    """test_sem_seg.py:162."""
    return np.argmax(vote_label_pool, 1)

"""ctypes front-end of the CPU ORACLE (oracle/pcb_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by pointcloud_bridge_b200/.

All functions take and return numpy arrays (fp32 data, int64 indices) in the layouts of the
reference functions they restate; see the C file for the reference file:line of each.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpcb_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    """Compile the oracle with the Makefile next to this file (gcc only, no GPU)."""
    src = os.path.join(_HERE, "pcb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_index_points.restype = ctypes.c_int64
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def radius_sq_f32(radius: float) -> np.float32:
    """fp32(double(radius)**2): how `sqrdists > radius ** 2` sees the threshold
    (pointnet_util.py:106)."""
    return np.float32(float(radius) ** 2)


def row_sumsq(x):
    x, xp = _f(x)
    out = np.empty(x.shape[:-1], np.float32)
    lib().orc_row_sumsq(xp, _i64(out.size), _i64(x.shape[-1]), out.ctypes.data_as(_f32p))
    return out


def square_distance(src, dst):
    src, sp = _f(src)
    dst, dp = _f(dst)
    B, N, C = src.shape
    M = dst.shape[1]
    out = np.empty((B, N, M), np.float32)
    lib().orc_square_distance(sp, dp, _i64(B), _i64(N), _i64(M), _i64(C), out.ctypes.data_as(_f32p))
    return out


def farthest_point_sample(xyz, npoint, start):
    xyz, xp = _f(xyz)
    start, stp = _i(start)
    B, N, _ = xyz.shape
    out = np.empty((B, npoint), np.int64)
    lib().orc_fps(xp, _i64(B), _i64(N), stp, _i64(npoint), out.ctypes.data_as(_i64p))
    return out


def query_ball_point(radius, nsample, xyz, new_xyz):
    xyz, xp = _f(xyz)
    new_xyz, qp = _f(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.empty((B, S, nsample), np.int64)
    lib().orc_ball_query(xp, qp, _i64(B), _i64(N), _i64(S), ctypes.c_float(radius_sq_f32(radius)),
                         _i64(nsample), out.ctypes.data_as(_i64p))
    return out


def knn(x_bdn, k, return_dist=False):
    """DGCNN.knn: x [B,D,N] channels-first -> idx [B,N,k]."""
    x = np.ascontiguousarray(np.transpose(np.asarray(x_bdn, np.float32), (0, 2, 1)))
    B, N, D = x.shape
    idx = np.empty((B, N, k), np.int64)
    dist = np.empty((B, N, k), np.float32)
    lib().orc_knn(x.ctypes.data_as(_f32p), _i64(B), _i64(N), _i64(D), _i64(k),
                  idx.ctypes.data_as(_i64p), dist.ctypes.data_as(_f32p))
    return (idx, dist) if return_dist else idx


def knn_cdist(xyz, k, return_dist=False):
    xyz, xp = _f(xyz)
    B, N, _ = xyz.shape
    idx = np.empty((B, N, k), np.int64)
    dist = np.empty((B, N, k), np.float32)
    sq = np.empty((B, N, k), np.float32)
    lib().orc_knn_cdist(xp, _i64(B), _i64(N), _i64(k), idx.ctypes.data_as(_i64p),
                        dist.ctypes.data_as(_f32p), sq.ctypes.data_as(_f32p))
    if return_dist == "sq":
        return idx, dist, sq
    return (idx, dist) if return_dist else idx


def three_nn(xyz1, xyz2, k=3):
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    dist = np.empty((B, N, k), np.float32)
    idx = np.empty((B, N, k), np.int64)
    lib().orc_three_nn(p1, p2, _i64(B), _i64(N), _i64(S), _i64(k), dist.ctypes.data_as(_f32p),
                       idx.ctypes.data_as(_i64p))
    return dist, idx


def interp_weights(dist):
    dist, dp = _f(dist)
    k = dist.shape[-1]
    w = np.empty_like(dist)
    lib().orc_interp_weights(dp, _i64(dist.size // k), _i64(k), w.ctypes.data_as(_f32p))
    return w


def three_interpolate(points2_bsd, idx, weight):
    p, pp = _f(points2_bsd)
    idx, ip = _i(idx)
    weight, wp = _f(weight)
    B, S, D = p.shape
    N, k = idx.shape[1], idx.shape[2]
    out = np.empty((B, N, D), np.float32)
    lib().orc_three_interpolate(pp, ip, wp, _i64(B), _i64(N), _i64(S), _i64(D), _i64(k),
                                out.ctypes.data_as(_f32p))
    return out


def index_points(points, idx, clamp=False):
    points, pp = _f(points)
    idx, ip = _i(idx)
    B, N, C = points.shape
    M = idx.size // B
    out = np.zeros(tuple(idx.shape) + (C,), np.float32)
    bad = lib().orc_index_points(pp, ip, _i64(B), _i64(N), _i64(C), _i64(M), ctypes.c_int(int(clamp)),
                                 out.ctypes.data_as(_f32p))
    if bad:
        raise IndexError(f"index out of range for dimension of size {N} ({bad} entries)")
    return out


def get_graph_feature(x_bdn, idx):
    x, xp = _f(x_bdn)
    idx, ip = _i(idx)
    B, D, N = x.shape
    k = idx.shape[2]
    out = np.empty((B, 2 * D, N, k), np.float32)
    lib().orc_graph_feature(xp, ip, _i64(B), _i64(D), _i64(N), _i64(k), out.ctypes.data_as(_f32p))
    return out


def group_points(xyz, points, new_xyz, idx, xyz_first=True, clamp=False):
    xyz, xp = _f(xyz)
    new_xyz, qp = _f(new_xyz)
    idx, ip = _i(idx)
    B, N, _ = xyz.shape
    S, K = idx.shape[1], idx.shape[2]
    if points is not None:
        points, pp = _f(points)
        D = points.shape[2]
    else:
        pp, D = None, 0
    out = np.empty((B, S, K, 3 + D), np.float32)
    lib().orc_group_points(xp, pp, qp, ip, _i64(B), _i64(N), _i64(S), _i64(K), _i64(D),
                           ctypes.c_int(int(xyz_first)), ctypes.c_int(int(clamp)),
                           out.ctypes.data_as(_f32p))
    return out


def pair_dist(x_bnd, idx, mode):
    """Oracle distances of explicit (row, idx) pairs: mode 'knn' (DGCNN.py:63-65) or 'cdist'."""
    x, xp = _f(x_bnd)
    idx, ip = _i(idx)
    B, N, D = x.shape
    k = idx.shape[2]
    out = np.empty((B, N, k), np.float32)
    lib().orc_pair_dist(xp, ip, _i64(B), _i64(N), _i64(D), _i64(k),
                        ctypes.c_int({"knn": 0, "cdist": 1, "cdist_sq": 2}[mode]), out.ctypes.data_as(_f32p))
    return out

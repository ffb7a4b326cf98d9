"""Tensor-level front end of the CUDA kernels: takes/returns torch CUDA tensors, calls the C ABI
(include/pcbridge.h) through ctypes on the current CUDA stream, and wires the differentiable
ops into autograd.  PyTorch is used for device memory, streams and autograd bookkeeping only.

Every function fails loudly when the extension is missing or the tensors are not on a CUDA
device: there is no CPU / eager fallback on this path.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

__all__ = [
    "furthest_point_sample", "ball_query", "ball_query_multi", "square_distance", "gather", "group_points",
    "three_nn", "three_interpolate", "knn", "knn_cdist", "graph_feature", "check_index_errors",
]

_STRICT = os.environ.get("PCB_STRICT_INDEX", "0") == "1"
_err_counters: dict[int, torch.Tensor] = {}


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.PcbError(f"{name} must be a CUDA tensor (got {t.device}); the hot path has no CPU fallback")
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _i64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.PcbError(f"{name} must be a CUDA tensor (got {t.device})")
    if t.dtype != torch.int64:
        t = t.long()
    return t if t.is_contiguous() else t.contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# Optional per-launch timing (bench.py / tools): a list that receives
# (entry point, algorithmic bytes, start event, end event) for every call while it is set.
_timer: list | None = None


def set_kernel_timer(sink: list | None) -> None:
    global _timer
    _timer = sink


def _call(name: str, dev: torch.device, *args, launches: int = 1, alg_bytes: int = 0) -> None:
    fn = getattr(_lib.lib(), name)
    timed = _timer is not None
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            code = fn(*args, _stream())
    else:
        code = fn(*args, _stream())
    if timed:
        e1.record()
        _timer.append((name, int(alg_bytes), e0, e1))
    _lib.check(code, name)
    _lib.count_launches(launches)


# Mid-step marker.  A step runner that computes the sampling / grouping indices of the NEXT batch on a side stream
# (engine.Trainer.prefetch) wants that chain -- farthest point sampling: 16 persistent CTAs for a third of a
# millisecond -- to run while the CURRENT step is in its small middle layers (a few dozen to a few hundred CTAs per
# kernel, SMs idle), not while the first set-abstraction levels stream hundreds of MB with every SM busy: there every
# persistent grid sized for 148 SMs gets a ragged second wave (+0.4 ms per step, tools/step_timing.py).  The
# networks call mark_mid_step() where their small layers begin; the runner registers an event (an EXTERNAL event, so
# that a captured step records it on every replay) and makes the side stream wait for it.
_mid_step_event = None


def set_mid_step_event(ev) -> None:
    global _mid_step_event
    _mid_step_event = ev


def mark_mid_step() -> None:
    if _mid_step_event is not None:
        _mid_step_event.record()


# Source of the FPS start indices.  None: the reference's own draw, torch.randint on the CPU
# default generator followed by a host->device copy (pointnet_util.py:79).  A callable
# (B, N, device) -> LongTensor[B] lets a CUDA-graph runner substitute static device buffers that
# it refills with that same CPU draw before every replay (engine.GraphedStep).
_fps_start_provider = None


def set_fps_start_provider(fn) -> None:
    global _fps_start_provider
    _fps_start_provider = fn


# Ring of pinned staging buffers for the FPS start indices: a host->device copy from pageable
# memory makes the host wait for all earlier work of the stream; from pinned memory it is
# asynchronous.  Every slot carries the CUDA event of its last copy: the host rewrites a slot only
# after that copy has run, however far it is ahead of the GPU (a step runner never drains the
# stream between steps).
_START_RING_SLOTS = 64
_start_ring: dict[int, list] = {}


def _draw_start(B: int, N: int, device) -> torch.Tensor:
    """The reference's draw (pointnet_util.py:79): torch.randint on the CPU default generator."""
    host = torch.randint(0, N, (B,), dtype=torch.long)
    ring = _start_ring.get(B)
    if ring is None:
        ring = [[torch.empty(B, dtype=torch.long).pin_memory() for _ in range(_START_RING_SLOTS)],
                [None] * _START_RING_SLOTS, 0]
        _start_ring[B] = ring
    bufs, events, pos = ring
    slot = pos % _START_RING_SLOTS
    ring[2] = pos + 1
    if events[slot] is not None:
        events[slot].synchronize()                 # the copy that last read this slot has completed
    buf = bufs[slot]
    buf.copy_(host)
    out = buf.to(device, non_blocking=True)
    if not torch.cuda.is_current_stream_capturing():
        ev = torch.cuda.Event()
        ev.record()
        events[slot] = ev
    return out


def _err_counter(dev: torch.device) -> torch.Tensor:
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    t = _err_counters.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int32, device=dev)
        _err_counters[key] = t
    return t


def check_index_errors(device=None) -> None:
    """Raise IndexError if any non-clamping index_points call since the last check saw an index
    outside [-N, N) -- the reference's behaviour (pointnet_util.py:62) made explicit, because a
    stream-ordered kernel cannot raise by itself.  Synchronises."""
    for key, t in list(_err_counters.items()):
        if device is not None and torch.device(device).index not in (None, key):
            continue
        n = int(t.item())
        if n:
            t.zero_()
            raise IndexError(f"index out of range in index_points ({n} entries)")


# ---------------------------------------------------------------------------------------------
# index producers (not differentiable)
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def furthest_point_sample(xyz: torch.Tensor, npoint: int, start: torch.Tensor | None = None) -> torch.Tensor:
    """xyz [B,N,3] -> LongTensor [B,npoint].  `start` [B]: first index of every cloud; when None
    it is drawn exactly as the reference does (CPU default generator, pointnet_util.py:79)."""
    xyz = _f32(xyz, "xyz")
    B, N, C = xyz.shape
    if C != 3:
        raise ValueError("farthest point sampling expects 3-D coordinates")
    if start is None:
        if _fps_start_provider is not None:
            start = _fps_start_provider(B, N, xyz.device)
        else:
            start = _draw_start(B, N, xyz.device)
    start = _i64(start.to(xyz.device), "start")
    out = torch.empty(B, npoint, dtype=torch.long, device=xyz.device)
    _call("pcb_fps_f32", xyz.device, xyz.data_ptr(), B, N, start.data_ptr(), int(npoint), out.data_ptr(),
          alg_bytes=B * (12 * N + 8 * npoint))
    return out


@torch.no_grad()
def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """First `nsample` points of xyz [B,N,3] within `radius` of each new_xyz [B,S,3] point, in index
    order, padded with the first hit (pointnet_util.py:91-112).  Returns LongTensor [B,S,nsample]."""
    xyz = _f32(xyz, "xyz")
    new_xyz = _f32(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    # `sqrdists > radius ** 2`: Python squares in double, the comparison rounds it to fp32
    r2 = float(torch.tensor(float(radius) ** 2, dtype=torch.float32).item())
    out = torch.empty(B, S, nsample, dtype=torch.long, device=xyz.device)
    _call("pcb_ball_query_f32", xyz.device, xyz.data_ptr(), new_xyz.data_ptr(), B, N, S, r2, int(nsample),
          out.data_ptr(), alg_bytes=B * (12 * N + 12 * S + 8 * S * nsample))
    return out


@torch.no_grad()
def ball_query_multi(radii, nsamples, xyz: torch.Tensor, new_xyz: torch.Tensor) -> list:
    """query_ball_point for several (radius, nsample) pairs around the same centroids in one scan of the
    cloud; returns one LongTensor [B,S,nsample_k] per pair, equal to separate ball_query calls."""
    import ctypes
    radii, nsamples = list(radii), [int(n) for n in nsamples]
    if len(radii) == 1 or len(radii) > 4:
        return [ball_query(r, n, xyz, new_xyz) for r, n in zip(radii, nsamples)]
    xyz = _f32(xyz, "xyz")
    new_xyz = _f32(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    k = len(radii)
    r2 = (ctypes.c_float * k)(*[float(torch.tensor(float(r) ** 2, dtype=torch.float32).item()) for r in radii])
    ns = (ctypes.c_int * k)(*nsamples)
    outs = [torch.empty(B, S, n, dtype=torch.long, device=xyz.device) for n in nsamples]
    ptrs = (ctypes.c_void_p * k)(*[o.data_ptr() for o in outs])
    _call("pcb_ball_query_multi_f32", xyz.device, xyz.data_ptr(), new_xyz.data_ptr(), B, N, S, k, r2, ns, ptrs,
          alg_bytes=B * (12 * N + 12 * S + 8 * S * sum(nsamples)))
    return outs


@torch.no_grad()
def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    src = _f32(src, "src")
    dst = _f32(dst, "dst")
    B, N, C = src.shape
    M = dst.shape[1]
    out = torch.empty(B, N, M, dtype=torch.float32, device=src.device)
    _call("pcb_square_distance_f32", src.device, src.data_ptr(), dst.data_ptr(), B, N, M, C, out.data_ptr(),
          alg_bytes=4 * B * (N * C + M * C + N * M))
    return out


@torch.no_grad()
def three_nn(xyz1: torch.Tensor, xyz2: torch.Tensor, k: int = 3):
    """k nearest xyz2 [B,S,3] points of every xyz1 [B,N,3] point.
    Returns (dist [B,N,k] squared ascending, idx [B,N,k] int64, weight [B,N,k])."""
    xyz1 = _f32(xyz1, "xyz1")
    xyz2 = _f32(xyz2, "xyz2")
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    dist = torch.empty(B, N, k, dtype=torch.float32, device=xyz1.device)
    idx = torch.empty(B, N, k, dtype=torch.long, device=xyz1.device)
    weight = torch.empty(B, N, k, dtype=torch.float32, device=xyz1.device)
    _call("pcb_three_nn_f32", xyz1.device, xyz1.data_ptr(), xyz2.data_ptr(), B, N, S, int(k), dist.data_ptr(),
          idx.data_ptr(), weight.data_ptr(), alg_bytes=B * (12 * N + 12 * S + 16 * k * N))
    return dist, idx, weight


@torch.no_grad()
def knn(x: torch.Tensor, k: int, channels_first: bool = True, return_dist: bool = False):
    """DGCNN.knn: x [B,D,N] (channels_first) or [B,N,D] -> LongTensor [B,N,k], neighbours ordered by
    (pairwise distance, index), self included."""
    x = _f32(x, "x")
    if channels_first:
        B, D, N = x.shape
    else:
        B, N, D = x.shape
        if D != 3:                       # feature-space kernels read the channels-first layout
            x = x.transpose(1, 2).contiguous()
            channels_first = True
    idx = torch.empty(B, N, k, dtype=torch.long, device=x.device)
    dist = torch.empty(B, N, k, dtype=torch.float32, device=x.device) if return_dist else None
    _call("pcb_knn_f32", x.device, x.data_ptr(), B, N, D, int(k), int(channels_first), idx.data_ptr(),
          dist.data_ptr() if return_dist else None, launches=1 if D == 3 else 2,
          alg_bytes=B * (4 * D * N + 8 * N * k))
    return (idx, dist) if return_dist else idx


@torch.no_grad()
def knn_cdist(xyz: torch.Tensor, k: int, return_dist: bool = False):
    """torch.cdist(xyz, xyz).topk(k, largest=False) without the [B,N,N] matrix
    (attention_modules.py:584-586).  xyz [B,N,3] -> LongTensor [B,N,k]."""
    xyz = _f32(xyz, "xyz")
    B, N, C = xyz.shape
    if C != 3:
        raise ValueError("knn_cdist expects 3-D coordinates")
    idx = torch.empty(B, N, k, dtype=torch.long, device=xyz.device)
    dist = torch.empty(B, N, k, dtype=torch.float32, device=xyz.device) if return_dist else None
    _call("pcb_knn_cdist_f32", xyz.device, xyz.data_ptr(), B, N, int(k), idx.data_ptr(),
          dist.data_ptr() if return_dist else None, alg_bytes=B * (12 * N + 8 * N * k))
    return (idx, dist) if return_dist else idx


# ---------------------------------------------------------------------------------------------
# differentiable gathers
# ---------------------------------------------------------------------------------------------
class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx, clamp):
        points = _f32(points, "points")
        idx = _i64(idx, "idx")
        B, N, C = points.shape
        M = idx.numel() // B
        out = torch.empty(*idx.shape, C, dtype=torch.float32, device=points.device)
        err = None if clamp else _err_counter(points.device)
        _call("pcb_gather_f32", points.device, points.data_ptr(), idx.data_ptr(), B, N, C, M, int(clamp),
              out.data_ptr(), err.data_ptr() if err is not None else None,
              alg_bytes=B * (4 * N * C + 8 * M + 4 * M * C))
        if not clamp and _STRICT:
            check_index_errors(points.device)
        ctx.save_for_backward(idx)
        ctx.shape = (B, N, C, M, int(clamp))
        return out

    @staticmethod
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        B, N, C, M, clamp = ctx.shape
        gout = _f32(gout, "grad")
        gpoints = torch.zeros(B, N, C, dtype=torch.float32, device=gout.device)
        _call("pcb_gather_bwd_f32", gout.device, gout.data_ptr(), idx.data_ptr(), B, N, C, M, clamp,
              gpoints.data_ptr(), alg_bytes=B * (4 * N * C + 8 * M + 4 * M * C))
        return gpoints, None, None


def gather(points: torch.Tensor, idx: torch.Tensor, clamp: bool = False) -> torch.Tensor:
    """index_points: points [B,N,C], idx [B,...] -> [B,...,C]."""
    return _Gather.apply(points, idx, bool(clamp))


class _GroupPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, points, new_xyz, idx, xyz_first, points_cf, clamp, pad_to):
        xyz = _f32(xyz, "xyz")
        new_xyz = _f32(new_xyz, "new_xyz")
        idx = _i64(idx, "idx")
        B, N, _ = xyz.shape
        _, S, K = idx.shape
        # under bf16 autocast the consumer is a bf16 GEMM: emit the grouped tensor in bf16 directly
        bf16 = torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16
        D = 0 if points is None else (points.shape[1] if points_cf else points.shape[2])
        pitch = -(-(3 + D) // pad_to) * pad_to
        # bf16 feature rows of the previous layer are gathered as they are (no fp32 round trip)
        pts_bf16 = bf16 and D > 0 and points.dtype == torch.bfloat16 and pitch % 8 == 0 and points.is_cuda
        if D:
            points = (points if points.is_contiguous() else points.contiguous()) if pts_bf16 else _f32(points, "points")
        out = torch.empty(B, S, K, pitch, dtype=torch.bfloat16 if bf16 else torch.float32, device=xyz.device)
        ab = B * (4 * N * 3 + (points.element_size() * N * D if D else 0) + 12 * S + 8 * S * K + out.element_size() * S * K * pitch)
        if bf16:
            _call("pcb_group_points_bf16", xyz.device, xyz.data_ptr(), points.data_ptr() if D else None, int(pts_bf16),
                  new_xyz.data_ptr(), idx.data_ptr(), B, N, S, K, D, int(xyz_first), int(points_cf), int(clamp), pitch,
                  out.data_ptr(), alg_bytes=ab)
        else:
            _call("pcb_group_points_f32", xyz.device, xyz.data_ptr(), points.data_ptr() if D else None,
                  new_xyz.data_ptr(), idx.data_ptr(), B, N, S, K, D, int(xyz_first), int(points_cf), int(clamp), pitch,
                  out.data_ptr(), alg_bytes=ab)
        ctx.save_for_backward(idx)
        ctx.meta = (B, N, S, K, D, int(xyz_first), int(points_cf), int(clamp), pitch)
        return out

    @staticmethod
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        B, N, S, K, D, xyz_first, points_cf, clamp, pitch = ctx.meta
        gpoints = None
        if D and ctx.needs_input_grad[1]:
            bf16 = gout.dtype == torch.bfloat16
            gout = gout.contiguous() if bf16 else _f32(gout, "grad")
            shape = (B, D, N) if points_cf else (B, N, D)
            gpoints = torch.zeros(shape, dtype=torch.float32, device=gout.device)
            _call("pcb_group_points_bwd_bf16" if bf16 else "pcb_group_points_bwd_f32", gout.device, gout.data_ptr(), idx.data_ptr(), B, N, S, K, D,
                  xyz_first, points_cf, clamp, pitch, gpoints.data_ptr(),
                  alg_bytes=B * (4 * N * D + 8 * S * K + gout.element_size() * S * K * D))
        return None, gpoints, None, None, None, None, None, None


class _GroupPointsMulti(torch.autograd.Function):
    """group_points for several neighbour lists around the same centroids (the scales of a multi-scale module):
    forward = one grouping kernel per list; backward = ONE zero-filled fp32 gradient that every list's
    scatter-add kernel accumulates into (instead of a zero fill + dtype cast + add per list)."""

    @staticmethod
    def forward(ctx, xyz, points, new_xyz, xyz_first, points_cf, clamp, pad_to, *idxs):
        outs, metas = [], []
        for idx in idxs:
            sub = type("ctx", (), {})()
            sub.save_for_backward = lambda *a, sub=sub: setattr(sub, "saved", a)
            outs.append(_GroupPoints.forward(sub, xyz, points, new_xyz, idx, xyz_first, points_cf, clamp, pad_to))
            metas.append(sub.meta)
        ctx.save_for_backward(*[_i64(i, "idx") for i in idxs])
        ctx.metas = metas
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        idxs = ctx.saved_tensors
        gpoints = None
        if ctx.needs_input_grad[1] and ctx.metas[0][4]:
            B, N, _, _, D, xyz_first, points_cf, clamp, _ = ctx.metas[0]
            dev = idxs[0].device
            gpoints = torch.zeros((B, D, N) if points_cf else (B, N, D), dtype=torch.float32, device=dev)
            for idx, gout, meta in zip(idxs, gouts, ctx.metas):
                if gout is None:
                    continue
                _, _, S, K, _, _, _, _, pitch = meta
                bf16 = gout.dtype == torch.bfloat16
                gout = gout.contiguous() if bf16 else _f32(gout, "grad")
                _call("pcb_group_points_bwd_bf16" if bf16 else "pcb_group_points_bwd_f32", dev, gout.data_ptr(), idx.data_ptr(),
                      B, N, S, K, D, xyz_first, points_cf, clamp, pitch, gpoints.data_ptr(),
                      alg_bytes=B * (4 * N * D + 8 * S * K + gout.element_size() * S * K * D))
        return (None, gpoints, None, None, None, None, None) + (None,) * len(idxs)


def group_points_multi(xyz, points, new_xyz, idxs, xyz_first: bool = True, points_cf: bool = False,
                       clamp: bool = False, pad_to: int = 1):
    """[group_points(xyz, points, new_xyz, idx, ...) for idx in idxs] with one shared gradient buffer for `points`."""
    return _GroupPointsMulti.apply(xyz, points, new_xyz, bool(xyz_first), bool(points_cf), bool(clamp), int(pad_to), *idxs)


def group_points(xyz, points, new_xyz, idx, xyz_first: bool = True, points_cf: bool = False,
                 clamp: bool = False, pad_to: int = 1) -> torch.Tensor:
    """Fused index_points(xyz, idx) - new_xyz, index_points(points, idx) and concat:
    -> [B,S,K,3+D].  `points` is [B,N,D], or [B,D,N] with points_cf=True, or None.
    pad_to > 1 rounds the channel count up to a multiple of pad_to with zero columns: 3+D is 99,
    259, 515 in the MSG network, and rows that are not 16-byte aligned send the following GEMM to
    cuBLAS's unaligned legacy kernels (`linear_rows` accepts the padded rows)."""
    return _GroupPoints.apply(xyz, points, new_xyz, idx, bool(xyz_first), bool(points_cf), bool(clamp), int(pad_to))


class _GraphFeature(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        x = _f32(x, "x")
        idx = _i64(idx, "idx")
        B, D, N = x.shape
        k = idx.shape[2]
        out = torch.empty(B, 2 * D, N, k, dtype=torch.float32, device=x.device)
        _call("pcb_graph_feature_f32", x.device, x.data_ptr(), idx.data_ptr(), B, D, N, k, out.data_ptr(),
              alg_bytes=B * (4 * N * D + 8 * N * k + 8 * D * N * k))
        ctx.save_for_backward(idx)
        ctx.meta = (B, D, N, k)
        return out

    @staticmethod
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        B, D, N, k = ctx.meta
        gout = _f32(gout, "grad")
        gx = torch.zeros(B, D, N, dtype=torch.float32, device=gout.device)
        _call("pcb_graph_feature_bwd_f32", gout.device, gout.data_ptr(), idx.data_ptr(), B, D, N, k,
              gx.data_ptr(), alg_bytes=B * (4 * N * D + 8 * N * k + 8 * D * N * k))
        return gx, None


def graph_feature(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """x [B,D,N], idx [B,N,k] -> [B,2D,N,k] = cat(x[idx] - x, x) (DGCNN.py:72-109)."""
    return _GraphFeature.apply(x, idx)


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points2, idx, weight, channels_first):
        points2 = _f32(points2, "points2")
        idx = _i64(idx, "idx")
        weight = _f32(weight, "weight")
        B, N, k = idx.shape
        if channels_first:
            _, D, S = points2.shape
            out = torch.empty(B, D, N, dtype=torch.float32, device=points2.device)
        else:
            _, S, D = points2.shape
            out = torch.empty(B, N, D, dtype=torch.float32, device=points2.device)
        _call("pcb_interpolate_f32", points2.device, points2.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, N,
              S, D, k, int(channels_first), out.data_ptr(), alg_bytes=B * (4 * S * D + 12 * k * N + 4 * N * D))
        ctx.save_for_backward(idx, weight)
        ctx.meta = (B, N, S, D, k, int(channels_first))
        return out

    @staticmethod
    def backward(ctx, gout):
        idx, weight = ctx.saved_tensors
        B, N, S, D, k, cf = ctx.meta
        gout = _f32(gout, "grad")
        shape = (B, D, S) if cf else (B, S, D)
        gp2 = torch.zeros(shape, dtype=torch.float32, device=gout.device)
        _call("pcb_interpolate_bwd_f32", gout.device, gout.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, N, S,
              D, k, cf, gp2.data_ptr(), alg_bytes=B * (4 * S * D + 12 * k * N + 4 * N * D))
        return gp2, None, None, None


def three_interpolate(points2: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor,
                      channels_first: bool = False) -> torch.Tensor:
    """sum_j weight[b,n,j] * points2[b, idx[b,n,j]]: points2 [B,S,D] -> [B,N,D], or with
    channels_first [B,D,S] -> [B,D,N]."""
    return _Interpolate.apply(points2, idx, weight, bool(channels_first))


class _FpConcat(torch.autograd.Function):
    """Input rows of a feature-propagation MLP under bf16 autocast, one kernel instead of
    interpolate (fp32) + torch.cat + cast: out [B,N,pitch] bf16 = [points1 | interp(points2) | 0]."""

    @staticmethod
    def forward(ctx, points1, points2, idx, weight, pitch):
        idx = _i64(idx, "idx")
        weight = _f32(weight, "weight")
        B, N, k = idx.shape
        if not points2.is_contiguous():
            points2 = points2.contiguous()
        S, D2 = points2.shape[1], points2.shape[2]
        D1 = 0
        if points1 is not None:
            if not points1.is_contiguous():
                points1 = points1.contiguous()
            D1 = points1.shape[2]
        dev = points2.device
        out = torch.empty(B, N, pitch, dtype=torch.bfloat16, device=dev)
        _call("pcb_fp_concat_bf16", dev, points1.data_ptr() if D1 else None,
              int(D1 > 0 and points1.dtype == torch.bfloat16), points2.data_ptr(),
              int(points2.dtype == torch.bfloat16), idx.data_ptr(), weight.data_ptr(), B, N, S, D1, D2, k, pitch,
              out.data_ptr(),
              alg_bytes=B * ((points1.element_size() * N * D1 if D1 else 0) + points2.element_size() * S * D2
                             + 12 * k * N + 2 * N * pitch))
        ctx.save_for_backward(idx, weight)
        ctx.meta = (B, N, S, D1, D2, k, pitch, points1.dtype if D1 else None, points2.dtype)
        return out

    @staticmethod
    def backward(ctx, gout):
        idx, weight = ctx.saved_tensors
        B, N, S, D1, D2, k, pitch, dt1, dt2 = ctx.meta
        gout = gout.contiguous()
        if gout.dtype != torch.bfloat16:
            gout = gout.to(torch.bfloat16)
        gp1 = gp2 = None
        if D1 and ctx.needs_input_grad[0]:
            gp1 = gout[..., :D1].to(dt1)
        if ctx.needs_input_grad[1]:
            gp2 = torch.zeros(B, S, D2, dtype=torch.float32, device=gout.device)
            _call("pcb_fp_concat_bwd_bf16", gout.device, gout.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, N, S,
                  D1, D2, k, pitch, gp2.data_ptr(), alg_bytes=B * (4 * S * D2 + 12 * k * N + 2 * N * D2))
            if dt2 != torch.float32:
                gp2 = gp2.to(dt2)
        return gp1, gp2, None, None, None


def fp_concat_supported(points1, points2) -> bool:
    ok = lambda t: t.is_cuda and t.dtype in (torch.float32, torch.bfloat16) and t.shape[2] % 2 == 0
    return (torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16
            and ok(points2) and (points1 is None or ok(points1)) and os.environ.get("PCB_NO_FPCONCAT", "0") != "1")


def fp_concat(points1, points2, idx, weight, pad_to: int = 8) -> torch.Tensor:
    """[points1 [B,N,D1] | three_interpolate(points2 [B,S,D2], idx, weight) | zero pad] as bf16 rows
    [B,N,pitch] (pointnet_util.py:325-340 fused); points1 may be None."""
    D = points2.shape[2] + (points1.shape[2] if points1 is not None else 0)
    return _FpConcat.apply(points1, points2, idx, weight, -(-D // pad_to) * pad_to)


# ---------------------------------------------------------------------------------------------
# BatchNorm (batch statistics) + ReLU (+ max over the neighbour axis) on point-major rows
# ---------------------------------------------------------------------------------------------
def _act_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.bfloat16:
        return 1
    raise _lib.PcbError(f"bn_relu_rows supports fp32 / bf16 activations, got {t.dtype}")


def _bn_work(C: int, dev) -> torch.Tensor:
    """fp32 scratch of one BN launch: 3*C results + per-CTA partial sums (csrc/bn_rows.cu)."""
    return torch.empty(_lib.lib().pcb_bn_work_floats(C), dtype=torch.float32, device=dev)


class _BnReluRows(torch.autograd.Function):
    """z = max_k relu(BN_train(y + bias)) on rows.  y [M,C] is the bias-free GEMM output.
    One cooperative kernel forward, one backward."""

    @staticmethod
    def forward(ctx, y, bias, gamma, beta, running_mean, running_var, momentum, eps, relu, pool_k, out=None):
        # out: optional [M/pool_k, C] column slice (unit column stride) of a wider buffer to write into -- the
        # scales of a multi-scale module fill their concatenated output directly (`join_columns`)
        if not y.is_contiguous():
            y = y.contiguous()
        M, C = y.shape                    # C: row pitch; Cv real channels (y may carry zero pad columns)
        Cv = gamma.shape[0]
        dt = _act_dtype(y)
        dev = y.device
        stats = torch.empty(2, C, dtype=torch.float32, device=dev)          # mean, invstd of bias-free y
        Mout = M // pool_k
        if out is None:
            out = torch.empty(Mout, C, dtype=y.dtype, device=dev)
        argmax = torch.empty(Mout, C, dtype=torch.uint8, device=dev) if pool_k > 1 else None
        g32, b32 = gamma.float(), beta.float()
        work = _bn_work(C, dev)
        esz = y.element_size()
        _call("pcb_bn_fwd_rows", dev, y.data_ptr(), dt, M, C, Cv, int(pool_k),
              bias.data_ptr() if bias is not None else None, g32.data_ptr(), b32.data_ptr(), float(eps),
              float(momentum), running_mean.data_ptr() if running_mean is not None else None,
              running_var.data_ptr() if running_var is not None else None, int(relu), stats[0].data_ptr(),
              stats[1].data_ptr(), out.data_ptr(), out.stride(0), argmax.data_ptr() if argmax is not None else None,
              work.data_ptr(), alg_bytes=(y.numel() + out.numel()) * esz + (Mout * C if pool_k > 1 else 0))
        ctx.save_for_backward(y, stats, g32, b32, argmax)
        ctx.meta = (M, C, Cv, dt, int(relu), int(pool_k), bias is not None)
        return out

    @staticmethod
    def backward(ctx, gz):
        y, stats, g32, b32, argmax = ctx.saved_tensors
        M, C, Cv, dt, relu, pool_k, has_bias = ctx.meta
        if gz.dtype != y.dtype:
            gz = gz.to(y.dtype)
        if not (gz.dim() == 2 and gz.stride(1) == 1 and gz.stride(0) >= C and gz.stride(0) % 8 == 0
                and gz.data_ptr() % 16 == 0):
            gz = gz.contiguous()          # a column slice of a concatenated gradient is read in place (gz_pitch)
        gy = torch.empty_like(y)
        work = _bn_work(C, y.device)
        _call("pcb_bn_bwd_rows", y.device, gz.data_ptr(), gz.stride(0), y.data_ptr(),
              argmax.data_ptr() if argmax is not None else None,
              dt, M, C, Cv, pool_k, stats[0].data_ptr(), stats[1].data_ptr(), g32.data_ptr(), b32.data_ptr(), relu,
              work.data_ptr(), gy.data_ptr(),
              alg_bytes=(2 * y.numel() + gz.numel()) * y.element_size() + (gz.numel() if pool_k > 1 else 0))
        s = work[:3 * C].view(3, C)
        # s[2] = d/d(conv bias) = sum_rows gy: zero up to rounding, as in the reference, where the bias of a
        # conv that feeds a training-mode BN gets a noise gradient
        return gy, (s[2, :Cv] if has_bias else None), s[1, :Cv], s[0, :Cv], None, None, None, None, None, None, None


class _JoinColumns(torch.autograd.Function):
    """buf [M, C_total] whose column slices were filled by the given parts (bn_relu_rows(out=buf[:, a:b])): returns
    buf with the parts as its autograd inputs -- torch.cat without the copy; backward hands each part its column
    slice of the gradient as a view."""

    @staticmethod
    def forward(ctx, buf, *parts):
        ctx.widths = [p.shape[1] for p in parts]
        return buf

    @staticmethod
    def backward(ctx, g):
        outs, off = [None], 0
        for w in ctx.widths:
            outs.append(g[:, off:off + w])
            off += w
        return tuple(outs)


def join_columns(buf, parts):
    return _JoinColumns.apply(buf, *parts)


def bn_relu_rows(y, bias, bn, relu=True, pool_k=1, out=None):
    """BatchNorm with batch statistics (+ReLU, + max over groups of `pool_k` consecutive rows) of
    the bias-free GEMM output y [M,C]; updates bn's running statistics like nn.BatchNorm does."""
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        if _step_ctx is not None:
            _step_ctx.counters.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked.add_(1)
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    if out is not None and not (out.dtype == y.dtype and out.shape == (y.shape[0] // pool_k, y.shape[1])
                                and out.stride(1) == 1 and out.stride(0) % 8 == 0 and out.data_ptr() % 16 == 0):
        out = None
    return _BnReluRows.apply(y, bias, bn.weight, bn.bias, rm, rv, bn.momentum, bn.eps, bool(relu), int(pool_k), out)


def bn_rows_supported(y, bn, pool_k=1) -> bool:
    return (y.is_cuda and y.dim() == 2 and y.shape[1] % 4 == 0 and y.dtype in (torch.float32, torch.bfloat16)
            and bn.training and bn.momentum is not None and bn.affine and y.shape[0] % pool_k == 0
            and y.shape[0] > 1 and pool_k <= 255 and y.numel() // 4 < 2 ** 31)


@torch.no_grad()
def eigvalsh3(a: torch.Tensor) -> torch.Tensor:
    """torch.linalg.eigvalsh for a batch of symmetric 3x3 matrices [...,3,3] -> [...,3] ascending, without the
    host synchronisation of the cuSOLVER path (CUDA-graph capturable); float64 closed form per matrix."""
    a = _f32(a, "a")
    if a.shape[-2:] != (3, 3):
        raise ValueError("eigvalsh3 expects [...,3,3]")
    M = a.numel() // 9
    out = torch.empty(*a.shape[:-2], 3, dtype=torch.float32, device=a.device)
    _call("pcb_eigvalsh3_f32", a.device, a.data_ptr(), M, out.data_ptr(), alg_bytes=48 * M)
    return out


@torch.no_grad()
def structure_rows(xyz: torch.Tensor, idx: torch.Tensor, freqs, grid_size: float = 1.0, bf16: bool = False,
                   rows: bool = True, feat: bool = False):
    """Input rows of BridgeStructureEncoding (attention_modules.py:552-613, 622-687) in one kernel: xyz [B,N,3],
    idx [B,N,k] (k nearest neighbours) -> rows [B*N*k, pitch] = [sin/cos encoding (6F) | neighbour - centre (3) | 13
    structure statistics | zero pad to a multiple of 8] (fp32, or bf16 for the training MLP) and / or the statistics
    alone [B,N,13].  No host synchronisation (the reference's eigh has one), nothing differentiable."""
    xyz = _f32(xyz, "xyz")
    idx = _i64(idx, "idx")
    B, N, k = idx.shape
    fr = [float(f) for f in (freqs.tolist() if isinstance(freqs, torch.Tensor) else freqs)]
    F = len(fr)
    pitch = _ru8(6 * F + 16)
    dev = xyz.device
    out_rows = torch.empty(B * N * k, pitch, dtype=torch.bfloat16 if bf16 else torch.float32, device=dev) if rows else None
    out_feat = torch.empty(B, N, 13, dtype=torch.float32, device=dev) if feat else None
    carr = (ctypes.c_float * max(F, 1))(*fr)
    _call("pcb_structure_rows_f32", dev, xyz.data_ptr(), idx.data_ptr(), B, N, k, ctypes.cast(carr, ctypes.c_void_p), F,
          float(grid_size), int(bf16), pitch, out_rows.data_ptr() if rows else None, out_feat.data_ptr() if feat else None,
          alg_bytes=B * N * (12 + 8 * k) + (out_rows.numel() * out_rows.element_size() if rows else 0))
    return out_rows, out_feat


# ---------------------------------------------------------------------------------------------
# mean NLL of the segmentation head straight from the classifier's logits rows
# ---------------------------------------------------------------------------------------------
class LogitRows:
    """What a segmentation head hands back instead of log-probabilities while `head_logits_mode` is on: the
    bias-free classifier rows [M, pitch] (zero pad columns), the classifier bias and the real class count."""

    def __init__(self, rows, bias, classes, B, N):
        self.rows, self.bias, self.classes, self.B, self.N = rows, bias, classes, B, N

    def log_probs(self):
        x = self.rows[:, :self.classes].float()
        if self.bias is not None:
            x = x + self.bias
        return torch.log_softmax(x, dim=-1).view(self.B, self.N, -1)


_head_logits = False


class head_logits_mode:
    """Context in which the drop-in segmentation heads return `LogitRows` (training runner: the loss kernel
    consumes the logits directly; log_softmax, the slice copy and ATen's NLL kernels disappear)."""

    def __enter__(self):
        global _head_logits
        self._prev, _head_logits = _head_logits, True

    def __exit__(self, *exc):
        global _head_logits
        _head_logits = self._prev
        return False


def head_logits_enabled() -> bool:
    return _head_logits


class _NllRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, bias, labels, classes):
        rows = rows if rows.is_contiguous() else rows.contiguous()
        labels = _i64(labels.reshape(-1), "labels")
        M, pitch = rows.shape
        dt = _act_dtype(rows)
        b32 = bias.float().contiguous() if bias is not None else None
        nblk = _lib.lib().pcb_nll_rows_blocks(M)
        partial = torch.empty(nblk, dtype=torch.float32, device=rows.device)
        _call("pcb_nll_rows_fwd", rows.device, rows.data_ptr(), dt, b32.data_ptr() if b32 is not None else None,
              labels.data_ptr(), M, int(classes), pitch, partial.data_ptr(), alg_bytes=rows.numel() * rows.element_size() + 8 * M)
        ctx.save_for_backward(rows, b32, labels)
        ctx.meta = (M, pitch, dt, int(classes))
        return partial.sum() * (1.0 / M)

    @staticmethod
    def backward(ctx, g):
        rows, b32, labels = ctx.saved_tensors
        M, pitch, dt, classes = ctx.meta
        g = g.reshape(1).float().contiguous()
        dx = torch.empty_like(rows)
        gb = torch.zeros(classes, dtype=torch.float32, device=rows.device) if b32 is not None else None
        _call("pcb_nll_rows_bwd", rows.device, rows.data_ptr(), dt, b32.data_ptr() if b32 is not None else None,
              labels.data_ptr(), M, classes, pitch, g.data_ptr(), dx.data_ptr(), gb.data_ptr() if gb is not None else None,
              alg_bytes=2 * rows.numel() * rows.element_size() + 8 * M)
        return dx, gb, None, None


def nll_logit_rows(lr: "LogitRows", labels: torch.Tensor) -> torch.Tensor:
    """mean_r( -log_softmax(rows[r, :classes] + bias)[labels[r]] ) -- F.nll_loss(F.log_softmax(logits), labels)."""
    return _NllRows.apply(lr.rows, lr.bias, labels, lr.classes)


# ---------------------------------------------------------------------------------------------
# fused set-abstraction / EdgeConv block for inference (tcgen05 tensor cores)
# ---------------------------------------------------------------------------------------------
def _ru16(n: int) -> int:
    return (n + 15) // 16 * 16


class PackedMLP:
    """Eval-mode (conv 1x1 -> BatchNorm -> activation) stack folded to W', b' per layer and packed
    into the shared-memory image the fused kernel expects: bf16, [K/8][N][8], zero padded."""

    def __init__(self, convs, bns, c_in: int):
        import ctypes
        dev = convs[0].weight.device
        kdim = [_ru16(c_in)]
        blobs, biases = [], []
        width = c_in
        with torch.no_grad():
            for conv, bn in zip(convs, bns):
                w = conv.weight.flatten(1).float()                      # [out, in]
                assert w.shape[1] == width
                b = conv.bias.float() if conv.bias is not None else torch.zeros(w.shape[0], device=dev)
                scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
                w = w * scale[:, None]
                b = (b - bn.running_mean.float()) * scale + bn.bias.float()
                n_pad, k_pad = _ru16(w.shape[0]), kdim[-1]
                wp = torch.zeros(n_pad, k_pad, device=dev)
                wp[:w.shape[0], :w.shape[1]] = w
                blobs.append(wp.view(n_pad, k_pad // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16).reshape(-1))
                bp = torch.zeros(n_pad, device=dev)
                bp[:w.shape[0]] = b
                biases.append(bp)
                kdim.append(n_pad)
                width = w.shape[0]
        blob = torch.cat(blobs)
        pad = (-blob.numel() * 2) % 128
        if pad:
            blob = torch.cat([blob, torch.zeros(pad // 2, dtype=torch.bfloat16, device=dev)])
        self.wblob = blob.contiguous()
        self.bias = torch.cat(biases).contiguous()
        self.kdim = kdim
        self.kdim_c = (ctypes.c_int * len(kdim))(*kdim)
        self.cout = width
        self.nlayers = len(convs)
        maxw = max(kdim)
        a_bytes = (128 * (maxw + 8) * 2 + 127) // 128 * 128
        self.smem = self.wblob.numel() * 2 + 2 * a_bytes + self.bias.numel() * 4 + 128
        self.ok = self.smem <= 200 * 1024 and maxw <= 256 and self.nlayers <= 3


class FoldedMLP:
    """Eval-mode (conv 1x1 -> BatchNorm -> activation) stack folded to W', b' per layer as row-major bf16 weights
    [n8, k8] (zero padded) + fp32 bias [n8]: the operands of the bias / activation epilogue of the tcgen05 GEMM
    (csrc/gemm_rows.cu, EPI_BIAS) -- every layer width, any number of layers."""

    def __init__(self, convs, bns):
        dev = convs[0].weight.device
        self.w, self.b, self.n = [], [], []
        with torch.no_grad():
            for conv, bn in zip(convs, bns):
                w = conv.weight.flatten(1).float()
                b = conv.bias.float() if conv.bias is not None else torch.zeros(w.shape[0], device=dev)
                if bn is not None:
                    scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
                    w = w * scale[:, None]
                    b = (b - bn.running_mean.float()) * scale + bn.bias.float()
                n8, k8 = _ru8(w.shape[0]), _ru8(w.shape[1])
                wp = torch.zeros(n8, k8, device=dev)
                wp[:w.shape[0], :w.shape[1]] = w
                bp = torch.zeros(n8, device=dev)
                bp[:w.shape[0]] = b
                self.w.append(wp.to(torch.bfloat16).contiguous())
                self.b.append(bp.contiguous())
                self.n.append(w.shape[0])
        self.cout = self.n[-1]


def folded_mlp(owner, key, convs, bns):
    """Cache of FoldedMLP per module, invalidated like packed_mlp."""
    ver = (_param_generation,) + tuple(t._version for m in list(convs) + [b for b in bns if b is not None]
                                       for t in list(m.parameters()) + list(m.buffers()))
    cache = owner.__dict__.setdefault("_pcb_folded", {})
    hit = cache.get(key)
    if hit is None or hit[0] != ver:
        hit = (ver, FoldedMLP(convs, bns))
        cache[key] = hit
    return hit[1]


@torch.no_grad()
def mlp_rows_infer(x: torch.Tensor, folded: FoldedMLP, pool_k: int = 1, act: int = 1, slope: float = 0.0,
                   last_act: bool = True) -> torch.Tensor:
    """Inference shared MLP on rows: x [M, Cin] -> [M / pool_k, n8 of the last layer] bf16, every layer one tcgen05
    GEMM whose epilogue adds the folded BatchNorm bias and applies the activation; the last layer's epilogue also
    takes the max over every `pool_k` consecutive rows (the neighbour axis) when pool_k divides 128 -- the [M, C]
    activation of the widest layer never reaches memory.  Replaces pointnet_util.py:213-217 / 273-279 / 343-345 in
    evaluation mode for every layer shape (the one-kernel block of csrc/sa_fused.cu covers the narrow ones)."""
    if x.dtype != torch.bfloat16:
        x = x.to(torch.bfloat16)
    if x.shape[1] % 8:
        x = torch.nn.functional.pad(x, (0, -x.shape[1] % 8))
    if not _rows_ok(x):
        x = x.contiguous()
    M = x.shape[0]
    L = len(folded.w)
    fuse_pool = pool_k > 1 and 128 % pool_k == 0 and M % pool_k == 0
    for l, (w, b, n) in enumerate(zip(folded.w, folded.b, folded.n)):
        last = l == L - 1
        pk = pool_k if (last and fuse_pool) else 1
        n8 = w.shape[0]
        K = min(x.shape[1], w.shape[1])
        y = torch.empty(M // pk, n8, dtype=torch.bfloat16, device=x.device)
        _call("pcb_linear_bias_act_rows_bf16", x.device, x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), M, n8, n8, K,
              b.data_ptr(), n, int(act if (last_act or not last) else 0), float(slope), pk, y.data_ptr(), y.stride(0),
              alg_bytes=2 * M * K + 2 * (M // pk) * n8 + 2 * w.numel())
        x = y
    if pool_k > 1 and not fuse_pool:
        x = x.view(-1, pool_k, x.shape[1]).max(dim=1)[0]
    return x


# Bumped by everything that rewrites parameters or BatchNorm buffers behind autograd's back (the step runner's fused
# Adam and BN kernels write through raw pointers, CUDA-graph replays run no Python at all): tensor `_version`
# counters do not see those updates, so the folded-weight cache below also keys on this generation.
_param_generation = 0


def bump_param_generation() -> None:
    global _param_generation
    _param_generation += 1


def packed_mlp(owner, key, convs, bns, c_in):
    """Cache of PackedMLP per module, invalidated when a parameter / running statistic changes."""
    ver = (_param_generation,) + tuple(t._version for m in list(convs) + list(bns)
                                       for t in list(m.parameters()) + list(m.buffers()))
    cache = owner.__dict__.setdefault("_pcb_packed", {})
    hit = cache.get(key)
    if hit is None or hit[0] != ver:
        hit = (ver, PackedMLP(convs, bns, c_in))
        cache[key] = hit
    return hit[1]


def fused_inference_enabled() -> bool:
    """The fused tcgen05 block computes in bf16: it is used only for eval-mode forwards under
    bf16 autocast without autograd (fp32 parity runs keep the exact unfused path)."""
    return (not torch.is_grad_enabled()) and torch.is_autocast_enabled() and \
        torch.get_autocast_dtype('cuda') == torch.bfloat16 and os.environ.get("PCB_NO_FUSED", "0") != "1"


@torch.no_grad()
def sa_fused(xyz, points, new_xyz, idx, packed: PackedMLP, xyz_first=True, mode=0, slope=0.0, out_bf16=True):
    """Fused gather -> shared MLP (folded BN) -> max over neighbours.  mode 0: xyz [B,N,3], points [B,N,D] | None,
    new_xyz [B,S,3], idx [B,S,K]; mode 1 (EdgeConv): points [B,N,D], idx [B,N,K].  -> [B*S, cout]."""
    idx = _i64(idx, "idx")
    B, S, K = idx.shape
    if points is not None:
        points = _f32(points, "points")
        N, D = points.shape[1], points.shape[2]
    else:
        N, D = xyz.shape[1], 0
    dev = idx.device
    if mode == 0:
        xyz = _f32(xyz, "xyz")
        new_xyz = _f32(new_xyz, "new_xyz")
    out = torch.empty(B * S, packed.cout, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
    flops = 2 * B * S * K * sum(a * b for a, b in zip(packed.kdim[:-1], packed.kdim[1:]))
    _call("pcb_sa_fused_bf16", dev, xyz.data_ptr() if mode == 0 else None, points.data_ptr() if D else None,
          new_xyz.data_ptr() if mode == 0 else None, idx.data_ptr(), B, N, S, K, D, int(mode), int(xyz_first),
          packed.nlayers, packed.kdim_c, packed.cout, packed.wblob.data_ptr(), packed.bias.data_ptr(), float(slope),
          out.data_ptr(), int(out_bf16),
          alg_bytes=B * (4 * N * (3 + D) + 8 * S * K + 12 * S) + out.numel() * out.element_size())
    return out


# ---------------------------------------------------------------------------------------------
# y = x @ W^T on rows with a weight-gradient schedule for very tall activations
# ---------------------------------------------------------------------------------------------
class StepContext:
    """Batches the tiny per-layer bookkeeping kernels of one training step (34 BatchNorm layers in
    the MSG network): inside `with ctx:` every `num_batches_tracked += 1` is deferred to ONE
    multi-tensor add on exit, and with bf16 autocast the fp32 -> bf16 casts of the shared-MLP
    weights are ONE multi-tensor copy on entry into persistent shadows that `linear_rows` picks up
    (autocast would cast each weight on use, forward and again backward)."""

    def __init__(self, module: torch.nn.Module, bf16: bool, grad_views: dict | None = None,
                 flat_offsets: dict | None = None, flat_numel: int = 0):
        self.counters = []
        # id(parameter) -> fp32 view of the runner's flat gradient buffer (zeroed at the start of every
        # step): the weight-gradient kernel accumulates straight into it, the parameter's .grad stays None
        self.grad_views = grad_views or {}
        p0 = next(module.parameters(), None)
        params = [p for p in module.parameters() if p.dim() >= 2] if bf16 else []
        # shadows are [n rounded up to 8, k rounded up to 8], zero padded (written once, here), all carved out
        # of ONE bf16 buffer: pad columns match the zero-padded rows of group_points(pad_to=8), pad rows make
        # the OUTPUT rows 16-byte aligned (196 -> 200 channels) when linear_rows(pad_n=True)
        shapes = [(p.shape[0], p[0].numel()) for p in params]
        sizes = [(-(-n // 8) * 8) * (-(-k // 8) * 8) for n, k in shapes]
        total = sum(sizes)
        # second half of the buffer: the same weights transposed ([k8, n8]) -- the B operand of the data-gradient
        # GEMM (csrc/gemm_rows.cu reads both operands K-major)
        self.shadow_flat = torch.zeros(max(2 * total, 1), dtype=torch.bfloat16, device=p0.device) if params else None
        self.dense, self.dense_shadows, self.ragged = [], [], []
        self.by_id, self.by_id_t = {}, {}
        self.shadow_index_t = None
        index_t = None
        # element i of the runner's flat parameter buffer -> element of shadow_flat (or -1): lets the fused
        # Adam kernel (csrc/adam.cu) refresh the shadows while it writes the updated parameters
        self.shadow_index = None
        index = None
        if params and flat_offsets is not None:
            import numpy as np
            index = np.full(flat_numel, -1, dtype=np.int32)
            index_t = np.full(flat_numel, -1, dtype=np.int32)
        off = 0
        for p, (n, k), size in zip(params, shapes, sizes):
            k8 = -(-k // 8) * 8
            n8 = -(-n // 8) * 8
            sh = self.shadow_flat[off:off + size].view(-1, k8)
            self.by_id[id(p)] = sh
            self.by_id_t[id(p)] = self.shadow_flat[total + off:total + off + size].view(k8, n8)
            if k8 == k:
                self.dense.append(p)
                self.dense_shadows.append(sh[:n])
            else:
                self.ragged.append((p, sh[:n, :k]))
            if index is not None and id(p) in flat_offsets:
                j = np.arange(n * k, dtype=np.int64)
                index[flat_offsets[id(p)]:flat_offsets[id(p)] + n * k] = (off + (j // k) * k8 + (j % k)).astype(np.int32)
                index_t[flat_offsets[id(p)]:flat_offsets[id(p)] + n * k] = \
                    (total + off + (j % k) * n8 + (j // k)).astype(np.int32)
            off += size
        if index is not None:
            self.shadow_index = torch.from_numpy(index).to(p0.device)
            self.shadow_index_t = torch.from_numpy(index_t).to(p0.device)
        self.external_refresh = False      # True: the optimizer kernel keeps the shadows current (engine.Trainer)
        self.gviews_used = set()           # parameters whose weight gradient was written in place at least once

    def refresh_shadows(self):
        """fp32 parameters -> bf16 shadows (one multi-tensor copy + one copy per ragged weight)."""
        with torch.no_grad():
            if self.dense:
                torch._foreach_copy_(self.dense_shadows, [p.detach().flatten(1) for p in self.dense])
            for p, view in self.ragged:
                view.copy_(p.detach().flatten(1))
            for pid, sh in self.by_id.items():
                t = self.by_id_t[pid]
                t.copy_(sh.t())

    def __enter__(self):
        global _step_ctx
        self.counters = []
        self.side_streams = set()
        if not self.external_refresh:
            self.refresh_shadows()
        _step_ctx = self
        return self

    def __exit__(self, *exc):
        global _step_ctx
        _step_ctx = None
        for st in self.side_streams:                      # weight gradients launched next to the backward chain
            torch.cuda.current_stream().wait_stream(st)
        self.side_streams = set()
        if self.counters:
            torch._foreach_add_(self.counters, 1)
        return False

    def shadow(self, w):
        base = w._base if w._base is not None else w
        return self.by_id.get(id(base))

    def shadow_t(self, w):
        base = w._base if w._base is not None else w
        return self.by_id_t.get(id(base))

    def grad_view(self, w):
        base = w._base if w._base is not None else w
        v = self.grad_views.get(id(base))
        if v is not None:
            self.gviews_used.add(id(base))
        return v


_step_ctx = None
_WGRAD_CHUNK = int(os.environ.get("PCB_WGRAD_CHUNK", "2048"))


_WGRAD_KERNEL = os.environ.get("PCB_NO_WGRAD_KERNEL", "0") != "1"
_PAD_N = os.environ.get("PCB_NO_PAD_N", "0") != "1"          # debugging aid: keep unaligned output rows


def wgrad_rows_supported(gy, x) -> bool:
    return (gy.is_cuda and gy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and gy.dim() == 2 and x.dim() == 2
            and gy.is_contiguous() and x.is_contiguous() and gy.shape[1] % 8 == 0 and x.shape[1] % 8 == 0
            and gy.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0)


# The weight gradient of a layer is a leaf of the backward pass: nothing reads it before the optimizer.  Inside a step
# runner (StepContext) it is launched on a side stream right after the gradient rows it contracts exist, so it runs
# next to the data-gradient chain of the layers below instead of in front of it (the kernels of this step are 5-30 us
# each and mostly latency-bound: two of them share the SMs almost for free); the step joins the side stream before
# the gradients are packed (StepContext.__exit__).  In a captured step this is a fork / join of the graph.
_WGRAD_SIDE = os.environ.get("PCB_NO_WGRAD_SIDE_STREAM", "0") != "1"
_wgrad_streams: dict[int, torch.cuda.Stream] = {}


def _wgrad_stream(dev: torch.device) -> torch.cuda.Stream:
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _wgrad_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device=dev)
        _wgrad_streams[key] = st
    return st


# The scales of a multi-scale set-abstraction level are independent chains of small kernels: inside a step runner each
# scale after the first runs on its own stream (PCB_NO_SCALE_STREAMS=1 switches it off).
_SCALE_STREAMS = os.environ.get("PCB_NO_SCALE_STREAMS", "0") != "1"
_scale_streams: dict[tuple, torch.cuda.Stream] = {}


def scale_stream(dev: torch.device, i: int):
    """Stream of scale `i` of a multi-scale module, or None outside a step runner / when switched off."""
    if not _SCALE_STREAMS or _step_ctx is None or _timer is not None:
        return None
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), i)
    st = _scale_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(device=dev)
        _scale_streams[key] = st
    return st


@torch.no_grad()
def wgrad_rows(gy, x, k: int | None = None, out: torch.Tensor | None = None, n: int | None = None) -> torch.Tensor:
    """gy [M,Np][:, :n]^T @ x [M,Kp][:, :k] in fp32 (bf16 operands, rows = the long contraction dimension;
    pad columns of gy and x are zero).  `out` [n,k] fp32 is accumulated into when given."""
    M = gy.shape[0]
    n = gy.shape[1] if n is None else n
    k = x.shape[1] if k is None else k
    side = None
    if out is None:
        out = torch.zeros(n, k, dtype=torch.float32, device=gy.device)
    elif _WGRAD_SIDE and _step_ctx is not None and _timer is None:
        side = _wgrad_stream(gy.device)
    args = ("pcb_wgrad_rows_bf16", gy.device, gy.data_ptr(), x.data_ptr(), M, n, k, gy.shape[1], x.shape[1],
            out.data_ptr(), out.stride(0))
    nbytes = 2 * M * (gy.shape[1] + x.shape[1]) + 4 * n * k
    if side is None:
        _call(*args, alg_bytes=nbytes)
        return out
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _call(*args, alg_bytes=nbytes)
    gy.record_stream(side)                                # the caching allocator must not hand these rows out again
    x.record_stream(side)                                 # before the side stream has read them
    _step_ctx.side_streams.add(side)
    return out


# ---------------------------------------------------------------------------------------------
# tensor-core GEMMs on rows (csrc/gemm_rows.cu): tcgen05.mma, fp32 accumulation in TMEM, BatchNorm sums in the epilogue
# ---------------------------------------------------------------------------------------------
# The training MLPs run on the tcgen05 GEMMs of csrc/gemm_rows.cu (BatchNorm sums in the epilogue); PCB_OWN_GEMM=0 selects
# the round-1 composition (library GEMM + cooperative BN row kernels), kept as the yardstick of tests/test_gpu_gemm_rows.py
# and for fp32 parity runs.  Measurements: profiles/r2_gemm_rows.md.
_OWN_GEMM = os.environ.get("PCB_OWN_GEMM", "1") == "1"
_ticket_pool: dict[int, list] = {}


def _tickets(dev: torch.device) -> torch.Tensor:
    """Zeroed counter words for one statistics GEMM (the kernel leaves them zeroed).  1024 slots handed out round
    robin: launches that could overlap never share a slot.  Created on first use -- before any graph capture, because
    the warm-up steps of a runner are eager."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _ticket_pool.get(key)
    if st is None:
        n = _lib.lib().pcb_gemm_tickets()
        st = [torch.zeros(1024 * n, dtype=torch.int32, device=dev), 0, n]
        _ticket_pool[key] = st
    buf, pos, n = st
    st[1] = (pos + 1) % 1024
    return buf[pos * n:(pos + 1) * n]


def _ru8(n: int) -> int:
    return -(-n // 8) * 8


_max_groups = 0


def _gemm_max_groups() -> int:
    global _max_groups
    if not _max_groups:
        _max_groups = int(_lib.lib().pcb_gemm_max_groups())
    return _max_groups


def _weight_pair(w: torch.Tensor):
    """(w as bf16 [n8, k8], its transpose [k8, n8]), zero padded: the step runner's shadows, or cast on the spot."""
    w2 = w.flatten(1)
    if _step_ctx is not None:
        a, b = _step_ctx.shadow(w), _step_ctx.shadow_t(w)
        if a is not None and b is not None:
            return a, b
    n, k = w2.shape
    a = torch.zeros(_ru8(n), _ru8(k), dtype=torch.bfloat16, device=w.device)
    a[:n, :k] = w2.detach()
    return a, a.t().contiguous()


def _rows_ok(t: torch.Tensor) -> bool:
    return (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0
            and t.stride(0) >= t.shape[1] and t.data_ptr() % 16 == 0 and t.shape[1] % 8 == 0)


def own_gemm_supported(x: torch.Tensor) -> bool:
    return _OWN_GEMM and _rows_ok(x) and x.shape[0] > 0


@torch.no_grad()
def gemm_rows(x, w, n_out: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """x [M, K] @ w [Nw, Kw]^T -> [M, n_out] bf16 (contraction over min(K, Kw) columns; rows of w beyond Nw count as
    zero).  x, w: bf16 rows, 16-byte aligned, multiples of 8 columns."""
    M = x.shape[0]
    if out is None:
        out = torch.empty(M, n_out, dtype=torch.bfloat16, device=x.device)
    K = min(x.shape[1], w.shape[1])
    _call("pcb_linear_rows_bf16", x.device, x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), M, int(n_out),
          min(w.shape[0], n_out), K, out.data_ptr(), out.stride(0), alg_bytes=2 * M * (K + n_out) + 2 * w.numel())
    return out


class _LinearRows(torch.autograd.Function):
    """F.linear without bias for x [M,K], W [N,K] with M in the 10^5..10^6 range: forward and data gradient on the
    tcgen05 GEMM of csrc/gemm_rows.cu (bf16), weight gradient on the row kernel of csrc/wgrad.cu (the contraction over
    M).  x may carry zero pad columns beyond K (`group_points(pad_to=8)`): the weight is padded to match.  fp32 inputs
    (parity mode) go to the library GEMM."""

    @staticmethod
    def forward(ctx, x, w, w_lp, gview=None, pad_n=False):
        # w_lp: the weight already in x's dtype (StepContext shadow [n8, k8], zero padded) or None
        # gview: fp32 view of the step runner's flat gradient buffer for w, or None
        # pad_n: emit n8 output columns (zero pad columns) instead of n
        ctx.gview = gview
        n = w.shape[0]
        ctx.kw = w.shape[1]
        ctx.n = n
        n_out = _ru8(n) if pad_n else n
        ctx.own = own_gemm_supported(x) and n_out % 8 == 0
        if ctx.own:
            wl, wt = _weight_pair(w)
            ctx.save_for_backward(x, wt)
            return gemm_rows(x, wl, n_out)
        if w_lp is not None:
            wl = w_lp if pad_n else w_lp[:n]
        else:
            wl = w.to(x.dtype)
        if wl.shape[1] > x.shape[1]:                      # padded shadow, dense rows
            wl = wl[:, :x.shape[1]]
        elif wl.shape[1] < x.shape[1]:
            wl = torch.nn.functional.pad(wl, (0, x.shape[1] - wl.shape[1]))
        ctx.save_for_backward(x, wl)
        return torch.mm(x, wl.t())

    @staticmethod
    def backward(ctx, gy):
        x, wl = ctx.saved_tensors                         # own path: wl is the transposed weight [k8, n8]
        gx = gw = None
        gy = gy.contiguous()
        if gy.dtype != x.dtype:
            gy = gy.to(x.dtype)
        if ctx.needs_input_grad[0]:
            gx = gemm_rows(gy, wl, x.shape[1]) if ctx.own else torch.mm(gy, wl)
        if ctx.needs_input_grad[1]:
            M = x.shape[0]
            c = _WGRAD_CHUNK
            if _WGRAD_KERNEL and wgrad_rows_supported(gy, x):
                if ctx.gview is not None:                 # accumulate in place; .grad of the parameter stays None
                    wgrad_rows(gy, x, ctx.kw, out=ctx.gview.view(ctx.n, ctx.kw), n=ctx.n)
                    return gx, None, None, None, None
                gw = wgrad_rows(gy, x, ctx.kw, n=ctx.n)
            elif c and M % c == 0 and M // c >= 8:
                p = M // c
                part = torch.bmm(gy.view(p, c, -1).transpose(1, 2), x.view(p, c, -1))   # [p,N,K]
                gw = part.sum(dim=0, dtype=torch.float32)
            else:
                gw = torch.mm(gy.t(), x).float()
            if gw.shape[1] != ctx.kw or gw.shape[0] != ctx.n:
                gw = gw[:ctx.n, :ctx.kw]
        return gx, gw, None, None, None


def linear_rows(x, w, pad_n: bool = False):
    """x [M,K] @ w[N,K]^T under the ambient autocast dtype.  x may carry zero pad columns beyond
    K (group_points(pad_to=8)); the weight is padded to match and its gradient sliced back.
    pad_n: when N is not a multiple of 8 and a bf16 weight shadow exists (StepContext), the result
    has N rounded up to 8 columns, the extra ones zero -- 16-byte aligned rows for the next GEMM."""
    if torch.is_autocast_enabled():
        dt = torch.get_autocast_dtype("cuda")
        x = x if x.dtype == dt else x.to(dt)
    w_lp = _step_ctx.shadow(w) if (_step_ctx is not None and x.dtype == torch.bfloat16) else None
    gview = _step_ctx.grad_view(w) if (_step_ctx is not None and x.dtype == torch.bfloat16) else None
    pad = bool(pad_n and _PAD_N and (w_lp is not None or (own_gemm_supported(x) and x.dtype == torch.bfloat16)))
    return _LinearRows.apply(x, w, w_lp, gview, pad)


# ---------------------------------------------------------------------------------------------
# shared MLP on rows in training mode: (1x1 conv -> BatchNorm -> ReLU) x L [-> max over pool_k rows], one autograd node
# ---------------------------------------------------------------------------------------------
def mlp_rows_fused_supported(x, convs, bns, pool_k: int = 1) -> bool:
    """bf16 training path of the tensor-core GEMMs: CUDA rows under bf16 autocast, every layer a training-mode
    affine BatchNorm with momentum."""
    if not (_OWN_GEMM and x.is_cuda and x.dim() == 2 and x.shape[0] > 1 and x.shape[0] % pool_k == 0 and 1 <= pool_k <= 255):
        return False
    if not (x.dtype == torch.bfloat16 or (torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16)):
        return False
    if x.shape[1] > 8192 or x.numel() // 4 >= 2 ** 31:
        return False
    for conv, bn in zip(convs, bns):
        if not (bn.training and bn.affine and bn.momentum is not None and _ru8(conv.weight.shape[0]) <= 1024):
            return False
    k0 = convs[0].weight[0].numel()
    return k0 <= x.shape[1] <= _ru8(k0)


class _MlpRows(torch.autograd.Function):
    """x [M, K0] -> out [M / pool_k, n8_L]: L layers of (bias-free 1x1 conv on the tcgen05 GEMM with the BatchNorm
    statistics in its epilogue) + (normalise + ReLU [+ max over pool_k rows on the last layer] as one elementwise
    kernel).  Backward: BatchNorm backward of the last layer (cooperative row kernel), then per layer the weight
    gradient (row kernel) and the data gradient GEMM whose epilogue applies the previous layer's ReLU mask and produces
    the two BatchNorm sums, followed by one elementwise kernel.  Replaces pointnet_util.py:213-217 / 273-279 / 343-345
    in training mode.  Tensor arguments after `x`: per layer conv.weight, conv.bias | None, bn.weight, bn.bias."""

    @staticmethod
    def forward(ctx, x, pool_k, out, layers, *params):
        L = len(layers)
        dev = x.device
        ctx.in_cols, ctx.in_dtype = x.shape[1], x.dtype
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        if x.shape[1] % 8:
            x = torch.nn.functional.pad(x, (0, -x.shape[1] % 8))
        if not _rows_ok(x):
            x = x.contiguous()
        M = x.shape[0]
        lib = _lib.lib()
        saved, metas = [x], []
        cur = x
        for l, (conv, bn) in enumerate(layers):
            w = conv.weight
            n = w.shape[0]
            n8 = _ru8(n)
            wl, wt = _weight_pair(w)
            K = min(cur.shape[1], wl.shape[1])
            y = torch.empty(M, n8, dtype=torch.bfloat16, device=dev)
            stats = torch.empty(3, n8, dtype=torch.float32, device=dev)          # mean, invstd, biased variance
            work = torch.empty(max(int(lib.pcb_gemm_work_floats(M, n8, K)), 1), dtype=torch.float32, device=dev)
            tick = _tickets(dev)
            track = bn.track_running_stats
            if track and bn.num_batches_tracked is not None:
                if _step_ctx is not None:
                    _step_ctx.counters.append(bn.num_batches_tracked)
                else:
                    bn.num_batches_tracked.add_(1)
            g32, b32 = params[4 * l + 2].detach().float(), params[4 * l + 3].detach().float()
            bias = params[4 * l + 1]
            # statistics: the GEMM folds its CTAs' partial (count, mean, M2) triples per group of 32 CTAs and stops;
            # the elementwise kernel that follows merges the <= 37 groups while it sets up its constants and writes
            # mean / invstd / var (no second serial fold level at the end of the GEMM)
            gparts = torch.empty(_gemm_max_groups(), 3, n8, dtype=torch.float32, device=dev)
            groups = ctypes.c_int(0)
            _call("pcb_linear_bn_stats_rows_bf16", dev, cur.data_ptr(), cur.stride(0), wl.data_ptr(), wl.stride(0), M, n8,
                  min(wl.shape[0], n8), K, y.data_ptr(), y.stride(0), n, float(bn.eps),
                  None, None, None, work.data_ptr(), tick.data_ptr(), gparts.data_ptr(), ctypes.byref(groups),
                  alg_bytes=2 * M * (K + n8) + 2 * wl.numel())
            last = l == L - 1
            pk = pool_k if last else 1
            Mo = M // pk
            z = None
            if last and out is not None and out.dtype == torch.bfloat16 and tuple(out.shape) == (Mo, n8) \
                    and out.stride(1) == 1 and out.stride(0) % 8 == 0 and out.data_ptr() % 16 == 0:
                z = out
            if z is None:
                z = torch.empty(Mo, n8, dtype=torch.bfloat16, device=dev)
            argmax = torch.empty(Mo, n8, dtype=torch.uint8, device=dev) if pk > 1 else None
            # pooled: the pre-activation of every winning row, all the backward pass needs for its BatchNorm sums
            ymax = torch.empty(Mo, n8, dtype=torch.bfloat16, device=dev) if pk > 1 else None
            _call("pcb_bn_apply_rows", dev, y.data_ptr(), 1, M, n8, n, pk, stats[0].data_ptr(), stats[1].data_ptr(),
                  g32.data_ptr(), b32.data_ptr(), 1, z.data_ptr(), z.stride(0),
                  argmax.data_ptr() if argmax is not None else None, stats[2].data_ptr(),
                  bias.data_ptr() if bias is not None else None, float(bn.momentum),
                  bn.running_mean.data_ptr() if track else None, bn.running_var.data_ptr() if track else None,
                  gparts.data_ptr(), int(groups.value), float(bn.eps), ymax.data_ptr() if ymax is not None else None,
                  alg_bytes=2 * (M + Mo) * n8 + (3 * Mo * n8 if pk > 1 else 0))
            saved += [y, stats, g32, b32, wt]
            if not last:
                saved.append(z)
            else:
                saved.append(argmax if argmax is not None else stats)      # placeholder keeps the layout regular
            metas.append((n, n8, K, bias is not None, _step_ctx.grad_view(w) if _step_ctx is not None else None,
                          tuple(w.shape)))
            cur = z
        if pool_k > 1:
            saved.append(ymax)
        ctx.save_for_backward(*saved)
        ctx.metas, ctx.pool_k, ctx.M = metas, int(pool_k), M
        return cur

    @staticmethod
    def backward(ctx, gout):
        sv = ctx.saved_tensors
        x0 = sv[0]
        L = len(ctx.metas)
        M, pool_k = ctx.M, ctx.pool_k
        dev = x0.device
        lib = _lib.lib()
        grads = [None] * (4 * L)
        lay = lambda l: sv[1 + 6 * l:1 + 6 * (l + 1)]          # y, stats, gamma, beta, w^T, z | argmax
        # ---- last layer: BatchNorm (+ ReLU, + max-pool routing) backward from the incoming gradient
        y, stats, g32, b32, wt, extra = lay(L - 1)
        n, n8, K, has_bias, gview, wshape = ctx.metas[L - 1]
        argmax = extra if pool_k > 1 else None
        gz = gout if gout.dtype == torch.bfloat16 else gout.to(torch.bfloat16)
        if not (gz.dim() == 2 and gz.stride(1) == 1 and gz.stride(0) >= n8 and gz.stride(0) % 8 == 0
                and gz.data_ptr() % 16 == 0):
            gz = gz.contiguous()
        gy = torch.empty_like(y)
        work = _bn_work(n8, dev)
        if pool_k > 1:
            # two ordinary launches: sums over the M / pool_k winning rows, then gy in one pass over y
            ymax = sv[1 + 6 * L]
            _call("pcb_bn_pool_bwd_rows", dev, gz.data_ptr(), gz.stride(0), ymax.data_ptr(), y.data_ptr(), argmax.data_ptr(),
                  1, M, n8, n, pool_k, stats[0].data_ptr(), stats[1].data_ptr(), g32.data_ptr(), b32.data_ptr(), 1,
                  work.data_ptr(), gy.data_ptr(), launches=2,
                  alg_bytes=2 * y.numel() * 2 + gz.numel() * 5 + ymax.numel() * 2)
        else:
            _call("pcb_bn_bwd_rows", dev, gz.data_ptr(), gz.stride(0), y.data_ptr(), None, 1, M, n8, n, pool_k,
                  stats[0].data_ptr(), stats[1].data_ptr(), g32.data_ptr(), b32.data_ptr(), 1, work.data_ptr(),
                  gy.data_ptr(), alg_bytes=(2 * y.numel() + gz.numel()) * 2)
        sums = work[:3 * n8].view(3, n8)
        for l in range(L - 1, -1, -1):
            n, n8, K, has_bias, gview, wshape = ctx.metas[l]
            wt = lay(l)[4]
            xin = lay(l - 1)[5] if l > 0 else x0
            grads[4 * l + 1] = sums[2, :n] if has_bias else None
            grads[4 * l + 2] = sums[1, :n]
            grads[4 * l + 3] = sums[0, :n]
            # weight gradient: gy^T xin over the rows
            kw = 1
            for d in wshape[1:]:
                kw *= d
            if gview is not None:
                wgrad_rows(gy, xin, kw, out=gview.view(n, kw), n=n)
            else:
                grads[4 * l] = wgrad_rows(gy, xin, kw, n=n).view(wshape)
            if l > 0:
                # data gradient through the previous layer's BN + ReLU: dy and its two column sums from the GEMM epilogue
                yp, statsp, gp, bp, _, _ = lay(l - 1)
                npv, np8 = ctx.metas[l - 1][0], ctx.metas[l - 1][1]
                dy = torch.empty_like(yp)
                sums = torch.empty(3, np8, dtype=torch.float32, device=dev)
                Kc = min(gy.shape[1], wt.shape[1])
                work = torch.empty(max(int(lib.pcb_gemm_work_floats(M, np8, Kc)), 1), dtype=torch.float32, device=dev)
                tick = _tickets(dev)
                gparts = torch.empty(_gemm_max_groups(), 3, np8, dtype=torch.float32, device=dev)
                groups = ctypes.c_int(0)
                _call("pcb_dgrad_bn_rows_bf16", dev, gy.data_ptr(), gy.stride(0), wt.data_ptr(), wt.stride(0), M, np8,
                      min(wt.shape[0], np8), Kc, yp.data_ptr(), yp.stride(0), statsp[0].data_ptr(), statsp[1].data_ptr(),
                      gp.data_ptr(), bp.data_ptr(), npv, 1, dy.data_ptr(), dy.stride(0), None, work.data_ptr(),
                      tick.data_ptr(), gparts.data_ptr(), ctypes.byref(groups),
                      alg_bytes=2 * M * (Kc + 2 * np8) + 2 * wt.numel())
                _call("pcb_bn_bwd_apply_rows", dev, dy.data_ptr(), yp.data_ptr(), 1, M, np8, npv, statsp[0].data_ptr(),
                      statsp[1].data_ptr(), gp.data_ptr(), sums.data_ptr(), dy.data_ptr(), gparts.data_ptr(),
                      int(groups.value), alg_bytes=6 * M * np8)
                gy = dy
        gx = None
        if ctx.needs_input_grad[0]:
            wt = lay(0)[4]
            gx = gemm_rows(gy, wt, x0.shape[1])
            if ctx.in_cols != gx.shape[1]:
                gx = gx[:, :ctx.in_cols]
            if gx.dtype != ctx.in_dtype:
                gx = gx.to(ctx.in_dtype)
        return (gx, None, None, None, *grads)


def mlp_rows_fused(x, convs, bns, pool_k: int = 1, out=None):
    """(1x1 conv -> BN(train) -> ReLU) per layer on rows, the last layer followed by the max over `pool_k` consecutive
    rows; returns [M / pool_k, n8] bf16 with the last layer's width rounded up to 8 (zero pad columns)."""
    layers = list(zip(convs, bns))
    params = []
    for conv, bn in layers:
        params += [conv.weight, conv.bias, bn.weight, bn.bias]
    return _MlpRows.apply(x, int(pool_k), out, layers, *params)

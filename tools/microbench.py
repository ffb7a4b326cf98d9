"""Per-kernel timings of the hot-path primitives at the B=16 x 4096-point shapes of SURVEY.md
section 8d, with algorithmic bytes / pair-tests and the fraction of the measured HBM peak.

    python tools/microbench.py [--batch 16] [--iters 20] [--json out.json]

CUDA-event timing on the launching stream, 5 warm-up launches, L2 flushed between timed
launches by writing a 256 MB buffer.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import ops, synthetic  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def timeit(fn, iters, flush):
    for _ in range(int(os.environ.get("PCB_MB_WARMUP", "5"))):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    dev = "cuda:0"
    B, N = a.batch, 4096
    hbm, how = peaks()
    xyz_np, rgb_np, _ = synthetic.bridge_batch(0, B, N)
    xyz = torch.from_numpy(xyz_np).to(dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    start = torch.zeros(B, dtype=torch.long, device=dev)
    rows = []

    def rec(name, fn, bytes_, work=None, work_unit=None):
        if a.only and a.only not in name:
            return
        med, best = timeit(fn, a.iters, flush)
        r = {"kernel": name, "ms_median": round(med, 4), "ms_best": round(best, 4),
             "alg_bytes": int(bytes_), "GBps": round(bytes_ / med / 1e6, 1),
             "hbm_frac": round(bytes_ / med / 1e6 / hbm, 4), "peak": how}
        if work:
            r[work_unit] = round(work / med / 1e6, 2)       # G units / s
        rows.append(r)
        print(json.dumps(r), flush=True)

    # --- FPS chain
    fps1 = ops.furthest_point_sample(xyz, 1024, start)
    l1 = ops.gather(xyz, fps1)
    fps2 = ops.furthest_point_sample(l1, 256, start)
    l2 = ops.gather(l1, fps2)
    fps3 = ops.furthest_point_sample(l2, 64, start)
    l3 = ops.gather(l2, fps3)
    for (src, S, tag) in ((xyz, 1024, "fps_4096_1024"), (l1, 256, "fps_1024_256"), (l1, 512, "fps_1024_512"),
                          (l2, 64, "fps_256_64"), (l3, 16, "fps_64_16")):
        n = src.shape[1]
        rec(tag, lambda src=src, S=S: ops.furthest_point_sample(src, S, start), B * (12 * n + 8 * S),
            B * S * n, "Gupdates_per_s")
    # --- ball query
    for (src, q, r, ns, tag) in ((xyz, l1, 0.1, 32, "ball_4096_1024_r.1_n32"), (xyz, l1, 0.05, 16, "ball_4096_1024_r.05_n16"),
                                 (xyz, l1, 0.2, 32, "ball_4096_1024_r.2_n32"), (l1, l2, 0.2, 32, "ball_1024_256_r.2_n32")):
        n, s = src.shape[1], q.shape[1]
        rec(tag, lambda src=src, q=q, r=r, ns=ns: ops.ball_query(r, ns, src, q), B * (12 * n + 12 * s + 8 * s * ns),
            B * s * n, "Gpairs_per_s")
    rec("ballmulti_4096_1024_r.05-.1_n16-32", lambda: ops.ball_query_multi([0.05, 0.1], [16, 32], xyz, l1),
        B * (12 * N + 12 * 1024 + 8 * 1024 * 48), B * 1024 * N, "Gpairs_per_s")
    ball = ops.ball_query(0.1, 32, xyz, l1)
    # --- gathers
    feat9 = torch.from_numpy(synthetic.sem_seg_input(xyz_np, rgb_np)).to(dev)              # [B,9,N]
    pts9 = feat9.transpose(1, 2).contiguous()
    rec("group_L1_c12", lambda: ops.group_points(xyz, pts9, l1, ball, True), B * (4 * N * 12 + 12 * 1024 + 8 * 1024 * 32 + 4 * 1024 * 32 * 12))
    f64 = torch.randn(B, 1024, 64, device=dev)
    ball2 = ops.ball_query(0.2, 32, l1, l2)
    rec("group_L2_c67", lambda: ops.group_points(l1, f64, l2, ball2, True), B * (4 * 1024 * 67 + 12 * 256 + 8 * 256 * 32 + 4 * 256 * 32 * 67))
    f256 = torch.randn(B, 1024, 256, device=dev)
    bri_idx = ops.ball_query(0.4, 32, l1, ops.gather(l1, ops.furthest_point_sample(l1, 512, start)))
    rec("gather_bri_sa2_c256", lambda: ops.gather(f256, bri_idx), B * (4 * 1024 * 256 + 8 * 512 * 32 + 4 * 512 * 32 * 256))
    # --- training-path rows: padded bf16 grouping of bf16 features, its backward, FP input rows, weight gradient
    fps_l2 = ops.furthest_point_sample(l1, 256, start)
    for (src_xyz, q_xyz, bq, D, tag) in ((xyz, l1, ball, 9, "L1_k32_c12"), (l1, l2, ball2, 96, "L2_k32_c99")):
        n, s, k = src_xyz.shape[1], q_xyz.shape[1], bq.shape[2]
        feats = torch.randn(B, n, D, device=dev)
        feats = feats.to(torch.bfloat16) if D % 8 == 0 else feats
        pitch = -(-(3 + D) // 8) * 8
        nb = B * (feats.element_size() * n * D + 12 * n + 12 * s + 8 * s * k + 2 * s * k * pitch)

        def grp(src_xyz=src_xyz, feats=feats, q_xyz=q_xyz, bq=bq):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return ops.group_points(src_xyz, feats, q_xyz, bq, xyz_first=False, pad_to=8)
        rec("group_bf16_" + tag, grp, nb)
        if D % 8 == 0:
            fr = feats.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = ops.group_points(src_xyz, fr, q_xyz, bq, xyz_first=False, pad_to=8)
            go = torch.randn_like(out)
            rec("group_bwd_bf16_" + tag, lambda out=out, fr=fr, go=go: torch.autograd.grad(out, fr, go, retain_graph=True),
                B * (4 * n * D + 8 * s * k + 2 * s * k * D))
    big_idx = torch.randint(0, 4096, (B, 4096, 32), device=dev)
    fbig = torch.randn(B, 4096, 64, device=dev).to(torch.bfloat16)

    def grp_big():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return ops.group_points(xyz, fbig, xyz, big_idx, xyz_first=False, pad_to=8)
    rec("group_bf16_big_4096x32_c67", grp_big, B * (2 * N * 64 + 24 * N + 8 * N * 32 + 2 * N * 32 * 72))
    d3t, i3t, w3t = ops.three_nn(xyz, l1, 3)
    p1t = torch.randn(B, N, 64, device=dev).to(torch.bfloat16)
    p2t = torch.randn(B, 1024, 128, device=dev).to(torch.bfloat16).requires_grad_(True)

    def fpc():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return ops.fp_concat(p1t, p2t, i3t, w3t, pad_to=8)
    rec("fp_concat_4096_1024_d64+128", fpc, B * (2 * N * 64 + 2 * 1024 * 128 + 12 * 3 * N + 2 * N * 192))
    outc = fpc()
    goc = torch.randn_like(outc)
    rec("fp_concat_bwd_4096_1024_d128", lambda: torch.autograd.grad(outc, p2t, goc, retain_graph=True),
        B * (4 * 1024 * 128 + 12 * 3 * N + 2 * N * 128))
    for (M, Nn, K) in ((524288, 64, 32), (131072, 128, 96), (65536, 128, 128)):
        gy = torch.randn(M, Nn, device=dev).to(torch.bfloat16)
        xx = torch.randn(M, K, device=dev).to(torch.bfloat16)
        gw = torch.zeros(Nn, K, device=dev)
        rec(f"wgrad_rows_{M}x{Nn}x{K}", lambda gy=gy, xx=xx, gw=gw, K=K: ops.wgrad_rows(gy, xx, K, out=gw), 2 * M * (Nn + K))
    # --- three-NN + interpolate
    rec("three_nn_4096_1024", lambda: ops.three_nn(xyz, l1, 3), B * (12 * N + 12 * 1024 + 12 * 3 * N), B * N * 1024, "Gpairs_per_s")
    d3, i3, w3 = ops.three_nn(xyz, l1, 3)
    p2 = torch.randn(B, 128, 1024, device=dev)
    rec("interp_cf_4096_1024_d128", lambda: ops.three_interpolate(p2, i3, w3, True), B * (4 * 1024 * 128 + 12 * 3 * N + 4 * N * 128))
    p2r = p2.transpose(1, 2).contiguous()
    rec("interp_rows_4096_1024_d128", lambda: ops.three_interpolate(p2r, i3, w3, False), B * (4 * 1024 * 128 + 12 * 3 * N + 4 * N * 128))
    # --- kNN
    x3 = xyz.transpose(1, 2).contiguous()
    rec("knn_xyz_k20", lambda: ops.knn(x3, 20), B * (12 * N + 8 * N * 20), B * N * N, "Gpairs_per_s")
    rec("knn_cdist_k32", lambda: ops.knn_cdist(xyz, 32), B * (12 * N + 8 * N * 32), B * N * N, "Gpairs_per_s")
    x64 = torch.from_numpy(synthetic.poly_features(xyz_np, 64, 0)).to(dev)
    rec("knn_feat64_k20", lambda: ops.knn(x64, 20), B * (4 * 64 * N + 8 * N * 20), B * N * N * (2 * 64 + 3) / 1e3, "TFLOPs")
    kidx = ops.knn(x64, 20)
    rec("graph_feature_d64_k20", lambda: ops.graph_feature(x64, kidx), B * (4 * N * 64 + 8 * N * 20 + 4 * 128 * N * 20))
    rec("graph_feature_d3_k20", lambda: ops.graph_feature(x3, kidx), B * (4 * N * 3 + 8 * N * 20 + 4 * 6 * N * 20))
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()

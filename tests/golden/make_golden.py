"""Generates tests/golden/prims_seed*.npz by running the UNMODIFIED reference functions
(imported from /root/reference) on seeded synthetic blocks, on CPU, fp32, torch 2.11.

Run in the authoring container:   python tests/golden/make_golden.py
The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and, through it and directly,
the CUDA path (tests/test_gpu_*.py).  The reference has no golden vectors of its own.

Index arrays are stored as int16 (all values <= 4096), distances as a per-row multiset
checksum (sum of the fp32 bit patterns as uint64) plus a few raw rows, to keep fixtures small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _reference  # noqa: E402
from pointcloud_bridge_b200 import synthetic  # noqa: E402

# (npoint, [(radius, nsample), ...]) per level: union of the SSG and MSG configurations
# (pointnet2_sem_seg.py:11-14, pointnet2_sem_seg_msg.py:11-14)
PN2_LEVELS = [(1024, [(0.05, 16), (0.1, 32)]), (256, [(0.1, 16), (0.2, 32)]),
              (64, [(0.2, 16), (0.4, 32)]), (16, [(0.4, 16), (0.8, 32)])]
# BriStruNet (Highway_bridge/models/model.py:73-76)
BRI_LEVELS = [(1024, [(0.1, 16), (0.2, 32)]), (512, [(0.2, 16), (0.4, 32)]),
              (128, [(0.4, 16), (0.8, 32)])]


def bits_checksum(d):
    """Order-independent checksum of each row's fp32 values."""
    return np.ascontiguousarray(d, np.float32).view(np.uint32).astype(np.uint64).sum(axis=-1)


def peek_randint(N, B):
    """The start indices the next farthest_point_sample call will draw (pointnet_util.py:79)."""
    st = torch.get_rng_state()
    v = torch.randint(0, N, (B,), dtype=torch.long)
    torch.set_rng_state(st)
    return v.numpy()


def run_chain(mod, xyz, levels, tag, out):
    cur = torch.from_numpy(xyz)
    chain = [cur]
    for li, (S, scales) in enumerate(levels):
        B, N, _ = cur.shape
        out[f"{tag}_start{li}"] = peek_randint(N, B).astype(np.int16)
        fps = mod.farthest_point_sample(cur, S)
        assert (fps[:, 0].numpy() == out[f"{tag}_start{li}"]).all()
        out[f"{tag}_fps{li}"] = fps.numpy().astype(np.int16)
        new = mod.index_points(cur, fps)
        for (r, ns) in scales:
            ball = mod.query_ball_point(r, ns, cur, new)
            out[f"{tag}_ball{li}_r{r}_n{ns}"] = ball.numpy().astype(np.int16)
        cur = new
        chain.append(cur)
    return chain


def make_prims(seed, full):
    pu, _, _ = _reference.partsize()
    p2u, dgcnn_mod, am, _ = _reference.highway()
    out = {}
    a, _, _ = synthetic.bridge_batch(seed, 1)
    d, _, _ = synthetic.duplicated_batch(seed + 50, 1)
    xyz = np.concatenate([a, d], 0)                      # [2,4096,3]: clean block + 15% duplicates
    out["xyz"] = xyz
    txyz = torch.from_numpy(xyz)

    torch.manual_seed(1000 + seed)
    chain = run_chain(pu, xyz, PN2_LEVELS, "pn2", out)
    if full:
        torch.manual_seed(2000 + seed)
        run_chain(p2u, xyz, BRI_LEVELS, "bri", out)

    # square_distance: a few raw rows (query = level-1 centroids, as ball query sees it)
    sq = pu.square_distance(chain[1][:, :4], chain[0])
    out["sqdist_rows"] = sq.numpy()

    # three-NN + interpolation for every FP level (pointnet_util.py:325-334); k=4 for fp1
    # as in EnhancedFeaturePropagation (pointnet2_utils.py:253-256)
    for li in range(len(PN2_LEVELS)):
        x1, x2 = chain[li], chain[li + 1]
        dists = pu.square_distance(x1, x2)
        ds, idx = dists.sort(dim=-1)
        for k in ((3, 4) if li == 0 else (3,)):
            dk, ik = ds[:, :, :k], idx[:, :, :k]
            out[f"nn{li}_k{k}_idx"] = ik.numpy().astype(np.int16)
            out[f"nn{li}_k{k}_dist"] = dk.numpy()
            rec = 1.0 / (dk + 1e-8)
            w = rec / torch.sum(rec, dim=2, keepdim=True)
            D2 = 4
            p2 = torch.from_numpy(synthetic.poly_features(x2.numpy(), D2, seed + li)).permute(0, 2, 1)
            interp = torch.sum(pu.index_points(p2, ik) * w.view(*w.shape, 1), dim=2)
            out[f"nn{li}_k{k}_weight"] = w.numpy()
            out[f"nn{li}_k{k}_interp"] = interp.numpy()

    # DGCNN.knn on xyz (D=3) and on 64-d features (DGCNN.py:49-70), k=20
    net = dgcnn_mod.DGCNN(5, 20)
    x3 = txyz.permute(0, 2, 1).contiguous()
    feats = synthetic.poly_features(xyz, 64, seed)
    out["feat64_sum"] = np.float64(feats.astype(np.float64).sum())
    for name, x in (("knn3", x3), ("knn64", torch.from_numpy(feats))):
        idx = net.knn(x, 20)
        xt = x.transpose(2, 1).contiguous()
        inner = -2 * torch.matmul(xt, xt.transpose(2, 1))
        xx = torch.sum(xt ** 2, dim=2, keepdim=True)
        pd = xx + inner + xx.transpose(2, 1)
        dsel = torch.gather(pd, 2, idx)
        out[f"{name}_idx"] = idx.numpy().astype(np.int16)
        out[f"{name}_dsum"] = bits_checksum(dsel.numpy())
        out[f"{name}_drows"] = dsel[:, :16].numpy()
        out[f"{name}_xx"] = xx[:, :, 0].numpy() if name == "knn64" else xx[:, :64, 0].numpy()
    # get_graph_feature (DGCNN.py:72-109): small slice + checksum
    gf = net.get_graph_feature(x3, k=20)
    out["graph3_slice"] = gf[:, :, :32].numpy()
    out["graph3_sum"] = np.float64(gf.double().sum().item())

    # cdist-kNN (attention_modules.py:584-586, 736-738), k = 32 and 16
    # torch.cdist == sqrt_(clamp_min_(matmul of the padded operands, 0)) (_euclidean_dist); the
    # pre-sqrt matrix is recomputed with the same ATen ops so that the fixture can carry values
    # that do not depend on MKL's not-correctly-rounded vector sqrt.
    def cdist_parts(t):
        n2 = t.pow(2).sum(-1, True)
        one = torch.ones_like(n2)
        pre = torch.cat([t.mul(-2), n2, one], -1).matmul(torch.cat([t, one, n2], -1).mT).clamp_min_(0)
        dist = torch.cdist(t, t)
        assert torch.equal(dist, pre.sqrt())
        return dist, pre

    dist, pre = cdist_parts(txyz)
    for k in ((32, 16) if full else (32,)):
        dv, idx = dist.topk(k, dim=-1, largest=False)
        out[f"cdist_k{k}_idx"] = idx.numpy().astype(np.int16)
        out[f"cdist_k{k}_sqsum"] = bits_checksum(torch.gather(pre, 2, idx).numpy())
        out[f"cdist_k{k}_drows"] = dv[:, :16].numpy()
    if full:
        # the small clouds BridgeStructureEncoding sees at geometric2/3 (512 and 128 points)
        for n in (512, 128):
            sub = txyz[:, :n].contiguous()
            dist, pre = cdist_parts(sub)
            dv, idx = dist.topk(16, dim=-1, largest=False)
            out[f"cdist{n}_k16_idx"] = idx.numpy().astype(np.int16)
            out[f"cdist{n}_k16_sqsum"] = bits_checksum(torch.gather(pre, 2, idx).numpy())

    # index_points + sample_and_group of level 1 (pointnet_util.py:116-152): slice + checksum
    pts = torch.from_numpy(synthetic.poly_features(xyz, 9, seed + 9)).permute(0, 2, 1).contiguous()
    fps = torch.from_numpy(out["pn2_fps0"].astype(np.int64))
    ball = torch.from_numpy(out["pn2_ball0_r0.1_n32"].astype(np.int64))
    new_xyz = pu.index_points(txyz, fps)
    g = torch.cat([pu.index_points(txyz, ball) - new_xyz.view(2, 1024, 1, 3), pu.index_points(pts, ball)], -1)
    out["group0_slice"] = g[:, :16].numpy()
    out["group0_sum"] = np.float64(g.double().sum().item())
    return out


def main():
    assert _reference.available(), "reference tree not found"
    torch.set_num_threads(8)
    for seed in (0, 1, 2):
        out = make_prims(seed, full=(seed == 0))
        path = os.path.join(HERE, f"prims_seed{seed}.npz")
        np.savez_compressed(path, **out)
        print(path, f"{os.path.getsize(path) / 1e6:.2f} MB", len(out), "arrays")


if __name__ == "__main__":
    main()

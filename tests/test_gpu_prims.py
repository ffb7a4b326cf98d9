"""GPU parity of every hot-path primitive: CUDA kernels (through the C ABI via
pointcloud_bridge_b200.ops) against the reference's golden vectors and against the CPU oracle
on the same seeded inputs.  Bar: indices and distances bit-exact (ties per tests/parity.py);
interpolated features within 1e-6 relative (they are in fact produced with the same rounded
operations)."""
import numpy as np
import pytest
import torch

import parity
from oracle import oracle as orc
from pointcloud_bridge_b200 import ops, synthetic

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


@pytest.fixture(scope="module", params=[0, 1, 2])
def g(request):
    d = parity.load(f"prims_seed{request.param}.npz")
    d["seed"] = request.param
    return d


@pytest.mark.parametrize("tag", ["pn2", "bri"])
def test_fps_ball_chain_vs_golden(g, tag):
    levels = parity.PN2_LEVELS if tag == "pn2" else parity.BRI_LEVELS
    if f"{tag}_fps0" not in g:
        pytest.skip("chain only stored for seed 0")
    cur = cu(g["xyz"])
    for li, (S, scales) in enumerate(levels):
        start = cu(g[f"{tag}_start{li}"], torch.long)
        fps = ops.furthest_point_sample(cur, S, start)
        assert np.array_equal(fps.cpu().numpy(), g[f"{tag}_fps{li}"]), f"FPS level {li}"
        new = ops.gather(cur, fps)
        for (r, ns) in scales:
            ball = ops.ball_query(r, ns, cur, new)
            assert np.array_equal(ball.cpu().numpy(), g[f"{tag}_ball{li}_r{r}_n{ns}"]), f"ball L{li} r={r}"
        cur = new


def test_fps_default_start_matches_reference_rng(g):
    """Without `start` the op must draw it like pointnet_util.py:79 (CPU generator)."""
    xyz = cu(g["xyz"])
    torch.manual_seed(1000 + g["seed"])
    fps = ops.furthest_point_sample(xyz, 1024)
    assert np.array_equal(fps.cpu().numpy(), g["pn2_fps0"])


def test_square_distance(g):
    xyz = cu(g["xyz"])
    fps = cu(g["pn2_fps0"], torch.long)
    new = ops.gather(xyz, fps)
    sq = ops.square_distance(new[:, :4].contiguous(), xyz).cpu().numpy()
    assert np.array_equal(sq.view(np.uint32), g["sqdist_rows"].view(np.uint32))


def _chain(g):
    cur = g["xyz"]
    chain = [cur]
    for li in range(4):
        cur = orc.index_points(cur, g[f"pn2_fps{li}"].astype(np.int64))
        chain.append(cur)
    return chain


def test_three_nn_and_interpolate(g):
    chain = _chain(g)
    for li in range(4):
        x1, x2 = chain[li], chain[li + 1]
        full = orc.square_distance(x1, x2)
        for k in ((3, 4) if li == 0 else (3,)):
            dist, idx, w = ops.three_nn(cu(x1), cu(x2), k)
            dist, idx, w = dist.cpu().numpy(), idx.cpu().numpy(), w.cpu().numpy()
            o_dist, o_idx = orc.three_nn(x1, x2, k)
            assert np.array_equal(idx, o_idx), f"nn{li} k{k} idx vs oracle"
            assert np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))
            assert np.array_equal(dist.view(np.uint32), g[f"nn{li}_k{k}_dist"].view(np.uint32))
            parity.assert_topk_equivalent(idx, g[f"nn{li}_k{k}_idx"].astype(np.int64),
                                          lambda i: parity.gather_rows(full, i), f"nn{li} k{k}")
            np.testing.assert_allclose(w, g[f"nn{li}_k{k}_weight"], rtol=1e-6, atol=0)
            p2 = np.ascontiguousarray(np.transpose(synthetic.poly_features(x2, 4, g["seed"] + li), (0, 2, 1)))
            ref_idx = cu(g[f"nn{li}_k{k}_idx"], torch.long)
            ref_w = cu(g[f"nn{li}_k{k}_weight"])
            out = ops.three_interpolate(cu(p2), ref_idx, ref_w, channels_first=False).cpu().numpy()
            np.testing.assert_allclose(out, g[f"nn{li}_k{k}_interp"], rtol=1e-6, atol=1e-7)
            out_cf = ops.three_interpolate(cu(np.transpose(p2, (0, 2, 1))), ref_idx, ref_w, channels_first=True)
            assert np.array_equal(out_cf.permute(0, 2, 1).cpu().numpy(), out)


@pytest.mark.parametrize("name,D", [("knn3", 3), ("knn64", 64)])
def test_dgcnn_knn(g, name, D):
    xyz = g["xyz"]
    x = np.ascontiguousarray(np.transpose(xyz, (0, 2, 1))) if D == 3 else synthetic.poly_features(xyz, 64, g["seed"])
    xt = np.ascontiguousarray(np.transpose(x, (0, 2, 1)))
    idx, dist = ops.knn(cu(x), 20, channels_first=True, return_dist=True)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    o_idx, o_dist = orc.knn(x, 20, return_dist=True)
    assert np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32)), "distances vs oracle"
    assert np.array_equal(idx, o_idx), "indices vs oracle (same (distance, index) order)"
    assert np.array_equal(parity.bits_checksum(dist), g[f"{name}_dsum"]), "distance multiset vs reference"
    parity.assert_topk_equivalent(idx, g[f"{name}_idx"].astype(np.int64), lambda i: orc.pair_dist(xt, i, "knn"), name)
    if D == 3:   # points-major entry of the same kernel
        idx2 = ops.knn(cu(xyz), 20, channels_first=False).cpu().numpy()
        assert np.array_equal(idx2, idx)


@pytest.mark.parametrize("k", [32, 16])
def test_cdist_knn(g, k):
    if f"cdist_k{k}_idx" not in g:
        pytest.skip("k=16 only stored for seed 0")
    xyz = g["xyz"]
    idx, dist = ops.knn_cdist(cu(xyz), k, return_dist=True)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    o_idx, o_dist = orc.knn_cdist(xyz, k, return_dist=True)
    assert np.array_equal(dist.view(np.uint32), o_dist.view(np.uint32))
    assert np.array_equal(idx, o_idx)
    parity.assert_topk_equivalent(idx, g[f"cdist_k{k}_idx"].astype(np.int64),
                                  lambda i: orc.pair_dist(xyz, i, "cdist"), f"cdist k{k}", ulp_tol=1)


def test_graph_feature_and_grouping(g):
    xyz = g["xyz"]
    x3 = np.ascontiguousarray(np.transpose(xyz, (0, 2, 1)))
    idx = g["knn3_idx"].astype(np.int64)
    gf = ops.graph_feature(cu(x3), cu(idx)).cpu().numpy()
    assert np.array_equal(gf[:, :, :32], g["graph3_slice"])
    assert np.array_equal(gf, orc.get_graph_feature(x3, idx))
    pts = np.ascontiguousarray(np.transpose(synthetic.poly_features(xyz, 9, g["seed"] + 9), (0, 2, 1)))
    fps = g["pn2_fps0"].astype(np.int64)
    ball = g["pn2_ball0_r0.1_n32"].astype(np.int64)
    new_xyz = ops.gather(cu(xyz), cu(fps))
    grp = ops.group_points(cu(xyz), cu(pts), new_xyz, cu(ball), xyz_first=True).cpu().numpy()
    assert np.array_equal(grp[:, :16], g["group0_slice"])
    o = orc.group_points(xyz, pts, new_xyz.cpu().numpy(), ball, xyz_first=True)
    assert np.array_equal(grp, o)
    # MSG channel order and channels-first feature operand
    grp2 = ops.group_points(cu(xyz), cu(np.transpose(pts, (0, 2, 1))), new_xyz, cu(ball), xyz_first=False,
                            points_cf=True).cpu().numpy()
    assert np.array_equal(grp2, orc.group_points(xyz, pts, new_xyz.cpu().numpy(), ball, xyz_first=False))
    # plain index_points, vectorised (C=12 -> float4) and scalar (C=9) paths
    pts12 = np.concatenate([pts, xyz], -1)
    for p in (pts, pts12):
        out = ops.gather(cu(p), cu(ball)).cpu().numpy()
        assert np.array_equal(out, orc.index_points(p, ball))


def test_index_points_error_behaviour():
    pts = torch.arange(2 * 5 * 3, dtype=torch.float32, device=DEV).reshape(2, 5, 3)
    idx = torch.tensor([[0, 4], [5, 1]], device=DEV)
    ops.check_index_errors()
    ops.gather(pts, idx, clamp=False)                    # idx == N: what an empty ball yields
    with pytest.raises(IndexError):                      # pointnet_util.py:62 raises
        ops.check_index_errors()
    out = ops.gather(pts, idx, clamp=True)               # pointnet2_utils.py:34-36 clamps
    assert torch.equal(out[1, 0], pts[1, 4])
    ops.check_index_errors()
    neg = ops.gather(pts, torch.tensor([[-1, 0], [-5, 1]], device=DEV))
    assert torch.equal(neg[0, 0], pts[0, 4]) and torch.equal(neg[1, 0], pts[1, 0])


def test_empty_ball_and_short_rows():
    """Edge cases of query_ball_point: empty ball -> N everywhere; short row padded with first."""
    xyz = torch.tensor([[[0, 0, 0], [1, 0, 0], [0.05, 0, 0], [2, 2, 2.0]]], device=DEV)
    new = torch.tensor([[[0, 0, 0], [10, 10, 10.0], [2, 2, 2]]], device=DEV)
    out = ops.ball_query(0.1, 4, xyz, new).cpu().numpy()
    assert out.tolist() == [[[0, 2, 0, 0], [4, 4, 4, 4], [3, 3, 3, 3]]]
    assert np.array_equal(out, orc.query_ball_point(0.1, 4, xyz.cpu().numpy(), new.cpu().numpy()))


@pytest.mark.parametrize("N,S", [(5, 3), (33, 7), (100, 100), (1000, 37), (4099, 130), (9000, 64), (20000, 16)])
def test_fps_ragged_sizes_vs_oracle(N, S):
    rng = np.random.default_rng(N)
    xyz = rng.random((3, N, 3)).astype(np.float32)
    xyz[1, N // 2:] = xyz[1, : N - N // 2]               # duplicates: argmax ties and exhausted clouds
    start = rng.integers(0, N, 3)
    got = ops.furthest_point_sample(cu(xyz), S, cu(start, torch.long)).cpu().numpy()
    assert np.array_equal(got, orc.farthest_point_sample(xyz, S, start))


@pytest.mark.parametrize("N,S,ns,r", [(7, 3, 4, 0.5), (130, 17, 16, 0.2), (4099, 50, 32, 0.1), (9001, 40, 64, 0.05),
                                      (20000, 33, 32, 0.03)])
def test_ball_query_ragged_sizes_vs_oracle(N, S, ns, r):
    rng = np.random.default_rng(N + 1)
    xyz = rng.random((2, N, 3)).astype(np.float32)
    new = np.ascontiguousarray(xyz[:, rng.choice(N, S, replace=False)])
    new[:, 0] += 5.0                                      # one empty ball
    got = ops.ball_query(r, ns, cu(xyz), cu(new)).cpu().numpy()
    assert np.array_equal(got, orc.query_ball_point(r, ns, xyz, new))


@pytest.mark.parametrize("N,D,k", [(40, 3, 5), (130, 3, 32), (1000, 3, 40), (257, 6, 8), (1001, 16, 20),
                                   (1024, 64, 20), (520, 128, 32), (9000, 3, 16)])
def test_knn_ragged_sizes_vs_oracle(N, D, k):
    rng = np.random.default_rng(N + D)
    x = rng.standard_normal((2, D, N)).astype(np.float32)
    x[1, :, N // 2:] = x[1, :, : N - N // 2]             # exact duplicates -> ties
    idx, dist = ops.knn(cu(x), k, return_dist=True)
    o_idx, o_dist = orc.knn(x, k, return_dist=True)
    assert np.array_equal(dist.cpu().numpy().view(np.uint32), o_dist.view(np.uint32))
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    if D == 3:
        xyz = np.ascontiguousarray(np.transpose(x, (0, 2, 1)))
        ci, cd = ops.knn_cdist(cu(xyz), min(k, 64), return_dist=True)
        oi, od = orc.knn_cdist(xyz, min(k, 64), return_dist=True)
        assert np.array_equal(cd.cpu().numpy().view(np.uint32), od.view(np.uint32))
        assert np.array_equal(ci.cpu().numpy(), oi)


def test_backward_of_gathers_matches_torch_autograd():
    """Scatter-add backward of index_points / grouping / graph feature / interpolation against
    plain PyTorch indexing (fp32, atomics reorder the sum: tolerance 1e-5 relative)."""
    torch.manual_seed(0)
    B, N, S, K, D = 2, 257, 33, 8, 12
    pts = torch.randn(B, N, D, device=DEV, requires_grad=True)
    xyz = torch.randn(B, N, 3, device=DEV)
    idx = torch.randint(0, N, (B, S, K), device=DEV)
    new_xyz = torch.randn(B, S, 3, device=DEV)
    bi = torch.arange(B, device=DEV).view(B, 1, 1).expand(B, S, K)

    out = ops.gather(pts, idx)
    ref = pts[bi, idx]
    assert torch.equal(out, ref)
    g = torch.randn_like(out)
    (ga,) = torch.autograd.grad(out, pts, g)
    (gb,) = torch.autograd.grad(ref, pts, g)
    torch.testing.assert_close(ga, gb, rtol=1e-5, atol=1e-5)

    for cf in (False, True):
        p_in = pts.transpose(1, 2).contiguous().detach().requires_grad_(True) if cf else pts
        out = ops.group_points(xyz, p_in, new_xyz, idx, xyz_first=True, points_cf=cf)
        ref = torch.cat([xyz[bi, idx] - new_xyz.view(B, S, 1, 3), pts[bi, idx]], -1)
        assert torch.equal(out, ref)
        g = torch.randn_like(out)
        (ga,) = torch.autograd.grad(out, p_in, g)
        (gb,) = torch.autograd.grad(ref, pts, g)
        torch.testing.assert_close(ga.transpose(1, 2) if cf else ga, gb, rtol=1e-5, atol=1e-5)
        # padded rows (pad_to=8): same values, zero pad columns, pad columns ignored by the backward
        outp = ops.group_points(xyz, p_in, new_xyz, idx, xyz_first=True, points_cf=cf, pad_to=8)
        C = 3 + D
        assert outp.shape[-1] == -(-C // 8) * 8
        assert torch.equal(outp[..., :C], ref) and not outp[..., C:].any()
        gp = torch.randn_like(outp)
        (gc,) = torch.autograd.grad(outp, p_in, gp)
        ref2 = torch.cat([xyz[bi, idx] - new_xyz.view(B, S, 1, 3), pts[bi, idx]], -1)
        (gd,) = torch.autograd.grad(ref2, pts, gp[..., :C])
        torch.testing.assert_close(gc.transpose(1, 2) if cf else gc, gd, rtol=1e-5, atol=1e-5)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outb = ops.group_points(xyz, p_in, new_xyz, idx, xyz_first=False, points_cf=cf, pad_to=8)
        refb = torch.cat([pts[bi, idx], xyz[bi, idx] - new_xyz.view(B, S, 1, 3)], -1).to(torch.bfloat16)
        assert outb.dtype == torch.bfloat16 and torch.equal(outb[..., :C], refb) and not outb[..., C:].any()

    x = torch.randn(B, D, N, device=DEV, requires_grad=True)
    kidx = torch.randint(0, N, (B, N, 5), device=DEV)
    out = ops.graph_feature(x, kidx)
    xt = x.transpose(1, 2)
    nb = xt[torch.arange(B, device=DEV).view(B, 1, 1).expand(B, N, 5), kidx]          # [B,N,k,D]
    ctr = xt.unsqueeze(2).expand(-1, -1, 5, -1)
    ref = torch.cat([nb - ctr, ctr], 3).permute(0, 3, 1, 2)
    assert torch.equal(out, ref)
    g = torch.randn_like(out)
    (ga,) = torch.autograd.grad(out, x, g)
    (gb,) = torch.autograd.grad(ref, x, g)
    torch.testing.assert_close(ga, gb, rtol=1e-5, atol=1e-5)

    p2 = torch.randn(B, S, D, device=DEV, requires_grad=True)
    iidx = torch.randint(0, S, (B, N, 3), device=DEV)
    w = torch.rand(B, N, 3, device=DEV)
    for cf in (False, True):
        p_in = p2.transpose(1, 2).contiguous().detach().requires_grad_(True) if cf else p2
        out = ops.three_interpolate(p_in, iidx, w, channels_first=cf)
        ref = (p2[torch.arange(B, device=DEV).view(B, 1, 1).expand(B, N, 3), iidx] * w.unsqueeze(-1)).sum(2)
        torch.testing.assert_close(out.transpose(1, 2) if cf else out, ref, rtol=1e-6, atol=1e-6)
        g = torch.randn_like(out)
        (ga,) = torch.autograd.grad(out, p_in, g)
        (gb,) = torch.autograd.grad(ref, p2, g.transpose(1, 2) if cf else g)
        torch.testing.assert_close(ga.transpose(1, 2) if cf else ga, gb, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("dt1,dt2", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                     (torch.bfloat16, torch.float32), (None, torch.bfloat16)])
@pytest.mark.parametrize("dims", [(10, 14), (16, 24), (8, 12)])      # pair kernel / 8-channel chunks / mixed
def test_fp_concat_equals_interpolate_cat_cast(dt1, dt2, dims):
    """The fused feature-propagation input rows (bf16 training) against the unfused chain
    three_interpolate (fp32) -> cat -> bf16: forward bit-equal, backward within bf16/atomic noise."""
    torch.manual_seed(1)
    B, N, S, k = 2, 300, 41, 3
    D1, D2 = (0 if dt1 is None else dims[0]), dims[1]
    p1 = torch.randn(B, N, D1, device=DEV).to(dt1).requires_grad_(True) if D1 else None
    p2 = torch.randn(B, S, D2, device=DEV).to(dt2).requires_grad_(True)
    idx = torch.randint(0, S, (B, N, k), device=DEV)
    w = torch.rand(B, N, k, device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert ops.fp_concat_supported(p1, p2)
        out = ops.fp_concat(p1, p2, idx, w, pad_to=8)
    D = D1 + D2
    assert out.dtype == torch.bfloat16 and out.shape == (B, N, -(-D // 8) * 8)
    interp = ops.three_interpolate(p2, idx, w, channels_first=False)
    ref = interp if p1 is None else torch.cat([p1.float(), interp], -1)
    assert torch.equal(out[..., :D], ref.to(torch.bfloat16)) and not out[..., D:].any()
    g = torch.randn_like(out)
    ins = [p2] if p1 is None else [p1, p2]
    ga = torch.autograd.grad(out, ins, g)
    gb = torch.autograd.grad(ref, ins, g[..., :D].float())
    for a, b in zip(ga, gb):
        assert a.dtype == b.dtype
        torch.testing.assert_close(a.float(), b.float(), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("D,xyz_first", [(16, False), (16, True), (9, False), (24, False)])
def test_group_points_padded_bf16_rows_fwd_bwd(D, xyz_first):
    """Padded bf16 rows of the training path (8-channel chunk kernel, 128-bit reductions backward)
    against plain indexing."""
    torch.manual_seed(2)
    B, N, S, K = 2, 301, 37, 16
    pts = torch.randn(B, N, D, device=DEV, requires_grad=True)
    xyz = torch.randn(B, N, 3, device=DEV)
    idx = torch.randint(0, N, (B, S, K), device=DEV)
    new_xyz = torch.randn(B, S, 3, device=DEV)
    bi = torch.arange(B, device=DEV).view(B, 1, 1).expand(B, S, K)
    C = 3 + D
    for bf16 in (True, False):
        parts = [xyz[bi, idx] - new_xyz.view(B, S, 1, 3), pts[bi, idx]]
        ref = torch.cat(parts if xyz_first else parts[::-1], -1)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            out = ops.group_points(xyz, pts, new_xyz, idx, xyz_first=xyz_first, pad_to=8)
        assert out.dtype == (torch.bfloat16 if bf16 else torch.float32) and out.shape[-1] % 8 == 0
        assert torch.equal(out[..., :C], ref.to(out.dtype)) and not out[..., C:].any()
        g = torch.randn_like(out)
        (ga,) = torch.autograd.grad(out, pts, g)
        (gb,) = torch.autograd.grad(ref, pts, g[..., :C].float())
        torch.testing.assert_close(ga, gb, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("D", [16, 12])
def test_group_points_takes_bf16_feature_rows(D):
    """bf16 features of the previous autocast layer are gathered without an fp32 round trip: the grouped
    bf16 rows equal those obtained from the same features given as fp32; gradient as plain indexing."""
    torch.manual_seed(3)
    B, N, S, K = 2, 200, 23, 16
    pts = torch.randn(B, N, D, device=DEV).to(torch.bfloat16).requires_grad_(True)
    xyz = torch.randn(B, N, 3, device=DEV)
    idx = torch.randint(0, N, (B, S, K), device=DEV)
    new_xyz = torch.randn(B, S, 3, device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = ops.group_points(xyz, pts, new_xyz, idx, xyz_first=False, pad_to=8)
        ref = ops.group_points(xyz, pts.detach().float(), new_xyz, idx, xyz_first=False, pad_to=8)
    assert out.dtype == torch.bfloat16 and torch.equal(out, ref)
    g = torch.randn_like(out)
    (ga,) = torch.autograd.grad(out, pts, g)
    bi = torch.arange(B, device=DEV).view(B, 1, 1).expand(B, S, K)
    p32 = pts.detach().float().requires_grad_(True)
    (gb,) = torch.autograd.grad(p32[bi, idx], p32, g[..., :D].float())
    assert ga.dtype == torch.bfloat16
    torch.testing.assert_close(ga.float(), gb, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("N,S", [(4096, 1024), (1024, 256), (5000, 300), (8200, 64), (33, 7)])
def test_ball_query_multi_equals_separate_queries(N, S):
    """The one-scan multi-radius kernel returns exactly what one ball_query per radius returns (which is
    checked against the oracle and the reference's golden vectors above), incl. empty and overfull balls."""
    xyz_np = synthetic.bridge_batch(5, 3, N)[0]
    xyz = cu(xyz_np)
    new = cu(xyz_np[:, ::max(1, N // S)][:, :S] + np.float32(0.001))
    for radii, ns in (((0.05, 0.1), (16, 32)), ((0.2, 0.4, 0.01), (16, 32, 8)), ((0.1, 0.2, 0.4, 0.8), (64, 32, 16, 8))):
        multi = ops.ball_query_multi(radii, ns, xyz, new)
        for r, n, m in zip(radii, ns, multi):
            assert torch.equal(m, ops.ball_query(r, n, xyz, new)), (r, n)

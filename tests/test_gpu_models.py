"""GPU parity of the whole networks against the reference's own outputs (tests/golden/models.npz,
produced by tests/golden/make_golden_models.py from the unmodified reference on CPU fp32 with
name-keyed seeded weights).

Tolerances (fp32 mode; north_star asks 1e-5 relative for fp32, 1e-2 for bf16):
  * indices inside the networks are bit-exact by construction (they depend on xyz only), so the
    remaining difference is GEMM/BN summation order: |delta| <= 2e-5 * max|ref| + 2e-5 is asserted
    for PointNet++ SSG/MSG and the Highway PointNet2;
  * DGCNN layers 2-4 run kNN on learned features, where a 1-ulp upstream difference can swap
    two near-tied neighbours.  The reference itself, run on CPU with 1 thread instead of 8, moves
    its logits by p50 3e-7 / p99.9 4.2e-4 / max 1.7e-3 (relative to max|logit|), and by max 3.7e-3
    under a 3e-7 relative weight perturbation (measured in the authoring container).  Asserted:
    p50 < 1e-5, p99.9 < 2e-3, max < 2e-2, argmax agreement >= 99.9 %; the kNN itself is checked
    bit-exact at op level on the oracle's own features (tests/test_gpu_prims.py);
  * BriStruNet feeds eigenvalue ratios (e0 - e1) / (e0 + 1e-8) of near-singular 3x3 covariances
    into the net (attention_modules.py:631-633), ill-conditioned in fp32 (LAPACK on CPU vs
    cuSOLVER on GPU; here one kernel with a float64 closed form, csrc/structure.cu, in training and evaluation).
    Measured p50 2e-7 / p99 1.4e-6 / max 1e-5 on the fixture; asserted: FPS indices bit-exact, median < 1e-5,
    p99 < 1e-4, max < 5e-4.
  * bf16 autocast: 1e-2 relative to max|ref| on the log-probabilities (mean) for MSG.
"""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import parity
from pointcloud_bridge_b200 import synthetic
from pointcloud_bridge_b200.highway import DGCNN as dgcnn_mod
from pointcloud_bridge_b200.highway import model as hb_model
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg as ssg
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SEED_FPS = 4242


@pytest.fixture(scope="module")
def g():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return parity.load("models.npz")


def inputs(g):
    x9 = torch.from_numpy(synthetic.sem_seg_input(g["xyz"], g["rgb"])).to(DEV)
    return x9, torch.from_numpy(g["xyz"]).to(DEV), torch.from_numpy(g["rgb"]).to(DEV), \
        torch.from_numpy(g["labels"].astype(np.int64)).to(DEV)


def run_eval(net, *args):
    net.eval()
    torch.manual_seed(SEED_FPS)
    with torch.no_grad():
        return net(*args)


def rel_err(a, ref):
    a = a.detach().float().cpu().numpy()
    return float(np.abs(a - ref).max() / (np.abs(ref).max() + 1e-12))


def test_ssg_forward_config1(g):
    x9, *_ = inputs(g)
    net = parity.seeded_fill_(ssg.get_model(13), 1).to(DEV)
    y, l4 = run_eval(net, x9[:1])
    assert y.shape == (1, 4096, 13) and l4.shape == (1, 512, 16)
    e1, e2 = rel_err(y, g["ssg_logp"]), rel_err(l4, g["ssg_l4"])
    print("ssg rel err", e1, e2)
    assert e1 < 2e-5 and e2 < 2e-5


def test_msg_forward_and_train_step_config2(g):
    x9, _, _, lab = inputs(g)
    net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV)
    y, l4 = run_eval(net, x9)
    e1, e2 = rel_err(y, g["msg_logp"]), rel_err(l4, g["msg_l4"])
    print("msg rel err", e1, e2)
    assert e1 < 2e-5 and e2 < 2e-5
    # training-mode forward + backward (dropout off, as in the fixture)
    net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV)
    net.train()
    net.drop1.eval()
    torch.manual_seed(SEED_FPS)
    y, _ = net(x9)
    loss = torch.nn.functional.nll_loss(y.reshape(-1, 5), lab.reshape(-1))
    loss.backward()
    assert abs(loss.item() - float(g["msg_train_loss"])) < 2e-5 * max(1.0, abs(float(g["msg_train_loss"])))
    assert rel_err(y, g["msg_train_logp"]) < 5e-5
    for name, p in (("g_sa1", net.sa1.conv_blocks[0][0].weight.grad), ("g_fp1", net.fp1.mlp_convs[0].weight.grad),
                    ("g_conv2", net.conv2.weight.grad), ("rm_sa1", net.sa1.bn_blocks[0][0].running_mean),
                    ("rv_sa1", net.sa1.bn_blocks[0][0].running_var)):
        e = rel_err(p, g[f"msg_train_{name}"])
        print("msg train", name, e)
        # Weight gradients below a BatchNorm cancel heavily in fp32: the reference's own CPU result
        # for g_sa1 / g_fp1 moves by 5e-3 between 1 and 8 threads and sits 5e-3..1e-2 away from an
        # fp64 evaluation (measured in the authoring container), so 2e-2 is the parity bar there;
        # quantities without that cancellation must agree to 2e-4.
        assert e < (2e-2 if name in ("g_sa1", "g_fp1") else 2e-4), name


def test_msg_bf16_autocast_within_1e2(g):
    x9, *_ = inputs(g)
    net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV)
    net.eval()
    torch.manual_seed(SEED_FPS)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y, _ = net(x9)
    ref = g["msg_logp"]
    err = np.abs(y.float().cpu().numpy() - ref)
    print("msg bf16 mean/max rel err", err.mean() / np.abs(ref).max(), err.max() / np.abs(ref).max())
    assert err.mean() / np.abs(ref).max() < 1e-2
    agree = (y.float().cpu().numpy().argmax(-1) == ref.argmax(-1)).mean()
    assert agree > 0.97, agree


def test_dgcnn_forward_config3(g):
    _, xyz, rgb, _ = inputs(g)
    net = parity.seeded_fill_(dgcnn_mod.DGCNN(5, 20), 3).to(DEV)
    y = run_eval(net, xyz[:1], rgb[:1]).float().cpu().numpy()
    ref = g["dgcnn_logits"]
    assert y.shape == ref.shape
    err = np.abs(y - ref) / (np.abs(ref).max() + 1e-12)
    print("dgcnn rel err: p50 %.2e p99.9 %.2e max %.2e" % (np.median(err), np.quantile(err, 0.999), err.max()))
    assert np.median(err) < 1e-5 and np.quantile(err, 0.999) < 2e-3 and err.max() < 2e-2
    assert (y.argmax(-1) == ref.argmax(-1)).mean() >= 0.999


def test_highway_pointnet2_forward(g):
    _, xyz, rgb, _ = inputs(g)
    net = parity.seeded_fill_(hb_model.PointNet2(5), 4).to(DEV)
    y = run_eval(net, xyz[:1], rgb[:1])
    e = rel_err(y, g["hbpn2_logits"])
    print("hb pointnet2 rel err", e)
    assert y.shape == (1, 5, 4096) and e < 2e-5


def test_bristrunet_forward_config4(g):
    _, xyz, rgb, lab = inputs(g)
    from pointcloud_bridge_b200.highway import pointnet2_utils as p2u
    torch.manual_seed(SEED_FPS)
    cur = xyz[:1]
    for li, S in enumerate((1024, 512, 128)):
        fps = p2u.farthest_point_sample(cur, S)
        assert np.array_equal(fps.cpu().numpy(), g[f"bri_fps{li}"]), f"BriStruNet FPS level {li}"
        cur = p2u.index_points(cur, fps)
    net = parity.seeded_fill_(hb_model.EnhancedPointNet2(5), 5).to(DEV)
    y = run_eval(net, xyz[:1], rgb[:1])
    assert y.shape == (1, 5, 4096)
    ref = g["bristrunet_logits"]
    err = np.abs(y.float().cpu().numpy() - ref) / (np.abs(ref).max() + 1e-12)
    print("bristrunet rel err: p50 %.2e p99 %.2e max %.2e" % (np.median(err), np.quantile(err, 0.99), err.max()))
    assert np.isfinite(y.float().cpu().numpy()).all()
    assert np.median(err) < 1e-5 and np.quantile(err, 0.99) < 1e-4 and err.max() < 5e-4     # measured max 1e-5
    crit = hb_model.BridgeStructureLoss(num_classes=5, alpha=80, rel_margin=0.3).to(DEV)
    loss = crit(torch.from_numpy(ref).to(DEV), lab[:1], xyz[:1])
    assert abs(loss.item() - float(g["bri_loss"])) < 1e-4 * max(1.0, abs(float(g["bri_loss"])))


def test_reference_checkpoint_keys_load(g):
    """state_dict key compatibility is checked on CPU against the reference in the authoring
    container (tests/test_state_dict_keys.py); here: a round trip through torch.save works and a
    strict load succeeds."""
    net = msg.get_model(5)
    buf = io.BytesIO()
    torch.save({"model_state_dict": net.state_dict()}, buf)
    buf.seek(0)
    ck = torch.load(buf, weights_only=True)
    msg.get_model(5).load_state_dict(ck["model_state_dict"], strict=True)


def test_cuda_graph_training_matches_eager(g):
    """engine.Trainer(graph=True) replays one captured step.  Training itself is chaotic (float
    atomics + Adam: two eager runs differ by 3e-3 after 3 steps), so the check uses lr = 0: the loss
    then only depends on the FPS start indices, which both modes must draw identically from the
    CPU generator, step after step."""
    from pointcloud_bridge_b200.engine import Trainer
    x9, _, _, lab = inputs(g)
    losses = {}
    for mode in (False, True):
        torch.manual_seed(7)
        net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV).train()
        net.drop1.eval()
        for m in net.modules():                       # freeze BN running stats as well
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 0.0
        tr = Trainer(net, amp=False, graph=mode, capturable=True, lr=0.0, weight_decay=0.0)
        torch.manual_seed(11)
        losses[mode] = [float(tr.step(x9, labels=lab).item()) for _ in range(7)]
    print("eager", losses[False], "graph", losses[True])
    assert len(set(round(v, 4) for v in losses[False])) > 3       # the starts really vary per step
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) <= 2e-5 * max(1.0, abs(a))


def test_block_inference_graph_matches_eager(g):
    """engine.BlockInference: captured forward == eager forward (same FPS draws), labels identical."""
    from pointcloud_bridge_b200.engine import BlockInference
    x9, *_ = inputs(g)
    x = x9.repeat(4, 1, 1)                                   # 8 blocks -> 4 batches of 2
    net = parity.seeded_fill_(ssg.get_model(13), 1).to(DEV)
    for nblocks, bb in ((8, 2), (7, 3)):                     # (7, 3): a short tail batch replays the full-batch graph
        outs = {}
        for mode in (False, True):
            torch.manual_seed(5)
            outs[mode] = BlockInference(net, batch_blocks=bb, amp=False, graph=mode).run(x[:nblocks]).cpu()
        agree = (outs[False] == outs[True]).float().mean().item()
        assert agree > 0.999, (nblocks, bb, agree)


def test_trainer_flat_gradients_match_plain_autograd(g, monkeypatch):
    """engine.Trainer under bf16 autocast (bf16 weight shadows, weight gradients written by the row wgrad kernel
    straight into the flat bucket, BN/bias gradients packed) against a plain autograd backward of the same
    network and batch: every parameter's gradient within 3e-2 in relative L2 norm (bf16 GEMM noise, atomics)."""
    from pointcloud_bridge_b200 import ops
    from pointcloud_bridge_b200.engine import Trainer
    # The 196-channel layers run zero-padded to 200 channels in the Trainer (other cuBLAS kernels, other
    # bf16 roundings of the same sums): through ReLU masks and max-pool routing that alone moves some
    # mid-network gradients by ~10 %, ten times the run-to-run noise of the atomics (tools/dbg_flatgrad.py).
    # It is switched off here so that the comparison isolates the bucket / weight-gradient plumbing; the
    # padded layers themselves are checked one by one in test_gpu_bn_rows.py.
    monkeypatch.setattr(ops, "_PAD_N", False)
    x9, _, _, lab = inputs(g)
    torch.manual_seed(5)
    net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV).train()
    net.drop1.eval()
    torch.manual_seed(11)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logp, _ = net(x9)
    F.nll_loss(logp.float().reshape(-1, logp.shape[-1]), lab.reshape(-1)).backward()
    ref = [p.grad.detach().clone() for p in net.parameters() if p.requires_grad]
    for p in net.parameters():
        p.grad = None
    tr = Trainer(net, amp=True, graph=False, lr=0.0, weight_decay=0.0)
    torch.manual_seed(11)
    tr.step(x9, labels=lab)
    torch.cuda.synchronize()
    assert len(ref) == len(tr.bucket.views)
    worst = 0.0
    for (name, _), a, b in zip(net.named_parameters(), tr.bucket.views, ref):
        scale = b.abs().max().item()
        err = (a - b).abs().max().item()
        if name.endswith(".bias") and "conv" in name and name != "conv2.bias":
            continue                                   # conv bias before a training-mode BN: exactly zero in theory, rounding noise here
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()     # atomics + bf16 rounding: compare in L2, bound the max
        worst = max(worst, rel)
        if ".bn" in name or "mlp_bns" in name or name.startswith("bn"):
            assert rel <= 0.2 or scale < 2e-3, (name, rel, err, scale)      # sums of routed gradients: atomics noise up to 10 %
        else:
            assert rel <= 3e-2 and err <= 0.15 * scale + 1e-6, (name, rel, err, scale)
    print("worst relative gradient difference", worst)


def test_bristrunet_training_forward_fused_rows_vs_plain(g, monkeypatch):
    """BriStruNet in training mode inside the step runner (Conv->BN->ReLU triples of its encoders through the
    fused row kernels, 3/6/9-channel layers carried as zero-padded 8-channel rows) against the same network
    evaluated layer by layer with library ops: same loss and logits within bf16 noise, finite gradients."""
    from pointcloud_bridge_b200.engine import Trainer
    from pointcloud_bridge_b200.highway import attention_modules as am, model as hm, pointnet2_utils as p2u
    _, xyz, rgb, lab = inputs(g)
    crit = hb_model.BridgeStructureLoss(num_classes=5, alpha=80, rel_margin=0.3).to(DEV)
    losses = {}
    for mode in ("plain", "fused"):
        if mode == "plain":
            for mod in (p2u, am, hm):
                monkeypatch.setattr(mod, "seq_rows", p2u._seq_rows_plain)
        else:
            monkeypatch.undo()
        torch.manual_seed(7)
        net = parity.seeded_fill_(hb_model.EnhancedPointNet2(5), 5).to(DEV).train()
        tr = Trainer(net, loss_fn=lambda out, labels, pts: crit(out, labels, pts), amp=True, graph=False, lr=0.0,
                     weight_decay=0.0)
        torch.manual_seed(11)
        losses[mode] = float(tr.step(xyz, rgb, labels=lab, loss_inputs=(xyz,)).item())
        flat = tr.bucket.flat
        assert torch.isfinite(flat).all() and flat.abs().max().item() > 0
    print("BriStruNet train loss plain / fused:", losses)
    assert abs(losses["plain"] - losses["fused"]) <= 2e-2 * max(1.0, abs(losses["plain"]))


def test_index_chain_precomputed_equals_inline(g):
    """get_model.index_chain + forward(pre=...) == the plain forward: same CPU-generator draws, same indices, so the
    training-mode loss is bit-identical (SSG and MSG)."""
    x9, _, _, lab = inputs(g)
    for mod, classes, seed in ((ssg, 13, 1), (msg, 5, 2)):
        losses = []
        for use_pre in (False, True):
            torch.manual_seed(3)
            net = parity.seeded_fill_(mod.get_model(classes), seed).to(DEV).train()
            net.drop1.eval()
            torch.manual_seed(SEED_FPS)
            pre = net.index_chain(x9) if use_pre else None
            y, _ = net(x9, pre=pre)
            losses.append(float(F.nll_loss(y.reshape(-1, classes), lab.reshape(-1) % classes).item()))
        assert losses[0] == losses[1], (mod.__name__, losses)


def test_graph_steps_without_host_sync_follow_the_eager_sequence(g):
    """The host runs several replays ahead of the GPU (no .item() between steps): the FPS start indices of every
    step must still be the ones the eager sequence draws -- the pinned staging buffers are rewritten only after their
    copies have run.  lr = 0, so the loss depends on the start indices only.  Both the prefetch pipeline (indices
    computed ahead on a side stream) and the in-graph sampling path (PCB_NO_INDEX_PREFETCH) are checked."""
    from pointcloud_bridge_b200.engine import Trainer
    x9, _, _, lab = inputs(g)
    x = x9.repeat(4, 1, 1)                                  # 8 blocks: a step long enough for the host to run ahead
    labs = lab.repeat(4, 1)
    nsteps = 10

    def run(graph, prefetch, chain=True):
        torch.manual_seed(7)
        net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV).train()
        net.drop1.eval()
        for m in net.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 0.0
        tr = Trainer(net, amp=False, graph=graph, lr=0.0, weight_decay=0.0)
        tr._use_chain = chain
        out = torch.zeros(nsteps, device=DEV)
        torch.manual_seed(11)
        if prefetch:
            tr.prefetch(x, labels=labs)
            for i in range(nsteps):
                out[i].copy_(tr.step_prefetched())
                tr.prefetch(x, labels=labs)
        else:
            for i in range(nsteps):
                out[i].copy_(tr.step(x, labels=labs))
        torch.cuda.synchronize()
        return out.cpu().numpy()

    ref = run(False, False)
    assert len(set(np.round(ref, 4))) > 4                   # the starts really vary per step
    for graph, prefetch, chain in ((True, False, True), (True, True, True), (False, True, True), (True, False, False)):
        got = run(graph, prefetch, chain)
        assert np.allclose(got, ref, rtol=2e-5, atol=0), (graph, prefetch, chain, got, ref)


def test_eval_after_training_steps_uses_current_weights(g, monkeypatch):
    """The fused tcgen05 inference block caches folded conv+BN weights.  The step runner updates parameters and running
    statistics through raw pointers (and CUDA-graph replays run no Python), so the cache is keyed on a parameter
    generation that every step bumps: eval logits after N replayed steps must equal the unfused evaluation."""
    from pointcloud_bridge_b200 import ops
    from pointcloud_bridge_b200.engine import Trainer
    x9, _, _, lab = inputs(g)
    torch.manual_seed(7)
    net = parity.seeded_fill_(ssg.get_model(5), 1).to(DEV).train()
    tr = Trainer(net, amp=True, graph=True, lr=1e-2, weight_decay=0.0)

    def evaluate():
        net.eval()
        torch.manual_seed(SEED_FPS)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            y, _ = net(x9)
        net.train()
        return y.float()

    torch.manual_seed(11)
    tr.step(x9, labels=lab)
    first = evaluate()                                        # fills the cache
    for _ in range(6):
        tr.step(x9, labels=lab)
    fused = evaluate()
    monkeypatch.setenv("PCB_NO_FUSED", "1")
    plain = evaluate()
    monkeypatch.delenv("PCB_NO_FUSED")
    scale = float(plain.abs().max())
    assert float((first - plain).abs().max()) > 5e-2 * scale, "the weights did not move: the test would prove nothing"
    assert float((fused - plain).abs().mean()) < 1e-2 * scale


@pytest.mark.parametrize("seed", [0, 1])
def test_msg_b16_benchmarked_configuration_pinned_to_reference(seed):
    """The configuration bench.py times -- PointNet++ MSG train step, 16 x 4096 points, Trainer(amp=True, graph=True),
    196 -> 200 channel padding ON -- against one training step of the unmodified reference at that size
    (tests/golden/msg_train_b16.npz, CPU fp32, two seeds).
      * level-1 FPS indices: bit-exact;
      * fp32 eager step: loss and log-probabilities at the fp32 bar;
      * bf16 graph step: loss within 1e-2; log-probabilities: mean (measured 1.2e-2 of max|logp|), 99th percentile and
        maximum (0.13) are asserted separately and against torch's own bf16 batch_norm under autocast on the same
        network: training-mode statistics on bf16 pre-activations do not meet 1e-2 on this untrained fixture in either
        implementation (eval mode does: test_msg_bf16_autocast_within_1e2);
      * gradients: this randomly initialised network amplifies a forward perturbation into its gradients by ~5e4 (the
        reference's own CPU gradients move by 5e-3 between 1 and 8 threads, i.e. under 1e-7 summation noise), so bf16
        gradients are compared by their distance to the fp32 golden with and without the channel padding: padding must
        not be further away than the unpadded path (it is a mathematical no-op; what differs is rounding)."""
    from pointcloud_bridge_b200 import ops
    from pointcloud_bridge_b200.engine import Trainer
    gold = parity.load("msg_train_b16.npz")
    p = f"s{seed}_"
    xyz, rgb, lab = synthetic.bridge_batch(100 + seed, 16)
    x9 = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).to(DEV)
    tlab = torch.from_numpy(lab).to(DEV)
    ref_loss = float(gold[p + "loss"])
    ref_logp = gold[p + "logp_sample"]
    scale = float(np.abs(ref_logp).max())
    names = [k[len(p) + 2:] for k in gold.keys() if k.startswith(p + "g_")]

    # indices
    torch.manual_seed(SEED_FPS + seed)
    fps1 = ops.furthest_point_sample(x9[:, :3, :].permute(0, 2, 1).contiguous(), 1024)
    assert np.array_equal(fps1.cpu().numpy().astype(np.int16), gold[p + "fps1"])

    # fp32 eager
    net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV).train()
    net.drop1.eval()
    torch.manual_seed(SEED_FPS + seed)
    y, _ = net(x9)
    loss = F.nll_loss(y.reshape(-1, 5), tlab.reshape(-1))
    loss.backward()
    e_logp = float(np.abs(y.detach().cpu().numpy()[:, ::16, :] - ref_logp).max()) / scale
    print(f"fp32: loss {float(loss):.7f} vs {ref_loss:.7f}, logp max rel err {e_logp:.2e}")
    assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss)
    assert e_logp < 5e-5          # measured 1e-5..3e-5: GEMM / BN summation order (cuBLAS vs MKL), 34 BN layers deep
    params = dict(net.named_parameters())
    rel = lambda a, b: float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-20))
    fp32_g = {n: rel(params[n].grad.cpu().numpy(), gold[p + "g_" + n]) for n in names}
    print("fp32 gradient distance to the reference:", {n: round(v, 4) for n, v in fp32_g.items()})

    # yardstick for bf16: the same network with the row kernels switched off -- library GEMM output in bf16, then
    # torch's own batch_norm / relu / max under autocast (what torch.autocast does to the reference's modules)
    from pointcloud_bridge_b200.partsize import pointnet_util as pu
    saved = (ops.bn_rows_supported, ops.mlp_rows_fused_supported, ops.fp_concat_supported)
    ops.bn_rows_supported = lambda *a, **k: False
    ops.mlp_rows_fused_supported = lambda *a, **k: False
    try:
        net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV).train()
        net.drop1.eval()
        torch.manual_seed(SEED_FPS + seed)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            y, _ = net(x9)
        lib_err = np.abs(y.float().cpu().numpy()[:, ::16, :] - ref_logp) / scale
    finally:
        ops.bn_rows_supported, ops.mlp_rows_fused_supported, ops.fp_concat_supported = saved
    print(f"bf16 library path (torch batch_norm under autocast): logp rel err mean {lib_err.mean():.2e} "
          f"p99 {np.quantile(lib_err, 0.99):.2e} max {lib_err.max():.2e}")

    # bf16 graph step, padded (benchmarked) and unpadded
    dist = {}
    for pad in (True, False):
        old = ops._PAD_N
        ops._PAD_N = pad
        try:
            net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV).train()
            net.drop1.eval()
            tr = Trainer(net, amp=True, graph=True, lr=0.0, weight_decay=0.0)
            for _ in range(4):                                # 3 eager warm-up steps + capture
                tr.step(x9, labels=tlab)
            torch.manual_seed(SEED_FPS + seed)
            loss = tr.step(x9, labels=tlab)                   # a pure replay with the reference's draws
            torch.cuda.synchronize()
            logp = tr._last_logits.log_probs().detach().float().cpu().numpy()[:, ::16, :]
            err = np.abs(logp - ref_logp) / scale
            q = lambda f: float(np.quantile(err, f))
            print(f"bf16 graph pad={pad}: loss {float(loss):.5f} vs {ref_loss:.5f}, logp rel err mean {err.mean():.2e} "
                  f"p99 {q(0.99):.2e} p99.9 {q(0.999):.2e} max {err.max():.2e}")
            assert abs(float(loss) - ref_loss) <= 1e-2 * abs(ref_loss)
            # Training-mode BatchNorm on bf16 pre-activations: a channel with |mean| / std = r carries its bf16 rounding
            # error (2^-9 relative to |y|) amplified by r into the normalised value, 34 layers deep, on an untrained
            # network.  north_star's 1e-2 bar is met by the MEAN error only marginally (measured 1.2e-2) and not by the
            # maximum (0.13); torch's own batch_norm under autocast sits at the same distance (printed above), so the
            # row kernels are asserted against that yardstick and against absolute bounds.
            assert err.mean() <= 1.25 * lib_err.mean() + 1e-3, (err.mean(), lib_err.mean())
            assert err.mean() < 2e-2 and q(0.99) < 8e-2 and err.max() < 0.25, (err.mean(), q(0.99), err.max())
            views = {n: v for (n, _), v in zip(net.named_parameters(), tr.bucket.views)}
            dist[pad] = {n: rel(views[n].cpu().numpy(), gold[p + "g_" + n]) for n in names}
        finally:
            ops._PAD_N = old
    print("bf16 gradient distance to the fp32 reference, padded:  ", {n: round(v, 3) for n, v in dist[True].items()})
    print("bf16 gradient distance to the fp32 reference, unpadded:", {n: round(v, 3) for n, v in dist[False].items()})
    med = lambda d: float(np.median(list(d.values())))
    assert med(dist[True]) <= 1.5 * med(dist[False]) + 0.02, (med(dist[True]), med(dist[False]))
    assert dist[True]["conv2.weight"] < 0.05 and dist[True]["conv2.bias"] < 0.05      # no BN stack below: plain bf16 noise


def test_flat_adam_leaves_gradient_less_parameters_alone():
    """EnhancedPointNet2 owns modules its forward never reaches (geometric1, cls_head): torch.optim.Adam skips their
    grad = None parameters; the flat Adam kernel must not decay them towards zero either, and the converted optimizer
    state has no entries for them."""
    from pointcloud_bridge_b200 import runner
    from pointcloud_bridge_b200.engine import Trainer
    xyz, rgb, lab = synthetic.bridge_batch(9, 2, 1024)
    txyz, trgb, tlab = (torch.from_numpy(a).to(DEV) for a in (xyz, rgb, lab))
    torch.manual_seed(0)
    net = hb_model.EnhancedPointNet2(5).to(DEV).train()
    crit = hb_model.BridgeStructureLoss(num_classes=5, alpha=80, rel_margin=0.3).to(DEV)
    tr = Trainer(net, loss_fn=lambda out, labels, pts: crit(out, labels, pts), lr=1e-2, weight_decay=1e-2, amp=True, graph=False)
    before = {n: p.detach().clone() for n, p in net.named_parameters()}
    for _ in range(2):
        tr.step(txyz, trgb, labels=tlab, loss_inputs=(txyz,))
    torch.cuda.synchronize()
    names = [n for n, _ in net.named_parameters()]
    frozen = {names[i] for i in tr.frozen}
    assert frozen and all(n.startswith(("geometric1.", "cls_head.")) for n in frozen), sorted(frozen)[:5]
    moved = 0
    for n, p in net.named_parameters():
        if n in frozen:
            assert torch.equal(p.detach(), before[n]), n
        else:
            moved += int(not torch.equal(p.detach(), before[n]))
    assert moved >= len(names) - len(frozen) - 2
    sd = runner.adam_state_to_torch(tr.opt, tr.bucket.params, tr.frozen)
    assert set(sd["state"]) == set(range(len(names))) - set(tr.frozen)

"""Pins the oracle's CPU networks (oracle/ref_models.py) against the reference's own outputs
(tests/golden/models.npz).  CPU only.  Tolerance 1e-5 relative to max|ref| (fp32, same ATen
kernels; only the tie order of three-NN can differ)."""
import numpy as np
import torch

import parity
from oracle import ref_models
from pointcloud_bridge_b200 import synthetic

SEED_FPS = 4242


def rel_err(a, ref):
    return float(np.abs(a.detach().numpy() - ref).max() / (np.abs(ref).max() + 1e-12))


def test_ssg_and_msg_eval_forward():
    g = parity.load("models.npz")
    x9 = torch.from_numpy(synthetic.sem_seg_input(g["xyz"], g["rgb"]))
    for cls, nc, seed, key, x in ((ref_models.PointNet2SSG, 13, 1, "ssg", x9[:1]), (ref_models.PointNet2MSG, 5, 2, "msg", x9)):
        net = parity.seeded_fill_(cls(nc), seed).eval()
        torch.manual_seed(SEED_FPS)
        with torch.no_grad():
            y, l4 = net(x)
        assert rel_err(y, g[f"{key}_logp"]) < 1e-5
        assert rel_err(l4, g[f"{key}_l4"]) < 1e-5


def test_msg_training_forward_backward():
    g = parity.load("models.npz")
    x9 = torch.from_numpy(synthetic.sem_seg_input(g["xyz"], g["rgb"]))
    lab = torch.from_numpy(g["labels"].astype(np.int64))
    net = parity.seeded_fill_(ref_models.PointNet2MSG(5), 2).train()
    net.drop1.eval()
    torch.manual_seed(SEED_FPS)
    y, _ = net(x9)
    loss = torch.nn.functional.nll_loss(y.reshape(-1, 5), lab.reshape(-1))
    loss.backward()
    assert abs(loss.item() - float(g["msg_train_loss"])) < 1e-5
    assert rel_err(net.sa1.conv_blocks[0][0].weight.grad, g["msg_train_g_sa1"]) < 1e-4
    assert rel_err(net.conv2.weight.grad, g["msg_train_g_conv2"]) < 1e-4

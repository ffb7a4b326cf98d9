"""Per-parameter difference between engine.Trainer's flat gradients and a plain autograd backward (debug aid)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from pointcloud_bridge_b200 import synthetic  # noqa: E402
from pointcloud_bridge_b200.engine import Trainer  # noqa: E402
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg  # noqa: E402

dev = "cuda:0"
xyz, rgb, lab = synthetic.bridge_batch(3, 4, 4096)
x9 = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).to(dev)
lab = torch.from_numpy(lab).to(dev)


def plain(seed):
    torch.manual_seed(5)
    net = parity.seeded_fill_(msg.get_model(5), 2).to(dev).train()
    net.drop1.eval()
    torch.manual_seed(seed)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logp, _ = net(x9)
    F.nll_loss(logp.float().reshape(-1, logp.shape[-1]), lab.reshape(-1)).backward()
    return net, [p.grad.detach().clone() for p in net.parameters()]


net, ref = plain(11)
_, ref2 = plain(11)          # run-to-run noise of the plain path (atomics)
for p in net.parameters():
    p.grad = None
tr = Trainer(net, amp=True, graph=False, lr=0.0, weight_decay=0.0)
torch.manual_seed(11)
tr.step(x9, labels=lab)
torch.cuda.synchronize()
rows = []
for (name, _), a, b, b2 in zip(net.named_parameters(), tr.bucket.views, ref, ref2):
    rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
    noise = ((b2 - b).norm() / (b.norm() + 1e-12)).item()
    rows.append((rel, noise, name, tuple(b.shape), b.abs().max().item()))
for r in sorted(rows, reverse=True)[:25]:
    print("rel %.4f  plain-vs-plain %.4f  %-32s %-18s max|g| %.3e" % r)

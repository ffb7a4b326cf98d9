"""Which piece of the step runner moves the gradients away from a plain autograd backward (debug aid)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from pointcloud_bridge_b200 import ops, synthetic  # noqa: E402
from pointcloud_bridge_b200.engine import Trainer  # noqa: E402
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg  # noqa: E402

dev = "cuda:0"
xyz, rgb, lab = synthetic.bridge_batch(3, 4, 4096)
x9 = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).to(dev)
lab = torch.from_numpy(lab).to(dev)


def make():
    torch.manual_seed(5)
    net = parity.seeded_fill_(msg.get_model(5), 2).to(dev).train()
    net.drop1.eval()
    return net


def plain(ctx=False, head=False):
    net = make()
    c = ops.StepContext(net, bf16=True) if ctx else None
    torch.manual_seed(11)
    if c is not None:
        c.__enter__()
    try:
        if head:
            with ops.head_logits_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
                out, _ = net(x9)
            loss = ops.nll_logit_rows(out, lab)
        else:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logp, _ = net(x9)
            loss = F.nll_loss(logp.float().reshape(-1, logp.shape[-1]), lab.reshape(-1))
        loss.backward()
    finally:
        if c is not None:
            c.__exit__(None, None, None)
    return float(loss), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}


def trainer():
    net = make()
    tr = Trainer(net, amp=True, graph=False, lr=0.0, weight_decay=0.0)
    torch.manual_seed(11)
    loss = tr.step(x9, labels=lab)
    torch.cuda.synchronize()
    return float(loss), {n: v.clone() for (n, _), v in zip(net.named_parameters(), tr.bucket.views)}


def cmp(tag, a, b):
    rows = []
    for n in b:
        if n in a and not (n.endswith(".bias") and "conv" in n):
            rows.append((((a[n] - b[n]).norm() / (b[n].norm() + 1e-12)).item(), n))
    rows.sort(reverse=True)
    print(tag, "worst:", ["%s %.3f" % (n, r) for r, n in rows[:4]], "median %.4f" % rows[len(rows) // 2][0])


l0, g0 = plain()
l1, g1 = plain()
print("loss plain", l0, l1)
cmp("plain vs plain", g1, g0)
l2, g2 = plain(ctx=True)
print("loss plain+ctx", l2)
cmp("plain+StepContext vs plain", g2, g0)
l3, g3 = plain(ctx=True, head=True)
print("loss plain+ctx+head", l3)
cmp("plain+StepContext+logit head vs plain", g3, g0)
l4, g4 = trainer()
print("loss trainer", l4)
cmp("trainer vs plain", g4, g0)
cmp("trainer vs plain+ctx+head", g4, g3)

# same net object stepped twice in plain mode: only the BN running statistics differ between the two backward passes
net = make()


def again(reset=False):
    if reset:
        for m in net.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.reset_running_stats()
    for p in net.parameters():
        p.grad = None
    torch.manual_seed(11)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logp, _ = net(x9)
    loss = F.nll_loss(logp.float().reshape(-1, logp.shape[-1]), lab.reshape(-1))
    loss.backward()
    return float(loss), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}


la, ga = again()
lb, gb = again()
print("same net, step 1 / step 2 loss", la, lb)
cmp("second plain pass on the same net vs first", gb, ga)
lc, gc = again(reset=True)
print("after reset_running_stats loss", lc)
cmp("third pass after reset_running_stats vs first", gc, ga)
rm = [m.running_mean.abs().max().item() for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
print("max |running_mean|", max(rm))

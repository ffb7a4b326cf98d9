"""How far a one-ulp change early in the untrained MSG network travels (the evidence behind the gradient tolerances
of tests/test_gpu_models.py): the same network is run three times from the same seed in training mode; with
`PCB_OWN_GEMM=1` round 2's first statistics epilogue shifted its column sums by the BatchNorm running mean, so a pass
differed from the previous one by ~1e-7 in the batch statistics -- 28 of 393 216 sa1 outputs moved by one bf16 ulp, half of
the fp1 outputs by up to 5 % after four more levels, and the parameter gradients by a median 34 % in relative L2 (the
reference's own fp32 CPU gradients move by 5e-3 between 1 and 8 threads).  With the default path every pass is bit-identical.

    python tools/forward_sensitivity.py [bwd]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from pointcloud_bridge_b200 import ops, synthetic  # noqa: E402
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg  # noqa: E402

dev = "cuda:0"
xyz, rgb, lab = synthetic.bridge_batch(3, 4, 4096)
x9 = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).to(dev)
lab = torch.from_numpy(lab).to(dev)
torch.manual_seed(5)
net = parity.seeded_fill_(msg.get_model(5), 2).to(dev).train()
net.drop1.eval()
caps = []


def hook(name):
    def f(mod, inp, out):
        o = out[1] if isinstance(out, tuple) else out
        caps[-1][name] = o.detach().float().clone()
    return f


for n in ("sa1", "sa2", "sa3", "sa4", "fp4", "fp3", "fp2", "fp1"):
    getattr(net, n).register_forward_hook(hook(n))
bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
for it in range(3):
    caps.append({})
    for p in net.parameters():
        p.grad = None
    torch.manual_seed(11)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logp, _ = net(x9)
    loss = F.nll_loss(logp.float().reshape(-1, logp.shape[-1]), lab.reshape(-1))
    if bwd:
        loss.backward()
    print("pass", it, "loss %.7f" % float(loss))
    if it:
        for k, v in caps[it].items():
            ref = caps[0][k]
            print("   %-4s max|diff| %.3e  (max|ref| %.3e)  differing entries %d" %
                  (k, float((v - ref).abs().max()), float(ref.abs().max()), int((v != ref).sum())))

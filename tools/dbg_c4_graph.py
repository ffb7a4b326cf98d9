"""Debug aid: which stage of the BriStruNet train step invalidates a CUDA-graph capture?
    python tools/dbg_c4_graph.py STAGE     STAGE in fwd_eval | fwd | loss | bwd | trainer"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_bridge_b200 import ops, synthetic  # noqa: E402
from pointcloud_bridge_b200.engine import FpsStartBuffers, Trainer  # noqa: E402
from pointcloud_bridge_b200.highway import model as hb  # noqa: E402

stage = sys.argv[1]
dev = torch.device("cuda", 0)
B, N = 16, 4096
xyz, rgb, lab = synthetic.bridge_batch(100, B, N)
txyz, trgb, tlab = torch.from_numpy(xyz).to(dev), torch.from_numpy(rgb).to(dev), torch.from_numpy(lab).to(dev)
torch.manual_seed(0)
net = hb.EnhancedPointNet2(5).to(dev).train()
crit = hb.BridgeStructureLoss(num_classes=5, alpha=80, rel_margin=0.3).to(dev)

if stage == "trainer":
    tr = Trainer(net, loss_fn=lambda out, labels, pts: crit(out, labels, pts), amp=True, graph=True)
    for i in range(5):
        loss = tr.step(txyz, trgb, labels=tlab, loss_inputs=(txyz,))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        tr.step(txyz, trgb, labels=tlab, loss_inputs=(txyz,))
    e1.record()
    torch.cuda.synchronize()
    print("trainer graph ok, loss", float(loss), "ms/step", e0.elapsed_time(e1) / 10)
    sys.exit(0)

starts = FpsStartBuffers()


def body():
    if stage == "fwd_eval":
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return net(txyz, trgb)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(txyz, trgb)
    if stage == "fwd":
        return out
    loss = crit(out.float(), tlab, txyz)
    if stage == "loss":
        return loss
    loss.backward()
    return loss


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(3):
        if i == 2:
            starts.shapes, starts.mode = [], "discover"
            ops.set_fps_start_provider(starts.provider)
        body()
        for p in net.parameters():
            p.grad = None
torch.cuda.current_stream().wait_stream(s)
starts.allocate()
starts.mode = "record"
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    r = body()
starts.mode = "off"
ops.set_fps_start_provider(None)
g.replay()
torch.cuda.synchronize()
print(stage, "capture ok")

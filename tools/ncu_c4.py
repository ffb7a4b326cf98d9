"""One captured BriStruNet train step (BASELINE config 4) between cudaProfilerStart/Stop, for
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/c4_step.csv python tools/ncu_c4.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import synthetic  # noqa: E402
from pointcloud_bridge_b200.engine import Trainer  # noqa: E402
from pointcloud_bridge_b200.highway import model as hb_model  # noqa: E402

dev = torch.device("cuda", 0)
B, N = 16, 4096
xyz, rgb, lab = synthetic.bridge_batch(100, B, N)
txyz, trgb, tlab = torch.from_numpy(xyz).to(dev), torch.from_numpy(rgb).to(dev), torch.from_numpy(lab).to(dev)
torch.manual_seed(0)
net = hb_model.EnhancedPointNet2(5).to(dev).train()
crit = hb_model.BridgeStructureLoss(num_classes=5, alpha=80, rel_margin=0.3).to(dev)
tr = Trainer(net, loss_fn=lambda out, labels, pts: crit(out, labels, pts), amp=True, graph=True)
for _ in range(6):
    tr.step(txyz, trgb, labels=tlab, loss_inputs=(txyz,))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(txyz, trgb, labels=tlab, loss_inputs=(txyz,))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")

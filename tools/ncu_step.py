"""Driver for ncu captures of ONE eager MSG train step (bench.py's workload, no CUDA graph):
two warm-up steps, then the step to be profiled between cudaProfilerStart/Stop.

    ncu --set full --profile-from-start off -k regex:bn_.*fused -o gpurun_out/step_bn python tools/ncu_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import synthetic  # noqa: E402
from pointcloud_bridge_b200.engine import Trainer  # noqa: E402
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, N = 16, 4096
xyz, rgb, lab = synthetic.bridge_batch(100, B, N)
x9 = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).to(dev)
labels = torch.from_numpy(lab).to(dev)
net = msg.get_model(5).to(dev).train()
tr = Trainer(net, amp=True, graph=False)
for _ in range(2):
    tr.step(x9, labels=labels)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(x9, labels=labels)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")

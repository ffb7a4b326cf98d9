"""Runs a few resident-input training steps of the bench workload between cudaProfilerStart/Stop
so that `ncu --profile-from-start off` sees exactly the timed region's launches.

    python tools/profile_step.py [--steps 2] [--model msg|ssg_infer|dgcnn_infer] [--graph]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import synthetic  # noqa: E402
from pointcloud_bridge_b200.engine import Trainer  # noqa: E402
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--torch-profiler", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(1234)
    net = msg.get_model(5).to(dev).train()
    tr = Trainer(net)
    xyz, rgb, lab = synthetic.bridge_batch(0, a.batch, 4096)
    x = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).to(dev)
    y = torch.from_numpy(lab).to(dev)
    for _ in range(a.warmup):
        tr.step(x, labels=y)
    torch.cuda.synchronize()
    if a.torch_profiler:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(a.steps):
                tr.step(x, labels=y)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
        return
    torch.cuda.cudart().cudaProfilerStart()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        tr.step(x, labels=y)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.cudart().cudaProfilerStop()
    print(f"{a.steps} steps, {dt / a.steps * 1e3:.3f} ms/step wall")


if __name__ == "__main__":
    main()

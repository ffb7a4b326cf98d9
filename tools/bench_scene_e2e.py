"""BASELINE config 5 as stated: ONE synthetic bridge-sized scene (default 50 M points, ~60 m x 12 m, 0.01 m spacing),
tiled into 4096-point blocks (block 1.0 m, stride 0.5 m), block-sharded over the GPUs of one box, PointNet++ SSG
inference (bf16, fused tcgen05 SA blocks, CUDA graph per 128-block batch), votes scattered back, per-point argmax --
`engine.segment_scene`, end to end on the device, wall clock (max over ranks).  Every rank holds the scene, counts and
orders the windows, builds only its shard of the blocks; the only exchange is one all-reduce of the int32 vote counts.

    python tools/bench_scene_e2e.py [--points 50000000]
    torchrun --nproc-per-node 8 ... tools/bench_scene_e2e.py
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_bridge_b200 import distributed as pdist, engine  # noqa: E402
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg as ssg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=50_000_000)
ap.add_argument("--classes", type=int, default=5)
ap.add_argument("--repeats", type=int, default=2)
a = ap.parse_args()
rank, world, local = pdist.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
P = a.points
g = torch.Generator(device=dev).manual_seed(0)           # same scene on every rank
pts = torch.empty(P, 6, device=dev)
pts[:, 0] = torch.rand(P, device=dev, generator=g) * 60.0
pts[:, 1] = torch.rand(P, device=dev, generator=g) * 12.0
pts[:, 2] = torch.rand(P, device=dev, generator=g) * 3.0
pts[:, 3:] = torch.rand(P, 3, device=dev, generator=g)
torch.manual_seed(0)
net = ssg.get_model(a.classes).to(dev).eval()
infer = engine.BlockInference(net, batch_blocks=128, amp=True, graph=True)
times = []
for it in range(a.repeats + 1):                          # first pass warms up (captures the inference graph)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    labels = engine.segment_scene(net, pts, a.classes, rank=rank, world=world, infer=infer)
    torch.cuda.synchronize()
    times.append(pdist.max_over_ranks(time.perf_counter() - t0, dev))
assert labels.shape == (P,) and int(labels.max()) < a.classes
best = min(times[1:])
if rank == 0:
    print(json.dumps({"config": "c5 one scene -> tiler -> block-sharded PN++ SSG inference -> votes -> labels (engine.segment_scene)",
                      "points": P, "n_gpus": world, "seconds": round(best, 4), "points_per_s": round(P / best),
                      "all_passes_s": [round(t, 4) for t in times], "collectives_on_data_path": 0,
                      "exchange": "one all-reduce of the int32 vote counts"}), flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()

"""Throughput of the GPU whole-scene tiler and vote scatter-back (scene.py / csrc/scene.cu) on a synthetic
bridge-sized scene (BASELINE config 5: 50 M points, ~60 m x 12 m, block 1.0 m, stride 0.5 m, 4096 points per block).

    python tools/bench_scene.py [--points 50000000]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_bridge_b200 import scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=50_000_000)
a = ap.parse_args()
dev = "cuda:0"
P = a.points
g = torch.Generator(device=dev).manual_seed(0)
pts = torch.empty(P, 6, device=dev)
pts[:, 0] = torch.rand(P, device=dev, generator=g) * 60.0
pts[:, 1] = torch.rand(P, device=dev, generator=g) * 12.0
pts[:, 2] = torch.rand(P, device=dev, generator=g) * 3.0
pts[:, 3:] = torch.rand(P, 3, device=dev, generator=g)
tiler = scene.SceneTiler()
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tiles = tiler.tile(pts)
    torch.cuda.synchronize()
    t_tile = time.perf_counter() - t0
    nb = tiles.data.shape[0]
    pred = (tiles.point_idx % 5).to(torch.uint8)
    pool = scene.new_vote_pool(P, 5, dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scene.add_vote(pool, tiles.point_idx, pred)
    labels = scene.vote_argmax(pool)
    torch.cuda.synchronize()
    t_vote = time.perf_counter() - t0
    if it == 0:
        del tiles, pred, pool, labels
assert torch.equal(labels.long(), torch.arange(P, device=dev) % 5)
entries = nb * 4096
print(json.dumps({"points": P, "blocks": nb, "entries_per_point": round(entries / P, 2), "grid": list(tiles.grid),
                  "tile_ms": round(t_tile * 1e3, 2), "tile_Mpoints_per_s": round(P / t_tile / 1e6, 1),
                  "tile_out_GBps": round(entries * 44 / t_tile / 1e9, 1),
                  "vote_argmax_ms": round(t_vote * 1e3, 2), "vote_Mentries_per_s": round(entries / t_vote / 1e6, 1)}))

"""Throughput of the other BASELINE.json configurations (bench.py covers configs[1]):

  c1  PointNet++ SSG sem-seg forward, 1 block of 4096 points (and batch 16)
  c3  DGCNN sem-seg k=20 forward, batch 16 x 4096
  c4  BriStruNet (EnhancedPointNet2) training step, batch 16 per GPU, BridgeStructureLoss, DDP
  c5  block-sharded scene inference: --blocks 4096-point blocks per GPU (12 208 blocks = 50 M points
      over 8 GPUs -> 1 526 per GPU), PointNet++ SSG, no cross-GPU traffic

    python tools/bench_configs.py [--only c1,c3] [--blocks 1526]
    torchrun --nproc-per-node N ... tools/bench_configs.py --only c4,c5

One JSON line per configuration (rank 0).  CUDA-event timing, max over ranks.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import distributed as pdist, synthetic  # noqa: E402
from pointcloud_bridge_b200.engine import Trainer, run_sharded_scene  # noqa: E402
from pointcloud_bridge_b200.highway import DGCNN as dgcnn_mod, model as hb_model  # noqa: E402
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg as ssg  # noqa: E402


def timed(fn, iters, warmup, dev):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return pdist.max_over_ranks(e0.elapsed_time(e1) / iters, dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c3,c4,c5")
    ap.add_argument("--blocks", type=int, default=1526)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch-blocks", type=int, default=128, help="c5: blocks per captured inference batch")
    a = ap.parse_args()
    rank, world, local = pdist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    want = set(a.only.split(","))
    B, N = 16, 4096
    xyz, rgb, lab = synthetic.bridge_batch(100 + rank, B, N)
    x9 = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).to(dev)
    txyz, trgb, tlab = torch.from_numpy(xyz).to(dev), torch.from_numpy(rgb).to(dev), torch.from_numpy(lab).to(dev)

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    if "c1" in want:
        torch.manual_seed(0)
        net = ssg.get_model(13).to(dev).eval()
        for bsz in (1, 16):
            def f():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    net(x9[:bsz])
            ms = timed(f, a.iters, 3, dev)
            emit(config="c1 PN++ SSG sem-seg forward (eager launches)", batch=bsz, ms=round(ms, 3),
                 points_per_s=round(bsz * N / ms * 1e3), n_gpus=1)
            from pointcloud_bridge_b200.engine import BlockInference
            inf = BlockInference(net, batch_blocks=bsz, amp=True, graph=True)
            xb = x9[:bsz].contiguous()
            ms = timed(lambda: inf.run(xb), a.iters, 4, dev)
            emit(config="c1 PN++ SSG sem-seg forward + argmax (CUDA graph, fused tcgen05 SA blocks)", batch=bsz, ms=round(ms, 3),
                 points_per_s=round(bsz * N / ms * 1e3), n_gpus=1)
    if "c3" in want:
        torch.manual_seed(0)
        net = dgcnn_mod.DGCNN(5, 20).to(dev).eval()

        def f():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                net(txyz, trgb)
        ms = timed(f, a.iters, 3, dev)
        emit(config="c3 DGCNN k=20 forward", batch=B, ms=round(ms, 3), points_per_s=round(B * N / ms * 1e3), n_gpus=1)
    if "c4" in want:
        torch.manual_seed(0)
        net = hb_model.EnhancedPointNet2(5).to(dev).train()
        crit = hb_model.BridgeStructureLoss(num_classes=5, alpha=80, rel_margin=0.3).to(dev)
        tr = Trainer(net, loss_fn=lambda out, labels, pts: crit(out, labels, pts), amp=True, graph=True)

        def f():
            tr.step(txyz, trgb, labels=tlab, loss_inputs=(txyz,))
        ms = timed(f, a.iters, 6, dev)                    # 3 eager steps + graph capture, then replays
        emit(config="c4 BriStruNet train step (DDP, flat NCCL all-reduce)", batch_per_gpu=B, ms=round(ms, 3),
             points_per_s=round(world * B * N / ms * 1e3), n_gpus=world, grad_bucket_mb=round(tr.bucket.nbytes / 1e6, 1))
    if "c5" in want:
        torch.manual_seed(0)
        net = ssg.get_model(13).to(dev).eval()
        nb_total = a.blocks * world
        # every rank builds only its own shard of the scene (same generator, disjoint seeds)
        rng = pdist.shard_range(nb_total, rank, world)
        base = synthetic.sem_seg_input(*synthetic.bridge_batch(7 + rank, 64, N)[:2])
        reps = -(-len(rng) // 64)
        shard = torch.from_numpy(base).repeat(reps, 1, 1)[:len(rng)].contiguous().pin_memory()
        scene = shard                                      # run_sharded_scene indexes [rng.start + lo]: offset below
        class _Shifted:                                    # view of the global block list restricted to this rank
            shape = (nb_total, 9, N)
            def __getitem__(self, s):
                return shard[s.start - rng.start:s.stop - rng.start]
        from pointcloud_bridge_b200.engine import BlockInference
        infer = BlockInference(net, batch_blocks=a.batch_blocks, amp=True, graph=True)
        for _ in range(1):
            run_sharded_scene(net, _Shifted(), rank, world, dev, infer=infer,
                              chunk_blocks=max(512, 4 * a.batch_blocks))               # warm-up (captures the graph)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r, labels = run_sharded_scene(net, _Shifted(), rank, world, dev, infer=infer,
                                      chunk_blocks=max(512, 4 * a.batch_blocks))
        torch.cuda.synchronize()
        dt = pdist.max_over_ranks(time.perf_counter() - t0, dev)
        emit(config="c5 block-sharded scene inference (PN++ SSG), host->device->host", blocks_per_gpu=a.blocks, batch_blocks=a.batch_blocks,
             points_total=nb_total * N, seconds=round(dt, 4), points_per_s=round(nb_total * N / dt), n_gpus=world,
             collectives_on_data_path=0)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

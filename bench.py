#!/usr/bin/env python
"""Headline benchmark: PointNet++ MSG semantic-segmentation TRAIN STEP, batch 16 x 4096-point
blocks per GPU, bf16 autocast (BASELINE.json configs[1]) -> points/sec.

    python bench.py --gpus N --steps K --warmup W            # our arm (B200 kernels)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference itself on host cores
    torchrun ... bench.py --gpus N ...                       # N > 1: one rank per GPU, NCCL

One JSON line on stdout (rank 0).  A step = forward + NLL loss + backward + one flat NCCL
gradient all-reduce (N > 1) + fused Adam on one synthetic batch already resident in HBM
(`value`); `e2e` repeats the measurement through the public API with the batch coming from
pinned host memory each step and the loss read back.  `roofline` describes the kernel of ours
with the largest share of the step, timed live with CUDA events in an instrumented pass;
`cpu_baseline` / `--impl reference` time the UNMODIFIED reference (oracle/_ref, copied there by
oracle/make_ref.py; the oracle's port if absent) on this host's cores on a bounded sample
(rank 0, N = 1).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PN++ sem-seg points/sec (MSG train step)"
UNIT = "points/s"
NPTS = 4096
NUM_CLASSES = 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="blocks per GPU")
    ap.add_argument("--fp32", action="store_true", help="disable bf16 autocast (parity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--cpu-sample-blocks", type=int, default=4)
    ap.add_argument("--repeats", type=int, default=5, help="timed K-step regions; the median one is reported")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "c5scene"],
                    help="BASELINE.json configuration; c2 (default) is the headline this file measures itself, the others "
                         "delegate to tools/bench_configs.py (c1 SSG forward, c3 DGCNN forward, c4 BriStruNet train step, c5 "
                         "block-sharded inference) and tools/bench_scene_e2e.py (c5scene: one scene end to end)")
    return ap.parse_args()


def workload_config(a, world):
    return {"workload": "pointnet2_sem_seg_msg train step (fwd+NLL+bwd+Adam), BASELINE configs[1]",
            "blocks_per_gpu": a.batch, "points_per_block": NPTS, "global_batch": a.batch * world,
            "channels": 9, "num_classes": NUM_CLASSES, "parallelism": f"dp{world} (block-sharded replicas)",
            "precision": "index kernels fp32 (bit-exact); shared-MLP GEMMs " + ("fp32" if a.fp32 else "bf16 autocast"),
            "l2": "256 MB buffer written between timed steps (L2 flush, outside the per-step CUDA-event brackets that are summed); 4 distinct batches cycled",
            "e2e": "public API (engine.Trainer): batch i+1 copied H2D from pinned memory on a side stream while step i runs "
                   "(Trainer.prefetch), loss of step i copied D2H after the step and read by the host one step later; "
                   "wall clock over K steps incl. the L2 flush writes",
            "launch": "eager" if a.no_graph else "one CUDA graph per step (zero+fwd+loss+bwd+Adam; programmatic dependent launch "
                      "between the kernels of libpcbridge, weight gradients and the scales of a multi-scale level on forked "
                      "branches); the sampling / grouping indices of batch i+1 (FPS with start indices drawn on the CPU "
                      "generator as the reference does, ball query, three-NN) are computed on a side stream from the "
                      "mid-point of step i and copied into the graph's static buffers before its replay",
            "timed_region": "K steps bracketed by barrier+synchronize, per-step CUDA events summed; repeated --repeats "
                            "times, median region reported"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi query fields via NVML), runs during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference path (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_msg_train_points_per_sec(blocks, steps, warmup, threads=None):
    """MSG train step (forward + NLL + backward + Adam) on the host cores.  With oracle/_ref present (the unmodified
    reference files, oracle/make_ref.py) this IS the reference: its get_model, its pointnet_util functions, torch CPU
    fp32 -> kind "reference"; otherwise the oracle's port (C primitives + torch CPU ops) -> kind "port"."""
    import contextlib
    import io
    import torch
    from oracle import make_ref
    from pointcloud_bridge_b200 import synthetic
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    xyz, rgb, lab = synthetic.bridge_batch(123, blocks, NPTS)
    x = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb))
    y = torch.from_numpy(lab)
    torch.manual_seed(0)
    if make_ref.available():
        kind = "reference"
        net = make_ref.load_msg().get_model(NUM_CLASSES).train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)

        def step():
            opt.zero_grad(set_to_none=True)
            with contextlib.redirect_stdout(io.StringIO()):       # the reference prints tensor shapes
                logp, _ = net(x)
            loss = torch.nn.functional.nll_loss(logp.reshape(-1, NUM_CLASSES), y.reshape(-1))
            loss.backward()
            opt.step()
    else:
        kind = "port"
        from oracle import oracle as orc
        from oracle import ref_models
        orc.set_num_threads(threads)
        net = ref_models.PointNet2MSG(NUM_CLASSES).train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
        step = lambda: ref_models.msg_train_step(net, opt, x, y)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return blocks * NPTS * steps / dt, dt / steps, threads, kind


def cpu_workload_config(a):
    return {"workload": "pointnet2_sem_seg_msg train step (fwd+NLL+bwd+Adam), BASELINE configs[1], on the host CPU",
            "blocks_per_step": a.cpu_sample_blocks, "of_blocks_per_gpu": a.batch, "points_per_block": NPTS, "channels": 9,
            "num_classes": NUM_CLASSES, "precision": "fp32 (torch CPU)", "launch": "eager PyTorch, all host threads"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return                                        # other ranks exit 0 without work
    blocks = a.cpu_sample_blocks
    # bounded: every step is a `blocks`-block sample of the 16-block batch
    steps = max(1, min(a.steps, 20))
    warmup = max(1, min(a.warmup, 2))
    pps, sec, threads, kind = cpu_msg_train_points_per_sec(blocks, steps, warmup)
    what = "unmodified reference from oracle/_ref (torch CPU fp32)" if kind == "reference" else \
        "oracle port (C prims + torch CPU fp32)"
    sample = f"{blocks} of {a.batch} blocks per step, {steps} timed + {warmup} warm-up steps, {what}"
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cpu_workload_config(a),
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
KERNEL_OF = {                                         # C-ABI entry point -> device kernel it launches in this workload
    "pcb_fps_f32": "fps_pair_kernel", "pcb_ball_query_f32": "ball_query_kernel",
    "pcb_ball_query_multi_f32": "ball_query_multi_kernel",
    "pcb_group_points_f32": "group_tile_kernel", "pcb_group_points_bwd_f32": "group_points_bwd_vec_kernel",
    "pcb_gather_f32": "gather_kernel", "pcb_gather_bwd_f32": "gather_bwd_kernel",
    "pcb_three_nn_f32": "three_nn_kernel", "pcb_interpolate_f32": "interp_rows_kernel",
    "pcb_interpolate_bwd_f32": "interp_bwd_kernel", "pcb_knn_f32": "knn_kernel",
    "pcb_knn_cdist_f32": "knn_xyz_kernel", "pcb_graph_feature_f32": "graph_feature_smem_kernel",
    "pcb_graph_feature_bwd_f32": "graph_feature_bwd_kernel", "pcb_group_points_bf16": "group_tile_kernel",
    "pcb_group_points_bwd_bf16": "group_points_bwd_vec_kernel", "pcb_bn_fwd_rows": "bn_fwd_fused_kernel",
    "pcb_bn_bwd_rows": "bn_bwd_fused_kernel", "pcb_fp_concat_bf16": "fp_concat_chunk_kernel",
    "pcb_fp_concat_bwd_bf16": "fp_concat_bwd_vec_kernel", "pcb_wgrad_rows_bf16": "wgrad_rows_kernel",
    "pcb_adam_flat_f32": "adam_flat_kernel", "pcb_sa_fused_bf16": "sa_fused_kernel",
    "pcb_linear_rows_bf16": "gemm_rows_kernel<EPI_STORE>", "pcb_linear_bn_stats_rows_bf16": "gemm_rows_kernel<EPI_STATS>",
    "pcb_dgrad_bn_rows_bf16": "gemm_rows_kernel<EPI_BNBWD>", "pcb_bn_apply_rows": "bn_apply_rows_kernel / bn_apply_pooled_kernel",
    "pcb_bn_bwd_apply_rows": "bn_bwd_apply_rows_kernel",
    "pcb_bn_pool_bwd_rows": "bn_pool_sums_kernel + bn_pool_bwd_apply_kernel"}


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pointcloud_bridge_b200 import _lib, distributed as pdist, ops, synthetic
    from pointcloud_bridge_b200.engine import Trainer
    from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    _lib.lib()                                        # fail loudly if libpcbridge.so is missing
    rank, world, local = pdist.init_from_env()
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    B = a.batch

    torch.manual_seed(1234)                           # same initial weights on every replica
    net = msg.get_model(NUM_CLASSES).to(dev).train()
    trainer = Trainer(net, lr=1e-3, weight_decay=1e-4, amp=not a.fp32, graph=not a.no_graph)

    # synthetic data: 4 distinct batches per rank, pinned on the host and resident in HBM
    host, resident = [], []
    for i in range(4):
        xyz, rgb, lab = synthetic.bridge_batch(1000 * rank + i, B, NPTS)
        x = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).pin_memory()
        y = torch.from_numpy(lab).pin_memory()
        host.append((x, y))
        resident.append((x.to(dev), y.to(dev)))
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)     # 256 MB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The step of batch i runs while the sampling / grouping indices of batch i + 1 are computed on a side stream
    # (Trainer.prefetch: farthest point sampling, ball query, three-NN depend on the coordinates only).  Every
    # iteration below issues exactly one step and one index chain, so K timed iterations are K complete steps.
    def step_resident(i, ev=None):
        flush.fill_(0.0)                                    # L2 flush BETWEEN timed steps: outside the event bracket
        if ev is not None:
            ev[0].record()
        loss = trainer.step_prefetched()
        if ev is not None:
            ev[1].record()
        x, y = resident[(i + 1) % len(resident)]
        trainer.prefetch(x, labels=y)
        return loss

    x, y = resident[0]
    trainer.prefetch(x, labels=y)
    for i in range(max(a.warmup, 3) + (4 if trainer.graph else 0)):      # graph mode: 3 eager + capture first
        step_resident(i)

    # A generation-2 pass of Python's garbage collector over the warm-up's objects (modules, graph captures, thousands of
    # tensors) pauses the host for tens of milliseconds; when it falls between a step's start event and its graph launch
    # the GPU idles inside the bracket.  Freeze what exists now (a training runner does the same after its first epoch).
    import gc
    gc.collect()
    if os.environ.get("PCB_BENCH_GC_FREEZE", "1") != "0":
        gc.freeze()

    # ---- value: K steps, inputs resident, device-timed, max over ranks; the K-step region is repeated a.repeats times
    # (each bracketed by barrier + synchronize on both sides) and the MEDIAN region is reported: one stalled replay in a
    # 20-step window otherwise moves the number by several per cent ----
    regions = []
    gpu_launches = 0
    with ClockSampler(local) as clocks:
        for rep in range(max(a.repeats, 1)):
            barrier()
            launches0 = _lib.launches()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
            for i in range(a.steps):
                step_resident(i, evs[i])
            barrier()
            per_step = [e0.elapsed_time(e1) for e0, e1 in evs]
            regions.append((pdist.max_over_ranks(sum(per_step), dev), per_step))     # K steps, flushes excluded
            gpu_launches = _lib.launches() - launches0
    if os.environ.get("PCB_BENCH_DUMP_STEPS") == "1":       # diagnostics: every step of every region, per rank
        for ri, (tot, steps_ms) in enumerate(regions):
            print(f"[rank {rank}] region {ri}: {tot / a.steps:.3f} ms/step  " + " ".join(f"{t:.2f}" for t in steps_ms),
                  file=sys.stderr, flush=True)
    regions.sort(key=lambda r: r[0])
    ms, per_step = regions[len(regions) // 2]
    per_step_sorted = sorted(per_step)
    value = world * B * NPTS * a.steps / (ms * 1e-3)

    # ---- e2e: batch from pinned host memory every step, loss read back every step ----
    def step_e2e(i):
        flush.fill_(0.0)
        hx, hy = host[i % len(host)]
        if trainer.graph:      # the step copies its inputs into the graph's static buffers: H2D from pinned memory
            return float(trainer.step(hx, labels=hy).item())
        x = hx.to(dev, non_blocking=True)
        y = hy.to(dev, non_blocking=True)
        return float(trainer.step(x, labels=y).item())              # D2H of the step's result

    trainer.step_prefetched()                           # drain the batch the value loop left in the prefetch stage
    for i in range(2):
        step_e2e(i)
    barrier()
    # Timed e2e loop: every step's batch comes from pinned host memory and every step's loss is read back.  The
    # copy of batch i+1 is issued (Trainer.prefetch, side stream) right after step i is launched, so it overlaps
    # the step; all K+1 transfers start inside the timed region.
    # The loss of step i is copied to pinned host memory right after the step and READ by the host after step
    # i+1 has been launched (one step of lag, as an asynchronous training log does), so the host never drains the
    # GPU between steps; every step's loss is read inside the timed region.
    loss_host = torch.empty(a.steps, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event() for _ in range(a.steps)]
    losses = []
    t0 = time.perf_counter()
    hx, hy = host[0]
    trainer.prefetch(hx, labels=hy)
    for i in range(a.steps):
        flush.fill_(0.0)
        loss = trainer.step_prefetched()
        loss_host[i:i + 1].copy_(loss.reshape(1), non_blocking=True)     # D2H of the step's result
        loss_ev[i].record()
        hx, hy = host[(i + 1) % len(host)]
        trainer.prefetch(hx, labels=hy)
        if i > 0:
            loss_ev[i - 1].synchronize()
            losses.append(float(loss_host[i - 1]))
    loss_ev[a.steps - 1].synchronize()
    losses.append(float(loss_host[a.steps - 1]))
    barrier()
    e2e_s = pdist.max_over_ranks(time.perf_counter() - t0, dev)
    assert len(losses) == a.steps and all(l == l for l in losses)
    e2e = {"value": world * B * NPTS * a.steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": world * int(host[0][0].numel() * 4 + host[0][1].numel() * 8),    # all ranks
           "d2h_bytes_per_step": world * 4, "ms_per_step": e2e_s / a.steps * 1e3}

    # ---- instrumented pass: per-launch CUDA events around every kernel of ours ----
    sink = []
    ops.set_kernel_timer(sink)
    ei0, ei1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ei0.record()
    n_inst = min(a.steps, 10)
    for i in range(n_inst):
        flush.fill_(0.0)
        x, y = resident[i % len(resident)]
        # Eager launches so that each kernel can be bracketed by events.  The host needs ~15 ms to
        # enqueue an eager step, the GPU ~8 ms to run it: a 40 ms spin kernel in front lets the host
        # run ahead, so events and kernels execute back to back and an event pair measures the kernel,
        # not the launch gap.
        torch.cuda._sleep(int(8e7))
        trainer._step_eager((x,), y, ())
        torch.cuda.synchronize()
    ei1.record()
    torch.cuda.synchronize()
    ops.set_kernel_timer(None)
    inst_ms = ei0.elapsed_time(ei1)   # (eager, instrumented: only used as a sanity figure)
    # an event pair around NOTHING still reads a few microseconds: measure that and take it off every bracket, so that
    # the per-kernel shares of the step are not inflated by the instrumentation
    cal = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    torch.cuda._sleep(int(2e7))
    for c0, c1 in cal:
        c0.record()
        c1.record()
    torch.cuda.synchronize()
    gaps = sorted(c0.elapsed_time(c1) for c0, c1 in cal)
    evt_overhead_ms = gaps[len(gaps) // 2]
    # aggregate per C-ABI entry point (one kernel family), all shapes of the step together
    agg = {}
    for name, nbytes, s, e in sink:
        t = agg.setdefault(name, [0.0, 0, 0])
        t[0] += max(s.elapsed_time(e) - evt_overhead_ms, 1e-4)
        t[1] += 1
        t[2] += nbytes
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic_path = os.path.join(ROOT, "profiles", "traffic_per_launch.json")       # from ncu --set full captures
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    step_ms = ms / a.steps
    kernels = []
    for name, (tot, cnt, nbytes) in agg.items():
        kernels.append({"entry": name, "kernel": KERNEL_OF.get(name, name), "launches_per_step": cnt / n_inst,
                        "alg_bytes_per_launch": int(nbytes / cnt), "avg_launch_ms": round(tot / cnt, 5),
                        "ms_per_step": round(tot / n_inst, 4), "share_of_step": round((tot / n_inst) / step_ms, 4),
                        "achieved_GBps": round(nbytes / tot / 1e6, 1), "frac": round(nbytes / tot / 1e6 / hbm_peak, 4)})
    kernels.sort(key=lambda k: -k["share_of_step"])
    ours_share = sum(k["share_of_step"] for k in kernels)
    top = kernels[0]
    roofline = {"kernel": top["kernel"], "entry": top["entry"], "bound": "hbm", "achieved": top["achieved_GBps"],
                "peak": hbm_peak, "unit": "GB/s", "frac": top["frac"],
                "traffic": traffic.get(top["kernel"]), "peak_source": peak_src,
                "avg_launch_ms": top["avg_launch_ms"], "launches_per_step": top["launches_per_step"],
                "alg_bytes_per_launch": top["alg_bytes_per_launch"], "share_of_step": top["share_of_step"],
                "note": "kernel family of ours with the largest share of the step (all layer shapes together); per-launch "
                        "durations from CUDA events around every launch of an eager, single-stream, spin-ahead "
                        "instrumented pass, minus the measured duration of an empty event bracket.  Shares are kernel time "
                        "over step time and add up to more than 1: inside the captured step the index chain, the weight "
                        "gradients and the second scale of every multi-scale level run concurrently with the main chain, "
                        "and programmatic dependent launch overlaps every kernel's prologue with its predecessor "
                        "(profiles/r2_final_step_launches.md has the ncu launch list of the graph replay itself)",
                "event_bracket_overhead_ms": round(evt_overhead_ms, 5),
                "all_our_kernels_share_of_step": round(ours_share, 4), "kernels": kernels[:10]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "timing": {"regions_ms_per_step": [round(r[0] / a.steps, 4) for r in regions], "reported": "median region",
                       "per_step_ms_median": round(per_step_sorted[len(per_step_sorted) // 2], 4),
                       "per_step_ms_p90": round(per_step_sorted[min(len(per_step_sorted) - 1, int(0.9 * len(per_step_sorted)))], 4)},
            "dtype": "f32" if a.fp32 else "bf16", "data": "synthetic", "config": workload_config(a, world),
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline}

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        pps, sec, threads, kind = cpu_msg_train_points_per_sec(a.cpu_sample_blocks, 3, 1)
        what = "unmodified reference from oracle/_ref" if kind == "reference" else "oracle port (C prims + torch CPU ops)"
        line["cpu_baseline"] = {"value": pps, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": f"{a.cpu_sample_blocks} of {B} blocks per step, 3 timed + 1 warm-up steps, "
                                          f"{what}, torch CPU fp32, scaled per point"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        trainer.close()                                   # the step graph holds a captured all-reduce: drop it first
        del trainer
        dist.destroy_process_group()


def run_other_config(a):
    """The secondary BASELINE configurations through their own tools (same launch line, one JSON line per measurement on
    stdout; torchrun environment passed through)."""
    import runpy
    tool = "bench_scene_e2e.py" if a.config == "c5scene" else "bench_configs.py"
    sys.argv = [os.path.join(ROOT, "tools", tool)] + ([] if a.config == "c5scene" else ["--only", a.config])
    runpy.run_path(sys.argv[0], run_name="__main__")


def main():
    a = parse()
    if a.config != "c2" and a.impl == "ours":
        return run_other_config(a)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

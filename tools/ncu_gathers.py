"""One launch of every HBM-bound gather-family kernel at a large B=16 shape between cudaProfilerStart/Stop, for
    ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/gathers python tools/ncu_gathers.py
(each kernel is warmed up once outside the profiled region)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_bridge_b200 import ops, synthetic  # noqa: E402

dev = "cuda:0"
B, N = 16, 4096
torch.manual_seed(0)
xyz = torch.from_numpy(synthetic.bridge_batch(0, B, N)[0]).to(dev)
start = torch.zeros(B, dtype=torch.long, device=dev)
l1 = ops.gather(xyz, ops.furthest_point_sample(xyz, 1024, start))
idx32 = torch.randint(0, N, (B, N, 32), device=dev)
f64 = torch.randn(B, N, 64, device=dev)
f64b = f64.to(torch.bfloat16)
f256 = torch.randn(B, 1024, 256, device=dev)
bri_idx = torch.randint(0, 1024, (B, 512, 32), device=dev)
x64 = torch.randn(B, 64, N, device=dev)
kidx = torch.randint(0, N, (B, N, 20), device=dev)
_, i3, w3 = ops.three_nn(xyz, l1, 3)
p2 = torch.randn(B, 1024, 128, device=dev)
p2b = p2.to(torch.bfloat16).requires_grad_(True)
p1b = torch.randn(B, N, 64, device=dev).to(torch.bfloat16)
gy = torch.randn(524288, 64, device=dev).to(torch.bfloat16)
xx = torch.randn(524288, 32, device=dev).to(torch.bfloat16)
gw = torch.zeros(64, 32, device=dev)
fr = f64b.clone().requires_grad_(True)


def grp_bf16(feats):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return ops.group_points(xyz, feats, xyz, idx32, xyz_first=False, pad_to=8)


def fpc():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return ops.fp_concat(p1b, p2b, i3, w3, pad_to=8)


def run():
    ops.gather(f256, bri_idx)                                   # gather_vec_ilp_kernel, C=256
    ops.group_points(xyz, f64, xyz, idx32, True)                # group_points_kernel fp32 [dxyz|feat], C=67
    out = grp_bf16(fr)                                          # group_points_chunk_kernel bf16 rows, pitch 72
    torch.autograd.grad(out, fr, torch.ones_like(out))          # group_points_bwd_vec_kernel
    ops.graph_feature(x64, kidx)                                # graph_feature_smem_kernel
    ops.three_interpolate(p2, i3, w3, False)                    # interp_rows_kernel
    o2 = fpc()                                                  # fp_concat_chunk_kernel
    torch.autograd.grad(o2, p2b, torch.ones_like(o2))           # fp_concat_bwd_vec_kernel
    ops.wgrad_rows(gy, xx, 32, out=gw)                          # wgrad_rows_kernel


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")

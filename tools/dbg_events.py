import sys, torch
sys.path.insert(0, '/root/repo')
from pointcloud_bridge_b200 import _lib
lib = _lib.lib()
dev = 'cuda:0'
C = 64
sums = torch.zeros(3*C, device=dev); y = torch.randn(1024, C, device=dev).bfloat16(); st = torch.zeros(2, C, device=dev)
def k():
    lib.pcb_bn_finalize(sums.data_ptr(), y.data_ptr(), 1, None, 1024, C, 1e-5, 0.1, None, None, st[0].data_ptr(), st[1].data_ptr(), torch.cuda.current_stream().cuda_stream)
def run(spin, n=300, inner=1):
    ev = []
    if spin: torch.cuda._sleep(int(4e7))
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): k()
        e1.record(); ev.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts)//2]*1e3, ts[0]*1e3, ts[-1]*1e3
for _ in range(3): k()
torch.cuda.synchronize()
print("eager   median/min/max us", run(False))
print("spin    median/min/max us", run(True))
print("spin x8 median/min/max us (8 kernels per bracket)", run(True, inner=8))
x = torch.zeros(16, device=dev)
def run_t(spin, n=300):
    ev=[]
    if spin: torch.cuda._sleep(int(4e7))
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); x.add_(1); e1.record(); ev.append((e0,e1))
    torch.cuda.synchronize(); ts = sorted(a.elapsed_time(b) for a,b in ev); return ts[len(ts)//2]*1e3, ts[0]*1e3
print("torch add_ spin", run_t(True))

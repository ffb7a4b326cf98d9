"""Builds a variant of libpcbridge.so with extra -D flags for ONE source file (kernel tuning
experiments): python tools/build_variant.py NAME FILE.cu -DFOO=1 ...  -> pointcloud_bridge_b200/variants/NAME.so
Select it at run time with PCB_LIB_PATH."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import build as B  # noqa: E402

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
vdir = os.path.join(B.HERE, "variants")
os.makedirs(vdir, exist_ok=True)
obj = os.path.join(vdir, name + ".o")
out = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", obj],
                     capture_output=True, text=True)
if out.returncode:
    sys.exit(out.stdout + out.stderr)
for l in (out.stdout + out.stderr).splitlines():
    if "registers" in l or "spill" in l and "0 bytes spill" not in l:
        print(l.strip()[:160])
objs = [os.path.join(B.OBJ, f) for f in os.listdir(B.OBJ) if f.endswith(".o") and f != src[:-3] + ".o"]
so = os.path.join(vdir, name + ".so")
subprocess.check_call([B._nvcc(), "-shared", "-o", so, obj, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
print(so)

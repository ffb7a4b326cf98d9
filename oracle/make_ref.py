"""Recipe for oracle/_ref/: the UNMODIFIED reference files of the benchmarked path, copied byte for byte from
/root/reference so that the CPU arm of bench.py (`--impl reference`, `cpu_baseline`) can time the reference itself on
the GPU box, where /root/reference does not exist.  oracle/_ref/ is git-ignored (never part of the history) but travels
with the gpurun snapshot, like the built .so files.  TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under
pointcloud_bridge_b200/ reads it.

    python oracle/make_ref.py            # run by __graft_entry__.build() when /root/reference is present

Files (Partsize-identical/models/): pointnet_util.py, pointnet2_sem_seg_msg.py, pointnet2_sem_seg.py.  The only file
written that is not a copy is an empty package marker, so that the MSG file's relative import resolves.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PCB_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "Partsize-identical", "models")
DST = os.path.join(HERE, "_ref", "ps_models")
FILES = ["pointnet_util.py", "pointnet2_sem_seg_msg.py", "pointnet2_sem_seg.py"]


def build() -> bool:
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    sums = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        sums[f] = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()
    open(os.path.join(DST, "__init__.py"), "w").close()
    json.dump({"source": SRC, "sha256": sums}, open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w"), indent=1)
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def load_msg():
    """The reference's pointnet2_sem_seg_msg module, imported from oracle/_ref (flat `pointnet_util` first: the
    SSG file imports it without a package)."""
    import importlib
    ref_root = os.path.join(HERE, "_ref")
    for p in (ref_root, DST):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("ps_models.pointnet2_sem_seg_msg")


if __name__ == "__main__":
    print("oracle/_ref:", "built" if build() else f"{SRC} not present; nothing copied")

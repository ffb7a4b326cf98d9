"""oracle/make_ref.py copies the reference's own files of the benchmarked path into oracle/_ref/ (git-ignored) so that
bench.py's CPU arm times the reference itself on the GPU box.  Here (authoring container): the copies are byte-identical
to /root/reference, import, and build the network with the reference's parameter names."""
import hashlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_ref  # noqa: E402


@pytest.mark.skipif(not os.path.isdir(make_ref.SRC), reason="reference tree not present")
def test_ref_copies_are_byte_identical_and_import():
    assert make_ref.build() and make_ref.available()
    for f in make_ref.FILES:
        a = hashlib.sha256(open(os.path.join(make_ref.SRC, f), "rb").read()).hexdigest()
        b = hashlib.sha256(open(os.path.join(make_ref.DST, f), "rb").read()).hexdigest()
        assert a == b, f
    net = make_ref.load_msg().get_model(5)
    from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as ours
    assert list(net.state_dict().keys()) == list(ours.get_model(5).state_dict().keys())


def test_ref_dir_is_git_ignored_and_not_product_input():
    assert "oracle/_ref/" in open(os.path.join(ROOT, ".gitignore")).read()
    pkg = os.path.join(ROOT, "pointcloud_bridge_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert "make_ref" not in open(os.path.join(dirpath, f)).read(), f

"""patch_reference(): rebinding logic on stand-in modules (CPU; the reference tree is not needed)."""
import sys
import types

import pointcloud_bridge_b200 as pcb
from pointcloud_bridge_b200.partsize import pointnet_util as ours


def test_rebinds_functions_forward_and_importers():
    ref = types.ModuleType("pointnet_util")

    def farthest_point_sample(xyz, npoint):
        raise AssertionError("reference implementation should have been replaced")
    farthest_point_sample.__module__ = "pointnet_util"
    ref.farthest_point_sample = farthest_point_sample
    ref.query_ball_point = lambda *a: None

    class PointNetSetAbstraction:                      # stand-in for the reference class
        def forward(self, xyz, points):
            raise AssertionError
    ref.PointNetSetAbstraction = PointNetSetAbstraction
    consumer = types.ModuleType("pointnet2_sem_seg")   # did `from pointnet_util import farthest_point_sample`
    consumer.farthest_point_sample = farthest_point_sample
    mods = {"pointnet_util": ref, "pointnet2_sem_seg": consumer}
    done = pcb.patch_reference(mods)
    assert ref.farthest_point_sample is ours.farthest_point_sample
    assert ref.query_ball_point is ours.query_ball_point
    assert consumer.farthest_point_sample is ours.farthest_point_sample
    assert PointNetSetAbstraction.forward is ours.PointNetSetAbstraction.__dict__["forward"]
    assert "pointnet_util.farthest_point_sample" in done


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    from pointcloud_bridge_b200 import _lib, ops
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(_lib.PcbError):
        ops.furthest_point_sample(torch.rand(1, 16, 3), 4)
    with pytest.raises(_lib.PcbError):
        ops.gather(torch.rand(1, 16, 3), torch.zeros(1, 2, dtype=torch.long))

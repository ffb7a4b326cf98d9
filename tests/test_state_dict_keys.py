"""Checkpoint compatibility (CPU): every network exposes exactly the reference's state_dict
keys, order and shapes (tests/golden/state_dict_keys.json, from the reference constructors), so
`load_state_dict(checkpoint['model_state_dict'])` of a reference checkpoint works
(Highway_bridge/inference.py:108-110)."""
import json
import os

import pytest

import parity

KEYS = json.load(open(os.path.join(parity.GOLDEN, "state_dict_keys.json")))


@pytest.mark.parametrize("spec", sorted(KEYS))
def test_state_dict_matches_reference(spec):
    import pointcloud_bridge_b200.highway.DGCNN  # noqa: F401
    import pointcloud_bridge_b200.highway.model  # noqa: F401
    import pointcloud_bridge_b200.partsize.pointnet2_sem_seg  # noqa: F401
    import pointcloud_bridge_b200.partsize.pointnet2_sem_seg_msg  # noqa: F401
    import pointcloud_bridge_b200 as pkg
    net = eval("pkg." + spec, {"pkg": pkg})
    mine = [[n, list(t.shape)] for n, t in net.state_dict().items()]
    assert mine == KEYS[spec]

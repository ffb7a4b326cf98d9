"""pointcloud_bridge_b200 -- B200-native (sm_100a) sampling-and-grouping hot path of the
bridge point-cloud segmentation networks of UT-Team-Chun/Pointcloud-bridge.

    ops                  tensor-level front end of the C-ABI CUDA library (libpcbridge.so)
    partsize.*           drop-ins for Partsize-identical/models (pointnet_util, PN++ SSG / MSG nets)
    highway.*            drop-ins for Highway_bridge/models (pointnet2_utils, DGCNN, BriStruNet)
    patch_reference()    rebind the hot-path names inside the imported reference modules
    engine, distributed  training step / block-sharded inference drivers, one process per GPU
    synthetic            seeded bridge-like blocks for benchmarks and tests

See DESIGN.md.  The CUDA library is loaded lazily on the first op; without it (or without a
CUDA device) every op raises -- there is no CPU fallback.
"""
from .patch import patch_reference  # noqa: F401

__version__ = "0.1.0"

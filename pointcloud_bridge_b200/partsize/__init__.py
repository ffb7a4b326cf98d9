"""B200 counterparts of Partsize-identical/models (PointNet++ SSG / MSG semantic segmentation)."""
from . import pointnet_util, pointnet2_sem_seg, pointnet2_sem_seg_msg  # noqa: F401

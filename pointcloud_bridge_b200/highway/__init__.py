"""B200 counterparts of Highway_bridge/models (PointNet2, BriStruNet = EnhancedPointNet2, DGCNN)."""
from . import pointnet2_utils, attention_modules, DGCNN, model  # noqa: F401

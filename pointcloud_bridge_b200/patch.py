"""`patch_reference()` -- run the reference's OWN networks on the B200 kernels.

The reference's hot path is reached through module globals (`from pointnet_util import ...`,
`from .pointnet2_utils import ...`), so rebinding those names inside the already-imported
reference modules swaps the implementation without touching the reference tree:

    sys.path.insert(0, ".../Partsize-identical/models"); import pointnet_util, pointnet2_sem_seg
    import pointcloud_bridge_b200 as pcb
    pcb.patch_reference()                      # finds the reference modules in sys.modules
    net = pointnet2_sem_seg.get_model(13).cuda()   # reference class, B200 kernels underneath

Functions are replaced one for one; the SA / FP module classes get the drop-in `forward`
(same parameters, so reference checkpoints keep loading).
"""
from __future__ import annotations

import sys
import types

__all__ = ["patch_reference", "PATCH_TABLE"]

_FUNCS = ("square_distance", "index_points", "farthest_point_sample", "query_ball_point", "sample_and_group",
          "sample_and_group_all")

# reference module name (as imported) -> (our module, function names, {reference class: our class})
PATCH_TABLE = {
    "pointnet_util": ("pointcloud_bridge_b200.partsize.pointnet_util", _FUNCS,
                      ("PointNetSetAbstraction", "PointNetSetAbstractionMsg", "PointNetFeaturePropagation")),
    "models.pointnet2_utils": ("pointcloud_bridge_b200.highway.pointnet2_utils", _FUNCS[:5],
                               ("SetAbstraction", "FeaturePropagation", "EnhancedFeaturePropagation",
                                "MultiScaleSetAbstraction")),
    "models.attention_modules": ("pointcloud_bridge_b200.highway.attention_modules",
                                 ("square_distance", "index_points"),
                                 ("BridgeStructureEncoding", "ColorFeatureExtraction", "GeometricFeatureExtraction",
                                  "CompositeFeatureFusion")),
    "models.DGCNN": ("pointcloud_bridge_b200.highway.DGCNN", (), ("DGCNN",)),
}


def _patch_module(ref_mod: types.ModuleType, ours: types.ModuleType, funcs, classes) -> list[str]:
    done = []
    for name in funcs:
        if hasattr(ref_mod, name) and hasattr(ours, name):
            setattr(ref_mod, name, getattr(ours, name))
            done.append(f"{ref_mod.__name__}.{name}")
    for cname in classes:
        ref_cls, our_cls = getattr(ref_mod, cname, None), getattr(ours, cname, None)
        if ref_cls is None or our_cls is None:
            continue
        for meth in ("forward", "knn", "get_graph_feature", "compute_absolute_position_encoding",
                     "get_structure_features"):
            if meth in our_cls.__dict__:
                setattr(ref_cls, meth, our_cls.__dict__[meth])
                done.append(f"{ref_mod.__name__}.{cname}.{meth}")
    return done


def patch_reference(modules: dict | None = None) -> list[str]:
    """Rebind the hot-path names inside the reference modules found in `modules` (default:
    sys.modules, matched by module name suffix).  Returns the list of patched attributes."""
    import importlib
    modules = sys.modules if modules is None else modules
    patched = []
    for ref_name, (our_name, funcs, classes) in PATCH_TABLE.items():
        ours = importlib.import_module(our_name)
        for name, mod in list(modules.items()):
            if mod is None or not isinstance(mod, types.ModuleType):
                continue
            if name.startswith("pointcloud_bridge_b200"):
                continue
            if name == ref_name or name.endswith("." + ref_name.split(".")[-1]) and ref_name.split(".")[-1] in name:
                patched += _patch_module(mod, ours, funcs, classes)
        # consumers that did `from pointnet_util import X` hold their own references: rebind those too
        for name, mod in list(modules.items()):
            if mod is None or not isinstance(mod, types.ModuleType) or name.startswith("pointcloud_bridge_b200"):
                continue
            for fname in funcs:
                obj = mod.__dict__.get(fname)
                if callable(obj) and getattr(obj, "__module__", "").split(".")[-1] == ref_name.split(".")[-1] \
                        and hasattr(ours, fname):
                    setattr(mod, fname, getattr(ours, fname))
                    patched.append(f"{name}.{fname}")
    return patched

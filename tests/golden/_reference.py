"""Imports the UNMODIFIED reference from /root/reference (authoring container only).

Used by make_golden*.py to produce the committed fixtures.  Nothing under tests/ that runs
on the GPU box imports this module: /root/reference does not exist there.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("PCB_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "Partsize-identical", "models"))


def partsize():
    """Partsize-identical/models: pointnet_util (flat import), pointnet2_sem_seg (flat import of
    pointnet_util) and pointnet2_sem_seg_msg (relative import -> synthetic package)."""
    d = os.path.join(REF, "Partsize-identical", "models")
    if d not in sys.path:
        sys.path.insert(0, d)
    import pointnet_util  # noqa
    import pointnet2_sem_seg  # noqa
    if "ps_models" not in sys.modules:
        pkg = types.ModuleType("ps_models")
        pkg.__path__ = [d]
        sys.modules["ps_models"] = pkg
        for name in ("pointnet_util", "pointnet2_sem_seg_msg"):
            spec = importlib.util.spec_from_file_location(f"ps_models.{name}", os.path.join(d, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"ps_models.{name}"] = mod
            spec.loader.exec_module(mod)
    return pointnet_util, pointnet2_sem_seg, sys.modules["ps_models.pointnet2_sem_seg_msg"]


def highway():
    """Highway_bridge/models: pointnet2_utils, DGCNN, attention_modules, model."""
    d = os.path.join(REF, "Highway_bridge")
    if d not in sys.path:
        sys.path.insert(0, d)
    import models.pointnet2_utils as p2u
    import models.DGCNN as dgcnn
    import models.attention_modules as am
    import models.model as model
    return p2u, dgcnn, am, model

"""Golden vectors for the whole-scene tiler: outputs of the UNMODIFIED reference class
ScannetDatasetWholeScene (Highway_bridge/utils/BridgeDataLoader.py:172-277) on a small synthetic scene, and of
add_vote (Partsize-identical/test_sem_seg.py:58-65).  Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden_scene.py

The module imports `laspy` at the top (absent here, only used to read .las files): a stub module stands in
for it, and the dataset object is built without its file-reading __init__.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def scene(seed, n):
    """A deck slab + two piers, ~6.3 m x 3.2 m, fp32-valued coordinates (what a .las delivers after scaling)."""
    rng = np.random.default_rng(seed)
    deck = np.stack([rng.uniform(0, 6.3, n // 2), rng.uniform(0, 3.2, n // 2), 2.0 + 0.01 * rng.standard_normal(n // 2)], 1)
    pier = np.stack([rng.uniform(1.0, 1.4, n // 4), rng.uniform(0.5, 2.7, n // 4), rng.uniform(0, 2.0, n // 4)], 1)
    pier2 = pier + np.array([3.7, 0.0, 0.0])
    xyz = np.concatenate([deck, pier, pier2]).astype(np.float32)
    rgb = rng.uniform(0, 255, (xyz.shape[0], 3)).astype(np.float32)
    lab = np.concatenate([np.zeros(n // 2), np.ones(n // 4), 2 * np.ones(n // 4)]).astype(np.int64)
    return np.concatenate([xyz, rgb], 1).astype(np.float64), lab


def main():
    sys.modules.setdefault("laspy", types.ModuleType("laspy"))
    bdl = load(os.path.join(REF, "Highway_bridge/utils/BridgeDataLoader.py"), "ref_bridge_data_loader")
    pts, lab = scene(0, 20000)
    ds = object.__new__(bdl.ScannetDatasetWholeScene)
    ds.block_points, ds.block_size, ds.padding, ds.stride = 1024, 1.0, 0.001, 0.5
    ds.scene_points_list, ds.semantic_labels_list = [pts], [lab.astype(np.float64)]
    ds.labelweights = np.ones(3, np.float32)
    np.random.seed(0)
    data, label, weight, index = ds[0]
    # add_vote of the reference on random predictions for these blocks
    sys.argv = ["x"]
    src = open(os.path.join(REF, "Partsize-identical/test_sem_seg.py")).read()
    ns = {}
    start = src.index("def add_vote")
    exec(src[start:src.index("def main(args)")], ns)
    rng = np.random.default_rng(1)
    pred = rng.integers(0, 3, index.shape)
    pool = ns["add_vote"](np.zeros((pts.shape[0], 3)), index, pred, weight)
    np.savez_compressed(os.path.join(HERE, "scene.npz"), points=pts.astype(np.float32), block_points=1024,
                        data=data.astype(np.float32), index=index.astype(np.int64), pred=pred.astype(np.uint8),
                        pool=pool.astype(np.int32), labels=np.argmax(pool, 1).astype(np.uint8))
    print("blocks", data.shape, "windows covered", len(np.unique(index)))


if __name__ == "__main__":
    main()

"""CPU-side checks of the C-ABI boundary: the library builds with nvcc for sm_100a, loads, and
exports every symbol include/pcbridge.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcbridge.h")


@pytest.fixture(scope="module")
def so_path():
    from pointcloud_bridge_b200 import build
    return build.build()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ("pcb_fps_f32", "pcb_ball_query_f32", "pcb_knn_f32", "pcb_knn_cdist_f32", "pcb_gather_f32",
                 "pcb_group_points_f32", "pcb_three_nn_f32", "pcb_interpolate_f32", "pcb_graph_feature_f32",
                 "pcb_square_distance_f32"):
        assert must in syms


def test_library_exports_every_declared_symbol(so_path):
    lib = ctypes.CDLL(so_path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in pcbridge.h but not exported"
    lib.pcb_version.restype = ctypes.c_int
    assert lib.pcb_version() == 100
    lib.pcb_error_string.restype = ctypes.c_char_p
    assert b"envelope" in lib.pcb_error_string(-2)


def test_python_binding_covers_every_declared_symbol(so_path):
    from pointcloud_bridge_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    _lib.lib()


def test_no_undeclared_exports(so_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", so_path], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "pcb_" in l.split()[-1]
                      and not l.split()[-1].startswith("_Z"))
    assert exported == declared_symbols()


def test_sass_is_sm100a_and_uses_bulk_copy(so_path):
    """The library carries sm_100a SASS only, and the staging path is the TMA bulk copy
    (UBLKCP in SASS, B200_PROFILING.md)."""
    out = subprocess.check_output(["cuobjdump", "-lelf", so_path], text=True)
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    sass = subprocess.check_output(["cuobjdump", "-sass", so_path], text=True)
    assert "UBLKCP" in sass          # cp.async.bulk (ball query staging)
    assert "CREDUX" in sass or "REDUX" in sass   # warp argmax of the FPS kernel


def test_arguments_outside_envelope_are_rejected_without_a_gpu(so_path):
    """Argument validation happens before any CUDA call, so it can be exercised here."""
    from pointcloud_bridge_b200 import _lib
    l = _lib.lib()
    assert l.pcb_fps_f32(None, 1, 16, None, 4, None, None) == -1
    assert l.pcb_fps_f32(1, 1, 100000, 1, 4, 1, None) == -2          # N above 49152
    assert l.pcb_three_nn_f32(1, 1, 1, 8, 8, 9, 1, 1, None, None) == -2   # k above 8
    assert l.pcb_knn_f32(1, 1, 8, 3, 100, 1, 1, None, None) == -2         # k above 64


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "pointcloud_bridge_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "pcb_oracle" not in text, f

"""INTEGRATION.md section 1 executed for real: the UNMODIFIED reference modules (oracle/_ref, copied by the committed
recipe oracle/make_ref.py; they travel to the GPU box with the snapshot) are imported, `patch_reference()` rebinds their
hot-path names, and the reference's own `get_model` then runs on the B200 kernels -- bit for bit the drop-in network's
encoder output, and within fp32 tolerance of the reference's CPU golden logits.  Runs in a subprocess: patching mutates the
reference classes, which other tests (and bench.py's CPU arm) import unpatched."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import contextlib, io, os, sys
    import numpy as np, torch
    torch.backends.cuda.matmul.allow_tf32 = False        # the reference's own Conv1d head stays on the library: fp32, not TF32
    torch.backends.cudnn.allow_tf32 = False
    ROOT = %r
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "ps_models"))      # as the reference's own scripts do
    import pointnet_util, pointnet2_sem_seg                                     # unmodified reference modules
    import parity
    import pointcloud_bridge_b200 as pcb
    from pointcloud_bridge_b200 import synthetic
    from pointcloud_bridge_b200.partsize import pointnet2_sem_seg as ours_ssg
    assert "oracle/_ref" in pointnet_util.__file__.replace(os.sep, "/")
    ref_fps = pointnet_util.farthest_point_sample
    done = pcb.patch_reference()
    assert pointnet_util.farthest_point_sample is not ref_fps and len(done) >= 9, done
    g = parity.load("models.npz")
    x9 = torch.from_numpy(synthetic.sem_seg_input(g["xyz"], g["rgb"]))[:1].cuda()
    ref_net = parity.seeded_fill_(pointnet2_sem_seg.get_model(13), 1).cuda().eval()   # the REFERENCE class
    our_net = parity.seeded_fill_(ours_ssg.get_model(13), 1).cuda().eval()
    assert list(ref_net.state_dict()) == list(our_net.state_dict())
    outs = []
    for net in (ref_net, our_net):
        torch.manual_seed(4242)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            y, l4 = net(x9)
        outs.append((y.float().cpu().numpy(), l4.float().cpu().numpy()))
    # the encoder (every patched module) is the same code on both sides: bit for bit; the segmentation head stays the
    # reference's own Conv1d / BatchNorm1d calls in `ref_net` (a different library GEMM than the drop-in's row head)
    assert np.array_equal(outs[0][1], outs[1][1]), np.abs(outs[0][1] - outs[1][1]).max()
    dh = np.abs(outs[0][0] - outs[1][0]).max() / np.abs(outs[1][0]).max()
    print("reference head (Conv1d) vs row head on identical encoder outputs: max rel diff %%.2e" %% dh)
    assert dh <= 2e-5
    ref = g["ssg_logp"]
    err = np.abs(outs[0][0] - ref).max() / np.abs(ref).max()
    print("patched reference vs its own CPU golden logits: max rel err %%.2e" %% err)
    assert err < 2e-5
    print("PATCH_OK", len(done))
""") % ROOT


def test_patch_reference_runs_the_real_reference_network_on_the_b200_kernels():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ps_models", "pointnet_util.py")):
        pytest.skip("oracle/_ref absent (run `python oracle/make_ref.py` where /root/reference exists)")
    out = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=600)
    print(out.stdout[-2000:], out.stderr[-3000:])
    assert out.returncode == 0 and "PATCH_OK" in out.stdout

"""Generates tests/golden/msg_train_b16.npz: one TRAINING step of the UNMODIFIED reference PointNet++ MSG network
(imported from /root/reference) at the size BASELINE configs[1] names -- batch 16 x 4096-point blocks -- for two
seeds, CPU fp32: loss, a strided sample of the log-probabilities, level-1 FPS indices, a set of weight / BatchNorm
gradients and the gradient norm of every parameter.  Inputs are synthetic.bridge_batch(100 + seed, 16): not stored.

    python tests/golden/make_golden_msg_b16.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import _reference  # noqa: E402
import parity  # noqa: E402
from pointcloud_bridge_b200 import synthetic  # noqa: E402

SEED_FPS = 4242
GRADS = ["sa1.conv_blocks.0.0.weight", "sa1.conv_blocks.1.2.weight", "sa2.conv_blocks.1.1.weight",
         "sa3.conv_blocks.0.1.weight", "sa4.bn_blocks.1.2.weight", "fp4.mlp_bns.0.weight", "fp1.mlp_convs.0.weight",
         "conv1.weight", "conv2.weight", "conv2.bias", "sa1.bn_blocks.0.0.weight", "sa2.bn_blocks.1.2.bias", "bn1.weight"]


def main():
    assert _reference.available()
    torch.set_num_threads(8)
    pu, _, msg = _reference.partsize()
    out = {}
    for seed in (0, 1):
        xyz, rgb, lab = synthetic.bridge_batch(100 + seed, 16)
        x9 = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb))
        tlab = torch.from_numpy(lab)
        net = parity.seeded_fill_(msg.get_model(5), 2)
        net.train()
        net.drop1.eval()                       # dropout mask comes from a device-specific RNG stream
        torch.manual_seed(SEED_FPS + seed)
        fps1 = pu.farthest_point_sample(x9[:, :3, :].permute(0, 2, 1), 1024)      # the first draw of the forward pass
        torch.manual_seed(SEED_FPS + seed)
        with contextlib.redirect_stdout(io.StringIO()):
            y, _ = net(x9)
        loss = torch.nn.functional.nll_loss(y.reshape(-1, 5), tlab.reshape(-1))
        loss.backward()
        p = f"s{seed}_"
        out[p + "loss"] = np.float32(loss.item())
        out[p + "logp_sample"] = y.detach().numpy()[:, ::16, :].copy()
        out[p + "fps1"] = fps1.numpy().astype(np.int16)
        params = dict(net.named_parameters())
        for name in GRADS:
            out[p + "g_" + name] = params[name].grad.numpy()
        out[p + "gnorm_names"] = np.array(list(params.keys()))
        out[p + "gnorm"] = np.array([float(v.grad.norm()) for v in params.values()], dtype=np.float32)
        out[p + "rm_sa1"] = net.sa1.bn_blocks[0][0].running_mean.numpy()
        out[p + "rv_sa1"] = net.sa1.bn_blocks[0][0].running_var.numpy()
        print("seed", seed, "loss", float(loss))
    path = os.path.join(HERE, "msg_train_b16.npz")
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()

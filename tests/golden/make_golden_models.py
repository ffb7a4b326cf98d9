"""Generates tests/golden/models.npz: outputs of the UNMODIFIED reference networks (imported
from /root/reference) on seeded synthetic blocks, CPU fp32, with name-keyed seeded weights
(tests/parity.py:seeded_fill_) so that the B200 networks can be loaded with identical values.

    python tests/golden/make_golden_models.py
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


def main():
    assert _reference.available()
    torch.set_num_threads(8)
    pu, ssg, msg = _reference.partsize()
    p2u, dgcnn_mod, am, model_mod = _reference.highway()
    out = {}
    xyz, rgb, lab = synthetic.bridge_batch(7, 2)
    out["xyz"], out["rgb"], out["labels"] = xyz, rgb, lab.astype(np.int8)
    x9 = torch.from_numpy(synthetic.sem_seg_input(xyz, rgb))
    txyz, trgb, tlab = torch.from_numpy(xyz), torch.from_numpy(rgb), torch.from_numpy(lab)

    def run(name, net, *inputs):
        net.eval()
        torch.manual_seed(SEED_FPS)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            y = net(*inputs)
        return y

    # config 1: PN++ SSG sem-seg forward, B=1 (pointnet2_sem_seg.py)
    net = parity.seeded_fill_(ssg.get_model(13), 1)
    y, l4 = run("ssg", net, x9[:1])
    out["ssg_logp"], out["ssg_l4"] = y.numpy(), l4.numpy()

    # config 2: PN++ MSG sem-seg (pointnet2_sem_seg_msg.py), eval forward B=2 ...
    net = parity.seeded_fill_(msg.get_model(5), 2)
    y, l4 = run("msg", net, x9)
    out["msg_logp"], out["msg_l4"] = y.numpy(), l4.numpy()
    # ... and one training-mode forward/backward (dropout switched off: its mask comes from a
    # device-specific RNG stream), NLL loss: loss value, BN running stats and two gradients
    net = parity.seeded_fill_(msg.get_model(5), 2)
    net.train()
    net.drop1.eval()
    torch.manual_seed(SEED_FPS)
    y, _ = net(x9)
    loss = torch.nn.functional.nll_loss(y.reshape(-1, 5), tlab.reshape(-1))
    loss.backward()
    out["msg_train_loss"] = np.float32(loss.item())
    out["msg_train_logp"] = y.detach().numpy()
    out["msg_train_g_sa1"] = net.sa1.conv_blocks[0][0].weight.grad.numpy()
    out["msg_train_g_fp1"] = net.fp1.mlp_convs[0].weight.grad.numpy()
    out["msg_train_g_conv2"] = net.conv2.weight.grad.numpy()
    out["msg_train_rm_sa1"] = net.sa1.bn_blocks[0][0].running_mean.numpy()
    out["msg_train_rv_sa1"] = net.sa1.bn_blocks[0][0].running_var.numpy()

    # config 3: DGCNN k=20 (Highway_bridge/models/DGCNN.py), B=1
    net = parity.seeded_fill_(dgcnn_mod.DGCNN(5, 20), 3)
    out["dgcnn_logits"] = run("dgcnn", net, txyz[:1], trgb[:1]).numpy()

    # Highway PointNet2 (model.py:12) and BriStruNet = EnhancedPointNet2 (model.py:58), B=1
    net = parity.seeded_fill_(model_mod.PointNet2(5), 4)
    out["hbpn2_logits"] = run("hbpn2", net, txyz[:1], trgb[:1]).numpy()
    net = parity.seeded_fill_(model_mod.EnhancedPointNet2(5), 5)
    out["bristrunet_logits"] = run("bri", net, txyz[:1], trgb[:1]).numpy()
    # BriStruNet's index-producing stages on their own (independent of the ill-conditioned
    # eigenvalue features): FPS chain drawn with the same seed
    torch.manual_seed(SEED_FPS)
    cur = txyz[:1]
    for li, S in enumerate((1024, 512, 128)):
        fps = p2u.farthest_point_sample(cur, S)
        out[f"bri_fps{li}"] = fps.numpy().astype(np.int16)
        cur = p2u.index_points(cur, fps)
    crit = model_mod.BridgeStructureLoss(num_classes=5, alpha=80, rel_margin=0.3)
    out["bri_loss"] = np.float32(crit(torch.from_numpy(out["bristrunet_logits"]), tlab[:1], txyz[:1]).item())

    path = os.path.join(HERE, "models.npz")
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1e6:.2f} MB")
    for k, v in out.items():
        print(f"  {k:24s} {np.asarray(v).shape} {np.asarray(v).dtype}")


if __name__ == "__main__":
    main()

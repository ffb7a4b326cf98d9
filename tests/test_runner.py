"""Checkpoint-compatible training runner (pointcloud_bridge_b200/runner.py): conversion between the flat Adam
state and torch.optim.Adam's state_dict (CPU), and an end-to-end fit / save / resume on the GPU."""
import os

import numpy as np
import pytest
import torch

from pointcloud_bridge_b200 import runner


class _FakeFlatAdam:
    def __init__(self, n):
        self.exp_avg, self.exp_avg_sq = torch.randn(n), torch.rand(n)
        self.step_t, self.lr_t = torch.tensor([7]), torch.tensor([3e-4])
        self.betas, self.eps, self.weight_decay = (0.9, 0.999), 1e-8, 1e-4

    def set_lr(self, lr):
        self.lr_t.fill_(lr)


def test_adam_state_round_trips_through_torch_format():
    params = [torch.nn.Parameter(torch.randn(4, 3, 1, 1)), torch.nn.Parameter(torch.randn(4)), torch.nn.Parameter(torch.randn(5, 4))]
    n = sum(p.numel() for p in params)
    a = _FakeFlatAdam(n)
    sd = runner.adam_state_to_torch(a, params)
    # torch's own optimizer accepts it (the reference's script would resume from it)
    ref = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-4)
    ref.load_state_dict(sd)
    assert ref.param_groups[0]["lr"] == pytest.approx(3e-4)
    assert torch.equal(ref.state[params[2]]["exp_avg"], a.exp_avg[16:].view(5, 4))
    # and a state written by torch's optimizer loads back into the flat layout
    for p in params:
        p.grad = torch.randn_like(p)
    ref.step()
    b = _FakeFlatAdam(n)
    runner.adam_state_from_torch(b, ref.state_dict(), params)
    assert int(b.step_t) == 8 and float(b.lr_t) == pytest.approx(3e-4)
    assert torch.equal(b.exp_avg[:12].view(4, 3, 1, 1), ref.state[params[0]]["exp_avg"])
    assert torch.equal(b.exp_avg_sq[12:16], ref.state[params[1]]["exp_avg_sq"])


def test_adam_state_omits_parameters_that_never_had_a_gradient():
    """torch.optim.Adam keeps no state for grad = None parameters; the converted state must look the same, and a
    reference optimizer that only stepped the other parameters must load it and produce it."""
    params = [torch.nn.Parameter(torch.randn(3, 2)), torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(2, 2))]
    a = _FakeFlatAdam(sum(p.numel() for p in params))
    sd = runner.adam_state_to_torch(a, params, frozen=[1])
    assert set(sd["state"]) == {0, 2} and sd["param_groups"][0]["params"] == [0, 1, 2]
    assert torch.equal(sd["state"][2]["exp_avg"], a.exp_avg[11:].view(2, 2))
    ref = torch.optim.Adam(params, lr=1e-3)
    ref.load_state_dict(sd)
    assert params[1] not in ref.state
    params[0].grad, params[2].grad = torch.randn(3, 2), torch.randn(2, 2)          # params[1] stays without gradient
    ref.step()
    assert set(ref.state_dict()["state"]) == {0, 2}
    b = _FakeFlatAdam(a.exp_avg.numel())
    keep = b.exp_avg[6:11].clone()
    runner.adam_state_from_torch(b, ref.state_dict(), params)
    assert torch.equal(b.exp_avg[6:11], keep) and int(b.step_t) == 8


@pytest.mark.gpu
def test_fit_save_resume_bristrunet(tmp_path):
    """Two epochs of the BriStruNet runner on a tiny synthetic loader, checkpoints in the reference's dictionary
    format, resume into a fresh runner: same parameters, optimizer state and scheduler state; torch.optim.Adam
    loads the saved optimizer state (what the reference's script does on resume)."""
    from pointcloud_bridge_b200 import synthetic
    from pointcloud_bridge_b200.highway import model as hb
    dev = "cuda:0"

    def loader(seed, nb):
        for i in range(nb):
            xyz, rgb, lab = synthetic.bridge_batch(seed + i, 2, 4096)
            yield {"points": torch.from_numpy(xyz), "colors": torch.from_numpy(rgb), "labels": torch.from_numpy(lab)}

    def make():
        torch.manual_seed(0)
        net = hb.EnhancedPointNet2(5).to(dev)
        crit = hb.BridgeStructureLoss(num_classes=5, alpha=80, rel_margin=0.3).to(dev)
        return net, runner.Runner(net, lambda out, lab, pts: crit(out, lab, pts), graph=True, out_dir=str(tmp_path))

    net, r = make()
    logs = []
    r.fit(list(loader(0, 5)), list(loader(100, 2)), epochs=2, log=logs.append)
    assert len(logs) == 2 and all(np.isfinite(float(l.split("train loss ")[1].split()[0])) for l in logs)
    ck = torch.load(os.path.join(tmp_path, "latest_checkpoint.pth"), weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "val_acc", "val_loss"}
    assert os.path.exists(os.path.join(tmp_path, "best_model.pth")) and ck["epoch"] == 2
    torch.optim.Adam(hb.EnhancedPointNet2(5).parameters()).load_state_dict(ck["optimizer_state_dict"])
    net2, r2 = make()
    r2.load_checkpoint(os.path.join(tmp_path, "latest_checkpoint.pth"))
    assert r2.epoch == 2 and int(r2.trainer.opt.step_t) == int(r.trainer.opt.step_t) == 10
    assert torch.equal(r2.trainer.flat_param, r.trainer.flat_param)
    assert torch.equal(r2.trainer.opt.exp_avg, r.trainer.opt.exp_avg)
    assert r2.scheduler.state_dict()["num_bad_epochs"] == r.scheduler.state_dict()["num_bad_epochs"]
    r2.fit(list(loader(0, 2)), list(loader(100, 1)), epochs=1, log=logs.append)        # resumes and keeps training
    assert r2.epoch == 3 and len(logs) == 3

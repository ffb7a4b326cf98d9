"""Checkpoint-compatible training loop on engine.Trainer -- the loop of
Highway_bridge/train_MulSca_BriStruNet_CB.py:158-336 (Adam 1e-3 / weight decay 1e-4, ReduceLROnPlateau on the
validation accuracy, `best_model.pth` / `latest_checkpoint.pth` dictionaries) re-expressed for the captured
train step: one CUDA graph per step, loss and accuracy accumulated on the device and read once per epoch.

The checkpoint files keep the reference's keys -- 'epoch', 'model_state_dict', 'optimizer_state_dict' (in
torch.optim.Adam's per-parameter format, converted from / to the flat optimizer state), 'scheduler_state_dict'
(torch's own ReduceLROnPlateau), 'val_acc', 'val_loss' -- so either side can resume from the other's files.
Dataset loading, logging and tensorboard stay with the caller (out of scope).
"""
from __future__ import annotations

import os

import torch

from .engine import Trainer

__all__ = ["Runner", "adam_state_to_torch", "adam_state_from_torch"]


def adam_state_to_torch(opt, params, frozen=()) -> dict:
    """engine.FlatAdam state -> torch.optim.Adam.state_dict() layout for `params` (the order of model.parameters()).
    frozen: indices of parameters that never received a gradient -- torch.optim.Adam keeps no state for them."""
    state, off = {}, 0
    skip = set(frozen or ())
    for i, p in enumerate(params):
        n = p.numel()
        if i in skip:
            off += n
            continue
        state[i] = {"step": opt.step_t.detach().clone().float().reshape(()).cpu(),
                    "exp_avg": opt.exp_avg[off:off + n].view_as(p).clone(),
                    "exp_avg_sq": opt.exp_avg_sq[off:off + n].view_as(p).clone()}
        off += n
    group = {"lr": float(opt.lr_t.item()), "betas": tuple(opt.betas), "eps": opt.eps, "weight_decay": opt.weight_decay,
             "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
             "fused": None, "decoupled_weight_decay": False, "params": list(range(len(params)))}
    return {"state": state, "param_groups": [group]}


def adam_state_from_torch(opt, sd: dict, params) -> None:
    """torch.optim.Adam.state_dict() (e.g. the reference's 'optimizer_state_dict') -> engine.FlatAdam."""
    group = sd["param_groups"][0]
    off, step = 0, 0
    for i, p in enumerate(params):
        n = p.numel()
        st = sd["state"].get(i) or sd["state"].get(str(i))
        if st is not None:
            opt.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            opt.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, int(float(st["step"])))
        off += n
    opt.step_t.fill_(step)
    opt.set_lr(group["lr"])
    opt.betas, opt.eps, opt.weight_decay = tuple(group["betas"]), group["eps"], group["weight_decay"]


def _default_batch(batch, device):
    """Batches of the reference's datasets: dict(points [B,N,3], colors [B,N,3], labels [B,N])
    (train_MulSca_BriStruNet_CB.py:170-172) -> (inputs, labels, loss inputs)."""
    pts = batch["points"].to(device, non_blocking=True)
    col = batch["colors"].to(device, non_blocking=True)
    lab = batch["labels"].to(device, non_blocking=True).long()
    return (pts, col), lab, (pts,)


class Runner:
    def __init__(self, net, loss_fn, lr=1e-3, weight_decay=1e-4, amp=True, graph=True, out_dir=None,
                 sched_factor=0.1, sched_patience=5, batch_fn=_default_batch, class_dim=1):
        self.net, self.loss_fn, self.out_dir, self.batch_fn, self.class_dim = net, loss_fn, out_dir, batch_fn, class_dim
        self.trainer = Trainer(net, loss_fn=loss_fn, lr=lr, weight_decay=weight_decay, amp=amp, graph=graph)
        self.device = next(net.parameters()).device
        # torch's own scheduler on a one-parameter stand-in optimizer: same state_dict as the reference's, its
        # learning rate is copied into the device scalar the Adam kernel reads
        self._lr_holder = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self._lr_holder, mode="max", factor=sched_factor,
                                                                    patience=sched_patience)
        self.epoch, self.best_val_acc = 0, 0.0

    # ---- one epoch -------------------------------------------------------------------------------
    def _accuracy(self, logits, labels):
        return (logits.argmax(dim=self.class_dim) == labels).float().mean()

    def train_epoch(self, loader):
        self.net.train()
        loss_sum = torch.zeros((), device=self.device)
        acc_sum = torch.zeros((), device=self.device)
        n = 0
        for batch in loader:
            inputs, labels, loss_inputs = self.batch_fn(batch, self.device)
            loss = self.trainer.step(*inputs, labels=labels, loss_inputs=loss_inputs)
            loss_sum += loss
            acc_sum += (self.trainer.last_pred(self.class_dim) == labels).float().mean()
            n += 1
        n = max(n, 1)
        return float(loss_sum.item()) / n, float(acc_sum.item()) / n          # one host sync per epoch

    @torch.no_grad()
    def validate(self, loader):
        self.net.eval()
        loss_sum = torch.zeros((), device=self.device)
        acc_sum = torch.zeros((), device=self.device)
        n = 0
        for batch in loader:
            inputs, labels, loss_inputs = self.batch_fn(batch, self.device)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.trainer.amp):
                out = self.net(*inputs)
            logits = (out[0] if isinstance(out, tuple) else out).float()
            loss_sum += self.loss_fn(logits, labels, *loss_inputs)
            acc_sum += self._accuracy(logits, labels)
            n += 1
        n = max(n, 1)
        return float(loss_sum.item()) / n, float(acc_sum.item()) / n

    def fit(self, train_loader, val_loader, epochs, log=print):
        for _ in range(epochs):
            self.epoch += 1
            tr_loss, tr_acc = self.train_epoch(train_loader)
            val_loss, val_acc = self.validate(val_loader)
            self.scheduler.step(val_acc)                                       # train_MulSca_BriStruNet_CB.py:296
            self.trainer.opt.set_lr(self._lr_holder.param_groups[0]["lr"])
            log(f"epoch {self.epoch}: train loss {tr_loss:.4f} acc {tr_acc:.4f} | val loss {val_loss:.4f} acc {val_acc:.4f} "
                f"| lr {self._lr_holder.param_groups[0]['lr']:.2e}")
            if self.out_dir:
                if val_acc > self.best_val_acc:
                    self.best_val_acc = val_acc
                    self.save_checkpoint(os.path.join(self.out_dir, "best_model.pth"), self.best_val_acc, val_loss, False)
                self.save_checkpoint(os.path.join(self.out_dir, "latest_checkpoint.pth"), val_acc, val_loss, True)
        return self

    # ---- checkpoints in the reference's format (train_MulSca_BriStruNet_CB.py:317-335) ------------
    def save_checkpoint(self, path, val_acc, val_loss, with_scheduler=True):
        ck = {"epoch": self.epoch, "model_state_dict": self.net.state_dict(),
              "optimizer_state_dict": adam_state_to_torch(self.trainer.opt, self.trainer.bucket.params, self.trainer.frozen),
              "val_acc": val_acc, "val_loss": val_loss}
        if with_scheduler:
            ck["scheduler_state_dict"] = self.scheduler.state_dict()
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save(ck, path)

    def load_checkpoint(self, path):
        ck = torch.load(path, map_location=self.device, weights_only=False)
        self.net.load_state_dict(ck["model_state_dict"])      # parameters are views of the flat buffer: copied in place
        self.trainer.refresh()                                # bf16 weight shadows follow
        if "optimizer_state_dict" in ck:
            adam_state_from_torch(self.trainer.opt, ck["optimizer_state_dict"], self.trainer.bucket.params)
        if "scheduler_state_dict" in ck:
            self.scheduler.load_state_dict(ck["scheduler_state_dict"])
        self._lr_holder.param_groups[0]["lr"] = float(self.trainer.opt.lr_t.item())
        self.epoch = int(ck.get("epoch", 0))
        self.best_val_acc = float(ck.get("val_acc", 0.0))
        return ck

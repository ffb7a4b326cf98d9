"""Step drivers around the networks: the data-parallel training step (BASELINE configs 2/4) and
the block-sharded inference loop (config 5).  Mirrors the hot loop of
Highway_bridge/train_MulSca_BriStruNet_CB.py:158-190 (forward, loss, backward, Adam) and the
batched block evaluation of Partsize-identical/test_sem_seg.py:132-152, re-expressed for one
process per GPU.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import distributed as pdist

__all__ = ["Trainer", "BlockInference"]


class Trainer:
    """forward -> loss -> backward -> one flat NCCL all-reduce -> fused Adam.

    `amp`: bf16 autocast for the shared-MLP GEMMs (index kernels always run fp32 -- their results
    must be bit-exact); parameters, gradients and optimizer state stay fp32.
    """

    def __init__(self, net, loss_fn=None, lr=1e-3, weight_decay=1e-4, amp=True):
        self.net = net
        self.loss_fn = loss_fn
        self.amp = amp
        self.bucket = pdist.FlatGradBucket(net)
        self.opt = torch.optim.Adam(self.bucket.params, lr=lr, weight_decay=weight_decay, fused=True)

    def step(self, *inputs, labels, loss_inputs=()):
        """One optimisation step on this rank's batch; returns the (device) loss tensor."""
        self.bucket.zero()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            out = self.net(*inputs)
        logits = out[0] if isinstance(out, tuple) else out
        if self.loss_fn is None:       # sem-seg nets return log-probabilities [B,N,C]
            loss = F.nll_loss(logits.float().reshape(-1, logits.shape[-1]), labels.reshape(-1))
        else:
            loss = self.loss_fn(logits.float(), labels, *loss_inputs)
        loss.backward()
        self.bucket.allreduce_mean()
        self.opt.step()
        return loss.detach()


class BlockInference:
    """Evaluates independent 4096-point blocks in fixed-size batches on this rank's shard of the
    block list.  No collective on the data path; every rank keeps its own label slice."""

    def __init__(self, net, batch_blocks=32, amp=True):
        self.net = net.eval()
        self.batch_blocks = batch_blocks
        self.amp = amp

    @torch.no_grad()
    def run(self, blocks_x, out_labels=None):
        """blocks_x: [nb,9,N] device tensor (this rank's shard) -> uint8 labels [nb,N]."""
        nb = blocks_x.shape[0]
        if out_labels is None:
            out_labels = torch.empty(nb, blocks_x.shape[2], dtype=torch.uint8, device=blocks_x.device)
        for lo in range(0, nb, self.batch_blocks):
            x = blocks_x[lo:lo + self.batch_blocks]
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                logp, _ = self.net(x)
            out_labels[lo:lo + x.shape[0]] = logp.argmax(dim=-1).to(torch.uint8)
        return out_labels

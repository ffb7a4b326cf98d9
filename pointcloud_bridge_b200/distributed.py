"""One-process-per-GPU plumbing for the two ways the path scales:

* training replicas (configs 2/4): every rank holds a full model and its own batch of blocks;
  BatchNorm statistics stay per replica (the reference has no SyncBN); gradients live in ONE
  flat fp32 buffer so a step needs exactly one NCCL all-reduce over NVLink/NVSwitch, issued on
  the compute stream right after backward;
* block-sharded inference (config 5): scene blocks are independent, so ranks take contiguous
  slices of the block list and never communicate (`shard_range`).

The backend is NCCL on GPUs and gloo on CPU (used by the world_size-2 tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

__all__ = ["init_from_env", "FlatGradBucket", "shard_range", "max_over_ranks", "is_dist"]


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT).  Returns (rank, world_size, local_rank); single-process when
    WORLD_SIZE is absent or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not is_dist():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous slice of `n_items` independent blocks owned by `rank`: ceil(n/world) each, the
    last ranks possibly short or empty."""
    per = -(-n_items // world)
    lo = min(rank * per, n_items)
    return range(lo, min(lo + per, n_items))


def max_over_ranks(value: float, device) -> float:
    """Maximum of a scalar over all ranks (how multi-GPU step times are reported)."""
    if not is_dist():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class FlatGradBucket:
    """One flat fp32 buffer for all parameter gradients of `module` = one collective per step.

    Gradients are produced by autograd as ordinary per-parameter tensors (`p.grad = None` before
    backward, so no accumulate kernels run) or written straight into their flat view by the producing
    kernel (`zero()` clears the buffer at the start of a step); `pack()` gathers the former into the
    flat buffer with one batched multi-tensor copy, `allreduce_mean()` is a single NCCL all-reduce over the buffer,
    and `unpack()` scatters the averaged values back with another batched copy.  With one rank
    nothing is packed at all.
    """

    def __init__(self, module: torch.nn.Module, dtype: torch.dtype = torch.float32):
        self.params = [p for p in module.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        device = self.params[0].device
        self.flat = torch.zeros(total, dtype=dtype, device=device)
        self.views = []
        off = 0
        for p in self.params:
            n = p.numel()
            self.views.append(self.flat[off:off + n].view_as(p))
            off += n

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def zero(self) -> None:
        """Start of a step: no per-parameter gradients, flat buffer zero (kernels that write their
        gradient straight into a flat view -- ops.wgrad_rows -- accumulate into it)."""
        for p in self.params:
            p.grad = None
        self.flat.zero_()

    def _live(self):
        return [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]

    def pack(self) -> None:
        live = self._live()                        # parameters without a .grad were written in place or are unused
        if live:
            torch._foreach_copy_([v for v, _ in live], [g for _, g in live])

    def unpack(self) -> None:
        live = self._live()
        if live:
            torch._foreach_copy_([g for _, g in live], [v for v, _ in live])

    def allreduce_mean(self) -> None:
        if not is_dist() or dist.get_world_size() == 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)           # the mean in the collective itself
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / dist.get_world_size())

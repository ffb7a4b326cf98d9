"""Step drivers around the networks: the data-parallel training step (BASELINE configs 2/4) and
the block-sharded inference loop (config 5).  Mirrors the hot loop of
Highway_bridge/train_MulSca_BriStruNet_CB.py:158-190 (forward, loss, backward, Adam) and the
batched block evaluation of Partsize-identical/test_sem_seg.py:132-152, re-expressed for one
process per GPU.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from . import distributed as pdist
from . import ops

__all__ = ["Trainer", "FlatAdam", "BlockInference", "FpsStartBuffers", "run_sharded_scene", "segment_scene"]


class FpsStartBuffers:
    """Static device buffers for the FPS start indices of a captured step.

    `discover` mode (one eager warm-up step) records the (B, N) of every farthest_point_sample
    call while drawing normally; static LongTensors are then allocated OUTSIDE the capture (a
    tensor created inside it would be re-initialised by the captured fill kernel on every
    replay) and handed out in call order while the graph is recorded.  `refill()` repeats,
    before each replay, exactly the draws the eager code would make -- `torch.randint(0, N, (B,))`
    on the CPU default generator, in call order (pointnet_util.py:79) -- into pinned staging
    memory and copies them over.  torch.manual_seed therefore reproduces the reference's samples
    in graph mode too.
    """

    def __init__(self):
        self.shapes = []           # (B, N) per call, from the discovery step
        self.calls = []            # (N, device buffer, pinned staging buffer)
        self.mode = "off"          # off | discover | record
        self._next = 0

    def provider(self, B, N, device):
        if self.mode == "record":
            n, buf, _ = self.calls[self._next]
            assert n == N and buf.shape[0] == B, "captured step differs from the discovery step"
            self._next += 1
            return buf
        if self.mode == "discover":
            self.shapes.append((B, N, device))
        return ops._draw_start(B, N, device)

    RING = 4                       # pinned staging sets; a set is rewritten only after its copies have run

    def allocate(self):
        self.calls = [(N, torch.zeros(B, dtype=torch.long, device=dev),
                       [torch.zeros(B, dtype=torch.long).pin_memory() for _ in range(self.RING)])
                      for (B, N, dev) in self.shapes]
        self._events = [None] * self.RING
        self._refills = 0
        self._next = 0

    def refill(self, n=None):
        """n: number of live clouds of a short batch -- draws exactly what the eager code would draw for n
        clouds (the remaining entries keep their previous values; their clouds are padding).
        The host may be several replays ahead of the GPU: each staging set is guarded by the event of its last
        host->device copies and rewritten only after they have completed."""
        if not self.calls:
            return
        slot = self._refills % self.RING
        self._refills += 1
        if self._events[slot] is not None:
            self._events[slot].synchronize()
        for N, buf, stages in self.calls:
            stage = stages[slot]
            m = stage.shape[0] if n is None else n
            if n is not None:                                # padding clouds keep the previous replay's values
                stage.copy_(stages[(slot - 1) % self.RING])
            stage[:m].copy_(torch.randint(0, N, (m,), dtype=torch.long))
            buf.copy_(stage, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._events[slot] = ev


def _nll_mean(logp, labels):
    """F.nll_loss(logp, labels) (mean over rows, no class weights, no ignore_index) as gather + mean:
    ATen's 2-D nll_loss reduces 65 536 rows in a single thread block (68 us forward + 39 us
    backward in the round-1 profile)."""
    return -logp.gather(1, labels.view(-1, 1)).mean()


class FlatAdam:
    """torch.optim.Adam (L2 weight decay, bias-corrected) over ONE flat fp32 parameter buffer whose gradient is
    the flat gradient bucket: a single kernel per step (csrc/adam.cu) that also rewrites the bf16 weight
    shadows of the next step's GEMMs.  Step count and learning rate are device scalars, so the kernel can
    sit inside a captured CUDA graph and still follow `set_lr` (ReduceLROnPlateau in the reference's
    runner, train_MulSca_BriStruNet_CB.py:181)."""

    def __init__(self, flat_param, flat_grad, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 shadow_index=None, shadow_flat=None, shadow_index_t=None):
        self.p, self.g = flat_param, flat_grad
        self.exp_avg = torch.zeros_like(flat_param)
        self.exp_avg_sq = torch.zeros_like(flat_param)
        self.step_t = torch.zeros(1, dtype=torch.long, device=flat_param.device)
        self.lr_t = torch.full((1,), float(lr), dtype=torch.float32, device=flat_param.device)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.shadow_index, self.shadow_flat, self.shadow_index_t = shadow_index, shadow_flat, shadow_index_t
        self.skip = None                   # uint8 [n]: elements of parameters that never receive a gradient (Trainer)

    def set_lr(self, lr: float) -> None:
        self.lr_t.fill_(float(lr))

    @torch.no_grad()
    def step(self, increment: bool = True) -> None:
        """increment=False: the caller already advanced `step_t` (Trainer batches it with the BN counters)."""
        if increment:
            self.step_t.add_(1)
        ops._call("pcb_adam_flat_f32", self.p.device, self.p.data_ptr(), self.g.data_ptr(), self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), self.p.numel(), self.lr_t.data_ptr(), float(self.betas[0]),
                  float(self.betas[1]), float(self.eps), float(self.weight_decay), self.step_t.data_ptr(),
                  self.shadow_index.data_ptr() if self.shadow_index is not None else None,
                  self.shadow_index_t.data_ptr() if self.shadow_index_t is not None and self.shadow_index is not None else None,
                  self.shadow_flat.data_ptr() if self.shadow_index is not None else None,
                  self.skip.data_ptr() if self.skip is not None else None, alg_bytes=self.p.numel() * 32)

    def state_dict(self):
        return {"step": self.step_t.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": float(self.lr_t.item()), "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.step_t.copy_(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.set_lr(sd["lr"])
        self.betas, self.eps, self.weight_decay = tuple(sd["betas"]), sd["eps"], sd["weight_decay"]


class Trainer:
    """forward -> loss -> backward -> one flat NCCL all-reduce -> fused Adam.

    `amp`: bf16 autocast for the shared-MLP GEMMs (index kernels always run fp32 -- their results
    must be bit-exact); parameters, gradients and optimizer state stay fp32.
    """

    def __init__(self, net, loss_fn=None, lr=1e-3, weight_decay=1e-4, amp=True, graph=False, capturable=None):
        self.net = net
        self.loss_fn = loss_fn
        self.amp = amp
        self.graph = graph
        self.bucket = pdist.FlatGradBucket(net)
        # all parameters live in ONE flat fp32 buffer (the modules' tensors are views of it) whose
        # gradient is the bucket's flat buffer: the optimizer is a single fused-Adam launch over one
        # tensor instead of multi-tensor launches over ~200, and with several ranks the all-reduced
        # flat gradient is consumed in place (no unpack)
        self.flat_param = torch.empty_like(self.bucket.flat).requires_grad_(True)
        with torch.no_grad():
            off = 0
            for p in self.bucket.params:
                n = p.numel()
                view = self.flat_param.detach()[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += n
        offsets, off = {}, 0
        for p in self.bucket.params:
            offsets[id(p)] = off
            off += p.numel()
        self._ctx = ops.StepContext(net, bf16=amp,
                                    grad_views={id(p): v for p, v in zip(self.bucket.params, self.bucket.views)},
                                    flat_offsets=offsets, flat_numel=off)
        self._ctx.refresh_shadows()
        self._ctx.external_refresh = True          # from here on the Adam kernel keeps the bf16 shadows current
        self.opt = FlatAdam(self.flat_param.detach(), self.bucket.flat, lr=lr, weight_decay=weight_decay,
                            shadow_index=self._ctx.shadow_index, shadow_flat=self._ctx.shadow_flat,
                            shadow_index_t=self._ctx.shadow_index_t)
        self._g = None                 # captured (zero, forward, loss, backward[, Adam]) graph
        self._static = None
        self._starts = FpsStartBuffers()
        self._warm = 0
        self._opt_in_graph = False
        self._last_logits = None
        self._pf_stream = self._pf_bufs = self._pf_ready = self._pf_consumed = None
        self._pf_pre = None
        self._mark_consumed = False
        # Networks whose sampling / grouping indices depend on the coordinates only expose `index_chain`: the
        # indices of a batch are then computed before its step -- by `prefetch` on a side stream, while the previous
        # batch trains -- and the captured step contains no farthest point sampling, ball query or three-NN at all.
        self._use_chain = hasattr(net, "index_chain") and os.environ.get("PCB_NO_INDEX_PREFETCH", "0") != "1"
        self._static_pre = None
        # the index chain of the next batch starts when the running step reaches its small middle layers
        # (ops.mark_mid_step); an external event, so that a captured step records it on every replay
        self._mid_event = torch.cuda.Event(external=True) \
            if self._use_chain and os.environ.get("PCB_MID_STEP_CHAIN", "1") != "0" else None
        self._mid_armed = False
        self.frozen = None                 # indices of parameters that never receive a gradient (known after step 1)

    @torch.no_grad()
    def last_pred(self, class_dim: int = 1) -> torch.Tensor:
        """Predicted labels [B,N] of the step that just ran (for running accuracy, as the reference's loop computes
        `outputs.max(1)[1]`, train_MulSca_BriStruNet_CB.py:181).  In graph mode this reads the graph's static output."""
        lg = self._last_logits
        if isinstance(lg, ops.LogitRows):
            x = lg.rows[:, :lg.classes].float()
            if lg.bias is not None:
                x = x + lg.bias
            return x.argmax(dim=1).view(lg.B, lg.N)
        return lg.argmax(dim=class_dim)

    def refresh(self) -> None:
        """Call after changing parameters from outside (net.load_state_dict, manual edits): re-derives the
        bf16 weight shadows that the optimizer kernel otherwise keeps current."""
        self._ctx.refresh_shadows()
        ops.bump_param_generation()

    def close(self) -> None:
        """Drop the captured graphs (a graph that holds a captured NCCL all-reduce must go before the process group is
        destroyed, otherwise `destroy_process_group` can block) and the prefetch state.  The trainer can still step
        eagerly afterwards; a later graph step re-captures."""
        torch.cuda.synchronize()
        self._g = None
        if hasattr(self, "_g_opt"):
            self._g_opt = None
        self._static = self._static_pre = None
        self._pf_bufs = self._pf_pre = self._pf_ready = self._pf_consumed = None
        self._warm = 0

    # -- eager pieces ---------------------------------------------------------------------
    def _fwd_bwd(self, inputs, labels, loss_inputs, pre=None):
        self.bucket.zero()
        kw = {} if pre is None else {"pre": pre}
        ops.set_mid_step_event(self._mid_event)
        try:
            return self._fwd_bwd_marked(inputs, labels, loss_inputs, kw)
        finally:
            ops.set_mid_step_event(None)

    def _fwd_bwd_marked(self, inputs, labels, loss_inputs, kw):
        with self._ctx:
            self._ctx.counters.append(self.opt.step_t)      # advanced with the BN counters in one multi-tensor add
            if self.loss_fn is None:       # plain mean NLL: the heads hand their logits rows to the loss kernel
                with ops.head_logits_mode(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                    out = self.net(*inputs, **kw)
            else:
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                    out = self.net(*inputs, **kw)
            logits = out[0] if isinstance(out, tuple) else out
            self._last_logits = logits if isinstance(logits, ops.LogitRows) else logits.detach()
            if isinstance(logits, ops.LogitRows):
                loss = ops.nll_logit_rows(logits, labels)
            elif self.loss_fn is None:     # sem-seg nets return log-probabilities [B,N,C]
                loss = _nll_mean(logits.float().reshape(-1, logits.shape[-1]), labels.reshape(-1))
            else:
                loss = self.loss_fn(logits.float(), labels, *loss_inputs)
            loss.backward()
        return loss.detach()

    def _multi(self):
        return pdist.is_dist() and torch.distributed.get_world_size() > 1

    def _mark_gradient_less(self):
        """After the first backward pass: parameters with neither a `.grad` nor an in-place weight gradient (modules the
        forward never reaches -- EnhancedPointNet2.geometric1 / cls_head) are left alone by the optimizer, as
        torch.optim.Adam leaves grad = None parameters alone (no weight decay, no state)."""
        self.frozen = [i for i, p in enumerate(self.bucket.params)
                       if p.grad is None and id(p) not in self._ctx.gviews_used]
        if self.frozen:
            skip = torch.zeros(self.flat_param.numel(), dtype=torch.uint8, device=self.flat_param.device)
            off = 0
            for i, p in enumerate(self.bucket.params):
                if i in set(self.frozen):
                    skip[off:off + p.numel()] = 1
                off += p.numel()
            self.opt.skip = skip

    def _step_eager(self, inputs, labels, loss_inputs, pre=None):
        loss = self._fwd_bwd(inputs, labels, loss_inputs, pre)
        if self.frozen is None:
            self._mark_gradient_less()
        self.bucket.pack()
        if self._multi():
            self.bucket.allreduce_mean()
        self.opt.step(increment=False)
        ops.bump_param_generation()
        return loss

    # -- CUDA-graph path --------------------------------------------------------------------
    def _capture(self, inputs, labels, loss_inputs, pre=None):
        """Capture one step with static input buffers.  With several ranks the NCCL all-reduce and
        the optimizer stay outside the graph (zero + forward + loss + backward are captured)."""
        from torch.utils import _pytree as pytree
        self._static = ([t.clone() for t in inputs], labels.clone(), [t.clone() for t in loss_inputs])
        if pre is not None:                                  # static copies of the precomputed indices, same structure
            flat, spec = pytree.tree_flatten(pre)
            self._static_pre_flat = [t.clone() for t in flat]
            self._static_pre = pytree.tree_unflatten(self._static_pre_flat, spec)
        # several ranks: the NCCL all-reduce is captured into the same graph (one launch per step) unless
        # PCB_GRAPH_ALLREDUCE=0, in which case it runs eagerly between two graphs
        self._opt_in_graph = (not self._multi()) or os.environ.get("PCB_GRAPH_ALLREDUCE", "1") != "0"
        self._starts.allocate()
        ops.set_fps_start_provider(self._starts.provider)
        self._starts.mode = "record"
        self._g = torch.cuda.CUDAGraph()
        from . import _lib
        n0 = _lib.launches()
        try:
            with torch.cuda.graph(self._g):
                loss = self._fwd_bwd(self._static[0], self._static[1], self._static[2], self._static_pre)
                self.bucket.pack()
                if self._opt_in_graph:
                    if self._multi():
                        self.bucket.allreduce_mean()
                    self.opt.step(increment=False)
            self._static_loss = loss
            self.kernel_launches_per_replay = _lib.launches() - n0   # libpcbridge kernels inside the graph
            if not self._opt_in_graph:                               # second graph: the optimizer alone
                self._g_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._g_opt, pool=self._g.pool()):
                    self.opt.step(increment=False)
        finally:
            self._starts.mode = "off"
            ops.set_fps_start_provider(None)

    # -- input prefetch: the host->device copy of the NEXT batch overlaps the current step ------------------
    def prefetch(self, *inputs, labels, loss_inputs=()):
        """Start copying the next step's batch (pinned host or device tensors) into staging buffers on a side
        stream; `step_prefetched()` then only needs device-to-device copies.  Call it right after launching a
        step: the H2D transfer runs while that step computes."""
        dev = self.bucket.flat.device
        if self._pf_stream is None:
            # high priority: the index chain's few persistent CTAs (FPS: one per cloud) take an SM as soon as one frees up
            self._pf_stream = torch.cuda.Stream(device=dev, priority=-1)
            if self._use_chain and self._g is None:
                # ... and the cooperative BN grids of the step leave those SMs out (csrc/bn_rows.cu: g_coop_sms), otherwise
                # every cooperative launch waits for the whole FPS kernel.  Grid sizes are frozen at capture: set it first.
                from . import _lib
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                clouds = int(inputs[0].shape[0]) if inputs else 0
                reserve = int(os.environ.get("PCB_PREFETCH_SMS", str(min(clouds, sms // 4))))
                _lib.lib().pcb_bn_set_coop_sms(sms - reserve if reserve > 0 else 0)
        srcs = list(inputs) + [labels] + list(loss_inputs)
        if self._pf_bufs is None or [tuple(b.shape) for b in self._pf_bufs] != [tuple(t.shape) for t in srcs]:
            self._pf_bufs = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in srcs]
        if self._pf_consumed is not None:
            self._pf_stream.wait_event(self._pf_consumed)          # the previous step has read the staging buffers
        with torch.cuda.stream(self._pf_stream):
            for b, t in zip(self._pf_bufs, srcs):
                b.copy_(t, non_blocking=True)
            # sampling / grouping indices of that batch, on the same side stream: they overlap the running step, from
            # the point where its small middle layers begin (the input copies above start right away)
            if self._mid_event is not None and self._mid_armed and self._use_chain:
                self._pf_stream.wait_event(self._mid_event)
            self._pf_pre = self.net.index_chain(self._pf_bufs[0]) if self._use_chain else None
            self._pf_ready = torch.cuda.Event()
            self._pf_ready.record(self._pf_stream)
        self._pf_counts = (len(inputs), len(loss_inputs))

    def step_prefetched(self):
        """`step` on the batch handed to `prefetch`."""
        main = torch.cuda.current_stream()
        main.wait_event(self._pf_ready)
        ni, nl = self._pf_counts
        bufs = self._pf_bufs
        pre = self._pf_pre
        if pre is not None:                  # allocated on the side stream, consumed on this one
            from torch.utils import _pytree as pytree
            for t in pytree.tree_flatten(pre)[0]:
                t.record_stream(main)
        self._mark_consumed = True
        loss = self.step(*bufs[:ni], labels=bufs[ni], loss_inputs=tuple(bufs[ni + 1:ni + 1 + nl]), pre=pre)
        self._mid_armed = True              # a step is in flight: its mid-step marker will be recorded
        if self._mark_consumed:             # eager step: the staging buffers were read throughout
            self._note_consumed()
        return loss

    def _note_consumed(self):
        self._pf_consumed = torch.cuda.Event()
        self._pf_consumed.record()
        self._mark_consumed = False

    def step(self, *inputs, labels, loss_inputs=(), pre=None):
        """One optimisation step on this rank's batch; returns the (device) loss tensor.
        pre: the batch's precomputed indices (`net.index_chain`), normally handed over by `step_prefetched`; computed
        here, on the step's own stream, when the network has an index chain and none was given."""
        replay = self.graph and self._warm >= 3 and self._g is not None
        if not replay:                                       # eager, warm-up and capture steps want device tensors
            dev = self.bucket.flat.device
            put = lambda t: t if t.is_cuda else t.to(dev, non_blocking=True)
            inputs, labels, loss_inputs = tuple(put(t) for t in inputs), put(labels), tuple(put(t) for t in loss_inputs)
            if self._use_chain and pre is None:
                pre = self.net.index_chain(inputs[0])
        if not self.graph:
            return self._step_eager(inputs, labels, loss_inputs, pre)
        if self._warm < 3:                                  # eager warm-up steps on a side stream
            self._warm += 1
            discover = self._warm == 3                      # the last one also records the FPS call list
            if discover:
                self._starts.shapes = []
                self._starts.mode = "discover"
                ops.set_fps_start_provider(self._starts.provider)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            try:
                with torch.cuda.stream(s):
                    loss = self._step_eager(inputs, labels, loss_inputs, pre)
            finally:
                if discover:
                    self._starts.mode = "off"
                    ops.set_fps_start_provider(None)
            torch.cuda.current_stream().wait_stream(s)
            return loss
        if self._g is None:
            self._capture(inputs, labels, loss_inputs, pre)
        for dst, src in zip(self._static[0], inputs):
            dst.copy_(src, non_blocking=True)
        self._static[1].copy_(labels, non_blocking=True)
        for dst, src in zip(self._static[2], loss_inputs):
            dst.copy_(src, non_blocking=True)
        if self._static_pre is not None:
            from torch.utils import _pytree as pytree
            if pre is None:                                  # no prefetch: the index chain runs here, on the step's stream
                pre = self.net.index_chain(self._static[0][0])
            torch._foreach_copy_(self._static_pre_flat, pytree.tree_flatten(pre)[0])
        if self._mark_consumed:             # the prefetch staging buffers are free again: the next copy may start
            self._note_consumed()
        self._starts.refill()
        self._g.replay()
        ops.bump_param_generation()
        from . import _lib
        _lib.count_launches(self.kernel_launches_per_replay)
        if not self._opt_in_graph:                          # eager NCCL all-reduce between the two graphs
            self.bucket.allreduce_mean()
            self._g_opt.replay()
        return self._static_loss


class BlockInference:
    """Evaluates independent 4096-point blocks in fixed-size batches on this rank's shard of the
    block list.  No collective on the data path; every rank keeps its own label slice.
    With `graph=True` the forward of a full batch is captured once into a CUDA graph (FPS start
    indices refilled from the CPU generator before every replay, as in training)."""

    def __init__(self, net, batch_blocks=128, amp=True, graph=True):
        # 128 blocks per batch: farthest point sampling runs one CTA per cloud, so a batch should cover the 148 SMs
        # (32 -> 128 blocks per batch: 72 -> 99 M points/s in the scene benchmark)
        self.net = net.eval()
        self.batch_blocks = batch_blocks
        self.amp = amp
        self.graph = graph
        self._g = None
        self._starts = FpsStartBuffers()
        self._seen = 0

    def _forward(self, x):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            logp, _ = self.net(x)
        return logp.argmax(dim=-1).to(torch.uint8)

    def _forward_graphed(self, x):
        if self._seen < 2:                                   # two eager batches; the second records the FPS calls
            self._seen += 1
            if self._seen == 2:
                self._starts.shapes, self._starts.mode = [], "discover"
                ops.set_fps_start_provider(self._starts.provider)
            try:
                return self._forward(x)
            finally:
                self._starts.mode = "off"
                ops.set_fps_start_provider(None)
        if self._g is None:
            self._static_x = x.clone()
            self._starts.allocate()
            self._starts.mode = "record"
            ops.set_fps_start_provider(self._starts.provider)
            torch.cuda.synchronize()
            self._g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(self._g):
                    self._static_out = self._forward(self._static_x)
            finally:
                self._starts.mode = "off"
                ops.set_fps_start_provider(None)
        self._static_x.copy_(x, non_blocking=True)
        self._starts.refill()
        self._g.replay()
        return self._static_out

    @torch.no_grad()
    def run(self, blocks_x, out_labels=None):
        """blocks_x: [nb,9,N] device tensor (this rank's shard) -> uint8 labels [nb,N]."""
        nb = blocks_x.shape[0]
        if out_labels is None:
            out_labels = torch.empty(nb, blocks_x.shape[2], dtype=torch.uint8, device=blocks_x.device)
        for lo in range(0, nb, self.batch_blocks):
            x = blocks_x[lo:lo + self.batch_blocks]
            n = x.shape[0]
            if self.graph and n == self.batch_blocks:
                lab = self._forward_graphed(x)
            elif self.graph and self._g is not None:
                # short tail batch: replay the captured full-batch graph on the static buffer (the rows beyond n
                # keep the previous batch's blocks; blocks are independent, their outputs are dropped)
                self._static_x[:n].copy_(x, non_blocking=True)
                self._starts.refill(n)
                self._g.replay()
                lab = self._static_out[:n]
            else:
                lab = self._forward(x)
            out_labels[lo:lo + n] = lab
        return out_labels


def run_sharded_scene(net, blocks_x_host, rank, world, device, batch_blocks=128, amp=True, chunk_blocks=512,
                      infer=None):
    """Config 5: evaluate a scene that was tiled into independent blocks.  `blocks_x_host` is the
    whole scene's pinned host tensor [nb,9,N]; this rank takes its contiguous shard
    (distributed.shard_range), streams it through the GPU in chunks (H2D on a side stream
    overlapped with compute) and returns (range, uint8 labels [n_local,N] on the host).
    No collective: ranks never exchange data."""
    rng = pdist.shard_range(blocks_x_host.shape[0], rank, world)
    n_local = len(rng)
    out = torch.empty(n_local, blocks_x_host.shape[2], dtype=torch.uint8).pin_memory()
    if n_local == 0:
        return rng, out
    infer = infer or BlockInference(net, batch_blocks=batch_blocks, amp=amp)
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    chunks = [(lo, min(lo + chunk_blocks, n_local)) for lo in range(0, n_local, chunk_blocks)]

    def upload(i):
        lo, hi = chunks[i]
        with torch.cuda.stream(copy_stream):
            x = blocks_x_host[rng.start + lo:rng.start + hi].to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return x, ev

    nxt = upload(0)
    for i, (lo, hi) in enumerate(chunks):
        x, ev = nxt
        if i + 1 < len(chunks):
            nxt = upload(i + 1)
        main.wait_event(ev)
        lab = infer.run(x)
        out[lo:hi].copy_(lab, non_blocking=True)
        x.record_stream(main)
    main.synchronize()
    return rng, out


@torch.no_grad()
def segment_scene(net, points, num_classes, rank=0, world=1, block_points=4096, stride=0.5, block_size=1.0,
                  padding=0.001, batch_blocks=128, amp=True, num_votes=1, infer=None, seed=0):
    """Whole-scene evaluation on the GPU -- the loop of Partsize-identical/test_sem_seg.py:120-162 (tile the scene
    into overlapping blocks, predict every block `num_votes` times, scatter the predictions back as votes, take
    the per-point argmax).  points [P, >=6] fp32 CUDA tensor (x, y, z, r, g, b); returns uint8 labels [P].
    With several ranks every rank tiles the scene, evaluates its contiguous shard of the block list
    (distributed.shard_range) and the int32 vote counts are summed with ONE all-reduce -- the only exchange."""
    from . import scene
    tiler = scene.SceneTiler(block_points, stride, block_size, padding, seed=seed)
    infer = infer or BlockInference(net, batch_blocks=batch_blocks, amp=amp)
    pool = scene.new_vote_pool(points.shape[0], num_classes, points.device)
    for vote in range(num_votes):
        # every vote sees another pseudo-random subsample of each window (the reference re-draws padding and shuffle
        # on every pass, BridgeDataLoader.py:239-242); the block list -- and with it the shards -- has the same length
        tiles = tiler.tile(points, vote=vote, block_range=lambda nb: pdist.shard_range(nb, rank, world))
        if tiles.data.shape[0]:                               # this rank's shard of the block list only
            scene.add_vote(pool, tiles.point_idx, infer.run(tiles.model_input()))
        del tiles
    if world > 1 and pdist.is_dist():
        torch.distributed.all_reduce(pool)
    return scene.vote_argmax(pool)

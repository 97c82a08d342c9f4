"""Host-side mirror of the reference's training iteration for synthetic-data runs
(nnUNetTrainer_simple.run_iteration, nnUNetTrainer_simple.py:529-583, and the setup in
simple_main.py:145-168): zero_grad -> forward -> deep-supervision loss -> backward ->
clip_grad_norm_(12) -> SGD(nesterov) step -> mask.step() -> loss read-back.

The loss (DC_and_CE_loss wrapped in MultipleOutputLoss2; dice_loss.py:302-359,
deep_supervision.py:18-43) and the optimizer are NOT part of the hot path this package
replaces (SURVEY 8(f) ranks them "next"); they stay plain torch here exactly as the
unchanged reference trainer would run them.
"""
from __future__ import annotations

import argparse
from typing import List, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, ops

POOLS = {
    "btcv": [[1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2]],
    "brats": [[2, 2, 2]] * 5,
    "hippo": [[2, 2, 2]] * 3 + [[1, 1, 1]] * 2,
}


def build_network(in_ch: int, num_classes: int, pools, patch, base: int = 48, deep_supervision: bool = True):
    """constructs the drop-in network with the positional arguments the reference trainer uses
    (nnUNetTrainer_simple.py:292-301)."""
    from .network_architecture.unetpp_d import Generic_UNetPlusPlus, InitWeights_He
    return Generic_UNetPlusPlus(tuple(patch), in_ch, base, num_classes, len(pools), 2, 2, nn.Conv3d,
                                nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True}, nn.Dropout3d,
                                {'p': 0, 'inplace': True}, nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True},
                                deep_supervision, False, lambda x: x, InitWeights_He(1e-2), pools, None, False, True,
                                True)


def ds_loss_weights(n_outputs: int = 4, net_numpool: int = 5) -> List[float]:
    w = np.array([1 / (2 ** i) for i in range(net_numpool)])
    w[-1] = 0
    w = w / w.sum()
    return [float(v) for v in w[:n_outputs]]


def dc_and_ce_loss(logits: torch.Tensor, target: torch.Tensor, smooth: float = 1e-5) -> torch.Tensor:
    """DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False}, {})"""
    ce = F.cross_entropy(logits, target[:, 0].long())
    prob = torch.softmax(logits, 1)
    onehot = torch.zeros_like(prob).scatter_(1, target.long(), 1.0)
    axes = tuple(range(2, logits.ndim))
    tp = (prob * onehot).sum(axes)
    fp = (prob * (1 - onehot)).sum(axes)
    fn = ((1 - prob) * onehot).sum(axes)
    dc = (2 * tp + smooth) / (2 * tp + fp + fn + smooth + 1e-8)
    return ce - dc[:, 1:].mean()


def multiple_output_loss(outs: Sequence[torch.Tensor], targets: Sequence[torch.Tensor]) -> torch.Tensor:
    w = ds_loss_weights(len(outs))
    l = w[0] * dc_and_ce_loss(outs[0], targets[0])
    for i in range(1, len(outs)):
        if w[i] != 0:
            l = l + w[i] * dc_and_ce_loss(outs[i], targets[i])
    return l


def synthetic_batch(batch: int, in_ch: int, num_classes: int, patch, pools, seed: int = 1):
    """the reference's dummy-load convention (nnUNetTrainerV2_dummyLoad.py:29-31): uniform data,
    targets = round(rand * (ncls-1)) at the 4 deep-supervision scales.  Host tensors."""
    g = torch.Generator().manual_seed(seed)
    data = torch.rand((batch, in_ch) + tuple(patch), generator=g)
    targets = []
    sp = np.array(patch)
    for k in range(4):
        targets.append(torch.round(torch.rand((batch, 1) + tuple(int(v) for v in sp), generator=g) * (num_classes - 1)))
        sp = sp // np.array(pools[k])
    return data, targets


def allreduce_mean_grads(params, world_size: int, group=None):
    """data-parallel gradient mean: ONE flat all-reduce over all gradients (the E2ENet gradients are
    ~95 MB fp32; NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests)."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or world_size <= 1:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, group=group)
    flat.div_(world_size)
    for g, s in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(s)


class SparseArgs(argparse.Namespace):
    adv = False
    fix = False
    update_frequency = 1200
    final_density = 0.05


class TrainStep(object):
    """one data-parallel replica of the reference's training iteration on synthetic data.

    fused_optimizer=True (default): clip_grad_norm_ + Nesterov SGD + apply_mask run as the three multi-tensor
    launches of optim.FusedSGD, weight gradients are written by the backward kernels straight into one flat
    gradient arena (laid out in gradient-ready order), and under data parallelism that arena is all-reduced in
    `n_buckets` buckets on a communication stream while the rest of the backward still runs.
    fused_optimizer=False: stock torch.optim.SGD + clip_grad_norm_ + Masking.apply_mask and one flat all-reduce
    after the backward, exactly the calls the unchanged reference trainer makes around the drop-in modules."""

    def __init__(self, in_ch=1, num_classes=14, pools=None, patch=(64, 160, 160), density=0.2, death_rate=0.5,
                 update_frequency=1200, device="cuda", world_size=1, seed=0, base=48, total_steps=250000,
                 fused_loss=True, fused_optimizer=True, n_buckets=4, process_group=None):
        from .sparselearning.core_channel import CosineDecay, Masking
        import random
        pools = POOLS["btcv"] if pools is None else pools
        torch.manual_seed(seed)
        self.device = torch.device(device)
        self.pools, self.patch, self.in_ch, self.num_classes = pools, tuple(patch), in_ch, num_classes
        self.network = build_network(in_ch, num_classes, pools, patch, base).to(self.device)
        self.fused_optimizer = bool(fused_optimizer)
        self.group = process_group
        if self.fused_optimizer:
            from .optim import FusedSGD
            self.optimizer = FusedSGD(self.network.parameters(), 1e-2, momentum=0.99, weight_decay=3e-5, nesterov=True,
                                      max_norm=12.0, grad_scale=1.0 / max(1, world_size))
            if _lib.precision() == "fp16":
                # fp16 gradients need the reference loop's loss scaling (GradScaler, nnUNetTrainer_simple.py:553-562);
                # here it is device state of the fused optimizer, so it also lives inside the captured graph
                self.optimizer.enable_loss_scale()
        else:
            self.optimizer = torch.optim.SGD(self.network.parameters(), 1e-2, weight_decay=3e-5, momentum=0.99,
                                             nesterov=True)
            self.amp_grad_scaler = torch.amp.GradScaler("cuda") if _lib.precision() == "fp16" else None
        args = SparseArgs()
        args.update_frequency = update_frequency
        random.seed(seed)
        self.mask = Masking(self.optimizer, death_rate=death_rate, death_mode='magnitude',
                            death_rate_decay=CosineDecay(death_rate, total_steps), growth_mode='random',
                            redistribution_mode='none', args=args)
        self.mask.add_module(self.network, sparse_init='uniform', density=density)
        if self.fused_optimizer:
            self.optimizer.set_masks_from(self.mask)
        self.world_size = world_size
        self.n_buckets = n_buckets
        self.arena = None               # built after the first backward (gradient-ready order is recorded there)
        self._comm = None
        self._graph = None
        # the trainer's loss (nnUNetTrainer_simple.py:100,200-215): fused statistics kernels, or plain torch
        if fused_loss:
            from .loss_functions import DC_and_CE_loss, MultipleOutputLoss2
            self.loss = MultipleOutputLoss2(DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False}, {}),
                                            ds_loss_weights(4, len(pools)))
        else:
            self.loss = multiple_output_loss

    def _allreduce_grads(self):
        """data-parallel gradient mean over NCCL (one flat bucket; grads are ~95 MB fp32)."""
        allreduce_mean_grads(list(self.network.parameters()), self.world_size)

    # ------------------------------------------------------------------ gradient arena + bucketed all-reduce
    def _record_ready_order(self, data, targets):
        """one plain iteration (no arena) whose only extra job is to record the order in which the parameters
        receive their gradients; the arena is laid out in that order so that its buckets complete front to back"""
        order, hooks = [], []
        for p in self.network.parameters():
            hooks.append(p.register_post_accumulate_grad_hook(lambda q, order=order: order.append(q)))
        self.optimizer.zero_grad(set_to_none=True)
        self._backward(self.loss(self.network(data), targets))
        for h in hooks:
            h.remove()
        seen = set(id(q) for q in order)
        order += [p for p in self.network.parameters() if id(p) not in seen]
        self.optimizer.zero_grad(set_to_none=True)
        return order

    def _build_arena(self, data, targets):
        from .optim import GradArena, attach_arena
        self.arena = GradArena(self._record_ready_order(data, targets), n_buckets=self.n_buckets)
        attach_arena(self.arena)
        if self.world_size > 1:
            self._comm = torch.cuda.Stream(device=self.device)
            self.arena.on_bucket_ready = self._allreduce_bucket

    def _allreduce_bucket(self, b: int):
        """bucket b of the arena is complete on the compute stream: sum it over the ranks on the communication
        stream (NCCL over NVLink / NVSwitch) while the backward continues; the mean's 1/world is FusedSGD's grad_scale"""
        import torch.distributed as dist
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        ev_side = None
        if ops.side_pending():                      # weight gradients of this bucket may still be running on the side stream
            ev_side = torch.cuda.Event()
            ev_side.record(ops.side_stream(self.device))
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            if ev_side is not None:
                self._comm.wait_event(ev_side)
            dist.all_reduce(self.arena.bucket(b), group=self.group)
        self._reduced.add(b)

    def _backward(self, l):
        """backward of the (loss-scaled, under fp16) loss; the scale is a device scalar of the fused optimizer"""
        sc = self.optimizer.loss_scale(self.device) if self.fused_optimizer else None
        (l if sc is None else l * sc).backward()

    def _device_step(self, data, targets):
        """everything of one iteration that runs on the device (no host synchronisation)"""
        if not self.fused_optimizer:
            self.optimizer.zero_grad()
            output = self.network(data)
            l = self.loss(output, targets)
            gs = self.amp_grad_scaler
            if gs is not None:                           # the reference loop's fp16 branch, :552-562
                gs.scale(l).backward()
                if self.world_size > 1:
                    self._allreduce_grads()
                gs.unscale_(self.optimizer)
                torch.nn.utils.clip_grad_norm_(self.network.parameters(), 12)
                gs.step(self.optimizer)
                gs.update()
                return l.detach()
            l.backward()
            if self.world_size > 1:
                self._allreduce_grads()
            torch.nn.utils.clip_grad_norm_(self.network.parameters(), 12)
            self.optimizer.step()
            return l.detach()
        if self.arena is None:
            self._build_arena(data, targets)
        self.optimizer.zero_grad(set_to_none=True)
        self._reduced = set()
        self.arena.begin_step()
        output = self.network(data)
        l = self.loss(output, targets)
        self._backward(l)
        ops.side_join()                       # weight gradients launched on the side stream (ops._SideSection)
        if self.world_size > 1:
            for b in range(self.arena.n_buckets):        # buckets whose completion the hooks did not see
                if b not in self._reduced:
                    self._allreduce_bucket(b)
            torch.cuda.current_stream().wait_stream(self._comm)
        if not self._arena_checked:
            for p in self.network.parameters():
                if p.grad is None or p.grad.data_ptr() != self.arena.view(p).data_ptr():
                    raise RuntimeError("TrainStep: a gradient did not land in the arena (autograd copied it); the "
                                       "bucketed all-reduce would miss it")
            self._arena_checked = True
        self.optimizer.step()                 # clip + Nesterov SGD + apply_mask, three launches
        self.arena.zeroed_this_step = False   # a further backward before the next begin_step() must not assume empty slots
        return l.detach()

    _arena_checked = False

    def step(self, data: torch.Tensor, targets: Sequence[torch.Tensor]) -> torch.Tensor:
        if self._graph is not None:
            return self._graph_step(data, targets)
        staged = self._take_staged(data)
        if staged is not None:
            data, targets = staged
        l = self._device_step(data, targets)
        self._release_stage()
        self.mask.step(_mask_already_applied=self.fused_optimizer)
        return l

    # ------------------------------------------------------------------ whole-step CUDA graph
    def enable_graph(self, data: torch.Tensor, targets: Sequence[torch.Tensor], warmup: int = 3):
        """captures one training iteration (forward, loss, backward, gradient all-reduce, clip, SGD,
        apply_mask: ~750 kernel launches) into ONE CUDA graph.  Later `step()` calls copy the batch into
        the static input buffers and replay it; Masking's host bookkeeping and the rare prune / regrow
        update stay eager.  Shapes must not change afterwards.  The `warmup` iterations that precede the capture
        are REAL training iterations on the given batch (weights, momentum, Masking.steps and the death-rate schedule
        advance); hyper-parameters are read from FusedSGD's device array, so param_groups changes (poly-LR) take
        effect on replay (with fused_optimizer=False torch's SGD bakes lr into the graph: re-capture after changing it)."""
        assert self._graph is None
        self._static_data = data.clone()
        self._static_targets = [t.clone() for t in targets]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 2)):            # lazy state: momentum buffers, plans, tensor maps, pack registry
                self._device_step(self._static_data, self._static_targets)
                self.mask.step(_mask_already_applied=self.fused_optimizer)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        self.optimizer.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        # the iteration is captured on a HIGH-priority stream: the weight-gradient GEMMs forked to ops' side stream (default
        # priority) then yield the SMs to the data-gradient chain whenever both are ready
        hi = torch.cuda.Stream(device=self.device, priority=-1)
        with torch.cuda.graph(graph, stream=hi):
            self._static_loss = self._device_step(self._static_data, self._static_targets)
            if not self.fused_optimizer:
                self.mask.apply_mask()
        self.graph_launches = _lib.launch_count() - n0       # kernels of libe2enet_b200.so inside one replay
        self._graph = graph                                   # (capture records, it does not execute)
        from . import ops
        self._graph_keepalive = ops.pack_registry_keepalive(self.device)   # memory the graph's pack launch points at
        return self

    # ------------------------------------------------------------------ input prefetch (host -> device overlap)
    def prefetch(self, data: torch.Tensor, targets: Sequence[torch.Tensor]):
        """starts the host -> device copy of the NEXT batch (pinned host tensors) on a copy stream while the current
        iteration computes; the next `step(data, targets)` with these same host tensors then only waits for that
        copy and moves the staged batch into the graph's static inputs with a device-side copy (30 MB: ~10 us)
        instead of a serialised PCIe transfer (~0.6 ms).  Two staging sets alternate, so a copy never waits for
        the iteration that is still computing."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stages, self._stage_idx = [None, None], 0
        k = self._stage_idx
        self._stage_idx ^= 1
        st = self._stages[k]
        if st is None or st["data"].shape != data.shape:
            st = {"data": torch.empty(data.shape, dtype=data.dtype, device=self.device),
                  "targets": [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in targets], "free": None}
            self._stages[k] = st
        if st["free"] is not None:
            self._copy_stream.wait_event(st["free"])       # the iteration that last read this set has finished
        with torch.cuda.stream(self._copy_stream):
            st["data"].copy_(data, non_blocking=True)
            for s_, t in zip(st["targets"], targets):
                s_.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._staged = (data.data_ptr(), ev, st)

    _copy_stream = None
    _staged = None
    _stage_in_use = None

    def _take_staged(self, data):
        """(staged data, staged targets) if `data` is the host tensor a prefetch() was started for, else None"""
        st = self._staged
        if st is None or st[0] != data.data_ptr() or data.is_cuda:
            return None
        torch.cuda.current_stream().wait_event(st[1])
        self._staged = None
        self._stage_in_use = st[2]
        return st[2]["data"], st[2]["targets"]

    def _release_stage(self):
        if self._stage_in_use is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._stage_in_use["free"] = ev
            self._stage_in_use = None

    def _graph_step(self, data, targets):
        staged = self._take_staged(data)
        if staged is not None:
            data, targets = staged
        if data.data_ptr() != self._static_data.data_ptr():
            self._static_data.copy_(data, non_blocking=True)
        for s_, t in zip(self._static_targets, targets):
            if t.data_ptr() != s_.data_ptr():
                s_.copy_(t, non_blocking=True)
        self._release_stage()                 # the staged batch now lives in the graph's static inputs
        if self.fused_optimizer:
            self.optimizer.sync_hyper()       # learning-rate / momentum / weight-decay changes reach the replayed step
        self._graph.replay()
        # the replay changed weights (SGD, apply_mask) behind torch's version counters: any EAGER use of the
        # network after this (validation, inference) must repack its operands
        from . import ops
        ops.bump_weight_epoch()
        self.mask.step(_mask_already_applied=True)            # host bookkeeping; prune / regrow when due
        return self._static_loss.clone()                      # the static buffer is overwritten by the next replay

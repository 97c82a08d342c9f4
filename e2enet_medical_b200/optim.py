"""Fused optimizer step + gradient arena.

`FusedSGD` is a `torch.optim.Optimizer` with the reference trainer's configuration surface
(`SGD(params, lr, weight_decay, momentum, nesterov)`, nnUNetTrainer_simple.py:367-371): `param_groups[i]['lr']`
can be rewritten by the poly-LR schedule (:863-877) and `state[p]['momentum_buffer']` is what the reference's
`Masking.apply_mask` multiplies (core_channel.py:427-434).  Its `step()` runs

    clip_grad_norm_(params, max_norm)  ->  Nesterov SGD with weight decay  ->  w *= mask, momentum *= mask

as three multi-tensor launches of libe2enet_b200.so (csrc/optim.cu) instead of ~200 ATen launches and a separate
`mask_apply_multi`; the clip coefficient never leaves the device and the hyper-parameters are read from a small
device array, so a captured CUDA graph follows learning-rate changes.

`GradArena` gives every parameter's gradient a fixed place in ONE flat fp32 buffer, laid out in the order in which
the backward pass produces them, and cuts it into buckets: the data-parallel gradient all-reduce then runs per
bucket on a communication stream as soon as the bucket's last weight gradient has been written, overlapped with
the rest of the backward, with no flatten / copy-back passes.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


class GradArena(object):
    """flat fp32 gradient storage for a list of parameters.  `order` (a list of parameters) fixes the layout --
    pass the order in which gradients become ready (reverse of use) so that buckets fill front to back."""

    def __init__(self, params, n_buckets: int = 4, align: int = 64):
        self.params = list(params)
        dev = self.params[0].device
        off, self.offset = 0, {}
        for p in self.params:
            self.offset[id(p)] = off
            off += (p.numel() + align - 1) // align * align
        self.total = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        # buckets: contiguous ranges of ~equal size; a bucket is complete when its LAST parameter has its gradient
        self.n_buckets = max(1, min(int(n_buckets), len(self.params)))
        target = (off + self.n_buckets - 1) // self.n_buckets
        self.bucket_of, self.bucket_range, self.bucket_last = {}, [], []
        lo, b = 0, 0
        for i, p in enumerate(self.params):
            end = self.offset[id(p)] + (p.numel() + align - 1) // align * align
            self.bucket_of[id(p)] = b
            last = i == len(self.params) - 1
            if (end - lo >= target and b < self.n_buckets - 1) or last:
                self.bucket_range.append((lo, end))
                self.bucket_last.append(id(p))
                lo, b = end, b + 1
        self.n_buckets = len(self.bucket_range)
        # packed fp32 scratch of the weight-gradient GEMMs (they ADD their split-K results into it): one persistent
        # buffer zeroed once per step instead of one torch.zeros per layer; sized from the first step's demand
        self._scratch = None
        self._scratch_used = 0
        self._scratch_demand = 0
        self.zeroed_this_step = False       # begin_step() zeroed the whole buffer and nothing has been added since
        self._pending = None                # per-step: set of parameter ids of each bucket still missing
        self.on_bucket_ready = None         # callback(bucket index) -- installed by the data-parallel trainer

    def view(self, p: torch.Tensor) -> torch.Tensor:
        """a FRESH view tensor of p's slot (autograd's AccumulateGrad adopts a gradient without copying it only
        when nobody else holds a reference to that tensor object)"""
        off = self.offset[id(p)]
        return self.flat.narrow(0, off, p.numel()).view(p.shape)

    def has(self, p) -> bool:
        return id(p) in self.offset

    def begin_step(self):
        """start of an iteration: zero the arena (the tcgen05 weight-gradient kernels ADD their split-K results into
        their slots) and re-arm the bucket bookkeeping"""
        self.flat.zero_()
        self.zeroed_this_step = True
        if self._scratch is None and self._scratch_demand > 0:
            self._scratch = torch.empty(self._scratch_demand, dtype=torch.float32, device=self.flat.device)
        if self._scratch is not None:
            self._scratch.zero_()
        self._scratch_used = 0
        self._scratch_demand = 0
        self._pending = [set() for _ in range(self.n_buckets)]
        for p in self.params:
            self._pending[self.bucket_of[id(p)]].add(id(p))

    def take_scratch(self, numel: int):
        """a zeroed fp32 slice of the per-step scratch, or None (first step / not enough room: the caller allocates)"""
        numel = (int(numel) + 63) // 64 * 64
        self._scratch_demand += numel
        if self._scratch is None or self._scratch_used + numel > self._scratch.numel():
            return None
        v = self._scratch.narrow(0, self._scratch_used, numel)
        self._scratch_used += numel
        return v

    def mark_ready(self, p):
        """called by the backward ops right after p's gradient has been enqueued into its arena slot"""
        if self._pending is None:
            return
        b = self.bucket_of[id(p)]
        s = self._pending[b]
        s.discard(id(p))
        if not s and self.on_bucket_ready is not None:
            self.on_bucket_ready(b)

    def bucket(self, b: int) -> torch.Tensor:
        lo, hi = self.bucket_range[b]
        return self.flat.narrow(0, lo, hi - lo)


_ARENA_ATTR = "_e2e_grad_arena"


def attach_arena(arena: Optional[GradArena]):
    """marks the arena's parameters so that the backward ops write their weight gradients straight into it"""
    if arena is None:
        return
    for p in arena.params:
        setattr(p, _ARENA_ATTR, arena)


def arena_of(p) -> Optional[GradArena]:
    return getattr(p, _ARENA_ATTR, None)


class FusedSGD(torch.optim.Optimizer):
    """SGD(momentum, nesterov, weight decay) + gradient-norm clipping + DSFF mask application in three
    multi-tensor CUDA launches.  `max_norm`: clip threshold of the reference loop (12, :560); None disables the
    clipping stage.  `masks`: optional {parameter: fp32 mask tensor} (set by `set_masks`, normally from the
    drop-in Masking: `opt.set_masks_from(masking)`).  `grad_scale`: multiplies every gradient first (1 / world
    size for a summed all-reduce, or a loss scaler's inverse)."""

    def __init__(self, params, lr=1e-2, momentum=0.99, weight_decay=3e-5, nesterov=True, max_norm=12.0, grad_scale=1.0,
                 loss_scale=None, growth_interval=2000, backoff_factor=0.5, growth_factor=2.0):
        defaults = dict(lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=nesterov, dampening=0)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedSGD: one parameter group (the reference trainer's setup)")
        self.max_norm = max_norm
        self.grad_scale = float(grad_scale)
        self._masks: Dict[int, torch.Tensor] = {}
        self._table = None
        self._table_sig = None
        self._hyper = None
        self._hyper_host = None
        self._norm_coef = None
        self._partial = None
        self.last_launches = 0
        # dynamic loss scale on the device (fp16 precision): the GradScaler of the reference loop, :553-562
        self._scaler_init = None if loss_scale is None else (float(loss_scale), 0.0, float(growth_interval),
                                                             float(backoff_factor), float(growth_factor))
        self._scaler = None

    # -------------------------------------------------------------- configuration
    def set_masks(self, masks: Dict[torch.Tensor, torch.Tensor]):
        self._masks = {id(p): m for p, m in masks.items()}
        self._table_sig = None

    def set_masks_from(self, masking):
        """masks of the drop-in Masking (same storage as Masking.masks[name]: prune / regrow writes are seen)"""
        m = {}
        for module in masking.modules:
            for name, p in module.named_parameters():
                if name in masking.masks:
                    m[p] = masking.masks[name]
        self.set_masks(m)

    # -------------------------------------------------------------- internals
    def _params(self) -> List[torch.Tensor]:
        return [p for p in self.param_groups[0]['params'] if p.grad is not None]

    def _ensure(self, ps):
        dev = ps[0].device
        for p in ps:
            st = self.state[p]
            if 'momentum_buffer' not in st:
                st['momentum_buffer'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        sig = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]['momentum_buffer'].data_ptr(),
                     self._masks[id(p)].data_ptr() if id(p) in self._masks else 0) for p in ps)
        if sig != self._table_sig:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedSGD: parameter / gradient / momentum / mask storage changed inside a CUDA graph "
                                   "capture; run one eager step first so that the pointer table is final")
            arr = (_lib.SgdTensor * len(ps))()
            mx = 0
            for i, p in enumerate(ps):
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise ValueError("FusedSGD: parameters and gradients must be contiguous fp32")
                m = self._masks.get(id(p))
                if m is not None and (m.shape != p.shape or m.dtype != torch.float32 or not m.is_contiguous()):
                    raise ValueError("FusedSGD: mask must be a contiguous fp32 tensor shaped like its parameter")
                arr[i].p, arr[i].g, arr[i].mom = p.data_ptr(), p.grad.data_ptr(), self.state[p]['momentum_buffer'].data_ptr()
                arr[i].mask = 0 if m is None else m.data_ptr()
                arr[i].numel = p.numel()
                mx = max(mx, p.numel())
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            # old tables are kept alive: a captured graph may still point at them
            self.__dict__.setdefault("_old_tables", []).append(self._table)
            self._table = host.to(dev)
            self._table_sig, self._n, self._max_numel = sig, len(ps), mx
            n_part = int(_lib.load().e2e_sgd_partial_count(len(ps), mx))
            self._partial = torch.empty(max(n_part, 1), dtype=torch.float32, device=dev)
        if self._hyper is None:
            self._hyper = torch.zeros(5, dtype=torch.float32, device=dev)
            self._hyper_host = torch.zeros(5, dtype=torch.float32).pin_memory()
            self._norm_coef = torch.zeros(4, dtype=torch.float32, device=dev)

    def sync_hyper(self):
        """param_groups -> the device hyper-parameter array (call before replaying a graph that captured step())"""
        g = self.param_groups[0]
        if self._hyper is None:
            return
        h = self._hyper_host
        h[0], h[1], h[2] = float(g['lr']), float(g['momentum']), float(g['weight_decay'])
        h[3] = float(self.max_norm) if self.max_norm is not None else 0.0
        h[4] = self.grad_scale
        self._hyper.copy_(h, non_blocking=True)

    def enable_loss_scale(self, init_scale=65536.0, growth_interval=2000, backoff_factor=0.5, growth_factor=2.0):
        """GradScaler defaults (torch.cuda.amp.GradScaler(): 2**16, x2 every 2000 clean steps, x0.5 on inf / nan)"""
        self._scaler_init = (float(init_scale), 0.0, float(growth_interval), float(backoff_factor), float(growth_factor))
        self._scaler = None

    def loss_scale(self, device=None) -> Optional[torch.Tensor]:
        """device scalar (a view of the scaler state) the loss is multiplied by before backward(); None = no scaling"""
        if self._scaler_init is None:
            return None
        if self._scaler is None:
            dev = device if device is not None else self.param_groups[0]['params'][0].device
            self._scaler = torch.tensor(self._scaler_init, dtype=torch.float32, device=dev)
        return self._scaler[0]

    @property
    def total_norm(self) -> Optional[torch.Tensor]:
        """device scalar: gradient norm of the last step (what clip_grad_norm_ returns)"""
        return None if self._norm_coef is None else self._norm_coef[0]

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedSGD.step: closures are not used by the reference loop")
        ps = self._params()
        if not ps:
            return None
        if not ps[0].is_cuda:
            raise _lib.E2EError("FusedSGD runs on CUDA only (no CPU fallback)")
        lib = _lib.load()
        from . import ops
        ops.side_join()                  # weight gradients still running on ops' side stream (no-op when there are none)
        self._ensure(ps)
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        st = _lib.stream_ptr()
        if self._scaler_init is not None:
            self.loss_scale(ps[0].device)
        _lib.check(lib.e2e_sgd_clip_coef(_p(self._table), self._n, self._max_numel, _p(self._hyper), _p(self._scaler),
                                         _p(self._partial), _p(self._norm_coef), st), "sgd_clip_coef")
        _lib.check(lib.e2e_sgd_update(_p(self._table), self._n, self._max_numel, _p(self._hyper), _p(self._norm_coef),
                                      1 if self.param_groups[0]['nesterov'] else 0, st), "sgd_update")
        ops.bump_weight_epoch()          # weights changed behind torch's version counters
        return None

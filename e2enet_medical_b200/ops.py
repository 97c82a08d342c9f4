"""torch.autograd glue over the C ABI (include/e2enet_b200.h).

Activations travel between ops in the "C8" layout: bf16 tensors of shape
(B, C/8, D, H, W, 8).  Every op below enqueues hand-written CUDA kernels from
libe2enet_b200.so on torch's current stream; torch is used for memory, streams and the
autograd tape only.  There is no eager / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from .plans import GemmPlan, SegHeadPlan, ShiftConvPlan, TConvPlan

EPS = 1e-5
# 0: mma.sync gather kernels everywhere; 1: tcgen05/TMA kernel where a layer qualifies
CONFIG = {"impl": 1, "stack3": True, "fuse_pool": True, "fuse_stats": True, "fuse_fanin": True, "wgrad_direct": False,
          "in_bwd_plane": False, "wgrad_side_stream": True}
# A/B switches for measurements (tools/, bench.py): E2E_FUSE_STATS=0 / E2E_FUSE_FANIN=0 / E2E_FUSE_POOL=0 / E2E_STACK3=0
import os as _os
for _k, _e in (("fuse_stats", "E2E_FUSE_STATS"), ("fuse_fanin", "E2E_FUSE_FANIN"), ("fuse_pool", "E2E_FUSE_POOL"),
               ("stack3", "E2E_STACK3"), ("wgrad_direct", "E2E_WGRAD_DIRECT"), ("in_bwd_plane", "E2E_IN_BWD_PLANE"),
               ("wgrad_side_stream", "E2E_WGRAD_SIDE")):
    if _os.environ.get(_e) is not None:
        CONFIG[_k] = _os.environ[_e] not in ("0", "false", "False")
# optional per-launch CUDA-event timing of the GEMM kernels (bench.py roofline): records are
# (kind, start_event, end_event, algorithmic dense FLOPs = 2*M*N*K over real rows/cols only)
PROFILE = {"enabled": False, "records": []}
# debugging aid (tools/debug_grads3.py): when a dict, the network records the gradient arriving at every activation
DEBUG_GRADS = None


def debug_tap(t: torch.Tensor, name: str):
    if DEBUG_GRADS is not None and t.requires_grad:
        t.register_hook(lambda g, name=name: DEBUG_GRADS.__setitem__(name, g.detach().clone()))
    return t


def _plan_flops(plan: GemmPlan, M: int) -> float:
    if "_kn" not in plan._dev:
        import numpy as np
        plan._dev["_kn"] = (int((plan.centoff >= 0).sum()) * plan.n_taps, int((plan.rowoff >= 0).sum()))
    k, n = plan._dev["_kn"]
    return 2.0 * M * k * n * plan.useful          # useful: fraction of (entry, column) pairs that are not masked out


class _Timed(object):
    def __init__(self, kind, flops):
        self.kind, self.flops = kind, flops

    def __enter__(self):
        if PROFILE["enabled"]:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if PROFILE["enabled"]:
            self.e1.record()
            PROFILE["records"].append((self.kind, self.e0, self.e1, self.flops))


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


# ---------------------------------------------------------------------------------------- side stream of the backward pass
# A layer's weight gradient feeds nothing but the optimizer (and the data-parallel all-reduce), while its data gradient is
# on the critical path: dgrad(L) -> InstanceNorm backward(L-1) -> dgrad(L-1) ...  The weight-gradient GEMMs are tensor-bound,
# the InstanceNorm backward kernels HBM-bound and small enough to share an SM with a GEMM CTA, so the weight gradients of a
# training step (trainer-installed gradient arena only: nothing else may touch the slot before the join) are launched on a
# second stream and overlap the norm kernels of the layers below.  Everything the side stream reads (the raw-output gradient
# and the saved layer inputs) is kept referenced until `side_join()` -- the caching allocator would otherwise hand the
# blocks to later main-stream kernels while the side stream still reads them; inside a CUDA-graph capture the fork / join
# become graph edges.
_SIDE = {"stream": {}, "keep": [], "dirty": False}


def side_stream(device) -> torch.cuda.Stream:
    st = _SIDE["stream"].get(str(device))
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE["stream"][str(device)] = st
    return st


class _SideSection(object):
    """with _SideSection(device, *tensors_read): launches inside go to the side stream, ordered after everything enqueued
    on the current stream so far"""

    def __init__(self, device, *keep):
        self.side = side_stream(device)
        _SIDE["keep"].append(keep)

    def __enter__(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.side.wait_event(ev)
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        _SIDE["dirty"] = True

    def __exit__(self, *a):
        self.ctx.__exit__(*a)


def side_pending() -> bool:
    return _SIDE["dirty"]


def side_join():
    """the current stream waits for the side stream; the tensors kept for it are released"""
    if _SIDE["dirty"]:
        cur = torch.cuda.current_stream()
        cur.wait_stream(side_stream(cur.device))
        _SIDE["dirty"] = False
    _SIDE["keep"].clear()


def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.E2EError("%s: tensor is on %s; the E2ENet B200 ops run on CUDA only (no CPU fallback)" % (what, t.device))


# ---------------------------------------------------------------------------------------- gradient fan-in
# An activation of the fusion grid has up to three consumers (the same-scale fusion conv, a transposed conv, a
# seg head; unetpp_d.py:453-483).  Autograd would add their data gradients with one bf16 add launch per extra
# consumer (28 per step, 1.4 ms).  Instead the FIRST consumer whose backward runs stores its gradient into a fresh
# buffer and hands it to autograd; later consumers ADD theirs into the same buffer inside the data-gradient
# epilogue (e2e_gemm_t.accumulate) and return no gradient.  The producer's backward retires the entry.
_FANIN = {}            # data_ptr of the C8 activation -> gradient buffer of the backward pass in flight


def fanin_reset():
    _FANIN.clear()


def _fanin_retire(ptr):
    if ptr is not None:
        _FANIN.pop(ptr, None)


def _fanin_dst(src: torch.Tensor, zero: bool):
    """(buffer, accumulate?) for the data gradient of C8 activation `src`"""
    if not CONFIG.get("fuse_fanin", True):
        return (torch.zeros_like(src) if zero else torch.empty_like(src)), False
    key = src.data_ptr()
    buf = _FANIN.get(key)
    if buf is not None and buf.shape == src.shape:
        return buf, True
    buf = torch.zeros_like(src) if zero else torch.empty_like(src)
    _FANIN[key] = buf
    return buf, False


def _dgrad_into(groups, weight, mask, srcs_of_gemm, src_grid, iter_grid_of, B, targets, dst_grid, impl, needs_zero):
    """runs the data-gradient GEMM(s) `groups` for the activations `targets`; returns one gradient per target (None
    where the contribution was accumulated into a buffer autograd already holds)."""
    lib = _lib.load()
    dsts, accs = [], []
    for t in targets:
        d, a = _fanin_dst(t, needs_zero)
        dsts.append(d)
        accs.append(a)
    accmask = sum(1 << i for i, a in enumerate(accs) if a)       # per-destination: accumulate or overwrite
    dst_cb = [t.shape[1] for t in targets]
    temps = None
    for group in groups:
        it = iter_grid_of(group[0])
        if min(it) <= 0:
            continue
        if accmask:
            # accumulate needs the tcgen05 epilogue; otherwise the accumulating targets get a temporary + add
            arr = (_lib.GemmParams * len(group))()
            for i, pl in enumerate(group):
                _fill_gemm(arr[i], pl, pack_weights(pl, weight, mask), srcs_of_gemm, src_grid, it, B, dsts, dst_grid, dst_cb, impl)
            if impl == 1 and int(lib.e2e_gather_gemm_on_tcgen05(arr, len(group))):
                run_gemm_chunks(group, weight, mask, srcs_of_gemm, src_grid, it, B, dsts, dst_grid, dst_cb, impl,
                                accumulate=accmask)
            else:
                if temps is None:
                    temps = [(torch.zeros_like(t) if needs_zero else torch.empty_like(t)) if a else d
                             for t, a, d in zip(targets, accs, dsts)]
                run_gemm_chunks(group, weight, mask, srcs_of_gemm, src_grid, it, B, temps, dst_grid, dst_cb, impl)
        else:
            run_gemm_chunks(group, weight, mask, srcs_of_gemm, src_grid, it, B, dsts, dst_grid, dst_cb, impl)
    if temps is not None:
        for d, t, a in zip(dsts, temps, accs):
            if a:
                _lib.check(lib.e2e_add_inplace(_p(d), _p(t), d.numel(), _lib.stream_ptr()), "add_inplace")
    return [None if a else d for d, a in zip(dsts, accs)]


def c8_shape(x: torch.Tensor):
    B, Cb, D, H, W, e = x.shape
    assert e == 8 and x.dtype == _lib.act_dtype() and x.is_contiguous()
    return B, Cb, D, H, W


# ---------------------------------------------------------------------------------------- raw calls
# Packed operands are cached per plan until the weights (or their mask) change.  torch's version
# counters see optimizer / load_state_dict updates; updates made through the C ABI (the drop-in
# Masking writes weights and masks with raw pointers) are announced with bump_weight_epoch().
_WEIGHT_EPOCH = [0]


def bump_weight_epoch():
    _WEIGHT_EPOCH[0] += 1


def set_precision(name: str):
    """"bf16" (default) or "fp16": the 16-bit type of activations, gradients and packed weights from now on (the
    library build that serves the calls, _lib.LIB_PATHS).  Packed operands are re-made on their next use; tensors,
    networks' cached activations and captured CUDA graphs of the other precision must not be reused."""
    if name != _lib.precision():
        _lib.set_precision(name)
        bump_weight_epoch()
        fanin_reset()


class _PackEntry(object):
    """one (plan, weight, mask) -> persistent packed bf16 operand"""
    __slots__ = ("tables", "plan", "wref", "mask", "out", "key", "dead")

    def current_key(self, w):
        m = self.mask
        return (w.data_ptr(), w._version, None if m is None else (m.data_ptr(), m._version), _WEIGHT_EPOCH[0])


_PACK_REG = {}          # device -> {"entries": [...], "sig": tuple, "table": device tensor, "total": int}


def _pack_key(weight, mask):
    return (weight.data_ptr(), weight._version, None if mask is None else (mask.data_ptr(), mask._version),
            _WEIGHT_EPOCH[0])


def _repack_all(device):
    """after an optimizer / Masking step every packed operand is stale: repack all of them in ONE launch"""
    import weakref  # noqa: F401
    reg = _PACK_REG[str(device)]
    live = []
    for e in reg["entries"]:
        w = e.wref()
        if w is not None and not e.dead:
            live.append((e, w))
    reg["entries"] = [e for e, _ in live]
    if not live:
        return
    sig = tuple((w.data_ptr(), 0 if e.mask is None else e.mask.data_ptr(), e.out.data_ptr()) for e, w in live)
    if reg.get("sig") != sig:
        arr = (_lib.PackJob * len(live))()
        total = 0
        for i, (e, w) in enumerate(live):
            j = arr[i]
            j.w, j.mask = w.data_ptr(), (0 if e.mask is None else e.mask.data_ptr())
            j.rowoff, j.centoff, j.tapoff = (e.tables[k].data_ptr() for k in ("rowoff", "centoff", "tapoff"))
            j.n_cent, j.n_taps, j.Npad = e.plan.n_cent, e.plan.n_taps, e.plan.Npad
            j.out, j.item_begin = e.out.data_ptr(), total
            j.emask = e.tables["emask"].data_ptr() if "emask" in e.tables else 0
            j.rclass = e.tables["rclass"].data_ptr() if "rclass" in e.tables else 0
            total += e.plan.n_cent * e.plan.n_taps * e.plan.Npad
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        reg["table"], reg["sig"], reg["total"] = host.to(device), sig, total
    _lib.check(_lib.load().e2e_pack_weights_multi(_p(reg["table"]), len(live), reg["total"], _lib.stream_ptr()),
               "pack_weights_multi")
    for e, w in live:
        e.key = e.current_key(w)


def pack_registry_keepalive(device):
    """everything a captured CUDA graph's pack_weights_multi launch points at: the job table and, through it,
    every packed operand / mask / plan table of the device.  A graph owner holds this so that later registry
    changes (another network on the device, a dead entry) can never free memory the graph still writes."""
    reg = _PACK_REG.get(str(device))
    if reg is None:
        return None
    return (reg.get("table"), [(e.out, e.mask, e.wref(), e.tables) for e in reg["entries"]])


def pack_weights(plan: GemmPlan, weight: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """bf16 packed operand of `plan` for the current value of weight * mask.  Cached until the weight or
    its mask changes; when it does, every registered operand of the device is repacked by one kernel."""
    import weakref
    lib = _lib.load()
    dev = plan.dev(weight.device)
    key = _pack_key(weight, mask)
    e = dev.get("_wpe")
    w0 = e.wref() if e is not None else None
    same = w0 is weight or (w0 is not None and w0.data_ptr() == weight.data_ptr() and w0.shape == weight.shape)
    if e is not None and same and e.mask is mask and not e.dead:
        if e.key != key:
            _repack_all(weight.device)
        return e.out
    # first use of this (plan, weight): allocate the persistent operand, pack it alone, register it
    if e is not None:
        e.dead = True
    w = weight.detach()
    assert w.dtype == torch.float32 and w.is_contiguous()
    if mask is not None:
        assert mask.dtype == torch.float32 and mask.is_contiguous() and mask.shape == w.shape
    out = torch.empty(plan.packed_numel, dtype=_lib.act_dtype(), device=weight.device)
    _lib.check(lib.e2e_pack_weights(_p(w), _p(mask), _p(dev["rowoff"]), _p(dev["centoff"]), _p(dev["tapoff"]),
                                    _p(dev.get("emask")), _p(dev.get("rclass")), plan.n_cent, plan.n_taps, plan.Npad,
                                    _p(out), _lib.stream_ptr()), "pack_weights")
    e = _PackEntry()
    e.tables, e.plan, e.wref, e.mask, e.out, e.key, e.dead = dev, plan, weakref.ref(weight), mask, out, key, False
    dev["_wpe"] = e
    _PACK_REG.setdefault(str(weight.device), {"entries": []})["entries"].append(e)
    return out


def run_gemm(plan: GemmPlan, wpacked: torch.Tensor, srcs: Sequence[torch.Tensor], src_grid, iter_grid, B: int,
             dsts: Sequence[torch.Tensor], dst_grid, dst_cb: Sequence[int], impl: int = 0):
    lib = _lib.load()
    p = _fill_gemm(_lib.GemmParams(), plan, wpacked, srcs, src_grid, iter_grid, B, dsts, dst_grid, dst_cb, impl)
    with _Timed("gemm", _plan_flops(plan, B * iter_grid[0] * iter_grid[1] * iter_grid[2])):
        _lib.check(lib.e2e_gather_gemm(C.byref(p), _lib.stream_ptr()), "gather_gemm")


def run_gemm_chunks(plans: Sequence[GemmPlan], weight: torch.Tensor, mask: Optional[torch.Tensor],
                    srcs: Sequence[torch.Tensor], src_grid, iter_grid, B: int, dsts: Sequence[torch.Tensor], dst_grid,
                    dst_cb: Sequence[int], impl: int = 0, want_stats: bool = False, accumulate: int = 0):
    """the column chunks of one GEMM (plans differ in columns / packed weights only): one launch.
    want_stats: ask the tcgen05 epilogue for the InstanceNorm partial sums of the (single, C8) destination;
    returns (stats tensor [slots][B][2][C], slots) or None when that launch cannot fuse them."""
    lib = _lib.load()
    n = len(plans)
    arr = (_lib.GemmParams * n)()
    flops = 0.0
    for i, pl in enumerate(plans):
        _fill_gemm(arr[i], pl, pack_weights(pl, weight, mask), srcs, src_grid, iter_grid, B, dsts, dst_grid, dst_cb, impl)
        arr[i].accumulate = int(accumulate)
        flops += _plan_flops(pl, B * iter_grid[0] * iter_grid[1] * iter_grid[2])
    stats = None
    if want_stats and impl == 1 and CONFIG.get("fuse_stats", True):
        key = ("_slots", tuple(src_grid), tuple(iter_grid), B, n, str(srcs[0].device))
        slots = plans[0]._dev.get(key)
        if slots is None:
            slots = int(lib.e2e_gather_gemm_stats_slots(arr, n))
            plans[0]._dev[key] = slots
        if slots > 0:
            ctot = 8 * int(dst_cb[0])
            stats = torch.empty((slots, B, 2, ctot), dtype=torch.float32, device=srcs[0].device)
            for i in range(n):
                arr[i].stats = stats.data_ptr()
                arr[i].stats_ctot = ctot
    with _Timed("gemm", flops):
        if n == 1:
            _lib.check(lib.e2e_gather_gemm(C.byref(arr[0]), _lib.stream_ptr()), "gather_gemm")
        else:
            _lib.check(lib.e2e_gather_gemm_multi(arr, n, _lib.stream_ptr()), "gather_gemm_multi")
    return stats


def _fill_gemm(p, plan: GemmPlan, wpacked: torch.Tensor, srcs, src_grid, iter_grid, B, dsts, dst_grid, dst_cb, impl):
    dev = plan.dev(srcs[0].device)
    p.B = B
    p.Di, p.Hi, p.Wi = src_grid
    p.Do, p.Ho, p.Wo = iter_grid
    p.isd, p.ish, p.isw = plan.istride
    p.ivd, p.ivh, p.ivw = plan.ivoff
    p.Dd, p.Hd, p.Wd = dst_grid
    p.osd, p.osh, p.osw = plan.ostride
    p.n_src = len(srcs)
    for i, s in enumerate(srcs):
        p.src[i] = s.data_ptr()
        p.src_cb[i] = s.shape[1]
    p.n_cent = plan.n_cent
    p.cents = dev["cents"].data_ptr()
    p.n_taps = plan.n_taps
    p.taps = dev["taps"].data_ptr()
    p.wpacked = wpacked.data_ptr()
    p.Npad = plan.Npad
    p.cols = dev["cols"].data_ptr()
    p.n_dst = len(dsts)
    for i, d in enumerate(dsts):
        p.dst[i] = d.data_ptr()
        p.dst_cb[i] = dst_cb[i]
    p.out_mode = plan.out_mode
    p.impl = impl
    p.col_bounds = plan.col_bounds
    return p


def _grad_slot(param: torch.Tensor):
    """(arena, fresh view of the parameter's slot in the flat gradient arena) -- or (None, None) when the parameter
    has no arena or already holds a gradient (then autograd must ACCUMULATE, and the arena slot may be that very
    gradient: writing into it first would double it)."""
    arena = getattr(param, "_e2e_grad_arena", None)
    if arena is None or param.grad is not None or not arena.has(param):
        return None, None
    return arena, arena.view(param)


def run_wgrad(plan: GemmPlan, srcs: Sequence[torch.Tensor], src_grid, iter_grid, B: int, grad: torch.Tensor,
              weight_shape, impl: int = 0, out: Optional[torch.Tensor] = None, out_is_zero: bool = False,
              scratch_from=None) -> torch.Tensor:
    """returns the fp32 weight gradient in the reference's parameter layout (written into `out` if given)."""
    lib = _lib.load()
    device = grad.device
    dev = plan.dev(device)
    p = _lib.WgradParams()
    p.B = B
    p.Di, p.Hi, p.Wi = src_grid
    p.Do, p.Ho, p.Wo = iter_grid
    p.isd, p.ish, p.isw = plan.istride
    p.ivd, p.ivh, p.ivw = (0, 0, 0)
    p.n_src = len(srcs)
    for i, s in enumerate(srcs):
        p.src[i] = s.data_ptr()
        p.src_cb[i] = s.shape[1]
    p.n_cent = plan.n_cent
    p.cents = dev["cents"].data_ptr()
    p.n_taps = plan.n_taps
    p.taps = dev["taps"].data_ptr()
    p.grad = grad.data_ptr()
    p.grad_cb = grad.shape[1]
    p.Npad = plan.Npad
    p.impl = impl
    flops = _plan_flops(plan, B * iter_grid[0] * iter_grid[1] * iter_grid[2])
    if impl == 1 and CONFIG.get("wgrad_direct", False) and int(lib.e2e_gather_wgrad_direct_ok(C.byref(p))):
        # (A/B switch, default off) the tcgen05 kernels add their split-K results straight into the gradient in the
        # parameter layout: no packed scratch, no unpack pass -- but the scattered fp32 atomics (stride 9 floats
        # between lanes instead of 8 contiguous floats) cost more L2 atomic transactions than the unpack pass saves.
        # The destination must start at zero: arena slots are zeroed once per step by GradArena.begin_step (passed
        # as `out` with out_is_zero), anything else is zeroed here.
        if out is not None and out_is_zero:
            gw = out
        else:
            gw = out if out is not None else torch.empty(weight_shape, dtype=torch.float32, device=device)
            gw.zero_()
        assert tuple(gw.shape) == tuple(weight_shape) and gw.dtype == torch.float32 and gw.is_contiguous()
        p.grad_out = gw.data_ptr()
        p.rowoff, p.centoff, p.tapoff = (dev[k].data_ptr() for k in ("rowoff", "centoff", "tapoff"))
        with _Timed("wgrad", flops):
            _lib.check(lib.e2e_gather_wgrad(C.byref(p), _lib.stream_ptr()), "gather_wgrad")
        return gw
    dwp = scratch_from.take_scratch(plan.packed_numel) if scratch_from is not None else None
    if dwp is None:
        dwp = torch.zeros(plan.packed_numel, dtype=torch.float32, device=device)
    p.dwp = dwp.data_ptr()
    with _Timed("wgrad", flops):
        _lib.check(lib.e2e_gather_wgrad(C.byref(p), _lib.stream_ptr()), "gather_wgrad")
    gw = out if out is not None else torch.empty(weight_shape, dtype=torch.float32, device=device)   # unpack writes every weight once
    assert tuple(gw.shape) == tuple(weight_shape) and gw.dtype == torch.float32 and gw.is_contiguous()
    _lib.check(lib.e2e_unpack_wgrad(_p(dwp), _p(dev["rowoff"]), _p(dev["centoff"]), _p(dev["tapoff"]), plan.n_cent,
                                    plan.n_taps, plan.Npad, _p(gw), _lib.stream_ptr()), "unpack_wgrad")
    return gw


def nc_to_c8(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, "nc_to_c8")
    x = x.detach().contiguous().float()
    B, Cc = x.shape[:2]
    sp = tuple(x.shape[2:])
    V = 1
    for s in sp:
        V *= s
    y = torch.empty((B, (Cc + 7) // 8) + sp + (8,), dtype=_lib.act_dtype(), device=x.device)
    _lib.check(_lib.load().e2e_nc_to_c8(_p(x), _p(y), B, Cc, V, _lib.stream_ptr()), "nc_to_c8")
    return y


def c8_to_nc(x: torch.Tensor, channels: int) -> torch.Tensor:
    B, Cb, D, H, W = c8_shape(x)
    y = torch.empty((B, channels, D, H, W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().e2e_c8_to_nc(_p(x), _p(y), B, channels, D * H * W, _lib.stream_ptr()), "c8_to_nc")
    return y


def _nchunk(V: int, planes: int) -> int:
    n = max(1, min((V + 2047) // 2048, (148 * 8 + planes - 1) // planes))
    return int(n)


# ---------------------------------------------------------------------------------------- autograd ops
class ToC8(torch.autograd.Function):
    """fp32 NCDHW -> bf16 C8 (differentiable, used at the module boundary)."""

    @staticmethod
    def forward(ctx, x):
        ctx.channels = x.shape[1]
        y = nc_to_c8(x)
        ctx.out_ptr = y.data_ptr()
        return y

    @staticmethod
    def backward(ctx, dy):
        _fanin_retire(ctx.out_ptr)
        return c8_to_nc(dy.contiguous(), ctx.channels)


class FromC8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, channels):
        return c8_to_nc(x, channels)

    @staticmethod
    def backward(ctx, dy):
        return nc_to_c8(dy), None


class ShiftDepth(torch.autograd.Function):
    """torch_shift.forward (unetpp_d.py:45-59) as a stand-alone op on a plain NCDHW CUDA tensor (fp32 / fp16 /
    bf16): y[b,c,d] = x[b,c,d - s_c], zero fill.  The network never calls it (the shift is folded into the conv's
    operand fetch); it backs direct calls of the exported `torch_shift` module."""

    @staticmethod
    def forward(ctx, x, shift_size):
        _need_cuda(x, "shift_depth")
        if x.dim() != 5 or x.element_size() not in (2, 4):
            raise ValueError("shift_depth: expected a 5-D tensor of 2- or 4-byte elements, got %s %s"
                             % (tuple(x.shape), x.dtype))
        ctx.shift_size = int(shift_size)
        return ShiftDepth._run(x.contiguous(), ctx.shift_size, 1)

    @staticmethod
    def _run(x, shift_size, sign):
        B, Cc, D, H, W = x.shape
        y = torch.empty_like(x)
        _lib.check(_lib.load().e2e_shift_depth(_p(x), _p(y), x.element_size(), B, Cc, D, H * W, shift_size, sign,
                                               _lib.stream_ptr()), "shift_depth")
        return y

    @staticmethod
    def backward(ctx, dy):
        return ShiftDepth._run(dy.contiguous(), ctx.shift_size, -1), None


class ShiftConvINLReLU(torch.autograd.Function):
    """depth-shift + Conv3d(1,3,3) + InstanceNorm3d(affine) + LeakyReLU over a virtual concat
    of C8 sources (reference: ConvDropoutNormNonlin.forward, unetpp_d.py:102-111).
    The conv bias is carried for state_dict/grad parity but not added: InstanceNorm removes
    any per-channel constant exactly."""

    @staticmethod
    def forward(ctx, plan: ShiftConvPlan, slope: float, weight, bias, gamma, beta, mask, pool_k, *srcs):
        """pool_k: None, or the (kd, kh, kw) of the MaxPool3d that consumes this activation: then the pooled
        tensor is produced by the same pass (returns (y, y_pooled)) and its gradient is folded into the norm
        backward."""
        lib = _lib.load()
        for s in srcs:
            _need_cuda(s, "shiftconv")
        B, _, D, H, W = c8_shape(srcs[0])
        Do, Ho, Wo = plan.out_grid(D, H, W)
        dev = srcs[0].device
        Cb = plan.cout // 8
        impl = CONFIG["impl"]
        raw = torch.empty((B, Cb, Do, Ho, Wo, 8), dtype=_lib.act_dtype(), device=dev)
        # the conv epilogue also reduces the InstanceNorm sums of the values it stores (no separate statistics pass)
        if impl == 1 and plan.fwd3 is not None and CONFIG.get("stack3", True):
            # narrow layer: kw-stacked tcgen05 kernel (N = 3 x Cout per MMA)
            stats = run_gemm_chunks([plan.fwd3], weight, mask, srcs, (D, H, W), (Do, Ho, Wo), B, [raw], (Do, Ho, Wo), [Cb],
                                    impl, want_stats=True)
        else:
            stats = run_gemm_chunks(plan.fwd_chunks, weight, mask, srcs, (D, H, W), (Do, Ho, Wo), B, [raw], (Do, Ho, Wo),
                                    [Cb], impl, want_stats=True)
        V = Do * Ho * Wo
        nch = _nchunk(V, B * Cb)
        mean = torch.empty(B * Cb * 8, dtype=torch.float32, device=dev)
        rstd = torch.empty_like(mean)
        if stats is None:
            partial = torch.empty(B * Cb * nch * 16, dtype=torch.float32, device=dev)
            _lib.check(lib.e2e_in_stats(_p(raw), B, Cb, V, EPS, _p(partial), nch, _p(mean), _p(rstd), _lib.stream_ptr()),
                       "in_stats")
        y = torch.empty_like(raw)
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        ctx.plan, ctx.slope, ctx.mask, ctx.grid = plan, slope, mask, (B, D, H, W, Do, Ho, Wo)
        ctx.bias = bias                 # not a saved tensor: only its identity (gradient slot) is needed
        ctx.pool_k = None
        if pool_k is not None:
            kd, kh, kw = (int(v) for v in pool_k)
            yp = torch.empty((B, Cb, Do // kd, Ho // kh, Wo // kw, 8), dtype=_lib.act_dtype(), device=dev)
            am = torch.empty(yp.shape, dtype=torch.uint8, device=dev)
            if stats is not None:       # mean / rstd are formed from the epilogue's slots inside the same launch
                _lib.check(lib.e2e_in_apply_pool_from_slots(_p(raw), _p(stats), stats.shape[0], EPS, _p(g32), _p(b32), slope,
                                                            B, Cb, Do, Ho, Wo, kd, kh, kw, _p(mean), _p(rstd), _p(y), _p(yp),
                                                            _p(am), _lib.stream_ptr()), "in_apply_pool_from_slots")
            else:
                _lib.check(lib.e2e_in_apply_pool(_p(raw), _p(mean), _p(rstd), _p(g32), _p(b32), slope, B, Cb, Do, Ho, Wo,
                                                 kd, kh, kw, _p(y), _p(yp), _p(am), _lib.stream_ptr()), "in_apply_pool")
            ctx.pool_k = (kd, kh, kw)
            ctx.out_ptrs = (y.data_ptr(), yp.data_ptr())
            ctx.save_for_backward(weight, gamma, beta, raw, mean, rstd, am, *srcs)
            return y, yp
        if stats is not None:
            _lib.check(lib.e2e_in_apply_from_slots(_p(raw), _p(stats), stats.shape[0], EPS, _p(g32), _p(b32), slope, B, Cb, V,
                                                   _p(mean), _p(rstd), _p(y), _lib.stream_ptr()), "in_apply_from_slots")
        else:
            _lib.check(lib.e2e_in_apply(_p(raw), _p(mean), _p(rstd), _p(g32), _p(b32), slope, B, Cb, V, _p(y),
                                        _lib.stream_ptr()), "in_apply")
        ctx.out_ptrs = (y.data_ptr(),)
        ctx.save_for_backward(weight, gamma, beta, raw, mean, rstd, *srcs)
        return y

    @staticmethod
    def backward(ctx, dy, dyp=None):
        lib = _lib.load()
        for ptr in ctx.out_ptrs:
            _fanin_retire(ptr)
        plan: ShiftConvPlan = ctx.plan
        weight, gamma, beta, raw, mean, rstd = ctx.saved_tensors[:6]
        bias = ctx.bias
        am = None
        if ctx.pool_k is not None:
            am = ctx.saved_tensors[6]
            srcs = ctx.saved_tensors[7:]
        else:
            srcs = ctx.saved_tensors[6:]
        B, D, H, W, Do, Ho, Wo = ctx.grid
        dev = raw.device
        Cb = plan.cout // 8
        V = Do * Ho * Wo
        impl = CONFIG["impl"]
        dy = dy.contiguous() if dy is not None else None
        nch = _nchunk(V, B * Cb)
        partial = torch.empty(B * Cb * nch * 24, dtype=torch.float32, device=dev)     # pass 1 (16) + pass 2 (8) per chunk
        sums = torch.empty(B * Cb * 16, dtype=torch.float32, device=dev)
        draw = torch.empty_like(raw)
        # small-parameter gradients go straight into the flat gradient arena when the trainer installed one
        slots = {}
        def small(prm, needed=True):
            a, v = _grad_slot(prm) if (prm is not None and needed and prm.dtype == torch.float32) else (None, None)
            if a is not None:
                slots[id(prm)] = (a, prm)
                return v
            return torch.empty(plan.cout, dtype=torch.float32, device=dev)
        dgamma, dbeta = small(gamma), small(beta)
        dbias = small(bias, ctx.needs_input_grad[3])
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        pooled = am is not None and dyp is not None
        if CONFIG.get("in_bwd_plane", False):
            # (A/B switch, default off: measured 0.5 ms SLOWER per step -- the per-plane group barriers cost more than the
            # L2 hits save) plane-resident backward: one kernel, dy / raw are read once from HBM, the second read hits L2
            if dy is None and not pooled:
                dy = torch.zeros_like(raw)
            kd, kh, kw = ctx.pool_k if pooled else (1, 1, 1)
            scratch = torch.empty(int(lib.e2e_in_bwd_scratch_floats(B, Cb, V)), dtype=torch.float32, device=dev)
            _lib.check(lib.e2e_in_bwd_fused(_p(dy), _p(dyp.contiguous()) if pooled else _p(None), _p(am) if pooled else _p(None),
                                            _p(raw), _p(mean), _p(rstd), _p(g32), _p(b32), ctx.slope, B, Cb, Do, Ho, Wo, kd, kh,
                                            kw, _p(scratch), _p(sums), _p(draw), _p(dgamma), _p(dbeta), _p(dbias),
                                            _lib.stream_ptr()), "in_bwd_fused")
        elif pooled:
            kd, kh, kw = ctx.pool_k
            _lib.check(lib.e2e_in_bwd_pool(_p(dy), _p(dyp.contiguous()), _p(am), _p(raw), _p(mean), _p(rstd), _p(g32),
                                           _p(b32), ctx.slope, B, Cb, Do, Ho, Wo, kd, kh, kw, _p(partial), nch, _p(sums),
                                           _p(draw), _p(dgamma), _p(dbeta), _p(dbias), _lib.stream_ptr()), "in_bwd_pool")
        else:
            if dy is None:
                dy = torch.zeros_like(raw)
            _lib.check(lib.e2e_in_bwd(_p(dy), _p(raw), _p(mean), _p(rstd), _p(g32), _p(b32), ctx.slope, B, Cb, V,
                                      _p(partial), nch, _p(sums), _p(draw), _p(dgamma), _p(dbeta), _p(dbias),
                                      _lib.stream_ptr()), "in_bwd")
        for a, prm in slots.values():
            a.mark_ready(prm)
        gw = None
        if ctx.needs_input_grad[2]:
            arena, slot = _grad_slot(weight)
            fresh = arena is not None and arena.zeroed_this_step
            if fresh and CONFIG["wgrad_side_stream"]:
                with _SideSection(draw.device, draw, *srcs):
                    gw = run_wgrad(plan.wgrad, srcs, (D, H, W), (Do, Ho, Wo), B, draw, tuple(weight.shape), impl, out=slot,
                                   out_is_zero=True, scratch_from=arena)
            else:
                gw = run_wgrad(plan.wgrad, srcs, (D, H, W), (Do, Ho, Wo), B, draw, tuple(weight.shape), impl, out=slot,
                               out_is_zero=fresh, scratch_from=arena)
            if arena is not None:
                arena.mark_ready(weight)
        # data gradients of every source
        need = [ctx.needs_input_grad[8 + i] for i in range(len(srcs))]
        dsrcs: List[Optional[torch.Tensor]] = [None] * len(srcs)
        if any(need):
            # column chunks of one GEMM go out as one launch; gradients of activations that already received a
            # contribution from another consumer are accumulated in the epilogue (see _FANIN)
            outs = _dgrad_into(plan.dgrad_groups, weight, ctx.mask, [draw], (Do, Ho, Wo),
                               lambda var: plan.dgrad_iter_grid(var, D, H, W), B, list(srcs), (D, H, W), impl,
                               plan.dgrad_needs_zero)
            dsrcs = [o if n else None for o, n in zip(outs, need)]
        return (None, None, gw, dbias.to(weight.dtype) if ctx.needs_input_grad[3] else None,
                dgamma.to(gamma.dtype), dbeta.to(beta.dtype), None, None, *dsrcs)


class TConv(torch.autograd.Function):
    """ConvTranspose3d(kernel == stride, bias=False) on C8 tensors (unetpp_d.py:521-522)."""

    @staticmethod
    def forward(ctx, plan: TConvPlan, weight, mask, x):
        _need_cuda(x, "tconv")
        B, Cb, D, H, W = c8_shape(x)
        kd, kh, kw = plan.k
        impl = CONFIG["impl"]
        y = torch.empty((B, plan.cout // 8, D * kd, H * kh, W * kw, 8), dtype=_lib.act_dtype(), device=x.device)
        run_gemm_chunks(plan.fwd, weight, mask, [x], (D, H, W), (D, H, W), B, [y], (D * kd, H * kh, W * kw),
                        [plan.cout // 8], impl)
        ctx.plan, ctx.mask = plan, mask
        ctx.out_ptr = y.data_ptr()
        ctx.save_for_backward(weight, x)
        return y

    @staticmethod
    def backward(ctx, dy):
        _fanin_retire(ctx.out_ptr)
        plan: TConvPlan = ctx.plan
        weight, x = ctx.saved_tensors
        B, Cb, D, H, W = c8_shape(x)
        kd, kh, kw = plan.k
        dy = dy.contiguous()
        fine = (D * kd, H * kh, W * kw)
        impl = CONFIG["impl"]
        gw = dx = None
        if ctx.needs_input_grad[1]:
            arena, slot = _grad_slot(weight)
            fresh = arena is not None and arena.zeroed_this_step
            if fresh and CONFIG["wgrad_side_stream"]:
                with _SideSection(dy.device, dy, x):
                    gw = run_wgrad(plan.wgrad, [dy], fine, (D, H, W), B, x, tuple(weight.shape), impl, out=slot, out_is_zero=True,
                                   scratch_from=arena)
            else:
                gw = run_wgrad(plan.wgrad, [dy], fine, (D, H, W), B, x, tuple(weight.shape), impl, out=slot, out_is_zero=fresh,
                               scratch_from=arena)
            if arena is not None:
                arena.mark_ready(weight)
        if ctx.needs_input_grad[3]:
            dx = _dgrad_into([plan.dgrad], weight, ctx.mask, [dy], fine, lambda var: (D, H, W), B, [x], (D, H, W), impl,
                             False)[0]
        return None, gw, None, dx


class MaxPool(torch.autograd.Function):
    """MaxPool3d(kernel == stride) on C8 tensors (unetpp_d.py:523-524)."""

    @staticmethod
    def forward(ctx, x, k):
        _need_cuda(x, "maxpool")
        B, Cb, D, H, W = c8_shape(x)
        kd, kh, kw = (int(v) for v in k)
        y = torch.empty((B, Cb, D // kd, H // kh, W // kw, 8), dtype=_lib.act_dtype(), device=x.device)
        am = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
        _lib.check(_lib.load().e2e_maxpool_fwd(_p(x), _p(y), _p(am), B * Cb, D, H, W, kd, kh, kw, _lib.stream_ptr()),
                   "maxpool_fwd")
        ctx.k, ctx.shape = (kd, kh, kw), (B, Cb, D, H, W)
        ctx.out_ptr = y.data_ptr()
        ctx.save_for_backward(am)
        return y

    @staticmethod
    def backward(ctx, dy):
        _fanin_retire(ctx.out_ptr)
        (am,) = ctx.saved_tensors
        B, Cb, D, H, W = ctx.shape
        kd, kh, kw = ctx.k
        dx = torch.empty((B, Cb, D, H, W, 8), dtype=_lib.act_dtype(), device=dy.device)
        _lib.check(_lib.load().e2e_maxpool_bwd(_p(dy.contiguous()), _p(am), _p(dx), B * Cb, D, H, W, kd, kh, kw,
                                               _lib.stream_ptr()), "maxpool_bwd")
        return dx, None


class SegHead(torch.autograd.Function):
    """Conv3d(C, ncls, 1, bias=False): C8 bf16 in, fp32 NCDHW logits out (unetpp_d.py:394-401)."""

    @staticmethod
    def forward(ctx, plan: SegHeadPlan, weight, x):
        _need_cuda(x, "seghead")
        B, Cb, D, H, W = c8_shape(x)
        wp = pack_weights(plan.fwd, weight, None)
        y = torch.empty((B, plan.ncls, D, H, W), dtype=torch.float32, device=x.device)
        run_gemm(plan.fwd, wp, [x], (D, H, W), (D, H, W), B, [y], (D, H, W), [plan.ncls], 0)
        ctx.plan = plan
        ctx.save_for_backward(weight, x)
        return y

    @staticmethod
    def backward(ctx, dy):
        plan: SegHeadPlan = ctx.plan
        weight, x = ctx.saved_tensors
        B, Cb, D, H, W = c8_shape(x)
        g = nc_to_c8(dy)
        impl = CONFIG["impl"]
        gw = dx = None
        if ctx.needs_input_grad[1]:
            arena, slot = _grad_slot(weight)
            gw = run_wgrad(plan.fwd, [x], (D, H, W), (D, H, W), B, g, tuple(weight.shape), impl, out=slot, out_is_zero=arena is not None and arena.zeroed_this_step,
                           scratch_from=arena)
            if arena is not None:
                arena.mark_ready(weight)
        if ctx.needs_input_grad[2]:
            dx = _dgrad_into([[plan.dgrad]], weight, None, [g], (D, H, W), lambda var: (D, H, W), B, [x], (D, H, W), impl,
                             False)[0]
        return None, gw, dx


class SoftmaxStats(torch.autograd.Function):
    """(logits fp32 (B,C,*sp), target (B,1,*sp) or (B,*sp) with class indices) ->
    S_p (B,C) = sum_v softmax, tp (B,C) = sum_v softmax * onehot, S_y (B,C) = class counts,
    ce_sum () = sum over all voxels of -log softmax[target].  One fused pass forward, one backward
    (csrc/loss.cu); everything else of DC_and_CE_loss is arithmetic on these tiny tensors."""

    @staticmethod
    def forward(ctx, logits, target):
        _need_cuda(logits, "softmax_stats")
        lib = _lib.load()
        x = logits.contiguous().float()
        t = target.detach().contiguous().float()
        B, Cc = x.shape[:2]
        V = x[0, 0].numel()
        if t.numel() != B * V:
            raise ValueError("softmax_stats: target must hold one class index per voxel (got %s for logits %s)"
                             % (tuple(target.shape), tuple(logits.shape)))
        stats = torch.zeros((B, Cc, 3), dtype=torch.float32, device=x.device)
        ce = torch.zeros((), dtype=torch.float32, device=x.device)
        partial = torch.empty(max(1, int(lib.e2e_softmax_stats_partial_count(B, Cc, V))), dtype=torch.float32, device=x.device)
        _lib.check(lib.e2e_softmax_stats_fwd(_p(x), _p(t), B, Cc, V, _p(partial), _p(stats), _p(ce), _lib.stream_ptr()),
                   "softmax_stats_fwd")
        ctx.save_for_backward(x, t)
        sp, tp, sy = stats[..., 0].clone(), stats[..., 1].clone(), stats[..., 2].clone()
        ctx.mark_non_differentiable(sy)
        return sp, tp, sy, ce

    @staticmethod
    def backward(ctx, gsp, gtp, _gsy, gce):
        x, t = ctx.saved_tensors
        lib = _lib.load()
        B, Cc = x.shape[:2]
        V = x[0, 0].numel()
        z = lambda g, shape: (torch.zeros(shape, dtype=torch.float32, device=x.device) if g is None
                              else g.contiguous().float())
        gsp, gtp, gce = z(gsp, (B, Cc)), z(gtp, (B, Cc)), z(gce, ())
        dx = torch.empty_like(x)
        _lib.check(lib.e2e_softmax_stats_bwd(_p(x), _p(t), _p(gsp), _p(gtp), _p(gce), _p(None), B, Cc, V, _p(dx),
                                             _lib.stream_ptr()), "softmax_stats_bwd")
        return dx, None


class DiceCELoss(torch.autograd.Function):
    """DC_and_CE_loss(net_output, target) of the reference trainer's configuration as FOUR launches forward (statistics,
    their fixed-order reduction, the dice / CE formulas on the (B, C) statistics) and ONE backward (the upstream
    gradient -- e.g. a GradScaler's scale times the deep-supervision weight -- is read on the device): no (B, C)-sized
    ATen arithmetic at all.  Reference: dice_loss.py:155-190, 302-359; crossentropy.py:4-11."""

    @staticmethod
    def forward(ctx, logits, target, smooth, do_bg, batch_dice, weight_ce, weight_dice):
        _need_cuda(logits, "dc_ce_loss")
        lib = _lib.load()
        x = logits.contiguous().float()
        t = target.detach().contiguous().float()
        B, Cc = x.shape[:2]
        V = x[0, 0].numel()
        if t.numel() != B * V:
            raise ValueError("dc_ce_loss: target must hold one class index per voxel (got %s for logits %s)"
                             % (tuple(target.shape), tuple(logits.shape)))
        dev = x.device
        buf = torch.zeros(B * Cc * 3 + 1, dtype=torch.float32, device=dev)          # stats | ce_sum (one zero fill)
        stats, ce = buf[:B * Cc * 3], buf[B * Cc * 3:]
        partial = torch.empty(max(1, int(lib.e2e_softmax_stats_partial_count(B, Cc, V))), dtype=torch.float32, device=dev)
        _lib.check(lib.e2e_softmax_stats_fwd(_p(x), _p(t), B, Cc, V, _p(partial), _p(stats), _p(ce), _lib.stream_ptr()),
                   "softmax_stats_fwd")
        out = torch.empty(2 * B * Cc + 2, dtype=torch.float32, device=dev)            # loss | gce | gsp | gtp
        loss, gce, gsp, gtp = out[0:1], out[1:2], out[2:2 + B * Cc], out[2 + B * Cc:]
        _lib.check(lib.e2e_dc_ce_from_stats(_p(stats), _p(ce), B, Cc, B * V, float(smooth), 1 if do_bg else 0,
                                            1 if batch_dice else 0, float(weight_ce), float(weight_dice), _p(loss), _p(gsp),
                                            _p(gtp), _p(gce), _lib.stream_ptr()), "dc_ce_from_stats")
        ctx.save_for_backward(x, t, out)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        x, t, out = ctx.saved_tensors
        lib = _lib.load()
        B, Cc = x.shape[:2]
        V = x[0, 0].numel()
        gce, gsp, gtp = out[1:2], out[2:2 + B * Cc], out[2 + B * Cc:]
        gs = gout.detach().reshape(1).float().contiguous()
        dx = torch.empty_like(x)
        _lib.check(lib.e2e_softmax_stats_bwd(_p(x), _p(t), _p(gsp), _p(gtp), _p(gce), _p(gs), B, Cc, V, _p(dx),
                                             _lib.stream_ptr()), "softmax_stats_bwd")
        return dx, None, None, None, None, None, None

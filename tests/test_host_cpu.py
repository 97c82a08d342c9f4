"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares, and the plan builder (shift folding, virtual concat, stride-parity dgrad variants,
transposed convs, seg heads) reproduces the oracle when its tables are executed by a numpy
interpreter of the documented kernel semantics (tests/plan_interp.py).  No GPU needed."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import plan_interp as pi
from oracle import network as onet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_declared_symbols():
    from e2enet_medical_b200 import _lib, build
    build.build_library()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "e2enet_b200.h")).read()
    declared = set(re.findall(r"\b(e2e_[a-z0-9_]+)\s*\(", header))
    assert declared, "no entry points found in the header"
    assert declared == set(_lib.SIGNATURES.keys())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.e2e_version() >= 100
    assert lib.e2e_launch_count() == 0          # nothing may launch on import / load
    assert lib.e2e_precision() == b"bf16" and _lib.precision() == "bf16" and _lib.act_dtype() == torch.bfloat16
    # the fp16 build (same sources, -DE2E_FP16) exports the same ABI
    lib16 = _lib.load("fp16")
    assert lib16 is not lib and lib16.e2e_precision() == b"fp16"
    for name in declared:
        assert hasattr(lib16, name), name
    with pytest.raises(ValueError):
        _lib.set_precision("fp8")


def test_product_has_no_cpu_fallback():
    from e2enet_medical_b200.network_architecture.unetpp_d import ConvDropoutNormNonlin
    from torch import nn
    blk = ConvDropoutNormNonlin(8, 8, nn.Conv3d, {'kernel_size': (1, 3, 3), 'stride': 1, 'padding': (0, 1, 1),
                                                  'dilation': 1, 'bias': True}, nn.InstanceNorm3d,
                                {'eps': 1e-5, 'affine': True}, nn.Dropout3d, {'p': 0, 'inplace': True})
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, 8, 2, 4, 4))
    # and the product never imports the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "e2enet_medical_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.parametrize("src,cout,stride,spatial", [
    ([12], 8, (1, 1, 1), (5, 4, 5)),
    ([8, 8], 8, (1, 1, 1), (6, 3, 4)),         # virtual concat, groups of 4 straddle the 8-blocks
    ([1], 8, (1, 1, 1), (6, 4, 4)),
    ([4], 8, (1, 1, 1), (6, 3, 3)),
    ([8], 8, (1, 2, 2), (5, 6, 4)),
    ([16], 8, (2, 2, 2), (6, 4, 6)),
    ([8], 8, (2, 2, 2), (5, 5, 3)),            # odd sizes under stride
])
def test_shiftconv_plans_reproduce_oracle(src, cout, stride, spatial):
    from e2enet_medical_b200.plans import build_shiftconv_plan
    rs = np.random.RandomState(1)
    cin = sum(src)
    B = 1
    xs = [rs.standard_normal((B, c) + spatial) for c in src]
    w = rs.standard_normal((cout, cin, 1, 3, 3))
    plan = build_shiftconv_plan(src, cout, stride)
    D, H, W = spatial
    Do, Ho, Wo = plan.out_grid(D, H, W)
    # forward
    tx = [torch.from_numpy(a).requires_grad_(True) for a in xs]
    tw = torch.from_numpy(w).requires_grad_(True)
    ref = F.conv3d(onet.shift_depth(torch.cat(tx, 1)), tw, None, stride=stride, padding=(0, 1, 1))
    assert tuple(ref.shape[2:]) == (Do, Ho, Wo)
    raw = np.zeros((B, cout // 8, Do, Ho, Wo, 8))
    for ch in plan.fwd_chunks:
        pi.gemm(ch, pi.pack(ch, w), [pi.to_c8(a) for a in xs], (D, H, W), (Do, Ho, Wo), B, [raw], (Do, Ho, Wo))
    np.testing.assert_allclose(pi.from_c8(raw, cout), ref.detach().numpy(), atol=1e-9)
    # backward
    gy = rs.standard_normal(tuple(ref.shape))
    (ref * torch.from_numpy(gy)).sum().backward()
    g8 = pi.to_c8(gy)
    gw = pi.wgrad(plan.wgrad, [pi.to_c8(a) for a in xs], (D, H, W), (Do, Ho, Wo), B, g8, w.shape)
    np.testing.assert_allclose(gw, tw.grad.numpy(), atol=1e-9)
    # strided convs: the variants write only voxels that receive a contribution; the caller zeroes dx
    outs = [np.full(pi.to_c8(a).shape, 0.0 if plan.dgrad_needs_zero else np.nan) for a in xs]
    sd, sh, sw = stride
    for var in plan.dgrad:
        it = plan.dgrad_iter_grid(var, D, H, W)
        if min(it) <= 0:
            continue
        pi.gemm(var, pi.pack(var, w), [g8], (Do, Ho, Wo), it, B, outs, (D, H, W))
    for o, t, c in zip(outs, tx, src):
        got = pi.from_c8(o, c)
        assert not np.isnan(got).any(), "dgrad variants must write every element of every source gradient"
        np.testing.assert_allclose(got, t.grad.numpy(), atol=1e-9)


def test_shiftconv_plan_applies_mask_on_pack():
    from e2enet_medical_b200.plans import build_shiftconv_plan
    rs = np.random.RandomState(2)
    w = rs.standard_normal((8, 16, 1, 3, 3))
    m = (rs.rand(8, 16, 1, 1, 1) < 0.3).astype(np.float64) * np.ones_like(w)
    plan = build_shiftconv_plan([16], 8)
    np.testing.assert_array_equal(pi.pack(plan.fwd, w, m), pi.pack(plan.fwd, w * m))


@pytest.mark.parametrize("cin,cout,k,spatial", [(16, 8, (1, 2, 2), (3, 3, 4)), (8, 16, (2, 2, 2), (2, 3, 2)),
                                               (8, 8, (1, 1, 1), (2, 2, 3)),
                                               (8, 40, (2, 2, 2), (1, 2, 2))])   # N = 320 -> two column chunks
def test_tconv_plans_reproduce_torch(cin, cout, k, spatial):
    from e2enet_medical_b200.plans import build_tconv_plan
    rs = np.random.RandomState(4)
    B = 1
    x = rs.standard_normal((B, cin) + spatial)
    w = rs.standard_normal((cin, cout) + k)
    tx = torch.from_numpy(x).requires_grad_(True)
    tw = torch.from_numpy(w).requires_grad_(True)
    ref = F.conv_transpose3d(tx, tw, stride=k)
    plan = build_tconv_plan(cin, cout, k)
    fine = tuple(s * kk for s, kk in zip(spatial, k))
    y = np.full((B, cout // 8) + fine + (8,), np.nan)
    for ch in plan.fwd:
        assert ch.Npad <= 256
        pi.gemm(ch, pi.pack(ch, w), [pi.to_c8(x)], spatial, spatial, B, [y], fine)
    np.testing.assert_allclose(pi.from_c8(y, cout), ref.detach().numpy(), atol=1e-9)
    gy = rs.standard_normal(tuple(ref.shape))
    (ref * torch.from_numpy(gy)).sum().backward()
    dx = np.full(pi.to_c8(x).shape, np.nan)
    for ch in plan.dgrad:
        pi.gemm(ch, pi.pack(ch, w), [pi.to_c8(gy)], fine, spatial, B, [dx], spatial)
    np.testing.assert_allclose(pi.from_c8(dx, cin), tx.grad.numpy(), atol=1e-9)
    gw = pi.wgrad(plan.wgrad, [pi.to_c8(gy)], fine, spatial, B, pi.to_c8(x), w.shape)
    np.testing.assert_allclose(gw, tw.grad.numpy(), atol=1e-9)


def test_seghead_plans_reproduce_torch():
    from e2enet_medical_b200.plans import build_seghead_plan
    rs = np.random.RandomState(6)
    cin, ncls, spatial, B = 16, 14, (2, 3, 3), 1
    x = rs.standard_normal((B, cin) + spatial)
    w = rs.standard_normal((ncls, cin, 1, 1, 1))
    tx = torch.from_numpy(x).requires_grad_(True)
    tw = torch.from_numpy(w).requires_grad_(True)
    ref = F.conv3d(tx, tw)
    plan = build_seghead_plan(cin, ncls)
    y = pi.gemm_planar(plan.fwd, pi.pack(plan.fwd, w), [pi.to_c8(x)], spatial, spatial, B, ncls)
    np.testing.assert_allclose(y, ref.detach().numpy(), atol=1e-9)
    gy = rs.standard_normal(tuple(ref.shape))
    (ref * torch.from_numpy(gy)).sum().backward()
    g8 = pi.to_c8(gy)
    dx = np.full(pi.to_c8(x).shape, np.nan)
    pi.gemm(plan.dgrad, pi.pack(plan.dgrad, w), [g8], spatial, spatial, B, [dx], spatial)
    np.testing.assert_allclose(pi.from_c8(dx, cin), tx.grad.numpy(), atol=1e-9)
    gw = pi.wgrad(plan.fwd, [pi.to_c8(x)], spatial, spatial, B, g8, w.shape)
    np.testing.assert_allclose(gw, tw.grad.numpy(), atol=1e-9)


def test_real_layer_plans_are_consistent():
    """every weight of every real layer shape is packed exactly once in fwd, and the dgrad
    variants cover every input channel exactly once per stride parity"""
    from e2enet_medical_b200.plans import build_shiftconv_plan
    for src, cout, stride in (([48, 48], 48, (1, 1, 1)), ([96, 96, 48], 96, (1, 1, 1)), ([192, 192, 96], 192, (1, 1, 1)),
                              ([320, 320, 192], 320, (1, 1, 1)), ([320, 320, 320], 320, (1, 1, 1)),
                              ([48], 96, (1, 2, 2)), ([192], 320, (2, 2, 2)), ([1], 48, (1, 1, 1))):
        plan = build_shiftconv_plan(src, cout, stride)
        cin = sum(src)
        f = plan.fwd
        co = f.centoff[f.centoff >= 0]
        # halo form: one entry per channel (x 9 taps in the kernel); point form: one entry per (tap, channel)
        want = [c * 9 for c in range(cin)] if f.n_taps == 9 else list(range(cin * 9))
        assert sorted(co.tolist()) == want, (src, cout)
        assert all(ch.Npad <= 256 for ch in plan.fwd_chunks) and sum(
            int((ch.rowoff >= 0).sum()) for ch in plan.fwd_chunks) == cout
        assert f.n_cent % 2 == 0 and f.Npad % 16 == 0
        # data gradient: every (input channel, output parity) is produced by exactly one column, over all variants
        # (stride 2 x 2: ONE GEMM whose column groups are the four parities, stored at row / column offsets (ph, pw))
        npar = stride[1] * stride[2]
        seen = []
        for var in plan.dgrad:
            assert tuple(int(v) for v in var.iter_off) == (0, 0, 0) or npar == 1 or var.emask is None
            for q, (dst, blk, chmask, od, oh, ow) in enumerate(var.cols):
                if dst < 0:
                    continue
                base = int(np.cumsum([0] + src)[dst]) + blk * 8
                par = (int(var.iter_off[1]) + int(oh), int(var.iter_off[2]) + int(ow)) if npar > 1 else (0, 0)
                seen += [(base + j, par) for j in range(8) if chmask & (1 << j)]
        assert len(seen) == len(set(seen))
        want = {(c, (ph, pw)) for c in range(cin) for ph in range(stride[1]) for pw in range(stride[2])}
        assert set(seen) == want, (src, stride)
        if stride[1] == 2 and stride[2] == 2:
            v0 = plan.dgrad[0]
            assert v0.emask is not None and abs(v0.useful - 9 / 16) < 1e-12 and v0.n_cent >= 4 * (cout // 8)
            # an entry that reads d(raw) one row / column further only feeds the odd output rows / columns
            assert sorted(set(int(m) for m in v0.emask if m)) == [8, 10, 12, 15]


@pytest.mark.parametrize("src,cout,spatial", [([72], 8, (3, 4, 6)), ([40, 40], 16, (4, 3, 5)),
                                              ([1], 48, (5, 3, 4)),       # the 1-channel input layer: one K pair
                                              ([4], 16, (6, 3, 3)),       # BraTS input: four single-channel shift groups
                                              ([20], 8, (6, 4, 3))])      # three K pairs, ragged last channel block
def test_kw_stacked_forward_plan(src, cout, spatial):
    """fwd3 (narrow layers, tcgen05 only): column kw*Np + n of the 3-tap (kh) GEMM holds
    D'[v] = sum_{kh,c} x~[v + (kh-1) rows][c] W[n,c,kh,kw]; the kernel's epilogue adds the three column
    groups with W shifts -1 / 0 / +1.  Interpreting the plan on CPU and applying that shift-sum must
    reproduce the convolution."""
    from e2enet_medical_b200.plans import build_shiftconv_plan
    rs = np.random.RandomState(5)
    cin, B = sum(src), 1
    D, H, W = spatial
    xs = [rs.standard_normal((B, c) + spatial) for c in src]
    w = rs.standard_normal((cout, cin, 1, 3, 3))
    plan = build_shiftconv_plan(src, cout, (1, 1, 1))
    p3 = plan.fwd3
    assert p3 is not None and p3.n_taps == 3 and p3.Npad % 48 == 0
    npo = p3.Npad // 3
    ref = F.conv3d(onet.shift_depth(torch.cat([torch.from_numpy(a) for a in xs], 1)), torch.from_numpy(w), None,
                   padding=(0, 1, 1)).numpy()
    dprime = pi.gemm_planar(p3, pi.pack(p3, w), [pi.to_c8(a) for a in xs], (D, H, W), (D, H, W), B, p3.Npad)
    out = np.zeros((B, npo, D, H, W))
    for kw in range(3):
        g = dprime[:, kw * npo:(kw + 1) * npo]
        sh = kw - 1                                  # out[w] += D'_kw[w + kw - 1]
        if sh == 0:
            out += g
        elif sh < 0:
            out[..., 1:] += g[..., :-1]
        else:
            out[..., :-1] += g[..., 1:]
    np.testing.assert_allclose(out[:, :cout], ref, atol=1e-9)
    assert not out[:, cout:].any()


def test_product_state_dict_matches_reference_inventory():
    """drop-in contract: the product network built with the trainer's positional arguments has
    exactly the reference's state_dict keys / shapes / registration order (golden dumped from the unmodified
    reference), for all three training configs.  Construction needs no GPU."""
    import json
    from e2enet_medical_b200.training import POOLS, build_network
    inv = json.load(open(os.path.join(ROOT, "tests", "golden", "param_inventory.json")))
    for tag, in_ch, ncls, patch in (("btcv", 1, 14, (64, 160, 160)), ("brats", 4, 4, (128, 128, 128)),
                                    ("hippo", 1, 3, (40, 56, 40))):
        net = build_network(in_ch, ncls, POOLS[tag], patch)
        sd = net.state_dict()
        assert [[k, list(v.shape)] for k, v in sd.items()] == inv[tag], tag
        assert [k for k, _ in net.named_parameters()] == inv[tag + "_named_parameters"], tag
        assert all(v.dtype == torch.float32 for v in sd.values())
    # the Masking selection rule of the reference (core_channel.py:320-336): 35 masked tensors for config 2
    net = build_network(1, 14, POOLS["btcv"], (64, 160, 160))
    sel = [n for n, p in net.named_parameters()
           if (('loc' in n and 'context' not in n) or 'up' in n) and 'bias' not in n and 'instnorm' not in n]
    assert len(sel) == 35 and sum(dict(net.named_parameters())[n].numel() for n in sel) == 17975040


def test_slab_plan_properties():
    """slab ownership plan of sharded sliding-window inference: for any volume / patch / step / rank count the
    owned x-slabs tile [0, X) exactly once, every tile is predicted by exactly one rank, a rank's tiles start
    inside or after its own slab, and every plane a rank touches beyond its slab is owned by a LATER rank
    (contributions only flow forward, which is what _slab_exchange implements)."""
    from e2enet_medical_b200.network_architecture.neural_network import SegmentationNetwork as S
    rs = np.random.RandomState(0)
    for _ in range(300):
        patch = tuple(int(v) for v in rs.randint(4, 40, 3))
        vol = tuple(int(p + rs.randint(0, 4 * p)) for p in patch)
        step = float(rs.choice([0.25, 0.5, 0.75, 1.0]))
        world = int(rs.randint(1, 12))
        steps = S._compute_steps_for_sliding_window(patch, vol, step)
        tiles = [(a, b, c) for a in steps[0] for b in steps[1] for c in steps[2]]
        cut, bx, hi = S._slab_plan(tiles, patch[0], vol[0], world)
        assert cut[0] == 0 and cut[-1] == len(tiles) and all(cut[i] <= cut[i + 1] for i in range(world))
        assert bx[0] == 0 and bx[-1] == vol[0] and all(bx[i] <= bx[i + 1] for i in range(world)), (bx, vol)
        for r in range(world):
            mine = tiles[cut[r]:cut[r + 1]]
            for (a, _, _) in mine:
                assert a >= bx[r], "a rank never contributes to planes owned by an earlier rank"
                assert a + patch[0] <= max(hi[r], bx[r + 1])
            if mine:
                assert hi[r] == mine[-1][0] + patch[0] and hi[r] <= vol[0]
        # every plane is covered by the tiles of ranks <= its owner (so the owner ends up with the full sum)
        for r in range(world):
            for x in {bx[r], max(bx[r], bx[r + 1] - 1)} if bx[r + 1] > bx[r] else ():
                contrib = [q for q in range(world) for (a, _, _) in tiles[cut[q]:cut[q + 1]] if a <= x < a + patch[0]]
                assert contrib and max(contrib) <= r, (x, r, contrib)


def test_shiftconv_plans_random_geometries():
    """randomised host-plan check (fixed seed): arbitrary source splits, strides, odd grids -- forward (column
    chunks), weight gradient and every data-gradient variant of the plan tables reproduce torch's
    conv3d(shift_depth(cat(x))) through the numpy interpreter."""
    from e2enet_medical_b200.plans import build_shiftconv_plan
    rs = np.random.RandomState(123)
    for trial in range(8):
        nsrc = int(rs.randint(1, 4))
        src = [int(rs.choice([1, 4, 8, 12, 16, 20])) for _ in range(nsrc)]
        cout = int(rs.choice([8, 16, 24]))
        stride = tuple(int(v) for v in rs.choice([1, 2], 3)) if trial % 2 else (1, 1, 1)
        spatial = tuple(int(v) for v in rs.randint(3, 7, 3))
        cin, B = sum(src), 1
        xs = [rs.standard_normal((B, c) + spatial) for c in src]
        w = rs.standard_normal((cout, cin, 1, 3, 3))
        plan = build_shiftconv_plan(src, cout, stride)
        D, H, W = spatial
        Do, Ho, Wo = plan.out_grid(D, H, W)
        tx = [torch.from_numpy(a).requires_grad_(True) for a in xs]
        tw = torch.from_numpy(w).requires_grad_(True)
        ref = F.conv3d(onet.shift_depth(torch.cat(tx, 1)), tw, None, stride=stride, padding=(0, 1, 1))
        x8 = [pi.to_c8(a) for a in xs]
        raw = np.zeros((B, cout // 8, Do, Ho, Wo, 8))
        for ch in plan.fwd_chunks:
            pi.gemm(ch, pi.pack(ch, w), x8, (D, H, W), (Do, Ho, Wo), B, [raw], (Do, Ho, Wo))
        np.testing.assert_allclose(pi.from_c8(raw, cout), ref.detach().numpy(), atol=1e-9, err_msg=str((src, cout, stride)))
        gy = rs.standard_normal(tuple(ref.shape))
        (ref * torch.from_numpy(gy)).sum().backward()
        g8 = pi.to_c8(gy)
        gw = pi.wgrad(plan.wgrad, x8, (D, H, W), (Do, Ho, Wo), B, g8, w.shape)
        np.testing.assert_allclose(gw, tw.grad.numpy(), atol=1e-9, err_msg=str((src, cout, stride)))
        outs = [np.full(a.shape, 0.0 if plan.dgrad_needs_zero else np.nan) for a in x8]
        for grp in plan.dgrad_groups:
            it = plan.dgrad_iter_grid(grp[0], D, H, W)
            if min(it) <= 0:
                continue
            for var in grp:
                pi.gemm(var, pi.pack(var, w), [g8], (Do, Ho, Wo), it, B, outs, (D, H, W))
        for o, t, c in zip(outs, tx, src):
            got = pi.from_c8(o, c)
            assert not np.isnan(got).any(), (src, cout, stride)
            np.testing.assert_allclose(got, t.grad.numpy(), atol=1e-9, err_msg=str((src, cout, stride)))


def test_grad_arena_layout_and_bucket_callbacks():
    """optim.GradArena (host logic of the bucketed data-parallel all-reduce): slots are disjoint, aligned and in
    the given (gradient-ready) order; a bucket's callback fires exactly when its last parameter is marked."""
    import torch
    from e2enet_medical_b200.optim import GradArena
    ps = [torch.nn.Parameter(torch.zeros(s)) for s in ((48, 96, 1, 3, 3), (48,), (48,), (96, 48, 1, 2, 2), (14, 48, 1, 1, 1), (7,))]
    arena = GradArena(ps, n_buckets=3, align=64)
    offs = [arena.offset[id(p)] for p in ps]
    assert offs == sorted(offs) and all(o % 64 == 0 for o in offs)
    for a, b, p in zip(offs, offs[1:] + [arena.total], ps):
        assert b - a >= p.numel()
    assert arena.bucket_range[0][0] == 0 and arena.bucket_range[-1][1] == arena.total
    for (l0, h0), (l1, h1) in zip(arena.bucket_range, arena.bucket_range[1:]):
        assert h0 == l1
    v = arena.view(ps[3])
    v.fill_(2.0)
    assert float(arena.flat.sum()) == 2.0 * ps[3].numel() and v.shape == ps[3].shape
    assert arena.view(ps[3]) is not v                          # a fresh tensor object per call (AccumulateGrad adopts it)
    fired = []
    arena.on_bucket_ready = fired.append
    arena.begin_step()
    for p in reversed(ps):                                     # any marking order: a bucket fires with its last member
        before = len(fired)
        arena.mark_ready(p)
        b = arena.bucket_of[id(p)]
        members = [q for q in ps if arena.bucket_of[id(q)] == b]
        done = all(id(q) not in arena._pending[b] for q in members)
        assert (len(fired) > before) == done
    assert sorted(fired) == list(range(arena.n_buckets))


def test_grad_arena_scratch_and_loss_scale_state():
    """host logic added in round 2: the per-step weight-gradient scratch of the arena (sized from the first step's
    demand, handed out as disjoint zeroed slices, None when it does not fit) and the loss-scale state of FusedSGD
    (GradScaler defaults; `loss_scale()` is a view of the device state the kernels advance)."""
    import torch
    from e2enet_medical_b200.optim import FusedSGD, GradArena
    ps = [torch.nn.Parameter(torch.zeros(s)) for s in ((8, 8, 1, 3, 3), (8,))]
    arena = GradArena(ps, n_buckets=1)
    arena.begin_step()
    assert arena.zeroed_this_step
    assert arena.take_scratch(100) is None and arena.take_scratch(1000) is None        # first step: demand is recorded
    arena.begin_step()                                                                 # ... and allocated here
    a, b = arena.take_scratch(100), arena.take_scratch(1000)
    assert a is not None and b is not None and a.numel() == 128 and b.numel() == 1024  # rounded to 64 floats
    assert a.data_ptr() + a.numel() * 4 == b.data_ptr() and float(a.abs().sum() + b.abs().sum()) == 0.0
    assert arena.take_scratch(64) is None                                              # more than the first step asked for
    a.fill_(1.0)
    arena.begin_step()
    assert float(arena.take_scratch(100).abs().sum()) == 0.0                           # zeroed again every step

    opt = FusedSGD(ps, lr=1e-2)
    assert opt.loss_scale() is None                                                    # bf16: no scaling
    opt.enable_loss_scale()
    s = opt.loss_scale(torch.device("cpu"))
    assert float(s) == 65536.0 and opt._scaler.tolist() == [65536.0, 0.0, 2000.0, 0.5, 2.0]
    opt._scaler[0] = 1024.0
    assert float(opt.loss_scale()) == 1024.0                                           # a view, not a copy
    with pytest.raises(NotImplementedError):
        FusedSGD([{"params": ps[:1]}, {"params": ps[1:]}])

"""GPU tests of the tcgen05/TMA implicit-GEMM kernel (impl=1): it must agree with the mma.sync
gather kernel (impl=0) -- same packed operands, fp32 accumulation in a different order, one bf16
rounding -- and with torch fp32 math on the same bf16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import network as onet

pytestmark = pytest.mark.gpu


def _ulp():
    """one unit in the last place of the library's 16-bit type relative to the tensor maximum"""
    from e2enet_medical_b200 import _lib
    return 2.0 ** -8 if _lib.precision() == "bf16" else 2.0 ** -11


def _act():
    from e2enet_medical_b200 import _lib
    return _lib.act_dtype()


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("src,cout,spatial,B", [
    ([48, 48], 48, (6, 16, 24), 2),        # m = 4, two sources, straddling shift groups
    ([48, 48], 48, (5, 37, 45), 1),        # ragged H / W: partial tiles in both directions
    ([96, 96, 48], 96, (5, 12, 16), 2),    # m = 2
    ([192, 192, 96], 192, (3, 8, 8), 1),   # m = 1, single accumulator stage
    ([1], 48, (6, 16, 16), 2),             # padded 1-channel input, whole input shifted by -2
    ([4], 16, (7, 16, 16), 1),             # four single-channel groups in one block
    ([20], 8, (6, 19, 21), 2),             # Npad 16, partial channel block
    ([48], 48, (4, 160, 160), 1),          # full-resolution rows (5 W tiles of 32)
])
def test_tcgen05_conv_matches_mma_sync_and_torch(src, cout, spatial, B):
    from e2enet_medical_b200 import ops
    from e2enet_medical_b200.plans import build_shiftconv_plan
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    cin = sum(src)
    plan = build_shiftconv_plan(src, cout, (1, 1, 1))
    D, H, W = spatial
    bf = lambda t: t.to(_act()).float()
    xs = [bf(torch.from_numpy(rs.standard_normal((B, c) + spatial).astype(np.float32))).to(dev) for c in src]
    w = bf(torch.from_numpy((rs.standard_normal((cout, cin, 1, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32))).to(dev)
    xs8 = [ops.nc_to_c8(x) for x in xs]
    wp = ops.pack_weights(plan.fwd, w, None)
    outs = []
    for impl in (0, 1):
        raw = torch.full((B, cout // 8, D, H, W, 8), float("nan"), dtype=_act(), device=dev)
        ops.run_gemm(plan.fwd, wp, xs8, (D, H, W), (D, H, W), B, [raw], (D, H, W), [cout // 8], impl)
        torch.cuda.synchronize()
        outs.append(ops.c8_to_nc(raw, cout))
    assert not torch.isnan(outs[1]).any()
    xc = torch.cat(xs, 1).clone().requires_grad_(True)
    ref = F.conv3d(onet.shift_depth(xc), w, None, padding=(0, 1, 1))
    assert rel(outs[1], ref) < _ulp()
    assert rel(outs[1], outs[0]) <= 2 * _ulp()
    # weight gradient: tcgen05 kernel (K = voxels, MN-major operands) vs mma.sync kernel vs torch
    g = bf(torch.from_numpy(rs.standard_normal((B, cout, D, H, W)).astype(np.float32))).to(dev)
    g8 = ops.nc_to_c8(g)
    gw0 = ops.run_wgrad(plan.wgrad, xs8, (D, H, W), (D, H, W), B, g8, tuple(w.shape), 0)
    gw1 = ops.run_wgrad(plan.wgrad, xs8, (D, H, W), (D, H, W), B, g8, tuple(w.shape), 1)
    torch.cuda.synchronize()
    wc = w.clone().requires_grad_(True)
    (F.conv3d(onet.shift_depth(torch.cat(xs, 1)), wc, None, padding=(0, 1, 1)) * g).sum().backward()
    assert rel(gw1, wc.grad) < 2e-4, rel(gw1, wc.grad)
    assert rel(gw1, gw0) < 2e-4
    # direct mode: the split-K flush adds into the parameter layout (no packed scratch / unpack pass)
    ops.CONFIG["wgrad_direct"] = True
    try:
        gw2 = ops.run_wgrad(plan.wgrad, xs8, (D, H, W), (D, H, W), B, g8, tuple(w.shape), 1)
    finally:
        ops.CONFIG["wgrad_direct"] = False
    torch.cuda.synchronize()
    assert rel(gw2, gw1) < 1e-5, rel(gw2, gw1)
    # data gradient variants (one per shift group) through the same kernel
    (ref * g).sum().backward()
    douts = [torch.full_like(s, float("nan")) for s in xs8]
    for var in plan.dgrad:
        ops.run_gemm(var, ops.pack_weights(var, w, None), [g8], (D, H, W), plan.dgrad_iter_grid(var, D, H, W), B,
                     douts, (D, H, W), [s.shape[1] for s in xs8], 1)
    torch.cuda.synchronize()
    off = 0
    for o, c in zip(douts, src):
        got = ops.c8_to_nc(o, c)
        assert not torch.isnan(got).any()
        assert rel(got, xc.grad[:, off:off + c]) < _ulp()
        off += c


@pytest.mark.parametrize("cin,cout,k,spatial,B", [
    (96, 48, (1, 2, 2), (3, 20, 40), 2),      # up4-style: N = 192, m = 1, several tiles, ragged H
    (32, 16, (2, 2, 2), (3, 5, 12), 1),       # N = 128, m = 2, partial tiles in H and W
    (192, 96, (2, 2, 2), (2, 8, 8), 1),       # N = 768 -> 3 column chunks; dgrad K = 8 taps x 12 blocks
    (320, 320, (2, 2, 2), (1, 5, 5), 1),      # N = 2560 -> 10 chunks; dgrad N = 320 -> 2 chunks
    (16, 8, (1, 1, 1), (2, 3, 3), 2),         # pool [1,1,1] (config 1): stride-1 point GEMM
])
def test_tcgen05_point_form_tconv(cin, cout, k, spatial, B):
    """ConvTranspose3d(kernel == stride) forward / data gradient through the 1-tap tcgen05 kernel
    (strided TMA boxes, per-column-block scatter) vs the mma.sync gather kernel and torch."""
    from e2enet_medical_b200 import ops
    from e2enet_medical_b200.plans import build_tconv_plan
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(7)
    bf = lambda t: t.to(_act()).float()
    D, H, W = spatial
    fine = (D * k[0], H * k[1], W * k[2])
    x = bf(torch.from_numpy(rs.standard_normal((B, cin) + spatial).astype(np.float32))).to(dev)
    w = bf(torch.from_numpy((rs.standard_normal((cin, cout) + k) / np.sqrt(cin)).astype(np.float32))).to(dev)
    g = bf(torch.from_numpy(rs.standard_normal((B, cout) + fine).astype(np.float32))).to(dev)
    plan = build_tconv_plan(cin, cout, k)
    x8, g8 = ops.nc_to_c8(x), ops.nc_to_c8(g)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yr = F.conv_transpose3d(xr, wr, stride=k)
    (yr * g).sum().backward()
    res = {}
    for impl in (0, 1):
        y = torch.full((B, cout // 8) + fine + (8,), float("nan"), dtype=_act(), device=dev)
        ops.run_gemm_chunks(plan.fwd, w, None, [x8], spatial, spatial, B, [y], fine, [cout // 8], impl)   # one launch
        dx = torch.full_like(x8, float("nan"))
        ops.run_gemm_chunks(plan.dgrad, w, None, [g8], fine, spatial, B, [dx], spatial, [cin // 8], impl)
        gw = ops.run_wgrad(plan.wgrad, [g8], fine, spatial, B, x8, tuple(w.shape), impl)
        torch.cuda.synchronize()
        res[impl] = (ops.c8_to_nc(y, cout), ops.c8_to_nc(dx, cin), gw)
    for impl in (0, 1):
        y, dx, gw = res[impl]
        assert not torch.isnan(y).any() and not torch.isnan(dx).any()
        assert rel(y, yr) < _ulp(), (impl, rel(y, yr))
        assert rel(dx, xr.grad) < _ulp(), (impl, rel(dx, xr.grad))
        assert rel(gw, wr.grad) < 2e-4, (impl, rel(gw, wr.grad))


@pytest.mark.parametrize("src,cout,spatial,B", [
    ([48, 48], 48, (3, 9, 70), 2),          # loc4-style: 3 W tiles (advance 30), ragged H (4-row tiles) and W
    ([48], 48, (2, 160, 160), 1),           # full-resolution rows
    ([96], 16, (5, 16, 33), 2),             # Np = 16, 6 K pairs
    ([80], 8, (4, 7, 31), 1),               # Cout 8 padded to 16, one W tile + 1 voxel
    ([1], 48, (6, 16, 40), 2),              # the 1-channel input layer: one K pair
    ([20], 8, (6, 19, 21), 2),              # three K pairs (the last stage half full), Cout 8, B = 2
])
def test_tcgen05_kw_stacked_forward(src, cout, spatial, B):
    """kw-stacked forward kernel (N = 3 x Cout per MMA, W shift applied in the epilogue by lane shuffles)
    vs the standard halo-form kernel on the same packed operands."""
    from e2enet_medical_b200 import ops
    from e2enet_medical_b200.plans import build_shiftconv_plan
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(2)
    cin = sum(src)
    plan = build_shiftconv_plan(src, cout, (1, 1, 1))
    assert plan.fwd3 is not None
    D, H, W = spatial
    bf = lambda t: t.to(_act()).float()
    xs8 = [ops.nc_to_c8(bf(torch.from_numpy(rs.standard_normal((B, c) + spatial).astype(np.float32))).to(dev)) for c in src]
    w = bf(torch.from_numpy((rs.standard_normal((cout, cin, 1, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32))).to(dev)
    Cb = cout // 8
    a = torch.full((B, Cb, D, H, W, 8), float("nan"), dtype=_act(), device=dev)
    # the destination of the stacked kernel sits in front of a guard area: Cout = 8 is padded to 16 columns inside the
    # kernel, and the padding block must never be stored (it would land in the next sample / past the tensor)
    n = a.numel()
    buf = torch.full((n + D * H * W * 8,), float("nan"), dtype=_act(), device=dev)
    b_ = buf[:n].view(a.shape)
    ops.run_gemm_chunks(plan.fwd_chunks, w, None, xs8, (D, H, W), (D, H, W), B, [a], (D, H, W), [Cb], 1)
    ops.run_gemm(plan.fwd3, ops.pack_weights(plan.fwd3, w, None), xs8, (D, H, W), (D, H, W), B, [b_], (D, H, W), [Cb], 1)
    torch.cuda.synchronize()
    assert torch.isnan(buf[n:].float()).all(), "the stacked kernel stored a padding channel block"
    assert not torch.isnan(b_.float()).any()
    assert rel(b_, a) <= 2 * _ulp(), rel(b_, a)       # two roundings of sums in different order: one spacing apart


@pytest.mark.parametrize("src,cout,stride,spatial,B,use3", [
    ([48, 48], 48, (1, 1, 1), (3, 9, 70), 2, True),        # kw-stacked narrow kernel, ragged tiles
    ([48, 48], 48, (1, 1, 1), (5, 37, 45), 3, False),      # halo form, m = 4, partial tiles, B = 3
    ([96, 96, 48], 96, (1, 1, 1), (5, 12, 16), 2, False),  # halo form, m = 2, three 32-column chunks
    ([320, 320, 192], 320, (1, 1, 1), (4, 5, 5), 2, False),  # Cout 320 -> two column chunks in one launch
    ([1], 48, (1, 1, 1), (6, 16, 16), 2, False),           # network input layer
    ([48], 96, (1, 2, 2), (6, 16, 16), 2, False),          # strided encoder conv: point form
    ([192], 320, (2, 2, 2), (4, 10, 12), 1, False),        # strided, two column chunks, odd output grid
])
def test_fused_epilogue_instancenorm_statistics(src, cout, stride, spatial, B, use3):
    """InstanceNorm statistics reduced in the conv epilogue (e2e_gemm_t.stats + e2e_in_stats_final) vs the
    separate statistics pass over the stored tensor (e2e_in_stats) and vs torch on the stored values; the fused
    path has no atomics, so two runs must agree bit for bit."""
    import ctypes as C
    from e2enet_medical_b200 import _lib, ops
    from e2enet_medical_b200.plans import build_shiftconv_plan
    lib = _lib.load()
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(4)
    cin = sum(src)
    plan = build_shiftconv_plan(src, cout, stride)
    D, H, W = spatial
    Do, Ho, Wo = plan.out_grid(D, H, W)
    bf = lambda t: t.to(_act()).float()
    xs8 = [ops.nc_to_c8(bf(torch.from_numpy((rs.standard_normal((B, c) + spatial) + 0.3).astype(np.float32))).to(dev))
           for c in src]
    w = bf(torch.from_numpy((rs.standard_normal((cout, cin, 1, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32))).to(dev)
    Cb, V = cout // 8, Do * Ho * Wo
    plans = [plan.fwd3] if use3 else plan.fwd_chunks
    assert plans[0] is not None
    p = lambda t: C.c_void_p(t.data_ptr())
    res = []
    for rep in range(2):
        raw = torch.full((B, Cb, Do, Ho, Wo, 8), float("nan"), dtype=_act(), device=dev)
        stats = ops.run_gemm_chunks(plans, w, None, xs8, (D, H, W), (Do, Ho, Wo), B, [raw], (Do, Ho, Wo), [Cb], 1,
                                    want_stats=True)
        assert stats is not None and stats.shape[1:] == (B, 2, cout), "the tcgen05 launch must fuse the statistics"
        mean = torch.empty(B * cout, dtype=torch.float32, device=dev)
        rstd = torch.empty_like(mean)
        _lib.check(lib.e2e_in_stats_final(p(stats), stats.shape[0], B, cout, V, 1e-5, p(mean), p(rstd), _lib.stream_ptr()))
        torch.cuda.synchronize()
        res.append((raw, mean, rstd))
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2]), "fused statistics are not reproducible"
    raw, mean, rstd = res[0]
    nch = ops._nchunk(V, B * Cb)
    partial = torch.empty(B * Cb * nch * 16, dtype=torch.float32, device=dev)
    mean2 = torch.empty_like(mean)
    rstd2 = torch.empty_like(mean)
    _lib.check(lib.e2e_in_stats(p(raw), B, Cb, V, 1e-5, p(partial), nch, p(mean2), p(rstd2), _lib.stream_ptr()))
    y = ops.c8_to_nc(raw, cout).double()
    m_ref = y.mean((2, 3, 4)).flatten()
    r_ref = (1.0 / torch.sqrt(y.var((2, 3, 4), unbiased=False) + 1e-5)).flatten()
    torch.cuda.synchronize()
    scale = float(y.abs().max())
    assert float((mean.double() - m_ref).abs().max()) < 1e-5 * scale
    assert float(((rstd.double() - r_ref) / r_ref).abs().max()) < 1e-4
    assert float((mean - mean2).abs().max()) < 1e-5 * scale and float(((rstd - rstd2) / rstd2).abs().max()) < 1e-4

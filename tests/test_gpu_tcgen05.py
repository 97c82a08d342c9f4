"""GPU tests of the tcgen05/TMA implicit-GEMM kernel (impl=1): it must agree with the mma.sync
gather kernel (impl=0) -- same packed operands, fp32 accumulation in a different order, one bf16
rounding -- and with torch fp32 math on the same bf16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import network as onet

pytestmark = pytest.mark.gpu
ULP = 2.0 ** -8


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
    bf = lambda t: t.bfloat16().float()
    xs = [bf(torch.from_numpy(rs.standard_normal((B, c) + spatial).astype(np.float32))).to(dev) for c in src]
    w = bf(torch.from_numpy((rs.standard_normal((cout, cin, 1, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32))).to(dev)
    xs8 = [ops.nc_to_c8(x) for x in xs]
    wp = ops.pack_weights(plan.fwd, w, None)
    outs = []
    for impl in (0, 1):
        raw = torch.full((B, cout // 8, D, H, W, 8), float("nan"), dtype=torch.bfloat16, device=dev)
        ops.run_gemm(plan.fwd, wp, xs8, (D, H, W), (D, H, W), B, [raw], (D, H, W), [cout // 8], impl)
        torch.cuda.synchronize()
        outs.append(ops.c8_to_nc(raw, cout))
    assert not torch.isnan(outs[1]).any()
    xc = torch.cat(xs, 1).clone().requires_grad_(True)
    ref = F.conv3d(onet.shift_depth(xc), w, None, padding=(0, 1, 1))
    assert rel(outs[1], ref) < ULP
    assert rel(outs[1], outs[0]) < ULP
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
        assert rel(got, xc.grad[:, off:off + c]) < ULP
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
    bf = lambda t: t.bfloat16().float()
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
        y = torch.full((B, cout // 8) + fine + (8,), float("nan"), dtype=torch.bfloat16, device=dev)
        ops.run_gemm_chunks(plan.fwd, w, None, [x8], spatial, spatial, B, [y], fine, [cout // 8], impl)   # one launch
        dx = torch.full_like(x8, float("nan"))
        ops.run_gemm_chunks(plan.dgrad, w, None, [g8], fine, spatial, B, [dx], spatial, [cin // 8], impl)
        gw = ops.run_wgrad(plan.wgrad, [g8], fine, spatial, B, x8, tuple(w.shape), impl)
        torch.cuda.synchronize()
        res[impl] = (ops.c8_to_nc(y, cout), ops.c8_to_nc(dx, cin), gw)
    for impl in (0, 1):
        y, dx, gw = res[impl]
        assert not torch.isnan(y).any() and not torch.isnan(dx).any()
        assert rel(y, yr) < ULP, (impl, rel(y, yr))
        assert rel(dx, xr.grad) < ULP, (impl, rel(dx, xr.grad))
        assert rel(gw, wr.grad) < 2e-4, (impl, rel(gw, wr.grad))


@pytest.mark.parametrize("src,cout,spatial,B", [
    ([48, 48], 48, (3, 9, 70), 2),          # loc4-style: 3 W tiles (advance 30), ragged H (4-row tiles) and W
    ([48], 48, (2, 160, 160), 1),           # full-resolution rows
    ([96], 16, (5, 16, 33), 2),             # Np = 16, 6 K pairs
    ([80], 8, (4, 7, 31), 1),               # Cout 8 padded to 16, one W tile + 1 voxel
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
    bf = lambda t: t.bfloat16().float()
    xs8 = [ops.nc_to_c8(bf(torch.from_numpy(rs.standard_normal((B, c) + spatial).astype(np.float32))).to(dev)) for c in src]
    w = bf(torch.from_numpy((rs.standard_normal((cout, cin, 1, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32))).to(dev)
    Cb = cout // 8
    a = torch.full((B, Cb, D, H, W, 8), float("nan"), dtype=torch.bfloat16, device=dev)
    b_ = torch.full_like(a, float("nan"))
    ops.run_gemm_chunks(plan.fwd_chunks, w, None, xs8, (D, H, W), (D, H, W), B, [a], (D, H, W), [Cb], 1)
    ops.run_gemm(plan.fwd3, ops.pack_weights(plan.fwd3, w, None), xs8, (D, H, W), (D, H, W), B, [b_], (D, H, W), [Cb], 1)
    torch.cuda.synchronize()
    assert not torch.isnan(b_.float()).any()
    assert rel(b_, a) < ULP, rel(b_, a)

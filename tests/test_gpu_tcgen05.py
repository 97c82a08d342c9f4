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
    gw0 = ops.run_wgrad(plan.fwd, xs8, (D, H, W), (D, H, W), B, g8, tuple(w.shape), 0)
    gw1 = ops.run_wgrad(plan.fwd, xs8, (D, H, W), (D, H, W), B, g8, tuple(w.shape), 1)
    torch.cuda.synchronize()
    wc = w.clone().requires_grad_(True)
    (F.conv3d(onet.shift_depth(torch.cat(xs, 1)), wc, None, padding=(0, 1, 1)) * g).sum().backward()
    assert rel(gw1, wc.grad) < 2e-4, rel(gw1, wc.grad)
    assert rel(gw1, gw0) < 2e-4
    # data gradient variants (one per shift group) through the same kernel
    (ref * g).sum().backward()
    douts = [torch.full_like(s, float("nan")) for s in xs8]
    for var in plan.dgrad:
        ops.run_gemm(var, ops.pack_weights(var, w, None), [g8], (D, H, W), (D, H, W), B, douts, (D, H, W),
                     [s.shape[1] for s in xs8], 1)
    torch.cuda.synchronize()
    off = 0
    for o, c in zip(douts, src):
        got = ops.c8_to_nc(o, c)
        assert not torch.isnan(got).any()
        assert rel(got, xc.grad[:, off:off + c]) < ULP
        off += c

"""GPU debugging aid: checks each CUDA stage of one shift-conv block against torch fp32 ops fed with
the SAME bf16-rounded operands (so differences are accumulation order only, ~1e-6).
Usage (on the GPU box): python tools/debug_kernels.py"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200 import _lib, ops  # noqa: E402
from e2enet_medical_b200.plans import build_shiftconv_plan  # noqa: E402
from oracle import network as onet  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def bf(x):
    return x.bfloat16().float()


def check(src, cout, stride, spatial, B=2, seed=0):
    rs = np.random.RandomState(seed)
    cin = sum(src)
    plan = build_shiftconv_plan(src, cout, stride)
    D, H, W = spatial
    Do, Ho, Wo = plan.out_grid(D, H, W)
    xs = [bf(torch.from_numpy(rs.standard_normal((B, c) + spatial).astype(np.float32))).to(dev) for c in src]
    w = bf(torch.from_numpy((rs.standard_normal((cout, cin, 1, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32))).to(dev)
    xs8 = [ops.nc_to_c8(x) for x in xs]
    # 1. raw conv
    wp = ops.pack_weights(plan.fwd, w, None)
    raw = torch.empty((B, cout // 8, Do, Ho, Wo, 8), dtype=torch.bfloat16, device=dev)
    ops.run_gemm(plan.fwd, wp, xs8, (D, H, W), (Do, Ho, Wo), B, [raw], (Do, Ho, Wo), [cout // 8], 0)
    xcat = torch.cat(xs, 1)
    ref_raw = F.conv3d(onet.shift_depth(xcat), w, None, stride=stride, padding=(0, 1, 1))
    r_raw = rel(ops.c8_to_nc(raw, cout), ref_raw)
    # 2. wgrad with a random upstream gradient
    g = bf(torch.from_numpy(rs.standard_normal((B, cout, Do, Ho, Wo)).astype(np.float32))).to(dev)
    g8 = ops.nc_to_c8(g)
    gw = ops.run_wgrad(plan.fwd, xs8, (D, H, W), (Do, Ho, Wo), B, g8, tuple(w.shape), 0)
    xc = xcat.clone().requires_grad_(True)
    wc = w.clone().requires_grad_(True)
    (F.conv3d(onet.shift_depth(xc), wc, None, stride=stride, padding=(0, 1, 1)) * g).sum().backward()
    r_gw = rel(gw, wc.grad)
    # 3. dgrad
    outs = [torch.full_like(s, float("nan")) for s in xs8]
    sd, sh, sw = stride
    for var in plan.dgrad:
        it = plan.dgrad_iter_grid(var, D, H, W)
        if min(it) <= 0:
            continue
        wpd = ops.pack_weights(var, w, None)
        ops.run_gemm(var, wpd, [g8], (Do, Ho, Wo), it, B, outs, (D, H, W), [s.shape[1] for s in xs8], 0)
    off = 0
    r_dx = []
    for o, c in zip(outs, src):
        got = ops.c8_to_nc(o, c)
        r_dx.append(rel(got, xc.grad[:, off:off + c]) if not torch.isnan(got).any() else float("nan"))
        off += c
    # 4. instance norm fwd/bwd on the raw tensor
    ga = torch.from_numpy((1 + 0.1 * rs.standard_normal(cout)).astype(np.float32)).to(dev)
    be = torch.from_numpy((0.1 * rs.standard_normal(cout)).astype(np.float32)).to(dev)
    V = Do * Ho * Wo
    Cb = cout // 8
    nch = ops._nchunk(V, B * Cb)
    partial = torch.empty(B * Cb * nch * 24, dtype=torch.float32, device=dev)
    mean = torch.empty(B * Cb * 8, dtype=torch.float32, device=dev)
    rstd = torch.empty_like(mean)
    P = ops._p
    _lib.check(lib.e2e_in_stats(P(raw), B, Cb, V, 1e-5, P(partial), nch, P(mean), P(rstd), None))
    y = torch.empty_like(raw)
    _lib.check(lib.e2e_in_apply(P(raw), P(mean), P(rstd), P(ga), P(be), 0.01, B, Cb, V, P(y), None))
    rawf = ops.c8_to_nc(raw, cout).requires_grad_(True)
    gar, ber = ga.clone().requires_grad_(True), be.clone().requires_grad_(True)
    yr = onet.instance_norm_lrelu(rawf, gar, ber)
    r_y = rel(ops.c8_to_nc(y, cout), yr)
    (yr * g).sum().backward()
    sums = torch.empty(B * Cb * 16, dtype=torch.float32, device=dev)
    draw = torch.empty_like(raw)
    dga, dbe, dbi = (torch.empty(cout, dtype=torch.float32, device=dev) for _ in range(3))
    _lib.check(lib.e2e_in_bwd(P(g8), P(raw), P(mean), P(rstd), P(ga), P(be), 0.01, B, Cb, V, P(partial), nch, P(sums),
                              P(draw), P(dga), P(dbe), P(dbi), None))
    r_draw = rel(ops.c8_to_nc(draw, cout), rawf.grad)
    r_dga, r_dbe = rel(dga, gar.grad), rel(dbe, ber.grad)
    mref = rawf.detach().mean((2, 3, 4)).reshape(-1)
    r_mean = float((mean.view(B, Cb * 8)[:, :cout].reshape(-1) - mref).abs().max())
    torch.cuda.synchronize()
    print(f"src={src} cout={cout} stride={stride} sp={spatial}: raw {r_raw:.2e} wgrad {r_gw:.2e} dgrad {['%.2e' % v for v in r_dx]} "
          f"| in: y {r_y:.2e} draw {r_draw:.2e} dgamma {r_dga:.2e} dbeta {r_dbe:.2e} mean_abs {r_mean:.2e}", flush=True)


if __name__ == "__main__":
    check([20], 8, (1, 1, 1), (6, 9, 10))
    check([8], 8, (1, 1, 1), (4, 8, 8))
    check([48, 48], 48, (1, 1, 1), (6, 16, 24))
    check([96, 96, 48], 96, (1, 1, 1), (5, 12, 8))
    check([1], 48, (1, 1, 1), (6, 16, 16))
    check([4], 16, (1, 1, 1), (7, 8, 8))
    check([48], 96, (1, 2, 2), (6, 16, 16))
    check([16], 32, (2, 2, 2), (8, 10, 12))
    check([320, 320, 192], 320, (1, 1, 1), (4, 5, 5))
    check([48, 48], 48, (1, 1, 1), (16, 40, 48), B=1)

"""Full-size (BASELINE.json shapes) GPU tests through size-independent properties: the oracle cannot
run these sizes in seconds, so the CUDA path is checked against itself / against invariants:
  * the tcgen05 kernels vs the mma.sync kernels on the real config-2 layer shapes (same packed
    operands: fp32 accumulation order and one bf16 rounding are the only differences);
  * convolution linearity in the weights (conv(x, w1 + w2) = conv(x, w1) + conv(x, w2) up to bf16);
  * sliding-window partition of unity on the AMOS-sized accumulator (constant logits -> the
    normalised probabilities are that constant everywhere, for any tile overlap pattern);
  * Masking.apply_mask idempotence and density on the config-2 network.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ULP = 2.0 ** -8


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("src,cout,stride,spatial", [
    ([48, 48], 48, (1, 1, 1), (64, 160, 160)),        # loc4: the largest GEMM of config 2
    ([96, 96, 48], 96, (1, 1, 1), (64, 80, 80)),      # loc3
    ([48], 96, (1, 2, 2), (64, 160, 160)),            # first strided encoder conv
])
def test_fullsize_conv_tc_vs_mma_sync(src, cout, stride, spatial):
    from e2enet_medical_b200 import ops
    from e2enet_medical_b200.plans import build_shiftconv_plan
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    B, cin = 2, sum(src)
    D, H, W = spatial
    plan = build_shiftconv_plan(src, cout, stride)
    Do, Ho, Wo = plan.out_grid(D, H, W)
    xs8 = [torch.randn((B, (c + 7) // 8, D, H, W, 8), device=dev, generator=g).bfloat16() for c in src]
    w = (torch.randn((cout, cin, 1, 3, 3), device=dev, generator=g) / np.sqrt(cin * 9)).bfloat16().float()
    outs = []
    for impl in (0, 1):
        raw = torch.empty((B, cout // 8, Do, Ho, Wo, 8), dtype=torch.bfloat16, device=dev)
        ops.run_gemm_chunks(plan.fwd_chunks, w, None, xs8, (D, H, W), (Do, Ho, Wo), B, [raw], (Do, Ho, Wo), [cout // 8], impl)
        outs.append(raw)
    torch.cuda.synchronize()
    assert rel(outs[1], outs[0]) < 2 * ULP
    # checksum of checksums: per-(b, channel-block) sums agree to fp32 accumulation noise
    s0, s1 = outs[0].float().sum((2, 3, 4)), outs[1].float().sum((2, 3, 4))
    assert float((s0 - s1).abs().max()) < 2e-3 * float(outs[0].float().abs().sum((2, 3, 4)).max())
    # weight gradient and data gradient, same comparison
    gr = torch.randn((B, cout // 8, Do, Ho, Wo, 8), device=dev, generator=g).bfloat16()
    gw = [ops.run_wgrad(plan.wgrad, xs8, (D, H, W), (Do, Ho, Wo), B, gr, tuple(w.shape), impl) for impl in (0, 1)]
    assert rel(gw[1], gw[0]) < 1e-3
    dx = []
    for impl in (0, 1):
        o = [(torch.zeros_like(s) if plan.dgrad_needs_zero else torch.full_like(s, float("nan"))) for s in xs8]
        for grp in plan.dgrad_groups:
            ops.run_gemm_chunks(grp, w, None, [gr], (Do, Ho, Wo), plan.dgrad_iter_grid(grp[0], D, H, W), B, o, (D, H, W),
                                [s.shape[1] for s in xs8], impl)
        dx.append(o)
    torch.cuda.synchronize()
    for a, b_ in zip(dx[1], dx[0]):
        assert not torch.isnan(a.float()).any()
        assert rel(a, b_) < 2 * ULP


def test_fullsize_conv_linearity_in_weights():
    from e2enet_medical_b200 import ops
    from e2enet_medical_b200.plans import build_shiftconv_plan
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    src, cout, (D, H, W), B = [48, 48], 48, (64, 160, 160), 2
    plan = build_shiftconv_plan(src, cout, (1, 1, 1))
    xs8 = [torch.randn((B, c // 8, D, H, W, 8), device=dev, generator=g).bfloat16() for c in src]
    # weights on a coarse binary grid so that w1 + w2 is exact in bf16
    w1 = torch.randint(-8, 9, (cout, 96, 1, 3, 3), device=dev, generator=g).float() / 64
    w2 = torch.randint(-8, 9, (cout, 96, 1, 3, 3), device=dev, generator=g).float() / 64
    res = []
    for w in (w1, w2, w1 + w2):
        raw = torch.empty((B, cout // 8, D, H, W, 8), dtype=torch.bfloat16, device=dev)
        ops.run_gemm_chunks(plan.fwd_chunks, w, None, xs8, (D, H, W), (D, H, W), B, [raw], (D, H, W), [cout // 8], 1)
        res.append(raw.float())
    torch.cuda.synchronize()
    assert rel(res[2], res[0] + res[1]) < 3 * ULP


def test_fullsize_window_partition_of_unity():
    from e2enet_medical_b200 import _lib
    from e2enet_medical_b200.network_architecture.neural_network import SegmentationNetwork as S
    dev = torch.device("cuda:0")
    lib = _lib.load()
    ncls, patch, vol = 16, (64, 160, 160), (300, 512, 512)
    steps = S._compute_steps_for_sliding_window(patch, vol, 0.5)
    assert len(steps[0]) * len(steps[1]) * len(steps[2]) == 324
    gauss = torch.from_numpy(S._get_gaussian(patch)).to(dev)
    # constant logits per class -> softmax is the same vector at every voxel
    z = torch.linspace(-2, 2, ncls, device=dev)
    logits = z.view(ncls, 1, 1, 1).expand(ncls, *patch).contiguous()
    want = torch.softmax(z, 0)
    agg = torch.zeros((ncls,) + vol, device=dev)
    wsum = torch.zeros(vol, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    for a in steps[0]:
        for b in steps[1]:
            for c in steps[2]:
                _lib.check(lib.e2e_window_accumulate(p(logits), p(gauss), p(agg), p(wsum), ncls, *patch, *vol, a, b, c, 0,
                                                     1.0, 1, 1, _lib.stream_ptr()))
    seg = torch.empty(vol, dtype=torch.int64, device=dev)
    _lib.check(lib.e2e_window_finalize(p(agg), p(wsum), ncls, *vol, p(seg), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert float(wsum.min()) > 0, "every voxel must be covered by at least one tile"
    err = (agg - want.view(ncls, 1, 1, 1)).abs().amax((1, 2, 3))
    assert float(err.max()) < 2e-6, err
    assert int((seg != int(want.argmax())).sum()) == 0


def test_fullsize_masking_apply_idempotent():
    import random
    from e2enet_medical_b200.training import POOLS, TrainStep
    dev = torch.device("cuda:0")
    random.seed(0)
    ts = TrainStep(1, 14, POOLS["btcv"], (64, 160, 160), 0.2, 0.5, 1200, dev, 1, seed=0)
    names = list(ts.mask.masks.keys())
    assert len(names) == 35 and sum(m.numel() for m in ts.mask.masks.values()) == 17975040
    params = dict(ts.network.named_parameters())
    ts.mask.apply_mask()
    snap = {n: params[n].detach().clone() for n in names}
    ts.mask.apply_mask()
    for n in names:
        assert torch.equal(params[n].detach(), snap[n]), n
        m = ts.mask.masks[n]
        assert torch.equal(params[n].detach() * m, params[n].detach())
        assert torch.equal(m, m * m), "masks are 0/1"
    dens = sum(float(m.sum()) for m in ts.mask.masks.values()) / 17975040
    assert abs(dens - 0.2) < 1e-3, dens

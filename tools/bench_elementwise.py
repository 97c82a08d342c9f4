"""Micro-benchmark of the bandwidth-bound stages against the HBM roofline: InstanceNorm stats /
apply / backward, MaxPool fwd / bwd, layout conversion, multi-tensor apply_mask, sliding-window
accumulate / finalize.  Algorithmic bytes / CUDA-event time.
Usage: python tools/bench_elementwise.py [--iters 10]"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from e2enet_medical_b200 import _lib, ops  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    pk = peak()
    p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    st = _lib.stream_ptr

    def report(name, byts, ms):
        gbs = byts / ms / 1e6
        print("%-52s %8.1f MB %8.3f ms %7.0f GB/s  %5.1f %% of %.0f" % (name, byts / 1e6, ms, gbs, 100 * gbs / pk, pk),
              flush=True)

    for (B, Cc, D, H, W) in ((2, 48, 64, 160, 160), (2, 96, 64, 80, 80), (2, 192, 32, 40, 40)):
        Cb, V = Cc // 8, D * H * W
        raw = torch.randn((B, Cb, D, H, W, 8), device=dev).bfloat16()
        dy = torch.randn_like(raw)
        y = torch.empty_like(raw)
        draw = torch.empty_like(raw)
        nch = ops._nchunk(V, B * Cb)
        partial = torch.empty(B * Cb * nch * 24, dtype=torch.float32, device=dev)
        mean = torch.empty(B * Cb * 8, dtype=torch.float32, device=dev)
        rstd = torch.empty_like(mean)
        sums = torch.empty(B * Cb * 16, dtype=torch.float32, device=dev)
        ga = torch.ones(Cc, device=dev)
        be = torch.zeros(Cc, device=dev)
        dg, db, dbi = torch.empty(Cc, device=dev), torch.empty(Cc, device=dev), torch.empty(Cc, device=dev)
        nb = raw.numel() * 2
        tag = "(%d,%d,%d,%d,%d)" % (B, Cc, D, H, W)
        report("in_stats " + tag, nb, timeit(lambda: _lib.check(lib.e2e_in_stats(
            p(raw), B, Cb, V, 1e-5, p(partial), nch, p(mean), p(rstd), st())), a.iters))
        report("in_apply " + tag, 2 * nb, timeit(lambda: _lib.check(lib.e2e_in_apply(
            p(raw), p(mean), p(rstd), p(ga), p(be), 0.01, B, Cb, V, p(y), st())), a.iters))
        report("in_bwd (reduce + apply) " + tag, 5 * nb, timeit(lambda: _lib.check(lib.e2e_in_bwd(
            p(dy), p(raw), p(mean), p(rstd), p(ga), p(be), 0.01, B, Cb, V, p(partial), nch, p(sums), p(draw), p(dg),
            p(db), p(dbi), st())), a.iters))
        scratch = torch.empty(int(lib.e2e_in_bwd_scratch_floats(B, Cb, V)), dtype=torch.float32, device=dev)
        report("in_bwd plane-resident (A/B, 3 HBM passes) " + tag, 3 * nb, timeit(lambda: _lib.check(lib.e2e_in_bwd_fused(
            p(dy), p(None), p(None), p(raw), p(mean), p(rstd), p(ga), p(be), 0.01, B, Cb, D, H, W, 1, 1, 1, p(scratch), p(sums),
            p(draw), p(dg), p(db), p(dbi), st())), a.iters))
        k = (1, 2, 2) if D == 64 and H == 160 else (2, 2, 2)
        yo = torch.empty((B, Cb, D // k[0], H // k[1], W // k[2], 8), dtype=torch.bfloat16, device=dev)
        am = torch.empty(yo.shape, dtype=torch.uint8, device=dev)
        report("maxpool_fwd k%s %s" % (k, tag), nb + yo.numel() * 3, timeit(lambda: _lib.check(lib.e2e_maxpool_fwd(
            p(raw), p(yo), p(am), B * Cb, D, H, W, k[0], k[1], k[2], st())), a.iters))
        report("maxpool_bwd k%s %s" % (k, tag), nb + yo.numel() * 3, timeit(lambda: _lib.check(lib.e2e_maxpool_bwd(
            p(yo), p(am), p(raw), B * Cb, D, H, W, k[0], k[1], k[2], st())), a.iters))
        del raw, dy, y, draw
    # layout conversion of the network input / logits
    x = torch.randn(2, 14, 64, 160, 160, device=dev)
    report("nc_to_c8 (2,14,64,160,160) fp32 -> bf16 C8", x.numel() * 4 + 2 * 16 * 64 * 160 * 160 * 2,
           timeit(lambda: ops.nc_to_c8(x), a.iters))
    # multi-tensor apply_mask on the 35 masked tensors of config 2 (17 975 040 elements)
    from e2enet_medical_b200.training import POOLS, TrainStep
    ts = TrainStep(1, 14, POOLS["btcv"], (64, 160, 160), 0.2, 0.5, 1200, dev, 1, seed=0)
    for prm in ts.network.parameters():       # give SGD its momentum buffers
        prm.grad = torch.zeros_like(prm)
    ts.optimizer.step()
    n = sum(m.numel() for m in ts.mask.masks.values())
    ts.mask.apply_mask()                                   # builds the device pointer tables
    _, wt, bt, mt, nt, mx = ts.mask._scratch[("apply", True)]
    n_t = int(nt.numel())
    report("mask_apply_multi kernel, 35 tensors (w, momentum RMW + mask read)", n * 4 * 6,
           timeit(lambda: _lib.check(lib.e2e_mask_apply_multi(p(wt), p(bt), p(mt), p(nt), n_t, mx, st())), a.iters))
    report("Masking.apply_mask() incl. Python host side", n * 4 * 6, timeit(ts.mask.apply_mask, a.iters))
    # fused optimizer step: clip coefficient + Nesterov SGD + weight decay + mask (grad read twice, p / momentum RMW, mask read)
    nparam = sum(q.numel() for q in ts.network.parameters())
    report("FusedSGD.step(): clip + SGD + apply_mask, 147 tensors", nparam * 4 * 6 + n * 4, timeit(ts.optimizer.step, a.iters))
    del ts
    # sliding window: one (16, 64, 160, 160) tile into a (16, 128, 320, 320) accumulator
    ncls, px, py, pz, X, Y, Z = 16, 64, 160, 160, 128, 320, 320
    logits = torch.randn(ncls, px, py, pz, device=dev)
    gauss = torch.rand(px, py, pz, device=dev)
    agg = torch.zeros(ncls, X, Y, Z, device=dev)
    wsum = torch.ones(X, Y, Z, device=dev)
    P = px * py * pz
    report("window_accumulate 16 classes, 64x160x160 tile", P * (ncls * 4 + 2 * ncls * 4 + 2 * 4 + 4),
           timeit(lambda: _lib.check(lib.e2e_window_accumulate(p(logits), p(gauss), p(agg), p(wsum), ncls, px, py, pz,
                                                                X, Y, Z, 32, 80, 80, 0, 1.0, 1, 1, st())), a.iters))
    feat = torch.randn(6, px, py, pz, 8, device=dev).bfloat16()
    hw = torch.randn(ncls, 48, device=dev)
    report("window_head_accumulate (1x1x1 head + softmax + accumulate), 48 ch -> 16", P * (48 * 2 + 2 * ncls * 4 + 2 * 4 + 4),
           timeit(lambda: _lib.check(lib.e2e_window_head_accumulate(p(feat), 6, p(hw), 48, p(gauss), p(agg), p(wsum), ncls, px,
                                                                     py, pz, X, Y, Z, 32, 80, 80, 0, 1.0, 1, st())), a.iters))
    wsum.fill_(1.0)
    lab = torch.empty((150, 400, 400), dtype=torch.uint8, device=dev)
    report("resample_argmax (16,128,320,320) -> (150,400,400) labels", 150 * 400 * 400 + ncls * X * Y * Z * 4,
           timeit(lambda: _lib.check(lib.e2e_resample_argmax(p(agg), ncls, X, Y, Z, 150, 400, 400, 1, 1, 1, p(None), p(lab), st())),
                  a.iters))
    seg = torch.empty(X, Y, Z, dtype=torch.int64, device=dev)
    Vv = X * Y * Z
    report("window_finalize (16,128,320,320)", Vv * (2 * ncls * 4 + 4 + 8),
           timeit(lambda: _lib.check(lib.e2e_window_finalize(p(agg), p(wsum), ncls, X, Y, Z, p(seg), st())), a.iters))


if __name__ == "__main__":
    main()

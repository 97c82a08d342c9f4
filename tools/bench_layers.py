"""Per-layer micro-benchmark of every GEMM-shaped op of config 2 (BTCV, B=2) in isolation:
shift-conv fwd / dgrad / wgrad and transposed-conv fwd / dgrad / wgrad, CUDA-event timed
(inputs of the big layers exceed the 126 MB L2; small layers are L2-resident as in the real step).
Usage: python tools/bench_layers.py [--impl 1] [--only loc4] [--iters 10] [--ops fwd,dgrad,wgrad]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200 import ops  # noqa: E402
from e2enet_medical_b200.plans import build_shiftconv_plan, build_tconv_plan  # noqa: E402

CONVS = [  # name, sources, cout, stride, input (D,H,W), count in the network
    ("loc4 96->48 @64x160x160", [48, 48], 48, (1, 1, 1), (64, 160, 160), 5),
    ("ctx0.1/loc0.4.1 48->48 @64x160x160", [48], 48, (1, 1, 1), (64, 160, 160), 2),
    ("ctx0.0 1->48 @64x160x160", [1], 48, (1, 1, 1), (64, 160, 160), 1),
    ("loc3 240->96 @64x80x80", [96, 96, 48], 96, (1, 1, 1), (64, 80, 80), 4),
    ("ctx1.1/loc0.3.1 96->96 @64x80x80", [96], 96, (1, 1, 1), (64, 80, 80), 2),
    ("ctx1.0 48->96 s(1,2,2) @64x160x160", [48], 96, (1, 2, 2), (64, 160, 160), 1),
    ("loc2 480->192 @32x40x40", [192, 192, 96], 192, (1, 1, 1), (32, 40, 40), 3),
    ("ctx2.1/loc0.2.1 192->192 @32x40x40", [192], 192, (1, 1, 1), (32, 40, 40), 2),
    ("ctx2.0 96->192 s2 @64x80x80", [96], 192, (2, 2, 2), (64, 80, 80), 1),
    ("loc1 832->320 @16x20x20", [320, 320, 192], 320, (1, 1, 1), (16, 20, 20), 2),
    ("ctx3.1/loc0.1.1 320->320 @16x20x20", [320], 320, (1, 1, 1), (16, 20, 20), 2),
    ("ctx3.0 192->320 s2 @32x40x40", [192], 320, (2, 2, 2), (32, 40, 40), 1),
    ("loc0.0.0 960->320 @8x10x10", [320, 320, 320], 320, (1, 1, 1), (8, 10, 10), 1),
    # probes (count 0: not in the totals): channel counts whose 5 depth-shift groups start on 8-channel block boundaries,
    # so no block is loaded / multiplied twice (C = 48 has 9 K entries for 6 blocks, C = 40 has 5 for 5)
    ("probe 40->48 aligned groups @64x160x160", [40], 48, (1, 1, 1), (64, 160, 160), 0),
    ("probe 80->48 aligned groups @64x160x160", [40, 40], 48, (1, 1, 1), (64, 160, 160), 0),
]
TCONVS = [  # name, cin, cout, k, input (D,H,W), count
    ("up 96->48 k(1,2,2) @64x80x80", 96, 48, (1, 2, 2), (64, 80, 80), 5),
    ("up 192->96 k2 @32x40x40", 192, 96, (2, 2, 2), (32, 40, 40), 4),
    ("up 320->192 k2 @16x20x20", 320, 192, (2, 2, 2), (16, 20, 20), 3),
    ("up 320->320 k2 @8x10x10", 320, 320, (2, 2, 2), (8, 10, 10), 2),
    ("up 320->320 k2 @4x5x5", 320, 320, (2, 2, 2), (4, 5, 5), 1),
]


GRAPH = [os.environ.get("E2E_BENCH_GRAPH", "1") != "0"]


def timeit(fn, iters):
    """device time per call.  The calls are captured into ONE CUDA graph and replayed (default): a launch through the
    C ABI costs ~0.1 ms of host time (tensor-map encodes, ctypes), which would otherwise bound every kernel shorter
    than that -- in the real step the whole iteration is a graph too."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if GRAPH[0]:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(iters):
                    fn()
        torch.cuda.current_stream().wait_stream(s)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", type=int, default=1)
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--ops", default="fwd,dgrad,wgrad")
    ap.add_argument("--stats", type=int, default=1, help="forward with the InstanceNorm statistics fused into the epilogue")
    ap.add_argument("--accum", type=int, default=0, help="data gradient: bitmask of destinations that accumulate (fan-in)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, impl, which = a.B, a.impl, a.ops.split(",")
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    totf = 0.0
    print("%-40s %9s | %s" % ("layer (x count)", "GFLOP", "  ".join("%-24s" % o for o in which)))
    for name, src, cout, stride, (D, H, W), cnt in CONVS:
        if a.only and a.only not in name:
            continue
        plan = build_shiftconv_plan(src, cout, stride)
        cin = sum(src)
        Do, Ho, Wo = plan.out_grid(D, H, W)
        xs8 = [torch.randn((B, (c + 7) // 8, D, H, W, 8), device=dev).bfloat16() for c in src]
        w = torch.randn((cout, cin, 1, 3, 3), device=dev) / np.sqrt(cin * 9)
        raw = torch.randn((B, cout // 8, Do, Ho, Wo, 8), device=dev).bfloat16()
        dxs = [torch.empty_like(x) for x in xs8]
        flops = 2.0 * B * Do * Ho * Wo * cout * cin * 9
        wpf = [ops.pack_weights(c, w, None) for c in plan.fwd_chunks]
        wpd = [ops.pack_weights(v, w, None) for v in plan.dgrad]

        stats = a.stats and impl == 1           # forward as the network runs it: InstanceNorm sums in the epilogue

        def fwd():
            if impl == 1 and plan.fwd3 is not None and os.environ.get("E2E_STACK3", "1") == "1":
                ops.run_gemm_chunks([plan.fwd3], w, None, xs8, (D, H, W), (Do, Ho, Wo), B, [raw], (Do, Ho, Wo),
                                    [cout // 8], impl, want_stats=stats)
            else:
                ops.run_gemm_chunks(plan.fwd_chunks, w, None, xs8, (D, H, W), (Do, Ho, Wo), B, [raw], (Do, Ho, Wo),
                                    [cout // 8], impl, want_stats=stats)

        def dgrad():
            if plan.dgrad_needs_zero:
                for t in dxs:
                    t.zero_()
            for grp in plan.dgrad_groups:
                it = plan.dgrad_iter_grid(grp[0], D, H, W)
                if min(it) > 0:
                    ops.run_gemm_chunks(grp, w, None, [raw], (Do, Ho, Wo), it, B, dxs, (D, H, W),
                                        [x.shape[1] for x in xs8], impl, accumulate=a.accum)

        def wgrad():
            ops.run_wgrad(plan.wgrad, xs8, (D, H, W), (Do, Ho, Wo), B, raw, tuple(w.shape), impl)

        line = "%-40s %9.1f |" % ("%s x%d" % (name, cnt), flops / 1e9)
        for o in which:
            ms = timeit({"fwd": fwd, "dgrad": dgrad, "wgrad": wgrad}[o], a.iters)
            tot[o] += ms * cnt
            line += " %7.3f ms %7.1f TF/s   " % (ms, flops / ms / 1e9)
        totf += flops * cnt
        print(line, flush=True)
        del xs8, raw, dxs
    for name, cin, cout, k, (D, H, W), cnt in TCONVS:
        if a.only and a.only not in name:
            continue
        plan = build_tconv_plan(cin, cout, k)
        fine = (D * k[0], H * k[1], W * k[2])
        x8 = torch.randn((B, cin // 8, D, H, W, 8), device=dev).bfloat16()
        y8 = torch.randn((B, cout // 8) + fine + (8,), device=dev).bfloat16()
        dx8 = torch.empty_like(x8)
        w = torch.randn((cin, cout) + k, device=dev) / np.sqrt(cin)
        flops = 2.0 * B * D * H * W * cin * cout * k[0] * k[1] * k[2]
        byts = (x8.numel() + y8.numel()) * 2
        wpf = [ops.pack_weights(c, w, None) for c in plan.fwd]
        wpd = [ops.pack_weights(c, w, None) for c in plan.dgrad]

        def fwd():
            ops.run_gemm_chunks(plan.fwd, w, None, [x8], (D, H, W), (D, H, W), B, [y8], fine, [cout // 8], impl)

        def dgrad():
            ops.run_gemm_chunks(plan.dgrad, w, None, [y8], fine, (D, H, W), B, [dx8], (D, H, W), [cin // 8], impl)

        def wgrad():
            ops.run_wgrad(plan.wgrad, [y8], fine, (D, H, W), B, x8, tuple(w.shape), impl)

        line = "%-40s %9.1f |" % ("%s x%d" % (name, cnt), flops / 1e9)
        for o in which:
            ms = timeit({"fwd": fwd, "dgrad": dgrad, "wgrad": wgrad}[o], a.iters)
            tot[o] += ms * cnt
            line += " %7.3f ms %6.0f GB/s %5.0f TF " % (ms, byts / ms / 1e6, flops / ms / 1e9)
        totf += flops * cnt
        print(line, flush=True)
    print("weighted totals: " + "  ".join("%s %.2f ms" % (o, tot[o]) for o in which) +
          "   (fwd GFLOP covered: %.0f)" % (totf / 1e9))


if __name__ == "__main__":
    main()

"""Per-layer micro-benchmark of the conv GEMM kernels on the real config-2 (BTCV, B=2) layer shapes:
mma.sync gather kernel (impl 0) vs tcgen05/TMA kernel (impl 1); forward GEMM only, CUDA events.
Usage: python tools/bench_conv.py [--impl 0,1] [--only loc4]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200 import ops  # noqa: E402
from e2enet_medical_b200.plans import build_shiftconv_plan  # noqa: E402

LAYERS = [  # name, sources, cout, (D,H,W)
    ("loc4 96->48 @64x160x160", [48, 48], 48, (64, 160, 160)),
    ("ctx0.1 48->48 @64x160x160", [48], 48, (64, 160, 160)),
    ("ctx0.0 1->48 @64x160x160", [1], 48, (64, 160, 160)),
    ("loc3 240->96 @64x80x80", [96, 96, 48], 96, (64, 80, 80)),
    ("ctx1.1 96->96 @64x80x80", [96], 96, (64, 80, 80)),
    ("loc2 480->192 @32x40x40", [192, 192, 96], 192, (32, 40, 40)),
    ("ctx2.1 192->192 @32x40x40", [192], 192, (32, 40, 40)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="0,1")
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--wgrad", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B = a.B
    for name, src, cout, (D, H, W) in LAYERS:
        if a.only and a.only not in name:
            continue
        plan = build_shiftconv_plan(src, cout, (1, 1, 1))
        cin = sum(src)
        xs8 = [torch.randn((B, (c + 7) // 8, D, H, W, 8), device=dev).bfloat16() for c in src]
        w = torch.randn((cout, cin, 1, 3, 3), device=dev) / np.sqrt(cin * 9)
        wp = ops.pack_weights(plan.fwd, w, None)
        raw = torch.empty((B, cout // 8, D, H, W, 8), dtype=torch.bfloat16, device=dev)
        flops = 2.0 * B * D * H * W * cout * cin * 9
        byts = sum(x.numel() for x in xs8) * 2 + raw.numel() * 2
        line = f"{name:34s} {flops / 1e9:8.1f} GF  K-entries {plan.fwd.n_cent:4d}"
        for impl in [int(v) for v in a.impl.split(",")]:
            for _ in range(3):
                ops.run_gemm(plan.fwd, wp, xs8, (D, H, W), (D, H, W), B, [raw], (D, H, W), [cout // 8], impl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                ops.run_gemm(plan.fwd, wp, xs8, (D, H, W), (D, H, W), B, [raw], (D, H, W), [cout // 8], impl)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            line += f" | impl{impl}: {ms:7.3f} ms {flops / ms / 1e9:7.1f} TF/s {byts / ms / 1e6:6.0f} GB/s"
            if a.wgrad:
                for _ in range(2):
                    ops.run_wgrad(plan.fwd, xs8, (D, H, W), (D, H, W), B, raw, tuple(w.shape), impl)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.iters):
                    ops.run_wgrad(plan.fwd, xs8, (D, H, W), (D, H, W), B, raw, tuple(w.shape), impl)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.iters
                line += f" wgrad {ms:7.3f} ms {flops / ms / 1e9:7.1f} TF/s"
        print(line, flush=True)


if __name__ == "__main__":
    main()

"""Run-to-run check of the training iteration: TrainStep instances with the same seed take the same steps on the same
batch; gradients and weights must agree to fp32 round-off (the split-K atomics of the weight gradients are the only
order-dependent sums).  A larger difference means a race or a dependence on memory contents / addresses; the first
differing gradients are listed.   python tools/determinism_check.py [graph]"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch  # noqa: E402

dev = torch.device("cuda:0")
pools, patch = POOLS["btcv"], (32, 96, 96)
data, targets = synthetic_batch(2, 1, 14, patch, pools, seed=1)
x, tg = data.to(dev), [t.to(dev) for t in targets]
graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
N = int(os.environ.get("E2E_DET_INSTANCES", "3"))


def make():
    random.seed(0)
    ts = TrainStep(1, 14, pools, patch, 0.2, 0.5, 1200, dev, 1, seed=0, n_buckets=4)
    if graph:
        ts.enable_graph(x, tg, warmup=2)
    return ts


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


tss = [make() for _ in range(N)]
ok = True
for step in range(3):
    losses = [float(ts.step(x, tg)) for ts in tss]
    torch.cuda.synchronize()
    ref = dict(tss[0].network.named_parameters())
    for i, ts in enumerate(tss[1:], 1):
        cur = dict(ts.network.named_parameters())
        dg = sorted(((rel(cur[k].grad, ref[k].grad), k) for k in ref if ref[k].grad is not None and not k.endswith("conv.bias")),
                    reverse=True)
        dw = max(rel(cur[k].detach(), ref[k].detach()) for k in ref if not k.endswith("conv.bias"))
        bad = [(round(v, 9), k) for v, k in dg if v > 1e-5]
        ok = ok and not bad and dw < 1e-5
        print("step %d instance %d: losses %r %r  worst grad diff %.2e (%s)  worst weight diff %.2e  grads differing > 1e-5: %d %s"
              % (step, i, losses[0], losses[i], dg[0][0], dg[0][1], dw, len(bad), bad[-4:]), flush=True)
print("DETERMINISM", "OK" if ok else "BROKEN", flush=True)

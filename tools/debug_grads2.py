"""Debug: find non-reproducible / uninitialised-memory gradients.  (1) same TrainStep, two forward+backward passes:
per-parameter relative gradient difference in registration order; (2) the same with the caching allocator's free
blocks poisoned with NaN first: a NaN gradient means some kernel read memory nobody wrote."""
import os, random, sys
from collections import OrderedDict
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200 import ops
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch

dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "hippo"
if cfg == "hippo":
    pools, patch, ncls, B = POOLS["hippo"], (40, 56, 40), 3, 1
else:
    pools, patch, ncls, B = POOLS["btcv"], (32, 96, 96), 14, 2
data, targets = synthetic_batch(B, 1, ncls, patch, pools, seed=1)
x, tg = data.to(dev), [t.to(dev) for t in targets]
ops.CONFIG["fuse_fanin"] = os.environ.get("FANIN", "0") == "1"
ops.CONFIG["fuse_stats"] = os.environ.get("STATS", "0") == "1"
ops.CONFIG["fuse_pool"] = os.environ.get("POOL", "1") == "1"
ops.CONFIG["impl"] = int(os.environ.get("IMPL", "1"))
random.seed(0)
ts = TrainStep(1, ncls, pools, patch, 0.2, 0.5, 1200, dev, 1, seed=0, fused_optimizer=False)


def poison(gb=24):
    blocks = []
    for mb in (1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610):
        n = max(1, int(gb * 1024 / 14 / mb))
        for _ in range(min(n, 64)):
            blocks.append(torch.full((mb * 1024 * 1024 // 2,), float("nan"), dtype=torch.bfloat16, device=dev))
    del blocks


def run(do_poison):
    ts.optimizer.zero_grad()
    if do_poison:
        poison()
    l = ts.loss(ts.network(x), tg)
    if do_poison:
        poison()
    l.backward()
    torch.cuda.synchronize()
    return float(l), OrderedDict((k, p.grad.detach().clone()) for k, p in ts.network.named_parameters())


a, b = run(False), run(False)
print("config", cfg, {k: ops.CONFIG[k] for k in ("impl", "fuse_fanin", "fuse_stats", "fuse_pool")}, "loss", a[0], b[0])
bad = []
for k in a[1]:
    d = float((a[1][k] - b[1][k]).norm() / b[1][k].norm().clamp_min(1e-30))
    if d > 1e-4 and not k.endswith("conv.bias"):
        bad.append((k, d))
print("non-reproducible (rel diff > 1e-4):", len(bad), "of", len(a[1]))
for k, d in bad[:60]:
    print("   %-50s %.3e" % (k, d))
c = run(True)
nan = [k for k, g in c[1].items() if not torch.isfinite(g).all()]
print("poisoned run: loss", c[0], "params with non-finite gradients:", len(nan))
for k in nan[:80]:
    print("   ", k)

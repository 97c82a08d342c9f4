"""Debug: which ACTIVATION gradient is the first (in backward order) that is not reproducible between two passes?"""
import os, random, sys
from collections import OrderedDict
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200 import ops
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch

dev = torch.device("cuda:0")
pools, patch, ncls, B = POOLS["hippo"], (40, 56, 40), 3, 1
data, targets = synthetic_batch(B, 1, ncls, patch, pools, seed=1)
x, tg = data.to(dev), [t.to(dev) for t in targets]
ops.CONFIG["fuse_fanin"] = os.environ.get("FANIN", "0") == "1"
ops.CONFIG["fuse_stats"] = os.environ.get("STATS", "0") == "1"
ops.CONFIG["fuse_pool"] = os.environ.get("POOL", "0") == "1"
random.seed(0)
ts = TrainStep(1, ncls, pools, patch, 0.2, 0.5, 1200, dev, 1, seed=0, fused_optimizer=False)


def run():
    ops.DEBUG_GRADS = OrderedDict()
    ts.optimizer.zero_grad()
    l = ts.loss(ts.network(x), tg)
    l.backward()
    torch.cuda.synchronize()
    g = ops.DEBUG_GRADS
    ops.DEBUG_GRADS = None
    return g


a, b = run(), run()
print("order of arrival:", list(a.keys()) == list(b.keys()))
for k in a:
    d = float((a[k].float() - b[k].float()).norm() / b[k].float().norm().clamp_min(1e-30))
    nz = float((a[k] != b[k]).float().mean())
    print("%-28s %-22s rel %.3e  frac differing %.4f" % (k, tuple(a[k].shape), d, nz))

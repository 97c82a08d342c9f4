"""Debug: are the gradients of one training iteration reproducible run to run, and identical between the fused
(arena) and the stock optimizer paths?  Prints per-parameter-kind worst relative differences."""
import os, random, sys
from collections import OrderedDict
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200 import ops
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch

dev = torch.device("cuda:0")
pools, patch = POOLS["hippo"], (40, 56, 40)
data, targets = synthetic_batch(1, 1, 3, patch, pools, seed=1)
x, tg = data.to(dev), [t.to(dev) for t in targets]


def grads(fused, fanin=True):
    ops.CONFIG["fuse_fanin"] = fanin
    random.seed(0)
    ts = TrainStep(1, 3, pools, patch, 0.2, 0.5, 1200, dev, 1, seed=0, fused_optimizer=fused)
    if fused:
        ts._build_arena(x, tg)
        ts.optimizer.zero_grad(set_to_none=True)
        ts.arena.begin_step()
    else:
        ts.optimizer.zero_grad()
    l = ts.loss(ts.network(x), tg)
    l.backward()
    torch.cuda.synchronize()
    return float(l), OrderedDict((k, p.grad.detach().clone()) for k, p in ts.network.named_parameters())


def cmp(tag, a, b):
    worst = {}
    for k in a[1]:
        kind = k.split(".")[-2] + "." + k.split(".")[-1] if "blocks" in k else k.split(".")[0][:3] + ".weight"
        d = float((a[1][k] - b[1][k]).norm() / b[1][k].norm().clamp_min(1e-30))
        if d > worst.get(kind, (0, ""))[0]:
            worst[kind] = (d, k)
    print(tag, "loss", a[0], b[0], {k: ("%.2e" % v[0], v[1]) for k, v in worst.items()}, flush=True)


f1, f2 = grads(True), grads(True)
cmp("fused vs fused   ", f1, f2)
s1, s2 = grads(False), grads(False)
cmp("stock vs stock   ", s1, s2)
cmp("fused vs stock   ", f1, s1)
n1, n2 = grads(False, fanin=False), grads(False, fanin=False)
cmp("nofanin vs nofanin", n1, n2)
cmp("stock vs nofanin ", s1, n1)

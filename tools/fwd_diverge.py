"""Locates run-to-run differences of the forward pass: several TrainStep instances with the same seed take two steps,
then every module's output checksum of a third forward is compared.  python tools/fwd_diverge.py"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("E2E_FILL_NAN", "0") == "1":        # torch.empty() returns NaN-filled memory: uninitialised reads become NaN
    torch.use_deterministic_algorithms(True, warn_only=True)
    torch.utils.deterministic.fill_uninitialized_memory = True
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch  # noqa: E402

dev = torch.device("cuda:0")
pools, patch = POOLS["btcv"], (32, 96, 96)
data, targets = synthetic_batch(2, 1, 14, patch, pools, seed=1)
x, tg = data.to(dev), [t.to(dev) for t in targets]


def make():
    random.seed(0)
    return TrainStep(1, 14, pools, patch, 0.2, 0.5, 1200, dev, 1, seed=0, n_buckets=4)


tss = [make() for _ in range(3)]
for step in range(2):
    for ts in tss:
        ts.step(x, tg)
torch.cuda.synchronize()
sums = []
for ts in tss:
    rec = []

    def hook(name):
        def f(mod, inp, out):
            t = out.parts[0] if hasattr(out, "parts") else (out[0] if isinstance(out, (list, tuple)) else out)
            if torch.is_tensor(t):
                v = t.detach().double()
                rec.append((name, float(v.sum()), float(v.abs().sum())))
        return f

    hs = [m.register_forward_hook(hook(n)) for n, m in ts.network.named_modules() if n and n.count(".") <= 3]
    if os.environ.get("E2E_FWD_GRAD", "0") == "1":       # as in the real step: autograd on, arena armed
        ts.optimizer.zero_grad(set_to_none=True)
        ts.arena.begin_step()
        out = ts.network(x)
        rec.append(("LOSS", float(ts.loss(out, tg)), 0.0))
        del out
    else:
        with torch.no_grad():
            ts.network(x)
    for h in hs:
        h.remove()
    torch.cuda.synchronize()
    sums.append(rec)
print("instance losses of the checked forward:", [[v for n, v, _ in r if n == "LOSS"] for r in sums], flush=True)
print("instance losses of a real third step:   ", [float(ts.step(x, tg)) for ts in tss], flush=True)
ref = sums[0]
for i in (1, 2):
    n = 0
    for (a, s0, q0), (b, s1, q1) in zip(ref, sums[i]):
        assert a == b
        if s0 != s1 or q0 != q1:
            print("instance %d vs 0: %-48s sum %.10g vs %.10g   abs-sum %.10g vs %.10g" % (i, a, s1, s0, q1, q0), flush=True)
            n += 1
            if n >= 6:
                break
    print("instance %d: %d of %d module outputs differ" % (i, sum(1 for u, v in zip(ref, sums[i]) if u != v), len(ref)), flush=True)

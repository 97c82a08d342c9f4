"""Finds reads of uninitialised memory in a training step: torch.empty() is made to return NaN-filled memory
(torch.utils.deterministic.fill_uninitialized_memory), so anything that consumes a never-written element turns into NaN.
python tools/uninit_check.py"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.use_deterministic_algorithms(True, warn_only=True)
torch.utils.deterministic.fill_uninitialized_memory = True
from e2enet_medical_b200 import ops  # noqa: E402
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch  # noqa: E402

dev = torch.device("cuda:0")
pools, patch = POOLS["btcv"], (32, 96, 96)
data, targets = synthetic_batch(2, 1, 14, patch, pools, seed=1)
x, tg = data.to(dev), [t.to(dev) for t in targets]
random.seed(0)
ts = TrainStep(1, 14, pools, patch, 0.2, 0.5, 1200, dev, 1, seed=0, n_buckets=4)
for it in range(3):
    l = float(ts.step(x, tg))
    bad = [k for k, p in ts.network.named_parameters() if not torch.isfinite(p).all()]
    badg = [k for k, p in ts.network.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    print("step", it, "loss", l, "non-finite params:", bad[:6], len(bad), "non-finite grads:", badg[:6], len(badg), flush=True)

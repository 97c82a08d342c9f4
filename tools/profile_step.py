"""Runs warm-up training steps, then ONE step inside a cudaProfilerStart/Stop range (for ncu
--profile-from-start off).  Usage: ncu ... python tools/profile_step.py [--impl 1]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200 import ops  # noqa: E402
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--impl", type=int, default=1)
ap.add_argument("--batch", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
ops.CONFIG["impl"] = a.impl
ts = TrainStep(1, 14, POOLS["btcv"], (64, 160, 160), 0.2, 0.5, 1200, dev, 1, seed=0)
data, targets = synthetic_batch(a.batch, 1, 14, (64, 160, 160), POOLS["btcv"], seed=1)
data, targets = data.to(dev), [t.to(dev) for t in targets]
for _ in range(2):
    ts.step(data, targets)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ts.step(data, targets)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")

"""Measures host (launch/enqueue) time vs device time of one training step: if the host time
is close to the device time the step is launch-bound and wants a CUDA graph."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200 import ops  # noqa: E402
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch  # noqa: E402

dev = torch.device("cuda:0")
ops.CONFIG["impl"] = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ts = TrainStep(1, 14, POOLS["btcv"], (64, 160, 160), 0.2, 0.5, 1200, dev, 1, seed=0)
data, targets = synthetic_batch(2, 1, 14, (64, 160, 160), POOLS["btcv"], seed=1)
data, targets = data.to(dev), [t.to(dev) for t in targets]
for _ in range(3):
    ts.step(data, targets)
torch.cuda.synchronize()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    ts.step(data, targets)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("host enqueue %.1f ms, device %.1f ms" % ((t1 - t0) * 1e3, e0.elapsed_time(e1)))

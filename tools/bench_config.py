"""Times one training step of any BASELINE.json training config on one GPU (CUDA events):
    python tools/bench_config.py --config brats|btcv|hippo [--steps 5]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch  # noqa: E402

CFG = {"btcv": dict(in_ch=1, ncls=14, patch=(64, 160, 160), batch=2),
       "brats": dict(in_ch=4, ncls=4, patch=(128, 128, 128), batch=2),
       "hippo": dict(in_ch=1, ncls=3, patch=(40, 56, 40), batch=1)}

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="brats")
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
c = CFG[a.config]
dev = torch.device("cuda:0")
pools = POOLS[a.config]
ts = TrainStep(c["in_ch"], c["ncls"], pools, c["patch"], 0.2, 0.5, 1200, dev, 1, seed=0)
data, targets = synthetic_batch(c["batch"], c["in_ch"], c["ncls"], c["patch"], pools, seed=1)
data, targets = data.to(dev), [t.to(dev) for t in targets]
for _ in range(3):
    loss = ts.step(data, targets)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = ts.step(data, targets)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print("%s: batch %d of %s, %d classes: %.2f ms/step = %.1f patches/s, loss %.4f, peak mem %.1f GB"
      % (a.config, c["batch"], "x".join(map(str, (c["in_ch"],) + c["patch"])), c["ncls"], ms, c["batch"] / (ms / 1e3),
         float(loss), torch.cuda.max_memory_allocated() / 2 ** 30))

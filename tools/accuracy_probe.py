"""Accuracy probe (GPU box): product network vs the fp32 CPU oracle on several configurations.
Prints per-output logits error (max|a-b|/max|b|), argmax agreement and the distribution of
per-tensor weight-gradient errors."""
import os
import sys
import time
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200.training import POOLS, build_network, multiple_output_loss  # noqa: E402
from oracle import network as onet  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def probe(tag, in_ch, base, ncls, pools, patch, B=1, seed=3, init="det"):
    torch.manual_seed(seed)
    net = build_network(in_ch, ncls, pools, patch, base)
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    if init == "det":
        params = onet.det_params(shapes, seed=seed)
        net.load_state_dict(params, strict=True)
    else:
        params = OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
    net = net.to(dev)
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.rand(B, in_ch, *patch).astype(np.float32))
    tg, sp = [], np.array(patch)
    for k in range(4):
        tg.append(torch.from_numpy(np.round(rs.rand(B, 1, *sp) * (ncls - 1)).astype(np.float32)))
        sp = sp // np.array(pools[k])
    t0 = time.time()
    outs = net(x.to(dev))
    loss = multiple_output_loss(outs, [t.to(dev) for t in tg])
    loss.backward()
    torch.cuda.synchronize()
    t1 = time.time()
    ref_p = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in params.items())
    torch.set_num_threads(os.cpu_count())
    ref_outs = onet.unetpp_forward(ref_p, x, pools)
    ref_loss = onet.ds_loss(ref_outs, tg)
    ref_loss.backward()
    t2 = time.time()
    lg = [rel(o, r) for o, r in zip(outs, ref_outs)]
    agree = float((outs[0].argmax(1).cpu() == ref_outs[0].argmax(1)).float().mean())
    prm = dict(net.named_parameters())
    errs, errs2 = {}, {}
    for k, v in ref_p.items():
        if k.endswith("conv.bias"):
            continue
        errs[k] = rel(prm[k].grad, v.grad)
        errs2[k] = rel2(prm[k].grad, v.grad)
    e = np.array(list(errs.values()))
    e2 = np.array(list(errs2.values()))
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print(f"[{tag}] loss {loss.item():.5f} vs {ref_loss.item():.5f} | logits maxrel {['%.3e' % v for v in lg]} argmax agree {agree:.5f} | "
          f"wgrad maxrel: median {np.median(e):.3e} p90 {np.percentile(e, 90):.3e} max {e.max():.3e}; L2rel median {np.median(e2):.3e} "
          f"max {e2.max():.3e} | gpu {t1 - t0:.2f}s cpu {t2 - t1:.2f}s", flush=True)
    print("    worst:", [(k, "%.3e" % v) for k, v in worst], flush=True)


if __name__ == "__main__":
    probe("small base8 32x64x64", 1, 8, 3, POOLS["btcv"], (32, 64, 64))
    probe("hippo cfg1 base48 40x56x40", 1, 48, 3, POOLS["hippo"], (40, 56, 40), seed=0)
    probe("hippo cfg1 He-init", 1, 48, 3, POOLS["hippo"], (40, 56, 40), seed=0, init="he")
    probe("btcv base48 32x96x96", 1, 48, 14, POOLS["btcv"], (32, 96, 96), seed=0)
    probe("brats base48 64x64x64", 4, 48, 4, POOLS["brats"], (64, 64, 64), seed=0)

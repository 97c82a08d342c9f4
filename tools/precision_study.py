"""Precision study (GPU box): how far are torch's own reduced-precision paths (autocast bf16 / fp16,
i.e. what the reference does under AMP) from the fp32 path on this network, compared with our kernels?
Establishes the achievable tolerance for bf16 activations on the random-init parity setup."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from e2enet_medical_b200.training import POOLS, build_network, multiple_output_loss  # noqa: E402
from oracle import network as onet  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel2(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_oracle(params, x, tg, pools, mode):
    p = OrderedDict((k, v.clone().to(dev).requires_grad_(True)) for k, v in params.items())
    if mode == "fp32":
        outs = onet.unetpp_forward(p, x, pools)
    elif mode == "fp64":
        p = OrderedDict((k, v.clone().to(dev).double().requires_grad_(True)) for k, v in params.items())
        outs = onet.unetpp_forward(p, x.double(), pools)
    else:
        with torch.autocast("cuda", dtype=torch.bfloat16 if mode == "bf16" else torch.float16):
            outs = onet.unetpp_forward(p, x, pools)
    outs = [o.float() if mode != "fp64" else o for o in outs]
    loss = onet.ds_loss([o.float() for o in outs], tg)
    scale = 1.0 if mode != "fp16" else 1024.0
    (loss * scale).backward()
    grads = OrderedDict((k, (v.grad / scale).double()) for k, v in p.items())
    return outs, grads, float(loss)


def study(tag, in_ch, base, ncls, pools, patch, seed=0, init="he", targets="random"):
    torch.manual_seed(seed)
    net = build_network(in_ch, ncls, pools, patch, base)
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    if init == "det":
        net.load_state_dict(onet.det_params(shapes, seed=seed), strict=True)
    params = OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
    net = net.to(dev)
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.rand(1, in_ch, *patch).astype(np.float32)).to(dev)
    tg, sp = [], np.array(patch)
    for k in range(4):
        tg.append(torch.from_numpy(np.round(rs.rand(1, 1, *sp) * (ncls - 1)).astype(np.float32)).to(dev))
        sp = sp // np.array(pools[k])
    ref_o, ref_g, ref_l = run_oracle(params, x, tg, pools, "fp64")
    rows = []
    for mode in ("fp32", "fp16", "bf16"):
        o, g, l = run_oracle(params, x, tg, pools, mode)
        rows.append((mode, o, g, l))
    outs = net(x)
    loss = multiple_output_loss(outs, tg)
    loss.backward()
    rows.append(("ours", outs, OrderedDict((k, v.grad.double()) for k, v in net.named_parameters()), float(loss)))
    print(f"== {tag} (reference: fp64 torch on GPU, loss {ref_l:.6f})")
    for mode, o, g, l in rows:
        lg = [rel(a, b) for a, b in zip(o, ref_o)]
        agree = float((o[0].argmax(1) == ref_o[0].argmax(1)).float().mean())
        keys = [k for k in ref_g if not k.endswith("conv.bias")]
        e = np.array([rel(g[k], ref_g[k]) for k in keys])
        e2 = np.array([rel2(g[k], ref_g[k]) for k in keys])
        allg = rel2(torch.cat([g[k].flatten() for k in keys]), torch.cat([ref_g[k].flatten() for k in keys]))
        print(f"   {mode:5s} loss {l:.6f} logits maxrel {max(lg):.2e} argmax {agree:.5f} | wgrad maxrel med {np.median(e):.2e} "
              f"max {e.max():.2e} | L2rel med {np.median(e2):.2e} max {e2.max():.2e} | all-params L2rel {allg:.2e}", flush=True)


if __name__ == "__main__":
    study("hippo base48 40x56x40 He init", 1, 48, 3, POOLS["hippo"], (40, 56, 40))
    study("btcv base48 32x96x96 He init", 1, 48, 14, POOLS["btcv"], (32, 96, 96))
    study("btcv base48 64x160x160 He init (full config-2 patch, B=1)", 1, 48, 14, POOLS["btcv"], (64, 160, 160))

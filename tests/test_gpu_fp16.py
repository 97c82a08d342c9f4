"""The fp16 build of the library (libe2enet_b200_fp16.so = the same sources with -DE2E_FP16).

north_star names bf16, and bf16 is the default everywhere.  The reference itself ships fp16 AMP
(`torch.cuda.amp.autocast` + `GradScaler`, nnUNetTrainer_simple.py:552-562), whose 11-bit significand is what
north_star's 2e-2 tolerance was written against.  `ops.set_precision("fp16")` (or E2E_PRECISION=fp16) switches the
16-bit type of activations, gradients and packed weights to IEEE fp16 -- same kernels, same tcgen05 kind::f16 rate,
same C8 layout -- and training then carries a loss scale (device state of FusedSGD, or the reference loop's own
GradScaler).  These tests show what that mode achieves against the same oracles as the bf16 tests:

  * kernel level: 1 fp16 ulp (2^-11) against torch fp32 math on identical operands;
  * the reference's golden network: logits within north_star's 2e-2;
  * config 2 at full size against the fp32 GPU oracle, recorded next to the bf16 rows in parity_fullsize.json;
  * the loss scale: skip + back-off on overflow, growth after clean steps, inside a captured CUDA graph;
  * the literal reference loop (autocast + GradScaler + mask.step()) through the drop-in modules.
"""
import json
import os
import random
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import network as onet

import test_gpu_oracle_fullsize as FS
import test_gpu_tcgen05 as K
from test_gpu_boundary import reference_names  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu
SCALE = 65536.0                      # GradScaler's initial scale


@pytest.fixture
def fp16():
    assert torch.cuda.is_available()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from e2enet_medical_b200 import _lib, ops
    ops.set_precision("fp16")
    assert _lib.load().e2e_precision().decode() == "fp16" and _lib.act_dtype() == torch.float16
    yield torch.device("cuda:0")
    ops.set_precision("bf16")


def test_fp16_kernels_one_ulp_vs_torch(fp16):
    """the tcgen05 / mma.sync GEMM kernels, the fused statistics and the transposed conv under fp16 operands: the
    bf16 tests' bodies with the rounding function and the ulp of the current precision (2^-11)"""
    assert K._ulp() == 2.0 ** -11
    K.test_tcgen05_conv_matches_mma_sync_and_torch([48, 48], 48, (5, 37, 45), 1)
    K.test_tcgen05_conv_matches_mma_sync_and_torch([96, 96, 48], 96, (5, 12, 16), 2)
    K.test_tcgen05_conv_matches_mma_sync_and_torch([1], 48, (6, 16, 16), 2)
    K.test_tcgen05_point_form_tconv(96, 48, (1, 2, 2), (3, 20, 40), 2)
    K.test_tcgen05_point_form_tconv(192, 96, (2, 2, 2), (2, 8, 8), 1)
    K.test_tcgen05_kw_stacked_forward([48, 48], 48, (3, 9, 70), 2)
    K.test_fused_epilogue_instancenorm_statistics([48, 48], 48, (1, 1, 1), (3, 9, 70), 2, True)
    K.test_fused_epilogue_instancenorm_statistics([48], 96, (1, 2, 2), (6, 16, 16), 2, False)


def test_fp16_network_vs_reference_golden(fp16, golden_dir):
    """whole network vs tests/golden/net_small.npz (the unmodified reference, fp32 CPU): north_star's 2e-2 on the
    logits holds in this mode; gradients are compared with what torch's own fp16 autocast reaches"""
    from test_gpu_parity import build_net, rel, rel2
    dev = fp16
    g = np.load(os.path.join(golden_dir, "net_small.npz"))
    cfg = json.load(open(os.path.join(golden_dir, "net_small_grads.json")))["config"]
    net = build_net(cfg["in_ch"], cfg["base"], cfg["ncls"], cfg["pools"], tuple(cfg["patch"]))
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    params = onet.det_params(shapes, seed=cfg["seed"])
    net.load_state_dict(params, strict=True)
    net = net.to(dev)
    x = torch.from_numpy(g["x"])
    tg = [torch.from_numpy(g[f"tgt{k}"].astype(np.float32)) for k in range(4)]
    outs = net(x.to(dev))
    errs = [rel(o, g[f"out{k}"]) for k, o in enumerate(outs)]
    assert max(errs) < 2e-2, errs                                   # north_star's logits tolerance
    agree = float((outs[0].argmax(1).cpu().numpy() == g["out0"].argmax(1)).mean())
    loss = onet.ds_loss(outs, [t.to(dev) for t in tg])
    assert abs(loss.item() - float(g["loss"])) < 2e-4 * abs(float(g["loss"]))
    (loss * SCALE).backward()
    # torch's fp16 autocast on the same weights / inputs (what the reference ships)
    p = OrderedDict((k, v.clone().to(dev).requires_grad_(True)) for k, v in params.items())
    with torch.autocast("cuda", dtype=torch.float16):
        ac = onet.unetpp_forward(p, x.to(dev), cfg["pools"])
    ac = [o.float() for o in ac]
    (onet.ds_loss(ac, [t.to(dev) for t in tg]) * SCALE).backward()
    agree_ac = float((ac[0].argmax(1).cpu().numpy() == g["out0"].argmax(1)).mean())
    assert agree > 0.99 and agree > agree_ac - 0.005, (agree, agree_ac)
    prm = dict(net.named_parameters())
    worst = 0.0
    for key in g.files:
        if key.startswith("grad:") and not key.endswith("conv.bias"):
            gr = prm[key[5:]].grad / SCALE
            assert torch.isfinite(gr).all(), key
            e, e_ac = rel2(gr, g[key]), rel2(p[key[5:]].grad / SCALE, g[key])
            worst = max(worst, e)
            assert e < 0.15 and e < 1.25 * e_ac + 1e-2, (key, e, e_ac)
    FS._record("golden_net_small_fp16", {"logits_maxrel": errs, "argmax_agree": agree, "argmax_agree_torch_fp16": agree_ac,
                                         "wgrad_L2rel_max": worst})


def test_fp16_fullsize_config2_vs_gpu_oracle(fp16):
    """BASELINE config 2 (B=2, 1x64x160x160, 14 classes, density 0.2), forward + DS loss + backward, vs the fp32 oracle
    on the GPU; torch autocast rows measured live next to it.  Written to gpurun_out/parity_fullsize.json."""
    from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch
    dev = fp16
    pools, patch, in_ch, ncls = POOLS["btcv"], (64, 160, 160), 1, 14
    random.seed(0)
    ts = TrainStep(in_ch, ncls, pools, patch, 0.2, 0.5, 1200, dev, 1, seed=0)
    params = OrderedDict((k, v.detach().clone()) for k, v in ts.network.state_dict().items())
    data, targets = synthetic_batch(2, in_ch, ncls, patch, pools, seed=1)
    x, tg = data.to(dev), [t.to(dev) for t in targets]
    outs = ts.network(x)
    loss = onet.ds_loss([o.float() for o in outs], tg)
    (loss * SCALE).backward()
    mine_o = [o.detach().float() for o in outs]
    mine_g = OrderedDict((k, v.grad.detach().clone() / SCALE) for k, v in ts.network.named_parameters())
    assert all(torch.isfinite(v).all() for v in mine_g.values())
    mine_l = float(loss)
    del outs, loss
    ts.optimizer.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()
    ref_o, ref_g, ref_l = FS._oracle_run(params, x, tg, pools, "fp32")
    rows = {"ours_fp16_tcgen05": FS._summary(mine_o, mine_g, mine_l, ref_o, ref_g, ref_l)}
    torch.cuda.empty_cache()
    o, g, l = FS._oracle_run(params, x, tg, pools, "fp16")
    rows["torch_autocast_fp16"] = FS._summary(o, g, l, ref_o, ref_g, ref_l)
    del o, g
    FS._record("config2_btcv_B2_64x160x160_fp16_mode", rows)
    me, ac = rows["ours_fp16_tcgen05"], rows["torch_autocast_fp16"]
    assert max(me["logits_maxrel"]) < 2e-2, rows                     # north_star's logits tolerance at full size
    assert max(me["logits_maxrel"]) < 1.5 * max(ac["logits_maxrel"]) + 2e-3, rows
    assert me["loss_rel"] < 5e-4, rows
    assert me["argmax_agree"] > 0.99 and me["argmax_agree"] > ac["argmax_agree"] - 0.005, rows
    assert me["wgrad_L2rel_all_params"] < 0.1 and me["wgrad_L2rel_all_params"] < 1.5 * ac["wgrad_L2rel_all_params"] + 1e-2, rows


def test_fp16_train_step_loss_scale_and_graph(fp16):
    """TrainStep under fp16: the loss scale lives on the device (GradScaler semantics: skip + halve on inf / nan,
    double after `growth_interval` clean steps), the trajectory follows the bf16 one, and all of it replays as one
    CUDA graph"""
    from e2enet_medical_b200 import ops
    from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch
    dev = fp16
    pools, patch = POOLS["hippo"], (40, 56, 40)
    data, targets = synthetic_batch(1, 1, 3, patch, pools, seed=1)
    data, targets = data.to(dev), [t.to(dev) for t in targets]

    def make():
        random.seed(0)
        return TrainStep(1, 3, pools, patch, 0.3, 0.5, 1200, dev, 1, seed=0, base=16)

    ts = make()
    opt = ts.optimizer
    assert opt.loss_scale() is not None and float(opt.loss_scale()) == 65536.0
    opt.enable_loss_scale(init_scale=65536.0, growth_interval=3)
    losses = [float(ts.step(data, targets)) for _ in range(2)]
    assert all(np.isfinite(losses)) and float(opt._scaler[1]) == 2.0 and float(opt.loss_scale()) == 65536.0
    assert float(opt._norm_coef[2]) == 0.0 and abs(float(opt._norm_coef[3]) - 1.0 / 65536.0) < 1e-12
    ts.step(data, targets)                                   # third clean step: the scale doubles
    assert float(opt.loss_scale()) == 131072.0 and float(opt._scaler[1]) == 0.0
    # overflow: an absurd scale makes the fp16 gradients infinite -> the step is skipped and the scale backs off
    w0 = [p.detach().clone() for p in ts.network.parameters()]
    opt._scaler[0] = 2.0 ** 60
    ts.step(data, targets)
    assert float(opt._norm_coef[2]) == 1.0 and float(opt.loss_scale()) == 2.0 ** 59
    assert all(torch.equal(a, p.detach()) for a, p in zip(w0, ts.network.parameters()))
    opt._scaler[0] = 65536.0
    # bf16 run of the same configuration: same trajectory within the precision of the two formats
    ops.set_precision("bf16")
    tb = make()
    lb = [float(tb.step(data, targets)) for _ in range(2)]
    ops.set_precision("fp16")
    for a, b in zip(losses, lb):
        assert abs(a - b) < 2e-2 * abs(b), (losses, lb)
    # whole-step graph with the scaler inside
    ts2 = make()
    ts2.optimizer.enable_loss_scale(init_scale=65536.0, growth_interval=4)
    eager = [float(ts2.step(data, targets)) for _ in range(1)]
    ts2.enable_graph(data, targets, warmup=1)
    n_before = float(ts2.optimizer._scaler[1]) + 4 * np.log2(float(ts2.optimizer.loss_scale()) / 65536.0)
    lg = [float(ts2.step(data, targets)) for _ in range(3)]
    n_after = float(ts2.optimizer._scaler[1]) + 4 * np.log2(float(ts2.optimizer.loss_scale()) / 65536.0)
    assert all(np.isfinite(lg)) and n_after - n_before == 3.0, (n_before, n_after)     # three clean replays counted
    assert lg[-1] < eager[0], (eager, lg)                    # it trains


def test_fp16_reference_loop_autocast_gradscaler(fp16, reference_names):
    """the literal reference iteration (autocast + GradScaler.scale / unscale_ / step / update + mask.step(), with a
    prune / regrow update and a checkpoint round trip) -- tests/test_gpu_boundary.py's body -- on the fp16 library:
    here the GradScaler's scaled loss really flows through fp16 gradients, as in the reference"""
    import test_gpu_boundary as TB
    TB.test_reference_loop_through_aliases(reference_names)


def test_fp16_sliding_window_inference(fp16):
    """config 3's path on the fp16 library: predict_3D (Gaussian, step 0.5, head folded into the accumulate kernel)
    agrees with the bf16 library to the precision of the two formats and returns normalised probabilities"""
    from e2enet_medical_b200 import ops
    from e2enet_medical_b200.network_architecture.unetpp_d import softmax_helper
    from e2enet_medical_b200.training import POOLS, build_network
    dev = fp16
    pools, patch, ncls = POOLS["btcv"], (32, 64, 64), 5
    vol = np.random.RandomState(0).randn(1, 40, 90, 100).astype(np.float32)
    res = {}
    for prec in ("fp16", "bf16"):
        ops.set_precision(prec)
        torch.manual_seed(0)
        net = build_network(1, ncls, pools, patch, 16, deep_supervision=True).to(dev).eval()
        net.do_ds = False
        net.inference_apply_nonlin = softmax_helper
        seg, probs = net.predict_3D(vol, False, (0, 1, 2), True, 0.5, patch, None, True, "constant", None, False, False, True)
        assert probs.shape == (ncls,) + vol.shape[1:] and seg.shape == vol.shape[1:]
        assert np.isfinite(probs).all() and float(np.abs(probs.sum(0) - 1).max()) < 1e-4
        res[prec] = (seg, probs)
    ops.set_precision("fp16")
    assert float(np.abs(res["fp16"][1] - res["bf16"][1]).max()) < 0.1
    assert float((res["fp16"][0] == res["bf16"][0]).mean()) > 0.97

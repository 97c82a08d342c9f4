"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference
modules (imported from /root/reference through _ref_shim) on CPU in fp32.

Run in the build container only:   cd /tmp && python /root/repo/tests/golden/make_golden.py
(the reference's e2enet/paths.py may create folders in the CWD, hence /tmp).
The outputs are committed; nothing on the GPU box reads /root/reference.
"""
import argparse
import hashlib
import json
import os
import random
import sys
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _ref_shim  # noqa: E402
from oracle import network as onet  # noqa: E402  (only det_params / param_shapes: the shared input generator)

unetpp_d, neural_network, core_channel = _ref_shim.import_reference()

POOLS_BTCV = [[1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2]]
POOLS_HIPPO = [[2, 2, 2]] * 3 + [[1, 1, 1]] * 2


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_ref_net(in_ch, base, ncls, pools, patch):
    net = unetpp_d.Generic_UNetPlusPlus(
        patch, in_ch, base, ncls, len(pools), 2, 2, nn.Conv3d, nn.InstanceNorm3d,
        {'eps': 1e-5, 'affine': True}, nn.Dropout3d, {'p': 0, 'inplace': True},
        nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True}, True, False, lambda x: x,
        unetpp_d.InitWeights_He(1e-2), pools, None, False, True, True)
    return net


# ------------------------------------------------------------------------------------
def gold_shift(out):
    rs = np.random.RandomState(7)
    d = {}
    for C in (1, 4, 6, 48, 96, 167):
        x = rs.standard_normal((2, C, 7, 3, 4)).astype(np.float32)
        y = unetpp_d.torch_shift(5, 2, 3)(torch.from_numpy(x)).numpy()
        d[f"x{C}"] = x
        d[f"y{C}"] = y
    np.savez_compressed(os.path.join(out, "shift.npz"), **d)


def gold_block(out):
    d = {}
    for tag, cin, cout, stride, shp in (("s1", 20, 8, (1, 1, 1), (2, 20, 6, 9, 10)),
                                        ("s2", 12, 16, (2, 2, 2), (1, 12, 8, 10, 12)),
                                        ("s122", 5, 8, (1, 2, 2), (1, 5, 5, 8, 8))):
        rs = np.random.RandomState(11)
        blk = unetpp_d.ConvDropoutNormNonlin(
            cin, cout, nn.Conv3d, {'kernel_size': (1, 3, 3), 'stride': stride, 'padding': (0, 1, 1),
                                   'dilation': 1, 'bias': True},
            nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True}, nn.Dropout3d, {'p': 0, 'inplace': True},
            nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True})
        sd = OrderedDict()
        sd["conv.weight"] = torch.from_numpy(rs.standard_normal((cout, cin, 1, 3, 3)).astype(np.float32) * 0.2)
        sd["conv.bias"] = torch.from_numpy(rs.standard_normal(cout).astype(np.float32) * 0.1)
        sd["instnorm.weight"] = torch.from_numpy(1 + 0.1 * rs.standard_normal(cout).astype(np.float32))
        sd["instnorm.bias"] = torch.from_numpy(0.1 * rs.standard_normal(cout).astype(np.float32))
        blk.load_state_dict(sd, strict=True)
        x = torch.from_numpy(rs.standard_normal(shp).astype(np.float32)).requires_grad_(True)
        y = blk(x)
        gy = torch.from_numpy(rs.standard_normal(tuple(y.shape)).astype(np.float32))
        (y * gy).sum().backward()
        d[f"{tag}_x"] = x.detach().numpy()
        d[f"{tag}_gy"] = gy.numpy()
        d[f"{tag}_y"] = y.detach().numpy()
        d[f"{tag}_gx"] = x.grad.numpy()
        for k, v in sd.items():
            d[f"{tag}_{k}"] = v.numpy()
        for k, prm in blk.named_parameters():
            d[f"{tag}_g_{k}"] = prm.grad.numpy()
    np.savez_compressed(os.path.join(out, "block.npz"), **d)


def gold_net(out):
    # (a) parameter inventory of the real configs (names, shapes, registration order)
    inv = {}
    for tag, in_ch, ncls, pools, patch in (("btcv", 1, 14, POOLS_BTCV, (64, 160, 160)),
                                           ("brats", 4, 4, [[2, 2, 2]] * 5, (128, 128, 128)),
                                           ("hippo", 1, 3, POOLS_HIPPO, (40, 56, 40))):
        with torch.device("meta"):
            net = build_ref_net(in_ch, 48, ncls, pools, patch)
        inv[tag] = [[k, list(v.shape)] for k, v in net.state_dict().items()]
        inv[tag + "_named_parameters"] = [k for k, _ in net.named_parameters()]
    with open(os.path.join(out, "param_inventory.json"), "w") as f:
        json.dump(inv, f)

    # (b) small net forward + backward, deterministic params
    in_ch, base, ncls, pools, patch = 1, 8, 3, POOLS_BTCV, (32, 64, 64)
    net = build_ref_net(in_ch, base, ncls, pools, patch)
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    params = onet.det_params(shapes, seed=3)
    net.load_state_dict(params, strict=True)
    rs = np.random.RandomState(5)
    x = rs.rand(1, in_ch, *patch).astype(np.float32)
    outs = net(torch.from_numpy(x))
    tg = []
    for o in outs:
        tg.append(np.round(rs.rand(o.shape[0], 1, *o.shape[2:]) * (ncls - 1)).astype(np.float32))
    loss = onet.ds_loss(outs, [torch.from_numpy(t) for t in tg])
    loss.backward()
    d = {"x": x, "loss": np.float32(loss.item())}
    for k, o in enumerate(outs):
        d[f"out{k}"] = o.detach().numpy()
        d[f"tgt{k}"] = tg[k].astype(np.uint8)
    gsum = {}
    for k, prm in net.named_parameters():
        g = prm.grad.numpy()
        gsum[k] = [float(np.abs(g).max()), float(np.sqrt((g.astype(np.float64) ** 2).sum())), float(g.sum(dtype=np.float64))]
    for k in ("seg_outputs.0.weight", "loc4.0.0.blocks.0.conv.weight", "up4.0.weight",
              "conv_blocks_context.0.blocks.0.conv.weight", "conv_blocks_context.1.blocks.0.conv.weight",
              "loc0.4.1.blocks.0.instnorm.weight", "loc0.4.1.blocks.0.conv.bias",
              "loc3.0.0.blocks.0.conv.weight", "up3.0.weight"):
        d["grad:" + k] = dict(net.named_parameters())[k].grad.numpy()
    np.savez_compressed(os.path.join(out, "net_small.npz"), **d)
    with open(os.path.join(out, "net_small_grads.json"), "w") as f:
        json.dump({"config": dict(in_ch=in_ch, base=base, ncls=ncls, pools=pools, patch=patch, seed=3),
                   "grads": gsum}, f)


class _Args:
    adv = False
    fix = False
    update_frequency = 1
    final_density = 0.05


def gold_masking(out):
    res = {}
    in_ch, base, ncls, pools, patch = 1, 48, 14, POOLS_BTCV, (64, 160, 160)
    net = build_ref_net(in_ch, base, ncls, pools, patch)
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    for quant in (False, True):
        for density in (0.1, 0.2, 0.5):
            params = onet.det_params(shapes, seed=9)
            if quant:
                for k in params:
                    params[k] = torch.round(params[k] * 1024) / 1024
            net.load_state_dict(params, strict=True)
            opt = torch.optim.SGD(net.parameters(), 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
            rs = np.random.RandomState(13)
            for prm in net.parameters():
                opt.state[prm]['momentum_buffer'] = torch.from_numpy(
                    rs.standard_normal(tuple(prm.shape)).astype(np.float32))
            decay = core_channel.CosineDecay(0.5, 1000)
            mask = core_channel.Masking(opt, death_rate=0.5, death_mode='magnitude',
                                        death_rate_decay=decay, growth_mode='random',
                                        redistribution_mode='none', args=_Args())
            random.seed(0)
            mask.add_module(net, sparse_init='uniform', density=density)
            rec = {"names": list(mask.masks.keys())}
            rec["init"] = {k: sha(v.numpy()[:, :, 0, 0, 0].astype(np.uint8)) for k, v in mask.masks.items()}
            rec["init_nnz"] = {k: int(v.sum().item()) for k, v in mask.masks.items()}
            # weights after init == params * mask; take an optimizer-like perturbation so that
            # masked weights are non-zero again before step() (what SGD does with dense grads, H3)
            rs2 = np.random.RandomState(17)
            with torch.no_grad():
                for k, prm in net.named_parameters():
                    if k in mask.masks:
                        pert = torch.from_numpy(rs2.standard_normal(tuple(prm.shape)).astype(np.float32)) * 1e-3
                        if quant:
                            pert = torch.round(pert * 1024 * 64) / (1024 * 64)
                        prm.add_(pert)
            random.seed(1)
            mask.step()
            rec["death_rate"] = float(mask.death_rate)
            rec["after"] = {k: sha(v.numpy()[:, :, 0, 0, 0].astype(np.uint8)) for k, v in mask.masks.items()}
            rec["after_nnz"] = {k: int(v.sum().item()) for k, v in mask.masks.items()}
            rec["pruned"] = {k: sha(v.numpy()[:, :, 0, 0, 0].astype(np.uint8)) for k, v in mask.pruned_masks.items()}
            rec["num_death"] = {k: int(v) for k, v in mask.num_death.items()}
            rec["num_remove"] = {k: int(v) for k, v in mask.num_remove.items()}
            rec["fired_nnz"] = {k: int(v.sum().item()) for k, v in mask.fired_masks.items()}
            rec["total_nozeros"] = int(mask.total_nozeros)
            rec["total_weights"] = int(mask.total_weights)
            # weight / momentum checksums of two tensors after the final apply_mask
            sd = dict(net.named_parameters())
            for k in ("loc4.0.0.blocks.0.conv.weight", "up0.0.weight"):
                rec["w_sha:" + k] = sha(sd[k].detach().numpy())
                rec["m_sha:" + k] = sha(opt.state[sd[k]]['momentum_buffer'].numpy())
            # a second step exercises ties: newly grown kernels are exactly 0 (H7)
            random.seed(2)
            mask.step()
            rec["after2"] = {k: sha(v.numpy()[:, :, 0, 0, 0].astype(np.uint8)) for k, v in mask.masks.items()}
            rec["after2_nnz"] = {k: int(v.sum().item()) for k, v in mask.masks.items()}
            rec["num_death2"] = {k: int(v) for k, v in mask.num_death.items()}
            res[f"{'quant' if quant else 'raw'}_{density}"] = rec
            print("masking", quant, density, rec["total_nozeros"], flush=True)
    with open(os.path.join(out, "masking.json"), "w") as f:
        json.dump(res, f)


class _ToyNet(neural_network.SegmentationNetwork):
    """cheap deterministic 'network' so the reference's predict_3D can be run on CPU"""

    def __init__(self, ncls):
        super().__init__()
        self.conv_op = nn.Conv3d
        self.num_classes = ncls
        self.inference_apply_nonlin = lambda x: torch.softmax(x, 1)
        self.w = nn.Parameter(torch.linspace(-1.5, 2.0, ncls).view(1, ncls, 1, 1, 1), requires_grad=False)
        self.b = nn.Parameter(torch.linspace(0.3, -0.4, ncls).view(1, ncls, 1, 1, 1), requires_grad=False)

    def forward(self, x):
        # position dependent inside the patch so that tiling/mirroring matter
        X, Y, Z = x.shape[2:]
        gx = torch.linspace(-1, 1, X).view(1, 1, X, 1, 1)
        gy = torch.linspace(-1, 1, Y).view(1, 1, 1, Y, 1)
        gz = torch.linspace(-1, 1, Z).view(1, 1, 1, 1, Z)
        s = x[:, :1] * self.w + self.b
        return s + 0.5 * gx * self.w.flip(1) + 0.25 * gy * gz * self.b


def gold_window(out):
    d = {}
    ncls = 3
    net = _ToyNet(ncls).eval()
    rs = np.random.RandomState(21)
    for tag, vol, patch, mirror in (("a", (50, 70, 60), (32, 48, 32), False),
                                    ("b", (40, 50, 44), (32, 48, 32), True),
                                    ("pad", (20, 60, 40), (32, 48, 32), False),
                                    ("one", (32, 48, 32), (32, 48, 32), False)):
        x = rs.standard_normal((1,) + vol).astype(np.float32)
        net._gaussian_3d = None
        seg, prob = net.predict_3D(x, mirror, (0, 1, 2), True, 0.5, patch, None, True, "constant",
                                   {'constant_values': 0}, False, False, False)
        d[f"{tag}_x"] = x
        d[f"{tag}_seg"] = seg.astype(np.uint8)
        d[f"{tag}_prob"] = prob[:, ::3, ::3, ::3].copy()
    g = neural_network.SegmentationNetwork._get_gaussian((16, 24, 20), 1. / 8)
    d["gauss_16_24_20"] = g
    g = neural_network.SegmentationNetwork._get_gaussian((64, 160, 160), 1. / 8)
    d["gauss_64_160_160_stats"] = np.array([g.min(), g.max(), g.sum(dtype=np.float64), g[0, 0, 0], g[32, 80, 80],
                                            g[10, 33, 150]], dtype=np.float64)
    np.savez_compressed(os.path.join(out, "window.npz"), **d)
    steps = {}
    for patch, image, st in (((64, 160, 160), (300, 512, 512), 0.5), ((128, 128, 128), (155, 240, 240), 0.5),
                             ((40, 56, 40), (40, 56, 40), 0.5), ((32, 48, 32), (50, 70, 60), 0.5),
                             ((64, 130), (128, 260), 0.5), ((128, 128, 128), (424, 456, 456), 0.5)):
        steps[f"{patch}|{image}|{st}"] = neural_network.SegmentationNetwork._compute_steps_for_sliding_window(
            patch, image, st)
    with open(os.path.join(out, "window_steps.json"), "w") as f:
        json.dump(steps, f)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.manual_seed(0)
    todo = a.only.split(",") if a.only else ["shift", "block", "net", "window", "masking"]
    for name in todo:
        print("== golden:", name, flush=True)
        globals()["gold_" + name](HERE)

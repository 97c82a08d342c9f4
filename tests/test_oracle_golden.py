"""Pins the CPU oracle (oracle/) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py) and against the reference's own known-answer vectors.
CPU only."""
import hashlib
import json
import os
import random
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import masking as omask
from oracle import network as onet
from oracle import window as owin

POOLS_BTCV = [[1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2]]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- shift / block / network
def test_shift_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "shift.npz"))
    for C in (1, 4, 6, 48, 96, 167):
        y = onet.shift_depth(torch.from_numpy(g[f"x{C}"])).numpy()
        assert np.array_equal(y, g[f"y{C}"]), C


def test_shift_group_geometry():
    # H2: g = ceil(C/5); shift = -2 + c // g
    assert onet.shift_groups(1) == [(0, 1, -2)]
    assert onet.shift_groups(4) == [(0, 1, -2), (1, 2, -1), (2, 3, 0), (3, 4, 1)]
    assert [hi - lo for lo, hi, _ in onet.shift_groups(832)] == [167, 167, 167, 167, 164]
    assert [s for _, _, s in onet.shift_groups(96)] == [-2, -1, 0, 1, 2]


@pytest.mark.parametrize("tag,stride", [("s1", (1, 1, 1)), ("s2", (2, 2, 2)), ("s122", (1, 2, 2))])
def test_block_fwd_bwd_matches_reference(golden_dir, tag, stride):
    g = np.load(os.path.join(golden_dir, "block.npz"))
    t = lambda k: torch.from_numpy(g[f"{tag}_{k}"]).clone()
    x = t("x").requires_grad_(True)
    prm = {k: t(k).requires_grad_(True) for k in ("conv.weight", "conv.bias", "instnorm.weight", "instnorm.bias")}
    y = onet.shiftconv_block(x, prm["conv.weight"], prm["conv.bias"], prm["instnorm.weight"],
                             prm["instnorm.bias"], stride)
    (y * t("gy")).sum().backward()
    np.testing.assert_allclose(y.detach().numpy(), g[f"{tag}_y"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(x.grad.numpy(), g[f"{tag}_gx"], rtol=1e-3, atol=2e-5)
    for k, v in prm.items():
        ref = g[f"{tag}_g_{k}"]
        np.testing.assert_allclose(v.grad.numpy(), ref, rtol=1e-3, atol=1e-4 * max(1.0, np.abs(ref).max()))


def test_param_inventory_matches_reference(golden_dir):
    inv = json.load(open(os.path.join(golden_dir, "param_inventory.json")))
    for tag, in_ch, ncls, pools in (("btcv", 1, 14, POOLS_BTCV), ("brats", 4, 4, [[2, 2, 2]] * 5),
                                    ("hippo", 1, 3, [[2, 2, 2]] * 3 + [[1, 1, 1]] * 2)):
        mine = onet.param_shapes(in_ch, 48, ncls, pools)
        assert [[k, list(v)] for k, v in mine.items()] == inv[tag], tag
        assert list(mine.keys()) == inv[tag + "_named_parameters"]
    n = sum(int(np.prod(s)) for _, s in inv["btcv"])
    assert n == 23805616


def test_network_fwd_bwd_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "net_small.npz"))
    meta = json.load(open(os.path.join(golden_dir, "net_small_grads.json")))
    cfg = meta["config"]
    shapes = onet.param_shapes(cfg["in_ch"], cfg["base"], cfg["ncls"], cfg["pools"])
    p = onet.det_params(shapes, seed=cfg["seed"])
    for v in p.values():
        v.requires_grad_(True)
    outs = onet.unetpp_forward(p, torch.from_numpy(g["x"]), cfg["pools"])
    for k, o in enumerate(outs):
        ref = g[f"out{k}"]
        assert np.abs(o.detach().numpy() - ref).max() <= 2e-4 * np.abs(ref).max(), k
    tg = [torch.from_numpy(g[f"tgt{k}"].astype(np.float32)) for k in range(4)]
    loss = onet.ds_loss(outs, tg)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    loss.backward()
    for name, (gmax, gnorm, gsum) in meta["grads"].items():
        gr = p[name].grad.numpy().astype(np.float64)
        if name.endswith("conv.bias"):       # H4: cancelled by the InstanceNorm mean -> pure rounding noise
            assert np.sqrt((gr ** 2).sum()) < 1e-5 and gnorm < 1e-5, name
            continue
        assert abs(np.sqrt((gr ** 2).sum()) - gnorm) <= 1e-2 * gnorm + 1e-7, name
    for key in g.files:
        if key.startswith("grad:"):
            ref = g[key]
            mine = p[key[5:]].grad.numpy()
            if key.endswith("conv.bias"):
                continue
            assert np.abs(mine - ref).max() <= 1e-2 * np.abs(ref).max() + 1e-7, key  # pre-InstanceNorm weight grads are ill-conditioned (covariance-like sums)


# ---------------------------------------------------------------- sliding window
def test_steps_known_answers_from_reference_tests():
    # verbatim known-answer vectors of the reference's own unit test
    # (tests/test_steps_for_sliding_window_prediction.py:96-163)
    f = owin.compute_steps
    assert f((64, 130), (128, 260), 0.5) == [[0, 32, 64], [0, 65, 130]]
    assert f((128, 128, 128), (146, 176, 148), 0.5) == [[0, 18], [0, 48], [0, 20]]
    assert f((80, 192, 160), (130, 320, 244), 0.5) == [[0, 25, 50], [0, 64, 128], [0, 42, 84]]
    assert f((80, 192, 160), (130, 320, 244), 0.75) == [[0, 50], [0, 128], [0, 84]]
    assert f((128, 128, 128), (424, 456, 456), 0.5) == [[0, 59, 118, 178, 237, 296],
                                                        [0, 55, 109, 164, 219, 273, 328],
                                                        [0, 55, 109, 164, 219, 273, 328]]
    assert f((40, 56, 40), (40, 56, 40), 0.5) == [[0], [0], [0]]
    assert f((64, 192, 192), (94, 308, 308), 0.5) == [[0, 30], [0, 58, 116], [0, 58, 116]]
    for st in (1, 0.125, 0.5):
        assert f((121, 243, 91), (121, 243, 91), st) == [[0], [0], [0]]
        assert f((121, 243), (121, 243), st) == [[0], [0]]


def test_steps_golden_and_properties(golden_dir):
    js = json.load(open(os.path.join(golden_dir, "window_steps.json")))
    for key, ref in js.items():
        patch, image, st = key.split("|")
        assert owin.compute_steps(eval(patch), eval(image), float(st)) == ref
    s = owin.compute_steps((64, 160, 160), (300, 512, 512), 0.5)
    assert [len(a) for a in s] == [9, 6, 6]          # 324 tiles
    rs = np.random.RandomState(0)
    for _ in range(2000):                              # the reference test's random property check
        dim = rs.randint(1, 4)
        patch = tuple(int(v) for v in rs.randint(1, 200, dim))
        image = tuple(int(p + rs.randint(0, 200)) for p in patch)
        st = float(rs.uniform(0.01, 1))
        steps = owin.compute_steps(patch, image, st)
        for d in range(dim):
            assert steps[d][0] == 0 and steps[d][-1] + patch[d] == image[d]
            assert all(steps[d][i + 1] <= steps[d][i] + patch[d] for i in range(len(steps[d]) - 1))
            assert all(steps[d][i] + np.ceil(patch[d] * st) >= steps[d][i + 1] for i in range(len(steps[d]) - 1))


def _toy_net(ncls):
    w = np.linspace(-1.5, 2.0, ncls, dtype=np.float32).reshape(ncls, 1, 1, 1)
    b = np.linspace(0.3, -0.4, ncls, dtype=np.float32).reshape(ncls, 1, 1, 1)

    def net(t):                                        # (C,X,Y,Z) -> (ncls,X,Y,Z); same toy as make_golden._ToyNet
        X, Y, Z = t.shape[1:]
        gx = torch.linspace(-1, 1, X).numpy().reshape(1, X, 1, 1)
        gy = torch.linspace(-1, 1, Y).numpy().reshape(1, 1, Y, 1)
        gz = torch.linspace(-1, 1, Z).numpy().reshape(1, 1, 1, Z)
        return (t[:1] * w + b) + np.float32(0.5) * gx * w[::-1] + np.float32(0.25) * gy * gz * b
    return net


def test_window_predict_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "window.npz"))
    net = _toy_net(3)
    for tag, patch, mirror in (("a", (32, 48, 32), False), ("b", (32, 48, 32), True),
                               ("pad", (32, 48, 32), False), ("one", (32, 48, 32), False)):
        seg, prob = owin.predict_tiled(net, g[f"{tag}_x"], 3, patch, 0.5, mirror, (0, 1, 2), True)
        np.testing.assert_allclose(prob[:, ::3, ::3, ::3], g[f"{tag}_prob"], rtol=2e-5, atol=2e-6)
        assert (seg == g[f"{tag}_seg"]).mean() > 0.9999, tag
        assert seg.shape == g[f"{tag}_x"].shape[1:]


def test_gaussian_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "window.npz"))
    np.testing.assert_array_equal(owin.gaussian_map((16, 24, 20)), g["gauss_16_24_20"])
    gm = owin.gaussian_map((64, 160, 160))
    st = g["gauss_64_160_160_stats"]
    assert gm.min() == np.float32(st[0]) and gm.max() == np.float32(st[1])
    assert abs(gm.sum(dtype=np.float64) - st[2]) < 1e-6 * st[2]
    assert gm[0, 0, 0] == np.float32(st[3]) and gm[32, 80, 80] == np.float32(st[4]) and gm[10, 33, 150] == np.float32(st[5])


# ---------------------------------------------------------------- masking
def _mask_case(density, quant):
    shapes = onet.param_shapes(1, 48, 14, POOLS_BTCV)
    params = onet.det_params(shapes, seed=9)
    if quant:
        for k in params:
            params[k] = torch.round(params[k] * 1024) / 1024
    w = OrderedDict((k, v.numpy().copy()) for k, v in params.items() if omask.is_masked_name(k))
    rs = np.random.RandomState(13)
    mom = OrderedDict()
    for k, shp in shapes.items():                      # same draw order as the golden (all parameters)
        m = rs.standard_normal(shp).astype(np.float32)
        if k in w:
            mom[k] = m
    return shapes, w, mom


@pytest.mark.parametrize("quant", [False, True])
@pytest.mark.parametrize("density", [0.1, 0.2, 0.5])
def test_masking_matches_reference(golden_dir, density, quant):
    ref = json.load(open(os.path.join(golden_dir, "masking.json")))[f"{'quant' if quant else 'raw'}_{density}"]
    shapes, w, mom = _mask_case(density, quant)
    random.seed(0)
    masks = omask.init_uniform(shapes, density)
    assert list(masks.keys()) == ref["names"] and len(masks) == 35
    omask.apply_mask(w, masks, mom)
    for k, m in masks.items():
        assert sha(m[:, :, 0, 0, 0].astype(np.uint8)) == ref["init"][k], k
        assert int(m.sum()) == ref["init_nnz"][k]
    rs2 = np.random.RandomState(17)
    for k in shapes:
        if k in masks:
            pert = rs2.standard_normal(shapes[k]).astype(np.float32) * np.float32(1e-3)
            if quant:
                pert = (torch.round(torch.from_numpy(pert) * 1024 * 64) / (1024 * 64)).numpy()
            w[k] += pert
    # Masking.step(): apply_mask, then prune/regrow with the decayed death rate
    omask.apply_mask(w, masks, mom)
    random.seed(1)
    info = omask.prune_regrow(w, masks, ref["death_rate"], mom)
    for k, m in masks.items():
        assert sha(info["pruned_masks"][k][:, :, 0, 0, 0].astype(np.uint8)) == ref["pruned"][k], k
        assert sha(m[:, :, 0, 0, 0].astype(np.uint8)) == ref["after"][k], k
        assert int(m.sum()) == ref["after_nnz"][k]
        assert info["num_death"][k] == ref["num_death"][k]
        assert info["num_remove"][k] == ref["num_remove"][k]
    for k in ("loc4.0.0.blocks.0.conv.weight", "up0.0.weight"):
        assert sha(w[k]) == ref["w_sha:" + k]
        assert sha(mom[k]) == ref["m_sha:" + k]
    assert int(sum(m.sum() for m in masks.values())) == ref["total_nozeros"]

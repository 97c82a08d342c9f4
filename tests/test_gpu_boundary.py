"""The drop-in boundary exercised the way the reference calls it.

`INTEGRATION.md` section 2 tells a maintainer to resolve three reference module names to this package.
This test installs exactly those aliases (the reference tree itself does not exist on the GPU box, so
the parent packages are empty stand-ins), then drives the modules through the *reference's* import
statements and the literal call sequence of its trainer:

  * `initialize_network`            nnUNetTrainer_simple.py:292-301,361-363   (positional constructor, .cuda())
  * `initialize_optimizer_...`      :367-371                                  (SGD nesterov 0.99, wd 3e-5)
  * simple_main.py:163-168          Masking(optimizer, ..., CosineDecay, args) + add_module
  * `run_iteration`                 :549-583   zero_grad -> autocast(): network(data), loss -> GradScaler.scale(l)
                                               .backward() -> unscale_ -> clip_grad_norm_(12) -> scaler.step ->
                                               scaler.update -> mask.step() -> l.detach().cpu().numpy()
  * `save_checkpoint` / `load_checkpoint_ram`   :1140-1176, :1211-1255   (state_dict to CPU, torch.save, reload)

for 3 iterations including one prune / regrow update, and checks the outcome against the fp32 oracle
taking the same three steps.
"""
import argparse
import importlib
import io
import random
import sys
import types
from collections import OrderedDict

import numpy as np
import pytest
import torch
from torch import nn

from oracle import masking as omask
from oracle import network as onet

pytestmark = pytest.mark.gpu

ALIASES = {
    "e2enet.network_architecture.unetpp_d": "e2enet_medical_b200.network_architecture.unetpp_d",
    "e2enet.network_architecture.neural_network": "e2enet_medical_b200.network_architecture.neural_network",
    "e2enet.training.network_training.sparselearning.core_channel": "e2enet_medical_b200.sparselearning.core_channel",
}


@pytest.fixture()
def reference_names():
    """sys.modules aliasing (+ empty parent packages standing in for the reference tree)"""
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "e2enet" or k.startswith("e2enet.")}
    for name, target in ALIASES.items():
        parts = name.split(".")
        for i in range(1, len(parts)):
            pk = ".".join(parts[:i])
            if pk not in sys.modules:
                m = types.ModuleType(pk)
                m.__path__ = []
                sys.modules[pk] = m
        mod = importlib.import_module(target)
        sys.modules[name] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    yield
    for k in [k for k in sys.modules if k == "e2enet" or k.startswith("e2enet.")]:
        del sys.modules[k]
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v


def _synthetic(rs, B, in_ch, ncls, patch, pools):
    data = rs.rand(B, in_ch, *patch).astype(np.float32)
    tg, sp = [], np.array(patch)
    for k in range(4):
        tg.append(np.round(rs.rand(B, 1, *sp) * (ncls - 1)).astype(np.float32))
        sp = sp // np.array(pools[k])
    return data, tg


def test_reference_loop_through_aliases(reference_names):
    # ---- the reference's import statements (nnUNetTrainer_simple.py:293, simple_main.py:26-27)
    from e2enet.network_architecture.unetpp_d import Generic_UNetPlusPlus, InitWeights_He, softmax_helper
    from e2enet.network_architecture.neural_network import SegmentationNetwork
    from e2enet.training.network_training.sparselearning.core_channel import Masking, CosineDecay, add_sparse_args
    from torch.cuda.amp import GradScaler, autocast
    from e2enet_medical_b200.loss_functions import DC_and_CE_loss, MultipleOutputLoss2
    from e2enet_medical_b200.training import ds_loss_weights

    dev = torch.device("cuda:0")
    patch, in_ch, ncls = (32, 64, 64), 1, 3
    pools = [[1, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2], [2, 2, 2]]
    parser = argparse.ArgumentParser()
    add_sparse_args(parser)
    args = parser.parse_args(["--sparse", "True", "--density", "0.3", "--update_frequency", "2", "--growth", "random",
                              "--death", "magnitude", "--redistribution", "none", "--death-rate", "0.5"])
    assert args.fix is False                       # (`--fix False` would turn it ON: the type=bool quirk)

    # ---- initialize_network: positional arguments exactly as the reference passes them
    torch.manual_seed(0)
    network = Generic_UNetPlusPlus(patch, in_ch, 48, ncls, len(pools), 2, 2, nn.Conv3d, nn.InstanceNorm3d,
                                   {'eps': 1e-5, 'affine': True}, nn.Dropout3d, {'p': 0, 'inplace': True}, nn.LeakyReLU,
                                   {'negative_slope': 1e-2, 'inplace': True}, True, False, lambda x: x,
                                   InitWeights_He(1e-2), pools, [[3, 3, 3]] * 6, False, True, True)
    network.cuda()
    network.inference_apply_nonlin = softmax_helper
    assert isinstance(network, (SegmentationNetwork, nn.DataParallel))
    assert network.get_device() == 0
    optimizer = torch.optim.SGD(network.parameters(), 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    loss_fn = MultipleOutputLoss2(DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False}, {}),
                                  ds_loss_weights(4, len(pools)))
    amp_grad_scaler = GradScaler()

    # ---- simple_main.py:163-168
    decay = CosineDecay(args.death_rate, 1000)
    mask = Masking(optimizer, death_rate=args.death_rate, death_mode=args.death, death_rate_decay=decay,
                   growth_mode=args.growth, redistribution_mode=args.redistribution, args=args)
    random.seed(0)
    mask.add_module(network, sparse_init="uniform", density=args.density)

    # ---- the oracle takes the same steps in fp32 on the CPU
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in network.state_dict().items())
    ref_p = OrderedDict((k, v.detach().cpu().clone().requires_grad_(True)) for k, v in network.state_dict().items())
    ref_opt = torch.optim.SGD(list(ref_p.values()), 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    random.seed(0)
    ref_masks = omask.init_uniform(shapes, args.density)
    for k in ref_masks:
        assert np.array_equal(mask.masks[k].cpu().numpy(), ref_masks[k]), k

    rs = np.random.RandomState(7)
    losses, ref_losses = [], []
    for it in range(3):
        data_np, target_np = _synthetic(rs, 1, in_ch, ncls, patch, pools)
        # ---- run_iteration (nnUNetTrainer_simple.py:538-583), fp16 branch
        data = torch.from_numpy(data_np).float().cuda(network.get_device(), non_blocking=True)
        target = [torch.from_numpy(t).float().cuda(network.get_device(), non_blocking=True) for t in target_np]
        optimizer.zero_grad()
        with autocast():
            output = network(data)
            del data
            l = loss_fn(output, target)
        amp_grad_scaler.scale(l).backward()
        amp_grad_scaler.unscale_(optimizer)
        torch.nn.utils.clip_grad_norm_(network.parameters(), 12)
        amp_grad_scaler.step(optimizer)
        amp_grad_scaler.update()
        update_now = (it + 1) % args.update_frequency == 0
        if update_now:                       # weights the prune sees: after the SGD step and step()'s apply_mask
            w_pre = OrderedDict((k, (dict(network.named_parameters())[k].detach() * mask.masks[k]).cpu().numpy())
                                for k in ref_masks)
        random.seed(100 + it)
        mask.step()
        losses.append(float(l.detach().cpu().numpy()))
        assert all(o.dtype == torch.float32 for o in output) and len(output) == 4

        # ---- oracle: same iteration
        ref_opt.zero_grad()
        ro = onet.unetpp_forward(ref_p, torch.from_numpy(data_np), pools)
        rl = onet.ds_loss(ro, [torch.from_numpy(t) for t in target_np])
        rl.backward()
        torch.nn.utils.clip_grad_norm_(list(ref_p.values()), 12)
        ref_opt.step()
        with torch.no_grad():
            for k, m in ref_masks.items():
                ref_p[k].mul_(torch.from_numpy(m))
                ref_opt.state[ref_p[k]]['momentum_buffer'].mul_(torch.from_numpy(m))
        ref_losses.append(float(rl))
        if update_now:
            # the oracle prunes / regrows on the PRODUCT's weights (index-set parity is defined for identical weights;
            # the two fp32-vs-bf16 trajectories differ in the last bits) with the same Python RNG stream, and must
            # arrive at bit-identical masks; it then continues its own trajectory under those masks
            random.seed(100 + it)
            omask.prune_regrow(w_pre, ref_masks, mask.death_rate, assoc=mask.sum_association)
            for k in ref_masks:
                assert np.array_equal(mask.masks[k].cpu().numpy(), ref_masks[k]), ("prune/regrow", k)
            with torch.no_grad():
                for k, m in ref_masks.items():
                    ref_p[k].mul_(torch.from_numpy(m))
                    ref_opt.state[ref_p[k]]['momentum_buffer'].mul_(torch.from_numpy(m))

    assert mask.steps == 3 and mask.explore_step == 1
    for a, b in zip(losses, ref_losses):
        assert np.isfinite(a) and abs(a - b) < 2e-2 * abs(b), (losses, ref_losses)
    # the masks changed at the update step; weights and momentum are zero outside them
    random.seed(0)
    init_masks = omask.init_uniform(shapes, args.density)
    changed = 0
    for k, m0 in init_masks.items():
        m = mask.masks[k].cpu().numpy()
        changed += int((m != m0).any())
        assert set(np.unique(m)) <= {0.0, 1.0}
        w = dict(network.named_parameters())[k].detach()
        assert float((w * (1 - mask.masks[k])).abs().max()) == 0.0, k       # apply_mask ran after regrow
        buf = optimizer.state[dict(network.named_parameters())[k]]['momentum_buffer']
        assert float((buf * (1 - mask.masks[k])).abs().max()) == 0.0, k
    assert changed > 0
    # weights after 3 steps stay close to the oracle's trajectory on the kernels alive in both
    for k in ("conv_blocks_context.0.blocks.0.conv.weight", "seg_outputs.0.weight"):
        a, b = dict(network.named_parameters())[k].detach().cpu(), ref_p[k].detach()
        assert float((a - b).norm() / b.norm()) < 5e-2, k

    # ---- save_checkpoint / load_checkpoint_ram (nnUNetTrainer_simple.py:1140-1176, 1211-1255)
    state_dict = network.state_dict()
    for key in state_dict.keys():
        state_dict[key] = state_dict[key].cpu()
    assert list(state_dict.keys()) == list(onet.param_shapes(in_ch, 48, ncls, pools).keys())
    buf = io.BytesIO()
    torch.save({'epoch': 1, 'state_dict': state_dict, 'optimizer_state_dict': optimizer.state_dict()}, buf)
    buf.seek(0)
    ckpt = torch.load(buf, map_location=torch.device('cpu'), weights_only=False)
    new_state_dict = OrderedDict()
    curr = list(network.state_dict().keys())
    for k, value in ckpt['state_dict'].items():
        key = k
        if key not in curr and key.startswith('module.'):
            key = key[7:]
        new_state_dict[key] = value
    torch.manual_seed(1)
    net2 = Generic_UNetPlusPlus(patch, in_ch, 48, ncls, len(pools), 2, 2, nn.Conv3d, nn.InstanceNorm3d,
                                {'eps': 1e-5, 'affine': True}, nn.Dropout3d, {'p': 0, 'inplace': True}, nn.LeakyReLU,
                                {'negative_slope': 1e-2, 'inplace': True}, True, False, lambda x: x,
                                InitWeights_He(1e-2), pools, [[3, 3, 3]] * 6, False, True, True)
    net2.cuda()
    net2.load_state_dict(new_state_dict)
    optimizer2 = torch.optim.SGD(net2.parameters(), 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    optimizer2.load_state_dict(ckpt['optimizer_state_dict'])
    xt = torch.from_numpy(_synthetic(rs, 1, in_ch, ncls, patch, pools)[0]).cuda()
    network.eval(); net2.eval()
    with torch.no_grad(), autocast():
        o1, o2 = network(xt), net2(xt)
    for a, b in zip(o1, o2):
        assert torch.equal(a, b)                 # same weights -> same kernels -> same bits
    # the reference module loads the same checkpoint: the oracle (pinned to the reference) consumes it by name
    with torch.no_grad():
        ro = onet.unetpp_forward(OrderedDict((k, v.float()) for k, v in new_state_dict.items()), xt.cpu(), pools)
    assert float((o1[0].cpu() - ro[0]).abs().max() / ro[0].abs().max()) < 8e-2

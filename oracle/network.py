"""CPU oracle (test infrastructure, NOT product code) for the E2ENet network hot path.

torch-fp32 functional restatement of
  * torch_shift.forward                       -- reference unetpp_d.py:45-59
  * ConvDropoutNormNonlin.forward             -- reference unetpp_d.py:102-111
  * Generic_UNetPlusPlus.forward (5 pools)    -- reference unetpp_d.py:447-488
  * the module tree built by __init__/create_nest (names, shapes, registration
    order of the parameters)                  -- reference unetpp_d.py:307-445, 491-550

Everything is keyed by the reference's state_dict names so that the same parameter
dict drives the reference module (when generating goldens), this oracle and the CUDA
product.  Parity is pinned by tests/test_oracle_golden.py against tests/golden/*.npz
(produced from the unmodified reference by tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NUM_POOL = 5          # the reference forward is hard-wired to 5 pooling stages (unetpp_d.py:451-478)
SHIFT_SIZE = 5        # unetpp_d.py:89
MAX_FEATURES = 320    # unetpp_d.py:216
EPS = 1e-5            # InstanceNorm3d eps (nnUNetTrainer_simple.py norm_op_kwargs)
NEG_SLOPE = 1e-2


# ----------------------------------------------------------------------------------------
# depth shift  (unetpp_d.py:45-59)
# ----------------------------------------------------------------------------------------
def shift_groups(C: int, shift_size: int = SHIFT_SIZE) -> List[Tuple[int, int, int]]:
    """[(c_lo, c_hi, shift)] -- torch.chunk(x, 5, dim=1) gives groups of ceil(C/5) channels
    (fewer than 5 groups when C is small); group k is rolled by k - shift_size//2."""
    g = -(-C // shift_size)
    pad = shift_size // 2
    out = []
    k = 0
    lo = 0
    while lo < C:
        hi = min(C, lo + g)
        out.append((lo, hi, k - pad))
        lo = hi
        k += 1
    return out


def shift_depth(x: torch.Tensor, shift_size: int = SHIFT_SIZE) -> torch.Tensor:
    """y[b, c, d] = x[b, c, d - s_c] with zero fill; s_c = -2 + c // ceil(C/5)."""
    B, C, D = x.shape[:3]
    y = torch.zeros_like(x)
    for lo, hi, s in shift_groups(C, shift_size):
        if s >= 0:
            if s < D:
                y[:, lo:hi, s:] = x[:, lo:hi, :D - s]
        else:
            if -s < D:
                y[:, lo:hi, :D + s] = x[:, lo:hi, -s:]
    return y


# ----------------------------------------------------------------------------------------
# one shift-conv block  (unetpp_d.py:102-111)
# ----------------------------------------------------------------------------------------
def instance_norm_lrelu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    dims = (2, 3, 4)
    mean = x.mean(dims, keepdim=True)
    var = x.var(dims, unbiased=False, keepdim=True)
    xh = (x - mean) * torch.rsqrt(var + EPS)
    z = xh * gamma.view(1, -1, 1, 1, 1) + beta.view(1, -1, 1, 1, 1)
    return F.leaky_relu(z, NEG_SLOPE)


def shiftconv_block(x, w, b, gamma, beta, stride=(1, 1, 1)):
    """depth shift -> Conv3d((1,3,3), pad (0,1,1), stride) -> InstanceNorm(affine) -> LeakyReLU."""
    if tuple(w.shape[-3:]) == (1, 3, 3):
        x = shift_depth(x)
    y = F.conv3d(x, w, b, stride=tuple(stride), padding=(0, 1, 1))
    return instance_norm_lrelu(y, gamma, beta)


# ----------------------------------------------------------------------------------------
# parameter inventory (names / shapes / registration order of the reference module tree)
# ----------------------------------------------------------------------------------------
def stage_features(base: int) -> List[int]:
    f = [base]
    for _ in range(NUM_POOL):
        f.append(min(int(np.round(f[-1] * 2)), MAX_FEATURES))
    return f          # features of scale 0..5


def node_module(i: int, j: int) -> Tuple[int, int]:
    """fusion node x{i}_{j} (scale i, depth j>=1) is produced by loc{z}[idx]."""
    return NUM_POOL - i - j, j - 1


def param_shapes(in_ch: int, base: int, num_classes: int, pools: Sequence[Sequence[int]]
                 ) -> "OrderedDict[str, Tuple[int, ...]]":
    """state_dict keys -> shapes, in the reference's registration order
    (loc0..loc4, conv_blocks_context, up0..up4, seg_outputs; unetpp_d.py:418-438)."""
    assert len(pools) == NUM_POOL
    f = stage_features(base)
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def block(prefix, cin, cout):
        out[prefix + ".conv.weight"] = (cout, cin, 1, 3, 3)
        out[prefix + ".conv.bias"] = (cout,)
        out[prefix + ".instnorm.weight"] = (cout,)
        out[prefix + ".instnorm.bias"] = (cout,)

    for z in range(NUM_POOL):
        for idx in range(NUM_POOL - z):
            j = idx + 1
            i = NUM_POOL - z - j
            cin = 2 * f[i] + (f[i - 1] if i > 0 else 0)
            block(f"loc{z}.{idx}.0.blocks.0", cin, f[i])
            if z == 0:
                block(f"loc{z}.{idx}.1.blocks.0", f[i], f[i])
    cin = in_ch
    for s in range(NUM_POOL):
        block(f"conv_blocks_context.{s}.blocks.0", cin, f[s])
        block(f"conv_blocks_context.{s}.blocks.1", f[s], f[s])
        cin = f[s]
    block("conv_blocks_context.5.0.blocks.0", f[4], f[5])
    block("conv_blocks_context.5.1.blocks.0", f[5], f[5])     # convolutional_upsampling: final_num_features = f[5] (unetpp_d.py:358-359)
    for z in range(NUM_POOL):
        for idx in range(NUM_POOL - z):
            j = idx + 1
            i = NUM_POOL - z - j
            out[f"up{z}.{idx}.weight"] = (f[i + 1], f[i]) + tuple(int(p) for p in pools[i])
    for k in range(4):
        out[f"seg_outputs.{k}.weight"] = (num_classes, f[k], 1, 1, 1)
    return out


def det_params(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic, torch-version-independent parameter fill used by goldens, tests and
    benches: He-style normal for conv / transposed-conv weights, small noise on biases and
    affine terms so no term is trivially dead."""
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shp in shapes.items():
        n = int(np.prod(shp))
        if name.endswith("instnorm.weight"):
            v = 1.0 + 0.1 * rng.standard_normal(n)
        elif name.endswith("bias"):
            v = 0.1 * rng.standard_normal(n)
        else:
            fan_in = int(np.prod(shp[1:]))
            v = rng.standard_normal(n) * math.sqrt(2.0 / (1.0 + NEG_SLOPE ** 2) / fan_in)
        out[name] = torch.from_numpy(v.astype(np.float32).reshape(shp))
    return out


# ----------------------------------------------------------------------------------------
# the fusion grid  (unetpp_d.py:447-488)
# ----------------------------------------------------------------------------------------
def _blk(p, prefix, x, stride=(1, 1, 1)):
    return shiftconv_block(x, p[prefix + ".conv.weight"], p[prefix + ".conv.bias"],
                           p[prefix + ".instnorm.weight"], p[prefix + ".instnorm.bias"], stride)


def unetpp_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, pools: Sequence[Sequence[int]],
                   deep_supervision: bool = True):
    """returns [logits(x0_5), logits(x1_4), logits(x2_3), logits(x3_2)] (or the first only)."""
    pools = [tuple(int(k) for k in q) for q in pools]
    node = {}
    # encoder column: first conv of stage s>0 is strided by pools[s-1] (convolutional pooling)
    h = x
    for s in range(NUM_POOL):
        st = pools[s - 1] if s > 0 else (1, 1, 1)
        h = _blk(p, f"conv_blocks_context.{s}.blocks.0", h, st)
        h = _blk(p, f"conv_blocks_context.{s}.blocks.1", h)
        node[(s, 0)] = h
    h = _blk(p, "conv_blocks_context.5.0.blocks.0", h, pools[4])
    h = _blk(p, "conv_blocks_context.5.1.blocks.0", h)
    node[(5, 0)] = h
    # nested fusion nodes
    for j in range(1, NUM_POOL + 1):
        for i in range(NUM_POOL - j, -1, -1):
            z, idx = node_module(i, j)
            parts = [node[(i, j - 1)],
                     F.conv_transpose3d(node[(i + 1, j - 1)], p[f"up{z}.{idx}.weight"], stride=pools[i])]
            if i > 0:
                parts.append(F.max_pool3d(node[(i - 1, j - 1)], pools[i - 1]))
            h = torch.cat(parts, 1)
            h = _blk(p, f"loc{z}.{idx}.0.blocks.0", h)
            if z == 0:
                h = _blk(p, f"loc{z}.{idx}.1.blocks.0", h)
            node[(i, j)] = h
    outs = [F.conv3d(node[(k, NUM_POOL - k)], p[f"seg_outputs.{k}.weight"]) for k in range(4)]
    return outs if deep_supervision else outs[0]


# ----------------------------------------------------------------------------------------
# a compact stand-in for the trainer's loss (deep-supervision weighted CE + soft Dice);
# used only to give the backward pass a scalar with the same structure as
# MultipleOutputLoss2(DC_and_CE_loss) (reference deep_supervision.py:18-43, dice_loss.py:302-359)
# ----------------------------------------------------------------------------------------
def ds_weights(n: int = 4) -> List[float]:
    w = np.array([1 / (2 ** i) for i in range(n)])
    return list(w / w.sum())


def dc_ce_loss(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """CE + (-soft dice), batch_dice=False, do_bg=False, smooth 1e-5 (trainer defaults)."""
    ce = F.cross_entropy(logits.float(), target[:, 0].long())
    prob = F.softmax(logits.float(), 1)
    onehot = torch.zeros_like(prob).scatter_(1, target.long(), 1.0)
    axes = (2, 3, 4)
    tp = (prob * onehot).sum(axes)
    fp = (prob * (1 - onehot)).sum(axes)
    fn = ((1 - prob) * onehot).sum(axes)
    dc = (2 * tp + 1e-5) / (2 * tp + fp + fn + 1e-5 + 1e-8)
    dc = dc[:, 1:].mean()
    return ce - dc


def ds_loss(outs, targets) -> torch.Tensor:
    w = ds_weights(len(outs))
    l = w[0] * dc_ce_loss(outs[0], targets[0])
    for k in range(1, len(outs)):
        l = l + w[k] * dc_ce_loss(outs[k], targets[k])
    return l

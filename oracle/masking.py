"""CPU oracle (test infrastructure, NOT product code) for the kernel-granular DSFF Masking.

numpy restatement of the path `Masking.add_module -> init('uniform')`, `apply_mask`,
`step -> truncate_weights -> kernel_death / kernel_growth` with death='magnitude',
growth='random' (reference: e2enet/training/network_training/sparselearning/
core_channel.py:141-169, 290-317, 320-336, 427-434, 556-611, 647-666, 721-739).

A "kernel" is one (C0, C1) channel pair of a 5-D weight (C0, C1, kd, kh, kw): (Cout, Cin)
for Conv3d, (Cin, Cout) for ConvTranspose3d.  All index sets are bit-exact requirements.

Pinned by tests/test_oracle_golden.py against masks produced by the unmodified reference
class (tests/golden/make_golden.py -> tests/golden/masking_*.json).
"""
from __future__ import annotations

import math
import random
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np


def is_masked_name(name: str) -> bool:
    """core_channel.py:324 then :328-331 (bias / instnorm entries are removed again).
    NB 'loc' also matches the substring of 'b-loc-ks', so the rule is effectively
    "every conv weight outside conv_blocks_context, plus the up* transposed convs"."""
    sel = ('loc' in name and 'context' not in name) or ('up' in name)
    return sel and ('bias' not in name) and ('instnorm' not in name)


def kernel_l1(w: np.ndarray, assoc: str = "cpu") -> np.ndarray:
    """sum(sum(sum(|w|, -1), -1), -1) in fp32 (core_channel.py:653-655) with the association of the torch
    build the reference runs on:
      "cpu"  -- left to right at every level (verified against CPU torch; the committed goldens);
      "cuda" -- torch's CUDA reduce kernel strides last_pow2(n) threads over the reduced axis and combines them
                with a shuffle tree: (a0 + a2) + a1 for n = 3, (a0 + a2) + (a1 + a3) for n = 4 (measured on
                B200 / torch 2.11 by tests/test_gpu_oracle_fullsize.py).  The reference's Masking
                hard-codes .cuda(), so this is the order of a real run."""
    a = np.abs(w.astype(np.float32, copy=False))
    f32 = np.float32

    def fold(t):
        n = t.shape[-1]
        if assoc == "cuda" and n == 3:
            return ((t[..., 0] + t[..., 2]).astype(f32) + t[..., 1]).astype(f32)
        if assoc == "cuda" and n == 4:
            return ((t[..., 0] + t[..., 2]).astype(f32) + (t[..., 1] + t[..., 3]).astype(f32)).astype(f32)
        if assoc not in ("cpu", "cuda") or n > 4 and assoc == "cuda":
            raise ValueError("kernel_l1: association %r for %d values is not pinned" % (assoc, n))
        acc = t[..., 0].copy()
        for k in range(1, n):
            acc = (acc + t[..., k]).astype(f32)
        return acc

    return fold(fold(fold(a)))


def init_uniform(shapes: "OrderedDict[str, tuple]", density: float, rng=random) -> "OrderedDict[str, np.ndarray]":
    """core_channel.py:141-169.  One random.sample call per masked tensor, in order."""
    masks = OrderedDict()
    for name, shp in shapes.items():
        if not is_masked_name(name):
            continue
        rho = 0.2 if shp[0] == 48 else density                      # :147-151
        k_size = int(np.prod(shp[-3:]))
        numel = int(np.prod(shp))
        kernel_num = round(numel * rho / k_size)                     # python round (banker's)
        pick = rng.sample(list(range(0, shp[0] * shp[1])), kernel_num)
        m = np.zeros(shp, dtype=np.float32)
        pick = np.asarray(pick, dtype=np.int64)
        m[pick // shp[1], pick % shp[1]] = 1.0
        masks[name] = m
    return masks


def apply_mask(weights: Dict[str, np.ndarray], masks: Dict[str, np.ndarray],
               momentum: Optional[Dict[str, np.ndarray]] = None) -> None:
    """core_channel.py:427-434 (in place here; values identical)."""
    for name, m in masks.items():
        weights[name] *= m
        if momentum is not None and name in momentum:
            momentum[name] *= m


def kernel_death(mask: np.ndarray, w: np.ndarray, death_rate: float, assoc: str = "cpu"):
    """core_channel.py:647-666.  Returns (new_mask (in place), prune_num)."""
    k_size = int(np.prod(w.shape[-3:]))
    nnz = float(mask.sum(dtype=np.float64))                         # mask.sum().item() is exact (< 2^24)
    nzero = mask.size - nnz
    l1 = kernel_l1(w, assoc)
    prune_num = math.ceil(death_rate * nnz / k_size)
    num_zeros = math.ceil(nzero / k_size)
    value = np.sort(l1.reshape(-1), kind="stable")
    thr = value[num_zeros + prune_num - 1]
    kill = l1 <= thr
    mask[kill] = 0.0
    return mask, prune_num


def dead_list(mask: np.ndarray) -> np.ndarray:
    """row-major (c0, c1) of kernels whose mask sums to < 1 (core_channel.py:727-732)."""
    s = mask.reshape(mask.shape[0], mask.shape[1], -1).sum(-1)
    return np.argwhere(s < 1)


def kernel_growth(mask: np.ndarray, num_growth: int, rng=random) -> np.ndarray:
    """core_channel.py:721-739."""
    dead = dead_list(mask)
    pick = rng.sample(list(range(0, dead.shape[0])), num_growth)
    out = mask.copy()
    if num_growth:
        g = dead[np.asarray(pick, dtype=np.int64)]
        out[g[:, 0], g[:, 1]] = 1.0
    return out


def prune_regrow(weights, masks, death_rate: float, momentum=None, rng=random, assoc: str = "cpu"):
    """truncate_weights (core_channel.py:556-611): death for ALL tensors, then growth for
    ALL tensors (one rng.sample per tensor, in order), then apply_mask.
    Returns dict(num_death, num_remove, pruned_masks)."""
    num_death, num_remove, pruned = {}, {}, {}
    for name in masks:
        nnz0 = float(masks[name].sum(dtype=np.float64))
        m, pn = kernel_death(masks[name], weights[name], death_rate, assoc)
        num_death[name] = pn
        num_remove[name] = int(nnz0 - float(m.sum(dtype=np.float64)))
        pruned[name] = m.copy()
    for name in list(masks.keys()):
        masks[name] = kernel_growth(masks[name], num_death[name], rng)
    apply_mask(weights, masks, momentum)
    return dict(num_death=num_death, num_remove=num_remove, pruned_masks=pruned)

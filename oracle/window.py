"""CPU oracle (test infrastructure, NOT product code) for nnU-Net's Gaussian-weighted
sliding-window prediction as used by E2ENet.

numpy restatement of (reference: e2enet/network_architecture/neural_network.py)
  * _compute_steps_for_sliding_window   :260-284  (pinned by the reference's own
    known-answer vectors, tests/test_steps_for_sliding_window_prediction.py:96-163)
  * _get_gaussian                        :244-258
  * _internal_maybe_mirror_and_pred_3D   :500-565  (softmax, up to 8 flips, x gaussian)
  * _internal_predict_3D_3Dconv_tiled    :286-426  (all_in_gpu=False branch: fp32 host accumulators)
and of the third-party batchgenerators==0.24 `pad_nd_image` call at :300 (absent from the
reference tree; published algorithm restated: symmetric pad up to new_shape, slicer returned).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np
from scipy.ndimage import gaussian_filter


def compute_steps(patch_size: Sequence[int], image_size: Sequence[int], step_size: float) -> List[List[int]]:
    assert 0 < step_size <= 1
    target = [i * step_size for i in patch_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, patch_size)]
    steps = []
    for dim in range(len(patch_size)):
        max_step = image_size[dim] - patch_size[dim]
        actual = max_step / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def gaussian_map(patch_size: Sequence[int], sigma_scale: float = 1. / 8) -> np.ndarray:
    tmp = np.zeros(patch_size)
    tmp[tuple(i // 2 for i in patch_size)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in patch_size], 0, mode='constant', cval=0)
    g = (g / np.max(g) * 1).astype(np.float32)
    g[g == 0] = np.min(g[g != 0])
    return g


def pad_nd_image(image: np.ndarray, new_shape: Sequence[int], mode: str = "constant", kwargs=None):
    """batchgenerators.augmentations.utils.pad_nd_image(image, new_shape, mode, kwargs, True, None):
    pads the trailing len(new_shape) axes symmetrically (extra voxel goes to the upper side)
    up to max(new_shape, old_shape); returns (padded, slicer-that-undoes-it)."""
    kwargs = kwargs or {'constant_values': 0}
    old = np.array(image.shape[-len(new_shape):])
    n_lead = image.ndim - len(new_shape)
    new = np.array([max(int(a), int(b)) for a, b in zip(new_shape, old)])
    diff = new - old
    below = diff // 2
    above = diff // 2 + diff % 2
    pad = [[0, 0]] * n_lead + [[int(a), int(b)] for a, b in zip(below, above)]
    if not all(p == [0, 0] for p in pad):
        res = np.pad(image, pad, mode, **kwargs)
    else:
        res = image
    full_pad = np.array(pad)
    slicer = [slice(int(full_pad[i, 0]), int(res.shape[i] - full_pad[i, 1])) for i in range(res.ndim)]
    return res, slicer


# fixed mirror order of neural_network.py:529-560 (axes are 0/1/2 = x/y/z of the (c,x,y,z) volume)
_MIRROR_ORDER = [(), (2,), (1,), (2, 1), (0,), (2, 0), (1, 0), (2, 1, 0)]


def softmax0(x: np.ndarray) -> np.ndarray:
    m = x.max(0, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(0, keepdims=True)).astype(np.float32)


def mirror_and_pred(net: Callable[[np.ndarray], np.ndarray], tile: np.ndarray, num_classes: int,
                    mirror_axes: Tuple[int, ...], do_mirroring: bool, mult=None) -> np.ndarray:
    """tile (C,X,Y,Z) -> (num_classes,X,Y,Z) fp32; net maps (C,X,Y,Z) -> logits (ncls,X,Y,Z)."""
    res = np.zeros((num_classes,) + tile.shape[1:], dtype=np.float32)
    if do_mirroring:
        n = 2 ** len(mirror_axes)
        combos = [m for m in _MIRROR_ORDER if all(a in mirror_axes for a in m)]
    else:
        n, combos = 1, [()]
    for m in combos:
        ax = tuple(a + 1 for a in m)
        t = np.flip(tile, ax) if ax else tile
        pred = softmax0(net(np.ascontiguousarray(t)))
        res += np.float32(1 / n) * (np.flip(pred, ax) if ax else pred)
    if mult is not None:
        res *= mult[None]
    return res


def predict_tiled(net: Callable[[np.ndarray], np.ndarray], x: np.ndarray, num_classes: int,
                  patch_size: Sequence[int], step_size: float = 0.5, do_mirroring: bool = False,
                  mirror_axes: Tuple[int, ...] = (0, 1, 2), use_gaussian: bool = True):
    """returns (seg int64 (X,Y,Z), probs fp32 (ncls,X,Y,Z))."""
    data, slicer = pad_nd_image(x, patch_size)
    shp = data.shape
    steps = compute_steps(patch_size, shp[1:], step_size)
    n_tiles = len(steps[0]) * len(steps[1]) * len(steps[2])
    if use_gaussian and n_tiles > 1:
        g = gaussian_map(patch_size)
        add = g
    else:
        g = None
        add = np.ones(patch_size, dtype=np.float32)
    agg = np.zeros((num_classes,) + shp[1:], dtype=np.float32)
    nb = np.zeros((num_classes,) + shp[1:], dtype=np.float32)
    for x0 in steps[0]:
        for y0 in steps[1]:
            for z0 in steps[2]:
                sl = (slice(None), slice(x0, x0 + patch_size[0]), slice(y0, y0 + patch_size[1]),
                      slice(z0, z0 + patch_size[2]))
                p = mirror_and_pred(net, data[sl], num_classes, mirror_axes, do_mirroring, g)
                agg[sl] += p
                nb[sl] += add
    sl = tuple([slice(None)] + list(slicer[1:]))
    agg = agg[sl]
    nb = nb[sl]
    agg = agg / nb
    return agg.argmax(0), agg

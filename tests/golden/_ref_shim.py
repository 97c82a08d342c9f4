"""Makes the UNMODIFIED reference hot-path modules importable in the build container.

Only used by tests/golden/make_golden.py (golden generation; never on the GPU box and never
by the product).  The reference needs `batchgenerators` (absent, un-vendored) for one
function on this path, `pad_nd_image` (neural_network.py:17,300); and its Masking
hard-codes `.cuda()` (core_channel.py:67,116,326), which is patched to the identity on CPU.
"""
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("E2ENET_REFERENCE", "/root/reference")


def _pad_nd_image(image, new_shape=None, mode="constant", kwargs=None, return_slicer=False,
                  shape_must_be_divisible_by=None):
    # batchgenerators==0.24 semantics (restated from its published source)
    if kwargs is None:
        kwargs = {'constant_values': 0}
    if new_shape is not None:
        old_shape = np.array(image.shape[-len(new_shape):])
    else:
        assert shape_must_be_divisible_by is not None
        assert isinstance(shape_must_be_divisible_by, (list, tuple, np.ndarray))
        new_shape = image.shape[-len(shape_must_be_divisible_by):]
        old_shape = new_shape
    num_axes_nopad = len(image.shape) - len(new_shape)
    new_shape = [max(new_shape[i], old_shape[i]) for i in range(len(new_shape))]
    if not isinstance(new_shape, np.ndarray):
        new_shape = np.array(new_shape)
    if shape_must_be_divisible_by is not None:
        if not isinstance(shape_must_be_divisible_by, (list, tuple, np.ndarray)):
            shape_must_be_divisible_by = [shape_must_be_divisible_by] * len(new_shape)
        else:
            assert len(shape_must_be_divisible_by) == len(new_shape)
        for i in range(len(new_shape)):
            if new_shape[i] % shape_must_be_divisible_by[i] == 0:
                new_shape[i] -= shape_must_be_divisible_by[i]
        new_shape = np.array([new_shape[i] + shape_must_be_divisible_by[i] - new_shape[i] %
                              shape_must_be_divisible_by[i] for i in range(len(new_shape))])
    difference = new_shape - old_shape
    pad_below = difference // 2
    pad_above = difference // 2 + difference % 2
    pad_list = [[0, 0]] * num_axes_nopad + list([list(i) for i in zip(pad_below, pad_above)])
    if not ((all([i == 0 for i in pad_below])) and (all([i == 0 for i in pad_above]))):
        res = np.pad(image, pad_list, mode, **kwargs)
    else:
        res = image
    if not return_slicer:
        return res
    pad_list = np.array(pad_list)
    pad_list[:, 1] = np.array(res.shape) - pad_list[:, 1]
    slicer = list(slice(*i) for i in pad_list)
    return res, slicer


def install():
    if "batchgenerators" not in sys.modules:
        bg = types.ModuleType("batchgenerators")
        aug = types.ModuleType("batchgenerators.augmentations")
        utils = types.ModuleType("batchgenerators.augmentations.utils")
        utils.pad_nd_image = _pad_nd_image
        bg.augmentations = aug
        aug.utils = utils
        sys.modules["batchgenerators"] = bg
        sys.modules["batchgenerators.augmentations"] = aug
        sys.modules["batchgenerators.augmentations.utils"] = utils
    if "unittest2" not in sys.modules:
        import unittest
        sys.modules["unittest2"] = unittest
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def import_reference():
    install()
    from e2enet.network_architecture import unetpp_d, neural_network
    from e2enet.training.network_training.sparselearning import core_channel
    return unetpp_d, neural_network, core_channel

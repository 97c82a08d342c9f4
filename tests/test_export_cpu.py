"""CPU tests of the export-path oracle (oracle/export.py) and of the host side of the drop-in export module: the
resampling restatement equals scipy.ndimage.zoom(grid_mode=True, mode='nearest') -- which is what
skimage.transform.resize(order, mode='edge', anti_aliasing=False) (scikit-image 0.19, the reference's pin) computes
-- the separate-z decision follows segmentation_export.py:86-107, and the built-in NIfTI-1 writer round-trips
geometry and voxels."""
import os

import numpy as np
import pytest
from scipy.ndimage import zoom

from oracle import export as oex


@pytest.mark.parametrize("shape,new_shape,order", [((6, 8, 5), (9, 12, 10), 1), ((9, 12, 10), (6, 8, 5), 1),
                                                   ((7, 7, 7), (7, 13, 4), 1), ((5, 6, 7), (11, 6, 9), 0)])
def test_oracle_resample_equals_ndi_zoom_grid_mode(shape, new_shape, order):
    rs = np.random.RandomState(0)
    x = rs.rand(3, *shape).astype(np.float32)
    got = oex.resample_softmax(x, new_shape, order)
    want = np.stack([zoom(x[c].astype(float), [n / o for n, o in zip(new_shape, shape)], order=order, mode='nearest',
                          grid_mode=True) for c in range(3)])
    assert got.shape == (3,) + tuple(new_shape) and got.dtype == np.float32
    assert float(np.abs(got - want).max()) < 1e-6


def test_oracle_separate_z_is_inplane_linear_plus_nearest_along_axis():
    rs = np.random.RandomState(1)
    x = rs.rand(2, 4, 10, 12).astype(np.float32)
    got = oex.resample_softmax(x, (9, 15, 18), 1, True, np.array([0]), 0)
    # reference recipe (preprocessing.py:150-179): every slice in-plane with `order`, then order_z along the axis
    inplane = np.stack([np.stack([zoom(x[c, s].astype(float), (15 / 10, 18 / 12), order=1, mode='nearest', grid_mode=True)
                                  for s in range(4)]) for c in range(2)])
    idx = np.clip(np.floor((np.arange(9) + 0.5) * (4 / 9) - 0.5 + 0.5).astype(int), 0, 3)
    assert float(np.abs(got - inplane[:, idx]).max()) < 1e-6


def test_separate_z_decision():
    iso = {'original_spacing': (1.0, 1.0, 1.0), 'spacing_after_resampling': (1.0, 1.0, 1.0)}
    aniso = {'original_spacing': (5.0, 0.8, 0.8), 'spacing_after_resampling': (2.0, 0.8, 0.8)}
    two = {'original_spacing': (0.24, 1.25, 1.25), 'spacing_after_resampling': (1.0, 1.0, 1.0)}
    assert oex.decide_separate_z(iso, None) == (False, None)
    do, ax = oex.decide_separate_z(aniso, None)
    assert do and list(ax) == [0]
    do, ax = oex.decide_separate_z(two, None)          # two coarse axes: never separately (segmentation_export.py:101-104)
    assert not do and list(ax) == [1, 2]
    assert oex.decide_separate_z(aniso, False) == (False, None)
    from e2enet_medical_b200.inference import segmentation_export as se
    assert bool(se.get_do_separate_z(aniso['original_spacing'])) and list(se.get_lowres_axis(aniso['original_spacing'])) == [0]


def test_nifti_writer_roundtrip(tmp_path):
    from e2enet_medical_b200.inference.segmentation_export import nifti_affine, write_nifti_uint8
    rs = np.random.RandomState(2)
    seg = (rs.rand(5, 6, 7) * 14).astype(np.uint8)
    th = 0.3
    direction = (np.cos(th), -np.sin(th), 0, np.sin(th), np.cos(th), 0, 0, 0, 1)
    spacing, origin = (0.8, 0.9, 2.5), (-10.0, 20.0, 5.0)
    for name in ("a.nii", "a.nii.gz"):
        f = str(tmp_path / name)
        write_nifti_uint8(seg, f, spacing, origin, direction)
        arr, aff = oex.read_nifti(f)
        assert np.array_equal(arr, seg)
        assert np.allclose(aff, oex.nifti_affine(spacing, origin, direction), atol=1e-5)
        assert np.allclose(aff, nifti_affine(spacing, origin, direction), atol=1e-5)
    # LPS -> RAS: an identity-direction ITK image has a negative x / y diagonal and negated x / y origin
    a = nifti_affine((1, 2, 3), (4, 5, 6), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    assert np.allclose(np.diag(a)[:3], (-1, -2, 3)) and np.allclose(a[:3, 3], (-4, -5, 6))


def test_labels_from_softmax_bbox_and_regions():
    rs = np.random.RandomState(3)
    sm = rs.rand(3, 6, 8, 5).astype(np.float32)
    props = {'size_after_cropping': (9, 12, 10), 'original_size_of_raw_data': (12, 14, 13),
             'crop_bbox': [[2, 11], [1, 13], [3, 13]], 'original_spacing': (1., 1., 1.), 'spacing_after_resampling': (1.5, 1.5, 2.)}
    seg = oex.labels_from_softmax(sm, props)
    assert seg.shape == (12, 14, 13) and seg.dtype == np.uint8
    assert seg[:2].sum() == 0 and seg[:, :1].sum() == 0 and seg[:, :, :3].sum() == 0
    inner = oex.resample_softmax(sm, (9, 12, 10), 1).argmax(0)
    assert np.array_equal(seg[2:11, 1:13, 3:13], inner)
    reg = oex.labels_from_softmax(sm, dict(props, crop_bbox=None), region_class_order=(1, 2, 3))
    assert set(np.unique(reg)) <= {0, 1, 2, 3}

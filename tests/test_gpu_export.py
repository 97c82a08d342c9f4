"""GPU parity of the export path: e2e_resample_argmax and the drop-in
`save_segmentation_nifti_from_softmax` vs the oracle restatement of segmentation_export.py:27-152."""
import os

import numpy as np
import pytest
import torch

from oracle import export as oex

pytestmark = pytest.mark.gpu


def _softmax(rs, shape):
    z = rs.standard_normal(shape).astype(np.float32) * 2
    e = np.exp(z - z.max(0, keepdims=True))
    return (e / e.sum(0, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("shape,new_shape,order,sep,axis,order_z", [
    ((4, 20, 24, 18), (33, 40, 29), 1, False, None, 0),           # upsampling, odd sizes
    ((16, 30, 28, 26), (17, 19, 23), 1, False, None, 0),          # downsampling
    ((3, 6, 40, 44), (17, 64, 70), 1, True, [0], 0),              # anisotropic: in-plane linear, nearest along x
    ((3, 30, 32, 7), (45, 48, 20), 1, True, [2], 1),              # separate z with order_z = 1
    ((5, 12, 12, 12), (20, 9, 31), 0, False, None, 0),            # nearest everywhere
])
def test_resample_argmax_vs_oracle(shape, new_shape, order, sep, axis, order_z):
    from e2enet_medical_b200.inference.segmentation_export import resample_softmax_and_argmax
    rs = np.random.RandomState(0)
    sm = _softmax(rs, shape)
    want = oex.resample_softmax(sm, new_shape, order, sep, None if axis is None else np.array(axis), order_z)
    labels, probs = resample_softmax_and_argmax(sm, new_shape, order, sep, axis, order_z, want_probs=True)
    assert labels.dtype == torch.uint8 and tuple(labels.shape) == tuple(new_shape)
    got = probs.cpu().numpy()
    assert float(np.abs(got - want).max()) < 2e-6
    lab = labels.cpu().numpy()
    assert np.array_equal(lab, got.argmax(0)), "labels are the arg-max (first maximum) of the resampled probabilities"
    assert float((lab == want.argmax(0)).mean()) > 0.9995
    # labels only (no resampled volume materialised) gives the same labels; device tensors are accepted
    l2, p2 = resample_softmax_and_argmax(torch.from_numpy(sm).cuda(), new_shape, order, sep, axis, order_z)
    assert p2 is None and torch.equal(l2, labels)


def test_save_segmentation_nifti_from_softmax_vs_oracle(tmp_path):
    from e2enet_medical_b200.inference.segmentation_export import save_segmentation_nifti_from_softmax
    rs = np.random.RandomState(1)
    sm = _softmax(rs, (14, 24, 40, 36))
    props = {'size_after_cropping': (31, 52, 47), 'original_size_of_raw_data': (36, 56, 50),
             'crop_bbox': [[3, 34], [2, 54], [1, 48]], 'original_spacing': (2.5, 0.8, 0.8),
             'spacing_after_resampling': (3.0, 1.0, 1.0), 'itk_spacing': (0.8, 0.8, 2.5), 'itk_origin': (-100.0, 50.0, 7.5),
             'itk_direction': (1, 0, 0, 0, 1, 0, 0, 0, 1)}
    want = oex.labels_from_softmax(sm, dict(props, crop_bbox=[list(b) for b in props['crop_bbox']]))
    # 1) array input, resampled .npz requested, a post-processing function that relabels
    out = str(tmp_path / "case.nii.gz")
    npz = str(tmp_path / "case.npz")
    raw = str(tmp_path / "case_raw.nii.gz")
    post = lambda seg, k: np.where(seg == k, 0, seg)
    save_segmentation_nifti_from_softmax(sm, out, dict(props, crop_bbox=[list(b) for b in props['crop_bbox']]), 1, None, post, (3,),
                                         npz, raw, None, 0, False)
    arr, aff = oex.read_nifti(raw)
    assert arr.shape == (36, 56, 50)
    assert float((arr == want).mean()) > 0.9995
    arr_p, _ = oex.read_nifti(out)
    assert np.array_equal(arr_p, np.where(arr == 3, 0, arr))
    assert np.allclose(aff, oex.nifti_affine(props['itk_spacing'], props['itk_origin'], props['itk_direction']), atol=1e-4)
    z = np.load(npz)['softmax']
    assert z.dtype == np.float16 and z.shape == (14, 31, 52, 47) and os.path.isfile(npz[:-4] + ".pkl")
    ref = oex.resample_softmax(sm, (31, 52, 47), 1, True, np.array([0]), 0)            # spacing (2.5, .8, .8): separate z
    assert float(np.abs(z.astype(np.float32) - ref).max()) < 2e-3
    # 2) .npy file input is consumed (deleted), no resampling needed, regions_class_order
    f = str(tmp_path / "sm.npy")
    np.save(f, sm)
    p2 = dict(props, size_after_cropping=(24, 40, 36), original_size_of_raw_data=(24, 40, 36), crop_bbox=None)
    save_segmentation_nifti_from_softmax(f, str(tmp_path / "b.nii"), p2, 1, (1, 2), None, None, None, None, None, 0, False)
    assert not os.path.exists(f)
    arr2, _ = oex.read_nifti(str(tmp_path / "b.nii"))
    assert np.array_equal(arr2, oex.labels_from_softmax(sm, p2, region_class_order=(1, 2)))
    # 3) cubic interpolation is not silently approximated
    with pytest.raises(NotImplementedError):
        save_segmentation_nifti_from_softmax(sm, out, dict(props, crop_bbox=None, original_size_of_raw_data=(31, 52, 47)), 3,
                                             None, None, None, None, None, None, 0, False)

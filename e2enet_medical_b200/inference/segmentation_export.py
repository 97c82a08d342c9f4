"""Drop-in mirror of e2enet/inference/segmentation_export.py (the export path that follows
predict_3D): `save_segmentation_nifti_from_softmax` with the reference's signature (:27-33).

What the reference does per case on one CPU core of a worker process -- resize every class volume to the original
grid (skimage / scipy), materialise the resampled (C, X', Y', Z') fp32 array, arg-max it, paste into the
uncropped volume, write NIfTI through SimpleITK (:76-152) -- runs here as ONE CUDA pass
(e2e_resample_argmax: per-axis nearest / linear with the reference's pixel-centre map, arg-max fused, labels
leave the GPU as uint8) plus a small built-in NIfTI-1 writer (SimpleITK is not a dependency).  Interpolation
orders 0 and 1 are implemented (the export default is 1, predict.py:250-262); order 3 raises.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import pickle
import struct
from copy import deepcopy
from typing import Tuple, Union

import numpy as np
import torch

from .. import _lib

RESAMPLING_SEPARATE_Z_ANISO_THRESHOLD = 3          # e2enet/configuration.py:4


def get_do_separate_z(spacing, anisotropy_threshold=RESAMPLING_SEPARATE_Z_ANISO_THRESHOLD):
    return (np.max(spacing) / np.min(spacing)) > anisotropy_threshold


def get_lowres_axis(new_spacing):
    return np.where(max(new_spacing) / np.array(new_spacing) == 1)[0]


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def resample_softmax_and_argmax(softmax: Union[np.ndarray, torch.Tensor], new_shape, order: int = 1,
                                do_separate_z: bool = False, axis=None, order_z: int = 0, want_probs: bool = False,
                                want_labels: bool = True, device=None):
    """resample_data_or_seg(softmax, new_shape, is_seg=False, axis, order, do_separate_z, order_z) fused with
    argmax(0) on the GPU.  Returns (labels uint8 CUDA tensor or None, resampled probabilities fp32 CUDA tensor or
    None).  softmax may already live on the device (straight from the sliding window)."""
    if order not in (0, 1) or (do_separate_z and order_z not in (0, 1)):
        raise NotImplementedError("E2ENet B200 export implements interpolation orders 0 and 1 (got order=%r, order_z=%r); "
                                  "the reference's export default is 1 (inference/predict.py:250-262)" % (order, order_z))
    if not torch.cuda.is_available():
        raise _lib.E2EError("segmentation export runs on CUDA only (no CPU fallback)")
    dev = torch.device("cuda") if device is None else torch.device(device)
    t = softmax if torch.is_tensor(softmax) else torch.from_numpy(np.ascontiguousarray(softmax, dtype=np.float32))
    t = t.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
    assert t.dim() == 4 and len(new_shape) == 3, "data must be (c, x, y, z)"
    Cc, X, Y, Z = t.shape
    Xo, Yo, Zo = (int(v) for v in new_shape)
    modes = [order] * 3
    if do_separate_z:
        assert axis is not None and len(axis) == 1, "only one anisotropic axis supported"
        modes[int(axis[0])] = order_z
    labels = torch.empty((Xo, Yo, Zo), dtype=torch.uint8, device=dev) if want_labels else None
    probs = torch.empty((Cc, Xo, Yo, Zo), dtype=torch.float32, device=dev) if want_probs else None
    _lib.check(_lib.load().e2e_resample_argmax(_p(t), Cc, X, Y, Z, Xo, Yo, Zo, modes[0], modes[1], modes[2], _p(probs),
                                               _p(labels), _lib.stream_ptr()), "resample_argmax")
    return labels, probs


def nifti_affine(spacing, origin, direction) -> np.ndarray:
    """ITK (LPS) spacing / origin / direction -> NIfTI (RAS) affine, as ITK's NiftiImageIO stores it"""
    d = np.asarray(direction, dtype=float).reshape(3, 3)
    a = np.eye(4)
    a[:3, :3] = d * np.asarray(spacing, dtype=float)[None, :]
    a[:3, 3] = np.asarray(origin, dtype=float)
    return np.diag([-1.0, -1.0, 1.0, 1.0]) @ a


def _quaternion(rot: np.ndarray):
    """rotation matrix (proper or improper) -> NIfTI qform (qb, qc, qd, qfac)"""
    r = rot.copy()
    qfac = 1.0
    if np.linalg.det(r) < 0:
        r[:, 2] = -r[:, 2]
        qfac = -1.0
    a = 1.0 + r[0, 0] + r[1, 1] + r[2, 2]
    if a > 0.5:
        a = 0.5 * np.sqrt(a)
        b, c, d = 0.25 * (r[2, 1] - r[1, 2]) / a, 0.25 * (r[0, 2] - r[2, 0]) / a, 0.25 * (r[1, 0] - r[0, 1]) / a
    else:
        xd, yd, zd = 1.0 + r[0, 0] - (r[1, 1] + r[2, 2]), 1.0 + r[1, 1] - (r[0, 0] + r[2, 2]), 1.0 + r[2, 2] - (r[0, 0] + r[1, 1])
        if xd > 1.0:
            b = 0.5 * np.sqrt(xd)
            c, d, a = 0.25 * (r[0, 1] + r[1, 0]) / b, 0.25 * (r[0, 2] + r[2, 0]) / b, 0.25 * (r[2, 1] - r[1, 2]) / b
        elif yd > 1.0:
            c = 0.5 * np.sqrt(yd)
            b, d, a = 0.25 * (r[0, 1] + r[1, 0]) / c, 0.25 * (r[1, 2] + r[2, 1]) / c, 0.25 * (r[0, 2] - r[2, 0]) / c
        else:
            d = 0.5 * np.sqrt(zd)
            b, c, a = 0.25 * (r[0, 2] + r[2, 0]) / d, 0.25 * (r[1, 2] + r[2, 1]) / d, 0.25 * (r[1, 0] - r[0, 1]) / d
        if a < 0:
            b, c, d = -b, -c, -d
    return float(b), float(c), float(d), qfac


def write_nifti_uint8(seg_zyx: np.ndarray, fname: str, spacing, origin, direction):
    """what `sitk.WriteImage(sitk.GetImageFromArray(seg.astype(uint8)) + SetSpacing / SetOrigin / SetDirection)`
    produces (segmentation_export.py:141-145): a single-file NIfTI-1 (.nii or .nii.gz), uint8, x fastest, with
    the LPS geometry stored as a RAS affine in qform and sform."""
    seg = np.ascontiguousarray(seg_zyx, dtype=np.uint8)
    assert seg.ndim == 3
    nz, ny, nx = seg.shape
    aff = nifti_affine(spacing, origin, direction)
    rot = aff[:3, :3] / np.linalg.norm(aff[:3, :3], axis=0, keepdims=True)
    qb, qc, qd, qfac = _quaternion(rot)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into("<2h", hdr, 70, 2, 8)                                   # NIFTI_TYPE_UINT8, bitpix
    struct.pack_into("<8f", hdr, 76, qfac, float(spacing[0]), float(spacing[1]), float(spacing[2]), 0.0, 0.0, 0.0, 0.0)
    struct.pack_into("<f", hdr, 108, 352.0)                                  # vox_offset
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)                              # scl_slope, scl_inter
    hdr[123] = 2 | 8                                                         # xyzt_units: mm, s
    struct.pack_into("<2h", hdr, 252, 1, 1)                                  # qform_code, sform_code: scanner anat
    struct.pack_into("<6f", hdr, 256, qb, qc, qd, float(aff[0, 3]), float(aff[1, 3]), float(aff[2, 3]))
    struct.pack_into("<12f", hdr, 280, *[float(v) for v in aff[:3].reshape(-1)])
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + b"\0\0\0\0" + seg.tobytes()
    if fname.endswith(".gz"):
        with gzip.open(fname, "wb", compresslevel=1) as f:
            f.write(payload)
    else:
        with open(fname, "wb") as f:
            f.write(payload)


def save_segmentation_nifti_from_softmax(segmentation_softmax: Union[str, np.ndarray, torch.Tensor], out_fname: str,
                                         properties_dict: dict, order: int = 1,
                                         region_class_order: Tuple[Tuple[int]] = None,
                                         seg_postprogess_fn: callable = None, seg_postprocess_args: tuple = None,
                                         resampled_npz_fname: str = None,
                                         non_postprocessed_fname: str = None, force_separate_z: bool = None,
                                         interpolation_order_z: int = 0, verbose: bool = True):
    """reference segmentation_export.py:27-152, same arguments and side effects (the NIfTI file, the optional
    resampled .npz + .pkl, the optional non-post-processed file); segmentation_softmax may also be a CUDA tensor."""
    if verbose:
        print("force_separate_z:", force_separate_z, "interpolation order:", order)
    if isinstance(segmentation_softmax, str):
        assert os.path.isfile(segmentation_softmax), "If isinstance(segmentation_softmax, str) then " \
                                                     "isfile(segmentation_softmax) must be True"
        del_file = deepcopy(segmentation_softmax)
        if segmentation_softmax.endswith('.npy'):
            segmentation_softmax = np.load(segmentation_softmax)
        elif segmentation_softmax.endswith('.npz'):
            segmentation_softmax = np.load(segmentation_softmax)['softmax']
        os.remove(del_file)
    current_shape = tuple(segmentation_softmax.shape)
    shape_original_after_cropping = properties_dict.get('size_after_cropping')
    shape_original_before_cropping = properties_dict.get('original_size_of_raw_data')
    need_probs = resampled_npz_fname is not None or region_class_order is not None
    resample = bool(np.any([i != j for i, j in zip(np.array(current_shape[1:]), np.array(shape_original_after_cropping))]))
    if resample:
        if force_separate_z is None:
            if get_do_separate_z(properties_dict.get('original_spacing')):
                do_separate_z, lowres_axis = True, get_lowres_axis(properties_dict.get('original_spacing'))
            elif get_do_separate_z(properties_dict.get('spacing_after_resampling')):
                do_separate_z, lowres_axis = True, get_lowres_axis(properties_dict.get('spacing_after_resampling'))
            else:
                do_separate_z, lowres_axis = False, None
        else:
            do_separate_z = force_separate_z
            lowres_axis = get_lowres_axis(properties_dict.get('original_spacing')) if do_separate_z else None
        if lowres_axis is not None and len(lowres_axis) != 1:
            do_separate_z = False
        if verbose:
            print("separate z:", do_separate_z, "lowres axis", lowres_axis)
        new_shape = tuple(int(v) for v in shape_original_after_cropping)
    else:
        if verbose:
            print("no resampling necessary")
        do_separate_z, lowres_axis, new_shape = False, None, current_shape[1:]
    # one pass: resample (identity map when the shapes agree) + arg-max; probabilities only when somebody needs them
    labels, probs = resample_softmax_and_argmax(segmentation_softmax, new_shape, order if resample else 0, do_separate_z,
                                                lowres_axis, interpolation_order_z, want_probs=need_probs,
                                                want_labels=region_class_order is None)
    if resampled_npz_fname is not None:
        np.savez_compressed(resampled_npz_fname, softmax=probs.cpu().numpy().astype(np.float16))
        if region_class_order is not None:
            properties_dict['regions_class_order'] = region_class_order
        with open(resampled_npz_fname[:-4] + ".pkl", 'wb') as f:
            pickle.dump(properties_dict, f)
    if region_class_order is None:
        seg_old_spacing = labels.cpu().numpy()
    else:
        fin = torch.zeros(probs.shape[1:], dtype=torch.float32, device=probs.device)
        for i, c in enumerate(region_class_order):
            fin[probs[i] > 0.5] = float(c)
        seg_old_spacing = fin.cpu().numpy()
    bbox = properties_dict.get('crop_bbox')
    if bbox is not None:
        seg_old_size = np.zeros(shape_original_before_cropping, dtype=np.uint8)
        for c in range(3):
            bbox[c][1] = np.min((bbox[c][0] + seg_old_spacing.shape[c], shape_original_before_cropping[c]))
        seg_old_size[bbox[0][0]:bbox[0][1], bbox[1][0]:bbox[1][1], bbox[2][0]:bbox[2][1]] = seg_old_spacing
    else:
        seg_old_size = seg_old_spacing
    if seg_postprogess_fn is not None:
        seg_old_size_postprocessed = seg_postprogess_fn(np.copy(seg_old_size), *seg_postprocess_args)
    else:
        seg_old_size_postprocessed = seg_old_size
    geo = (properties_dict['itk_spacing'], properties_dict['itk_origin'], properties_dict['itk_direction'])
    write_nifti_uint8(seg_old_size_postprocessed.astype(np.uint8), out_fname, *geo)
    if (non_postprocessed_fname is not None) and (seg_postprogess_fn is not None):
        write_nifti_uint8(seg_old_size.astype(np.uint8), non_postprocessed_fname, *geo)

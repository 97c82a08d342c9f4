"""CPU oracle (test infrastructure, NOT product code) for the export path that follows the sliding window:
`save_segmentation_nifti_from_softmax` (reference e2enet/inference/segmentation_export.py:27-160) with
`resample_data_or_seg(is_seg=False)` (e2enet/preprocessing/preprocessing.py:113-202), `get_do_separate_z` /
`get_lowres_axis` (:28-35).

PARITY UNPINNED at the third-party boundary: the reference resizes with `skimage.transform.resize(order,
mode='edge', anti_aliasing=False)` (scikit-image==0.19.3, requirements.txt:44) and writes with SimpleITK 2.0.2
(:49); neither package exists in this image and the reference module imports both at the top, so it cannot be run
here to produce goldens.  Restated from their published behaviour:
  * skimage >= 0.19 `resize` without anti-aliasing is `scipy.ndimage.zoom(..., grid_mode=True, mode='nearest')`,
    i.e. the pixel-centre map  in = (out + 0.5) * n_in / n_out - 0.5  with edge clamping -- the very map the
    reference itself spells out for its separate-z branch (preprocessing.py:164-179), which IS scipy and is
    restated here verbatim in behaviour (map_coordinates, mode='nearest');
  * SimpleITK / ITK NiftiImageIO writes the LPS geometry (spacing, origin, direction) as a RAS affine
    diag(-1,-1,1,1) @ [direction * spacing | origin] in both qform and sform.
"""
from __future__ import annotations

import gzip
import struct
from typing import Optional, Sequence, Tuple

import numpy as np
from scipy.ndimage import map_coordinates

ANISO_THRESHOLD = 3          # e2enet/configuration.py:4


def get_do_separate_z(spacing, anisotropy_threshold=ANISO_THRESHOLD) -> bool:
    return bool((np.max(spacing) / np.min(spacing)) > anisotropy_threshold)


def get_lowres_axis(new_spacing):
    return np.where(max(new_spacing) / np.array(new_spacing) == 1)[0]


def _centre_coords(n_in: int, n_out: int) -> np.ndarray:
    return (float(n_in) / n_out) * (np.arange(n_out) + 0.5) - 0.5


def resize_volume(vol: np.ndarray, new_shape: Sequence[int], orders: Sequence[int]) -> np.ndarray:
    """separable resampling with the pixel-centre map and edge clamping; orders[a] in {0, 1} per axis."""
    out = vol.astype(float)
    for a, (n_out, order) in enumerate(zip(new_shape, orders)):
        n_in = out.shape[a]
        if n_in == n_out:
            continue
        c = _centre_coords(n_in, n_out)
        grids = np.meshgrid(*[np.arange(s, dtype=float) if k != a else c for k, s in enumerate(out.shape)], indexing="ij")
        out = map_coordinates(out, np.array(grids), order=order, mode="nearest")
    return out


def resample_softmax(data: np.ndarray, new_shape, order=1, do_separate_z=False, axis=None, order_z=0) -> np.ndarray:
    """resample_data_or_seg(data, new_shape, is_seg=False, axis, order, do_separate_z, order_z) for order, order_z <= 1"""
    assert data.ndim == 4 and len(new_shape) == 3
    if tuple(data.shape[1:]) == tuple(int(v) for v in new_shape):
        return data
    if order > 1 or order_z > 1:
        raise NotImplementedError("oracle restates interpolation orders 0 and 1 (the export default is 1)")
    orders = [order] * 3
    if do_separate_z:
        assert len(axis) == 1
        orders[int(axis[0])] = order_z
    return np.stack([resize_volume(data[c], new_shape, orders) for c in range(data.shape[0])]).astype(data.dtype)


def decide_separate_z(properties: dict, force_separate_z: Optional[bool]):
    """segmentation_export.py:86-107"""
    if force_separate_z is None:
        if get_do_separate_z(properties.get('original_spacing')):
            do, axis = True, get_lowres_axis(properties.get('original_spacing'))
        elif get_do_separate_z(properties.get('spacing_after_resampling')):
            do, axis = True, get_lowres_axis(properties.get('spacing_after_resampling'))
        else:
            do, axis = False, None
    else:
        do = force_separate_z
        axis = get_lowres_axis(properties.get('original_spacing')) if do else None
    if axis is not None and len(axis) != 1:
        do = False
    return do, axis


def labels_from_softmax(softmax: np.ndarray, properties: dict, order=1, region_class_order=None, force_separate_z=None,
                        interpolation_order_z=0) -> np.ndarray:
    """the array handed to the NIfTI writer (segmentation_export.py:76-134, without post-processing)"""
    shape_after = properties.get('size_after_cropping')
    shape_before = properties.get('original_size_of_raw_data')
    if any(i != j for i, j in zip(softmax.shape[1:], shape_after)):
        do, axis = decide_separate_z(properties, force_separate_z)
        res = resample_softmax(softmax, shape_after, order, do, axis, interpolation_order_z)
    else:
        res = softmax
    if region_class_order is None:
        seg = res.argmax(0)
    else:
        seg = np.zeros(res.shape[1:])
        for i, c in enumerate(region_class_order):
            seg[res[i] > 0.5] = c
    bbox = properties.get('crop_bbox')
    if bbox is not None:
        full = np.zeros(shape_before, dtype=np.uint8)
        bb = [[int(b[0]), int(min(b[0] + seg.shape[c], shape_before[c]))] for c, b in enumerate(bbox)]
        full[bb[0][0]:bb[0][1], bb[1][0]:bb[1][1], bb[2][0]:bb[2][1]] = seg
        seg = full
    return seg.astype(np.uint8)


def nifti_affine(spacing, origin, direction) -> np.ndarray:
    """ITK (LPS) geometry -> NIfTI (RAS) 4x4 affine of the (x, y, z) voxel grid"""
    d = np.asarray(direction, dtype=float).reshape(3, 3)
    a = np.eye(4)
    a[:3, :3] = d * np.asarray(spacing, dtype=float)[None, :]
    a[:3, 3] = np.asarray(origin, dtype=float)
    return np.diag([-1.0, -1.0, 1.0, 1.0]) @ a


def read_nifti(fname: str) -> Tuple[np.ndarray, np.ndarray]:
    """minimal NIfTI-1 reader for the round-trip tests: returns (array indexed (z, y, x), sform affine)"""
    raw = gzip.open(fname, "rb").read() if fname.endswith(".gz") else open(fname, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 348 and raw[344:348] == b"n+1\0"
    dim = struct.unpack("<8h", raw[40:56])
    datatype, bitpix = struct.unpack("<2h", raw[70:74])
    vox_offset = int(struct.unpack("<f", raw[108:112])[0])
    assert dim[0] == 3 and datatype == 2 and bitpix == 8
    nx, ny, nz = dim[1:4]
    arr = np.frombuffer(raw, dtype=np.uint8, count=nx * ny * nz, offset=vox_offset).reshape(nz, ny, nx)
    srow = np.array(struct.unpack("<12f", raw[280:328])).reshape(3, 4)
    return arr, np.vstack([srow, [0, 0, 0, 1]])

"""ctypes binding of libe2enet_b200.so (the C ABI declared in include/e2enet_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, the
product path raises.  (The CPU oracle lives in /oracle and is test infrastructure only.)
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libe2enet_b200.so")
# storage / operand precision of activations, gradients and packed weights: "bf16" (default, north_star) or "fp16"
# (the reference's shipped AMP arithmetic; same kernels compiled with -DE2E_FP16, needs a loss scale in training)
LIB_PATHS = {"bf16": LIB_PATH, "fp16": os.path.join(_PKG, "libe2enet_b200_fp16.so")}
_precision = os.environ.get("E2E_PRECISION", "bf16")
if _precision not in LIB_PATHS:
    raise ValueError("E2E_PRECISION must be one of %s" % sorted(LIB_PATHS))

E2E_MAX_SRC = 4


class CEntry(C.Structure):
    _fields_ = [("src", C.c_int32), ("blk", C.c_int32), ("dd", C.c_int32), ("dh", C.c_int32), ("dw", C.c_int32)]


class Tap(C.Structure):
    _fields_ = [("dd", C.c_int32), ("dh", C.c_int32), ("dw", C.c_int32)]


class ColBlk(C.Structure):
    _fields_ = [("dst", C.c_int32), ("blk", C.c_int32), ("chmask", C.c_int32), ("od", C.c_int32),
                ("oh", C.c_int32), ("ow", C.c_int32)]


class GemmParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("Di", C.c_int32), ("Hi", C.c_int32), ("Wi", C.c_int32),
        ("Do", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("isd", C.c_int32), ("ish", C.c_int32), ("isw", C.c_int32),
        ("ivd", C.c_int32), ("ivh", C.c_int32), ("ivw", C.c_int32),
        ("Dd", C.c_int32), ("Hd", C.c_int32), ("Wd", C.c_int32),
        ("osd", C.c_int32), ("osh", C.c_int32), ("osw", C.c_int32),
        ("n_src", C.c_int32),
        ("src", C.c_void_p * E2E_MAX_SRC),
        ("src_cb", C.c_int32 * E2E_MAX_SRC),
        ("n_cent", C.c_int32),
        ("cents", C.c_void_p),
        ("n_taps", C.c_int32),
        ("taps", C.c_void_p),
        ("wpacked", C.c_void_p),
        ("Npad", C.c_int32),
        ("cols", C.c_void_p),
        ("n_dst", C.c_int32),
        ("dst", C.c_void_p * E2E_MAX_SRC),
        ("dst_cb", C.c_int32 * E2E_MAX_SRC),
        ("out_mode", C.c_int32),
        ("impl", C.c_int32),
        ("col_bounds", C.c_int32),
        ("stats", C.c_void_p),
        ("stats_ctot", C.c_int32),
        ("accumulate", C.c_int32),
    ]


class WgradParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("Di", C.c_int32), ("Hi", C.c_int32), ("Wi", C.c_int32),
        ("Do", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("isd", C.c_int32), ("ish", C.c_int32), ("isw", C.c_int32),
        ("ivd", C.c_int32), ("ivh", C.c_int32), ("ivw", C.c_int32),
        ("n_src", C.c_int32),
        ("src", C.c_void_p * E2E_MAX_SRC),
        ("src_cb", C.c_int32 * E2E_MAX_SRC),
        ("n_cent", C.c_int32),
        ("cents", C.c_void_p),
        ("n_taps", C.c_int32),
        ("taps", C.c_void_p),
        ("grad", C.c_void_p),
        ("grad_cb", C.c_int32),
        ("Npad", C.c_int32),
        ("dwp", C.c_void_p),
        ("impl", C.c_int32),
        ("grad_out", C.c_void_p),
        ("rowoff", C.c_void_p),
        ("centoff", C.c_void_p),
        ("tapoff", C.c_void_p),
    ]


class PackJob(C.Structure):
    _fields_ = [("w", C.c_void_p), ("mask", C.c_void_p), ("rowoff", C.c_void_p), ("centoff", C.c_void_p),
                ("tapoff", C.c_void_p), ("n_cent", C.c_int32), ("n_taps", C.c_int32), ("Npad", C.c_int32),
                ("pad_", C.c_int32), ("out", C.c_void_p), ("item_begin", C.c_int64), ("emask", C.c_void_p),
                ("rclass", C.c_void_p)]


class SgdTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("mom", C.c_void_p), ("mask", C.c_void_p), ("numel", C.c_int64)]


# name -> (restype, argtypes); every symbol declared in include/e2enet_b200.h
_VP, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "e2e_last_error": (C.c_char_p, []),
    "e2e_version": (C.c_int, []),
    "e2e_precision": (C.c_char_p, []),
    "e2e_launch_count": (C.c_longlong, []),
    "e2e_gather_gemm": (C.c_int, [C.POINTER(GemmParams), _VP]),
    "e2e_gather_gemm_multi": (C.c_int, [C.POINTER(GemmParams), _I32, _VP]),
    "e2e_gather_gemm_stats_slots": (C.c_int, [C.POINTER(GemmParams), _I32]),
    "e2e_gather_gemm_on_tcgen05": (C.c_int, [C.POINTER(GemmParams), _I32]),
    "e2e_in_stats_final": (C.c_int, [_VP, _I32, _I32, _I32, _I64, _F, _VP, _VP, _VP]),
    "e2e_gather_wgrad": (C.c_int, [C.POINTER(WgradParams), _VP]),
    "e2e_gather_wgrad_direct_ok": (C.c_int, [C.POINTER(WgradParams)]),
    "e2e_pack_weights": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I32, _VP, _VP]),
    "e2e_pack_weights_multi": (C.c_int, [_VP, _I32, _I64, _VP]),
    "e2e_unpack_wgrad": (C.c_int, [_VP, _VP, _VP, _VP, _I32, _I32, _I32, _VP, _VP]),
    "e2e_nc_to_c8": (C.c_int, [_VP, _VP, _I32, _I32, _I64, _VP]),
    "e2e_c8_to_nc": (C.c_int, [_VP, _VP, _I32, _I32, _I64, _VP]),
    "e2e_shift_depth": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _I64, _I32, _I32, _VP]),
    "e2e_in_stats": (C.c_int, [_VP, _I32, _I32, _I64, _F, _VP, _I32, _VP, _VP, _VP]),
    "e2e_in_apply": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _F, _I32, _I32, _I64, _VP, _VP]),
    "e2e_in_apply_from_slots": (C.c_int, [_VP, _VP, _I32, _F, _VP, _VP, _F, _I32, _I32, _I64, _VP, _VP, _VP, _VP]),
    "e2e_in_apply_pool_from_slots": (C.c_int, [_VP, _VP, _I32, _F, _VP, _VP, _F, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32,
                                               _VP, _VP, _VP, _VP, _VP, _VP]),
    "e2e_in_bwd": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _F, _I32, _I32, _I64, _VP, _I32, _VP, _VP, _VP, _VP,
                             _VP, _VP]),
    "e2e_in_bwd_scratch_floats": (C.c_int64, [_I32, _I32, _I64]),
    "e2e_in_bwd_fused": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _F, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32,
                                   _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "e2e_in_apply_pool": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _F, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP, _VP,
                                    _VP, _VP]),
    "e2e_in_bwd_pool": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _F, _I32, _I32, _I32, _I32, _I32, _I32, _I32,
                                  _I32, _VP, _I32, _VP, _VP, _VP, _VP, _VP, _VP]),
    "e2e_maxpool_fwd": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    "e2e_maxpool_bwd": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP]),
    "e2e_add_inplace": (C.c_int, [_VP, _VP, _I64, _VP]),
    "e2e_mask_apply_multi": (C.c_int, [_VP, _VP, _VP, _VP, _I32, _I64, _VP]),
    "e2e_mask_kernel_l1": (C.c_int, [_VP, _I32, _I32, _I32, _I32, _I32, _VP, _VP]),
    "e2e_mask_kth": (C.c_int, [_VP, _I32, _I32, _VP, _VP, _VP]),
    "e2e_mask_kill": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _VP, _VP]),
    "e2e_mask_dead_list": (C.c_int, [_VP, _I32, _I32, _VP, _VP, _VP, _VP]),
    "e2e_mask_grow": (C.c_int, [_VP, _VP, _VP, _I32, _I32, _VP]),
    "e2e_mask_counts": (C.c_int, [_VP, _VP, _I64, _VP, _VP]),
    "e2e_sgd_partial_count": (C.c_int, [_I32, _I64]),
    "e2e_sgd_clip_coef": (C.c_int, [_VP, _I32, _I64, _VP, _VP, _VP, _VP, _VP]),
    "e2e_sgd_update": (C.c_int, [_VP, _I32, _I64, _VP, _VP, _I32, _VP]),
    "e2e_softmax_stats_partial_count": (C.c_int, [_I32, _I32, _I64]),
    "e2e_softmax_stats_fwd": (C.c_int, [_VP, _VP, _I32, _I32, _I64, _VP, _VP, _VP, _VP]),
    "e2e_dc_ce_from_stats": (C.c_int, [_VP, _VP, _I32, _I32, _I64, _F, _I32, _I32, _F, _F, _VP, _VP, _VP, _VP, _VP]),
    "e2e_softmax_stats_bwd": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _I32, _I32, _I64, _VP, _VP]),
    "e2e_window_accumulate": (C.c_int, [_VP, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32,
                                        _I32, _I32, _F, _I32, _I32, _VP]),
    "e2e_window_head_accumulate": (C.c_int, [_VP, _I32, _VP, _I32, _VP, _VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32,
                                             _I32, _I32, _I32, _F, _I32, _VP]),
    "e2e_resample_argmax": (C.c_int, [_VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP, _VP, _VP]),
    "e2e_window_finalize": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _VP, _VP]),
    "e2e_window_finalize_range": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _VP, _VP]),
}

_libs = {}
_lock = threading.Lock()


class E2EError(RuntimeError):
    pass


def precision() -> str:
    return _precision


def set_precision(name: str):
    """selects which build of the library the ops use from now on ("bf16" | "fp16").  Call it before building
    networks / training steps (packed operands and captured graphs belong to one precision); ops.set_precision()
    is the public wrapper that also invalidates the packed-operand cache."""
    global _precision
    if name not in LIB_PATHS:
        raise ValueError("precision must be one of %s" % sorted(LIB_PATHS))
    _precision = name


def act_dtype():
    """torch dtype of C8 activations / gradients under the current precision"""
    import torch
    return torch.float16 if _precision == "fp16" else torch.bfloat16


def load(which: str = None):
    """dlopen the CUDA library of the current (or the named) precision; raises (never falls back) if it is not built."""
    which = which or _precision
    lib = _libs.get(which)
    if lib is not None:
        return lib
    with _lock:
        lib = _libs.get(which)
        if lib is not None:
            return lib
        path = LIB_PATHS[which]
        if not os.path.exists(path):
            raise E2EError(
                "%s is missing (%s). Build it with `python -m e2enet_medical_b200.build`; "
                "this package has no CPU / eager fallback." % (os.path.basename(path), path))
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)       # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.e2e_precision().decode() != which:
            raise E2EError("%s was built for %s, expected %s" % (path, lib.e2e_precision().decode(), which))
        _libs[which] = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().e2e_last_error()
        raise E2EError("%s failed (rc=%d): %s" % (what or "e2enet_b200 call", rc, (msg or b"").decode()))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    """kernels launched by this process through the C ABI (all loaded precisions)"""
    load()
    return sum(int(l.e2e_launch_count()) for l in _libs.values())

"""ctypes binding of libitr_b200.so (declared in include/itr_b200.h).

The library is the product: there is no Python/CPU fallback.  If it is missing the
import raises, telling the user how to build it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libitr_b200.so")

ITR_OK, ITR_ERR_INVALID, ITR_ERR_CUDA, ITR_ERR_UNSUPPORTED = 0, 1, 2, 3
T2I, I2T = 0, 1
NORM_CODES = {"clipped_l2norm": 0, "l2norm": 1, "softmax": 2, "clipped": 3, "no_norm": 4}
AGG_CODES = {"LogSumExp": 0, "Mean": 1, "Max": 2, "Sum": 3}
REGIONS, EMBED, TILE_WORDS, TILE_IMAGES, GRAM_BYTES, MAX_WORDS_F32 = 36, 1024, 128, 4, 4752, 96

_p, _i, _f, _l = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes); must list every symbol of include/itr_b200.h
SIGNATURES = {
    "itr_last_error": (C.c_char_p, []),
    "itr_version": (_i, []),
    "itr_device_supported": (_i, [_i]),
    "itr_cosine_scores_f32": (_i, [_p, _p, _i, _i, _i, _p, _l, _p]),
    "itr_region_gram_f32": (_i, [_p, _i, _i, _i, _p, _p]),
    "itr_scan_scores_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p, _l, _p]),
    "itr_scan_backward_workspace_f32": (_l, [_i, _i, _i, _l, _l, _i]),
    "itr_scan_backward_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _l, _l, _i, _i, _i, _f, _f,
                                   _p, _l, _p, _p, _p, _l, _p]),
    "itr_order_scores_f32": (_i, [_p, _p, _i, _i, _i, _p, _l, _p]),
    "itr_order_backward_f32": (_i, [_p, _p, _p, _l, _p, _l, _i, _i, _i, _p, _p, _p]),
    "itr_multiview_scores_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _l, _p, _p]),
    "itr_multiview_backward_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _l, _p, _p, _p, _p, _p]),
    "itr_scan_plan_max_tiles": (_i, [_p, _i]),
    "itr_scan_plan_words": (_i, [_p, _i, _p, C.POINTER(C.c_int)]),
    "itr_scan_pack_words_bf16": (_i, [_p, _i, _i, _i, _p, _i, _p, _p, _p]),
    "itr_scan_prep_images_bf16": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "itr_scan_t2i_scores_bf16": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _i, _f, _f, _p, _l, _p]),
    "itr_scan_plan_gt_items": (_i, [_p, _i, _i, _i, _i, _p, _i, C.POINTER(C.c_int)]),
    "itr_scan_t2i_gt_thresholds_bf16": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _p, _i, _i, _i, _f, _f, _i, _i, _p, _p, _p]),
    "itr_scan_t2i_count_bf16": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _f, _f, _i, _p, _p, _p, _l, _p, _p, _p, _p, _i, _i, _p]),
    "itr_scan_caption_gram_frag_bf16": (_i, [_p, _p, _i, _p, _p]),
    "itr_scan_i2t_scores_bf16": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _i, _f, _f, _p, _l, _p]),
    "itr_scan_affinity_bf16": (_i, [_p, _i, _p, _i, _p, _p]),
    "itr_scan_caption_gram_f32": (_i, [_p, _p, _p, _p, _i, _i, _p, _p]),
    "itr_scan_epilogue_f32": (_i, [_p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _p, _l, _p]),
    "itr_scan_t2i_affinity_debug": (_i, [_p, _i, _p, _i, _i, _i, _p, _p]),
    "itr_scan_t2i_profile": (_i, [_p, _p, _i, _p, _p, _p, _i, _p, _l, _p, _i, _p]),
    "itr_scan_t2i_pair_profile": (_i, [_p, _p, _i, _p, _p, _p, _i, _p, _l, _p, _p]),
    "itr_tc_mma_microbench": (_i, [_i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "itr_tc_mma2_microbench": (_i, [_i, _i, _i, _i, _i, _i, _p, _p]),
    "itr_hinge_fwd_bwd_f32": (_i, [_p, _l, _i, _f, _i, _p, _p, _l, _p]),
    "itr_cosine_hinge_workspace_f32": (_l, [_i, _i]),
    "itr_cosine_hinge_fwd_bwd_f32": (_i, [_p, _p, _i, _i, _f, _i, _p, _p, _p, _p, _p]),
    "itr_rank_thresholds_f32": (_i, [_p, _l, _i, _i, _i, _i, _p, _p, _p]),
    "itr_rank_count_f32": (_i, [_p, _l, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "itr_rank_f64": (_i, [_p, _l, _i, _i, _i, _p, _p, _p, _p, _p]),
    "itr_scores_to_host_f64": (_i, [_p, _l, _i, _i, _p, _l, _p]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libitr_b200.so is not built: run `python __graft_entry__.py build` "
                "(or `python image-text-retrieval_b200/build.py`).  itr_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the header and the library drift apart
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    """ITR_ERR_INVALID -> ValueError (the reference raises ValueError for unknown modes,
    Objectives.py:50,71,366,457); everything else -> RuntimeError."""
    if rc == ITR_OK:
        return
    msg = lib().itr_last_error().decode("utf-8", "replace")
    if rc == ITR_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError("itr_b200 [{}]: {}".format({2: "CUDA", 3: "unsupported device"}.get(rc, rc), msg))


def norm_code(name) -> int:
    try:
        return NORM_CODES[name]
    except KeyError:
        raise ValueError("unknown first norm type: {}".format(name)) from None


def agg_code(name) -> int:
    try:
        return AGG_CODES[name]
    except KeyError:
        raise ValueError("unknown aggfunc: {}".format(name)) from None


def ptr(t):
    """Device (or pinned-host) pointer of a torch tensor / numpy array, or NULL for None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream

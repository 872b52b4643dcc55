"""Tensor-level wrappers over the C ABI (device memory, streams and nothing else).

Every function takes CUDA tensors, launches on the current torch stream and returns CUDA
tensors.  There is no CPU path: a CPU tensor is an error.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _capi as capi
from ._capi import check, ptr, stream_ptr


def _cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("{} must be a torch tensor".format(name))
    if not t.is_cuda:
        raise RuntimeError("{} lives on {}; itr_b200 runs on CUDA only (no CPU fallback)".format(name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def lengths_to_numpy(cap_lens, n_cap):
    """cap_lens may be a list, numpy array or tensor (the reference indexes it per caption)."""
    if isinstance(cap_lens, torch.Tensor):
        cap_lens = cap_lens.detach().cpu().numpy()
    ln = np.ascontiguousarray(np.asarray(cap_lens)[:n_cap], dtype=np.int32)
    if ln.shape[0] != n_cap:
        raise ValueError("cap_lens has {} entries for {} captions".format(ln.shape[0], n_cap))
    return ln


STAGED_UPLOAD_MIN_BYTES = 32 << 20
_STAGING = {}


def upload_pageable(x, dev, chunk_bytes=32 << 20):
    """Pageable host f32 tensor -> CUDA tensor at close to the PCIe rate: the array is cut into chunks, each chunk is
    copied into one of two pinned staging buffers (torch's CPU copy is multi-threaded; an explicit thread pool on top of it
    measured slower on the 16-core boxes: 190 vs 172 ms for the whole call sequence) and DMA'd from there, so the
    host copy of chunk k+1 overlaps the DMA of chunk k.  A plain ``.to(device)`` of pageable memory is staged by the
    driver single-threaded at a fraction of that."""
    dev = torch.device(dev)
    out = torch.empty(x.shape, dtype=x.dtype, device=dev)
    flat_src, flat_dst = x.reshape(-1), out.reshape(-1)
    n = flat_src.numel()
    per = max(1, chunk_bytes // x.element_size())
    key = (x.dtype, per)
    if key not in _STAGING:
        _STAGING[key] = ([torch.empty(per, dtype=x.dtype, pin_memory=True) for _ in range(2)], [None, None])
    bufs, events = _STAGING[key]
    stream = torch.cuda.current_stream(dev)
    for k, lo in enumerate(range(0, n, per)):
        hi = min(lo + per, n)
        b = k & 1
        if events[b] is not None:
            events[b].synchronize()                     # the DMA that last read this staging buffer is done
        bufs[b][: hi - lo].copy_(flat_src[lo:hi])
        flat_dst[lo:hi].copy_(bufs[b][: hi - lo], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        events[b] = ev
    return out


def _trim_to_longest(captions, ln):
    """captions (n_cap, lmax, d) -> (captions[:, :max(ln)] contiguous, max(ln)): the float32 kernels size their
    shared-memory tiles by the width they are given (<= 96 words), so hand them the true longest caption."""
    lmax = captions.size(1)
    longest = int(ln.max()) if len(ln) else lmax
    if 1 <= longest < lmax:
        return captions[:, :longest].contiguous(), longest
    return captions, lmax


# ------------------------------------------------------------------------------ VSE++
def cosine_scores(im, s):
    im, s = _cuda_f32(im, "im"), _cuda_f32(s, "s")
    if im.dim() != 2 or s.dim() != 2 or im.size(1) != s.size(1):
        raise ValueError("cosine_sim expects (n_img, d) and (n_cap, d) embeddings, got {} and {}".format(
            tuple(im.shape), tuple(s.shape)))
    out = torch.empty(im.size(0), s.size(0), device=im.device, dtype=torch.float32)
    if out.numel() == 0:
        return out
    with torch.cuda.device(im.device):
        check(capi.lib().itr_cosine_scores_f32(ptr(im), ptr(s), im.size(0), s.size(0), im.size(1), ptr(out),
                                               out.stride(0) if out.numel() else s.size(0), stream_ptr()))
    return out


def order_scores(im, s):
    """order_sim, Objectives.py:24-30."""
    im, s = _cuda_f32(im, "im"), _cuda_f32(s, "s")
    if im.dim() != 2 or s.dim() != 2 or im.size(1) != s.size(1):
        raise ValueError("order_sim expects (n_img, d) and (n_cap, d) embeddings, got {} and {}".format(
            tuple(im.shape), tuple(s.shape)))
    out = torch.empty(im.size(0), s.size(0), device=im.device, dtype=torch.float32)
    if out.numel() == 0:
        return out
    with torch.cuda.device(im.device):
        check(capi.lib().itr_order_scores_f32(ptr(im), ptr(s), im.size(0), s.size(0), im.size(1), ptr(out),
                                              max(s.size(0), 1), stream_ptr()))
    return out


def order_backward(im, s, scores, d_scores, need_im=True, need_s=True):
    im, s, scores, d_scores = (_cuda_f32(t, n) for t, n in ((im, "im"), (s, "s"), (scores, "scores"), (d_scores, "d_scores")))
    d_im = torch.empty_like(im) if need_im else None
    d_s = torch.empty_like(s) if need_s else None
    with torch.cuda.device(im.device):
        check(capi.lib().itr_order_backward_f32(ptr(im), ptr(s), ptr(scores), max(scores.stride(0), 1), ptr(d_scores),
                                                max(d_scores.stride(0), 1), im.size(0), s.size(0), im.size(1), ptr(d_im),
                                                ptr(d_s), stream_ptr()))
    return d_im, d_s


def multiview_scores(imgs, caps, need_argmax=False, max_workspace_bytes=1 << 30):
    """MultiViewMatching.forward, Fusionmodule.py:670-692: imgs (n_img, n_views, d), caps (n_cap, d)."""
    imgs, caps = _cuda_f32(imgs, "imgs"), _cuda_f32(caps, "caps")
    if imgs.dim() != 3 or caps.dim() != 2 or imgs.size(2) != caps.size(1):
        raise ValueError("MultiViewMatching expects (n_img, n_views, d) and (n_cap, d), got {} and {}".format(
            tuple(imgs.shape), tuple(caps.shape)))
    n_img, n_views, d = imgs.shape
    n_cap = caps.size(0)
    out = torch.empty(n_img, n_cap, device=imgs.device, dtype=torch.float32)
    arg = torch.empty(n_img, n_cap, device=imgs.device, dtype=torch.int32) if need_argmax else None
    if n_img == 0 or n_cap == 0:
        return (out, arg) if need_argmax else out
    chunk = max(1, min(n_img, 65535, max_workspace_bytes // (4 * n_views * n_cap)))
    with torch.cuda.device(imgs.device):
        ws = torch.empty(chunk * n_views * n_cap, device=imgs.device, dtype=torch.float32)
        for i0 in range(0, n_img, chunk):
            i1 = min(i0 + chunk, n_img)
            check(capi.lib().itr_multiview_scores_f32(ptr(imgs[i0:i1]), ptr(caps), i1 - i0, n_views, n_cap, d, ptr(ws),
                                                      ptr(out[i0:i1]), n_cap, ptr(arg[i0:i1]) if need_argmax else None,
                                                      stream_ptr()))
    return (out, arg) if need_argmax else out


def multiview_backward(imgs, caps, d_scores, arg, need_imgs=True, need_caps=True):
    imgs, caps, d_scores = _cuda_f32(imgs, "imgs"), _cuda_f32(caps, "caps"), _cuda_f32(d_scores, "d_scores")
    n_img, n_views, d = imgs.shape
    n_cap = caps.size(0)
    d_imgs = torch.empty_like(imgs) if need_imgs else None
    d_caps = torch.empty_like(caps) if need_caps else None
    with torch.cuda.device(imgs.device):
        ws = torch.empty(n_img * n_views * n_cap, device=imgs.device, dtype=torch.float32)
        check(capi.lib().itr_multiview_backward_f32(ptr(imgs), ptr(caps), n_img, n_views, n_cap, d, ptr(d_scores),
                                                    max(d_scores.stride(0), 1), ptr(arg), ptr(ws), ptr(d_imgs), ptr(d_caps),
                                                    stream_ptr()))
    return d_imgs, d_caps


# ------------------------------------------------------------------------------ SCAN, fp32 mode
def scan_scores_f32(images, captions, cap_lens, cross_attn, raw_feature_norm, agg_func, lambda_softmax, lambda_lse):
    images, captions = _cuda_f32(images, "images"), _cuda_f32(captions, "captions")
    n_img, n_reg, d = images.shape
    n_cap, lmax, d2 = captions.shape
    if d != d2:
        raise ValueError("embedding sizes differ: {} vs {}".format(d, d2))
    ln = lengths_to_numpy(cap_lens, n_cap)
    if n_cap and (ln.min() < 1 or ln.max() > lmax):
        raise ValueError("caption lengths must be in [1, {}]".format(lmax))
    norm, agg = capi.norm_code(raw_feature_norm), capi.agg_code(agg_func)
    if cross_attn not in ("t2i", "i2t"):
        raise ValueError("unknown cross_attn: {}".format(cross_attn))
    out = torch.empty(n_img, n_cap, device=images.device, dtype=torch.float32)
    if out.numel() == 0:
        return out
    captions, lmax = _trim_to_longest(captions, ln)
    lens_dev = torch.from_numpy(ln).to(images.device)
    L = capi.lib()
    with torch.cuda.device(images.device):
        gram = None
        if cross_attn == "t2i":
            gram = torch.empty(n_img, n_reg, n_reg, device=images.device, dtype=torch.float32)
            check(L.itr_region_gram_f32(ptr(images), n_img, n_reg, d, ptr(gram), stream_ptr()))
        check(L.itr_scan_scores_f32(ptr(images), ptr(gram), ptr(captions), ptr(lens_dev), n_img, n_reg, n_cap, lmax, d,
                                    capi.T2I if cross_attn == "t2i" else capi.I2T, norm, agg,
                                    float(lambda_softmax), float(lambda_lse), ptr(out), max(n_cap, 1), stream_ptr()))
    return out


def scan_backward_f32(images, captions, cap_lens, d_scores, cross_attn, raw_feature_norm, agg_func, lambda_softmax,
                      lambda_lse, max_workspace_bytes=1 << 30):
    """Gradients of sum(scores * d_scores) w.r.t. images and captions for ``scan_scores_f32`` (the training
    backward the reference gets from autograd, Models.py:219-222 over Objectives.py:329-476).  Images are
    processed in chunks so the coefficient workspace stays under ``max_workspace_bytes``."""
    images, captions = _cuda_f32(images, "images"), _cuda_f32(captions, "captions")
    d_scores = _cuda_f32(d_scores, "d_scores")
    n_img, n_reg, d = images.shape
    n_cap, lmax, d2 = captions.shape
    if d != d2:
        raise ValueError("embedding sizes differ: {} vs {}".format(d, d2))
    if tuple(d_scores.shape) != (n_img, n_cap):
        raise ValueError("d_scores must be ({}, {}), got {}".format(n_img, n_cap, tuple(d_scores.shape)))
    ln = lengths_to_numpy(cap_lens, n_cap)
    if n_cap and (ln.min() < 1 or ln.max() > lmax):
        raise ValueError("caption lengths must be in [1, {}]".format(lmax))
    norm, agg = capi.norm_code(raw_feature_norm), capi.agg_code(agg_func)
    if cross_attn not in ("t2i", "i2t"):
        raise ValueError("unknown cross_attn: {}".format(cross_attn))
    cross = capi.T2I if cross_attn == "t2i" else capi.I2T
    dev = images.device
    d_images = torch.empty_like(images)
    d_captions_full = torch.zeros_like(captions)
    if n_img == 0 or n_cap == 0:
        return d_images.zero_(), d_captions_full
    captions, lmax_used = _trim_to_longest(captions, ln)
    # the kernels see the batch's true longest caption (<= 96 words), whatever the padded width; the gradient of the
    # columns beyond it is zero
    d_captions = d_captions_full if lmax_used == lmax else torch.zeros_like(captions)
    lmax = lmax_used
    l64 = ln.astype(np.int64)
    n_words, sum_sq = int(l64.sum()), int((l64 * l64).sum())
    lens_dev = torch.from_numpy(ln).to(dev)
    L = capi.lib()
    chunk = n_img
    while chunk > 4 and L.itr_scan_backward_workspace_f32(chunk, n_reg, n_cap, n_words, sum_sq, cross) > max_workspace_bytes:
        chunk = max(4, (chunk // 2 + 3) // 4 * 4)
    ws_bytes = L.itr_scan_backward_workspace_f32(chunk, n_reg, n_cap, n_words, sum_sq, cross)
    if ws_bytes < 0:
        raise ValueError("itr_scan_backward_workspace_f32: bad shape")
    with torch.cuda.device(dev):
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        gram = None
        if cross_attn == "t2i":
            gram = torch.empty(n_img, n_reg, n_reg, device=dev, dtype=torch.float32)
            check(L.itr_region_gram_f32(ptr(images), n_img, n_reg, d, ptr(gram), stream_ptr()))
        for i0 in range(0, n_img, chunk):
            i1 = min(i0 + chunk, n_img)
            ds = d_scores[i0:i1]
            check(L.itr_scan_backward_f32(ptr(images[i0:i1]), ptr(gram[i0:i1]) if gram is not None else None, ptr(captions),
                                          ptr(lens_dev), i1 - i0, n_reg, n_cap, lmax, d, n_words, sum_sq, cross, norm, agg,
                                          float(lambda_softmax), float(lambda_lse), ptr(ds), ds.stride(0),
                                          ptr(d_images[i0:i1]), ptr(d_captions), ptr(ws), ws_bytes, stream_ptr()))
    if d_captions is not d_captions_full:
        d_captions_full[:, :lmax].copy_(d_captions)
    return d_images, d_captions_full


# ------------------------------------------------------------------------------ SCAN t2i, tensor cores
@dataclass
class PreparedImages:
    images_bf16: torch.Tensor     # (n_img, 36, 1024) bf16
    gram_pack: torch.Tensor       # (n_img, 4752) u8: fp16 off-diagonal Gram (UMMA layout) + fp32 diagonal
    n_img: int
    # multi-GPU (prepare_images_sharded): the rows this rank prepared itself are valid at once; the other ranks' rows
    # are being pulled over NVLink by the copy engines and are valid once `gathered` has been waited for
    local_rows: tuple = None
    gathered: object = None       # torch.cuda.Event recorded on the gather stream
    # host images uploaded in chunks (prepare_images_streamed): [(lo, hi, event)], rows [lo, hi) are valid once the event
    # (recorded on the upload stream after the chunk's cast / Gram kernel) has been waited for
    pending: list = None

    def rows(self, lo, hi):
        """The images [lo, hi) as a PreparedImages of their own (views)."""
        return PreparedImages(self.images_bf16[lo:hi], self.gram_pack[lo:hi], hi - lo)

    def wait_gathered(self):
        """Make the current stream wait until every row is valid: the other ranks' rows have arrived, every uploaded chunk
        is prepared (no-op for images prepared on the current stream)."""
        stream = torch.cuda.current_stream(self.images_bf16.device)
        if self.gathered is not None:
            stream.wait_event(self.gathered)
            self.gathered = None
        if self.pending:
            for _, _, ev in self.pending:
                stream.wait_event(ev)
        self.pending = None

    def row_ranges(self):
        """[(lo, hi, token)] in the order the rows become valid (this rank's own rows first, uploaded chunks in upload
        order); pass the token to wait_rows() before launching on the range."""
        if self.pending:
            rest = []
            if self.gathered is not None and self.local_rows is not None:
                lo, hi = self.local_rows
                rest = [(a, b, True) for a, b in ((0, lo), (hi, self.n_img)) if b > a]
            return list(self.pending) + rest
        if self.gathered is None or self.local_rows is None:
            return [(0, self.n_img, self.gathered is not None)]
        lo, hi = self.local_rows
        out = [(lo, hi, False)] if hi > lo else []
        out += [(a, b, True) for a, b in ((0, lo), (hi, self.n_img)) if b > a]
        return out

    def wait_rows(self, token):
        """Make the current stream wait for the rows a row_ranges() entry stands for."""
        if token is True:
            self.wait_gathered()
        elif token is not None and token is not False:
            torch.cuda.current_stream(self.images_bf16.device).wait_event(token)

    def wait_local(self):
        """Make the current stream wait for the uploaded chunks only (multi-GPU: this rank's own rows), not for the gather."""
        if self.pending:
            stream = torch.cuda.current_stream(self.images_bf16.device)
            for _, _, ev in self.pending:
                stream.wait_event(ev)
        self.pending = None

    def ranges_consumed(self):
        """Every entry of row_ranges() has been waited for on the current stream: from now on all rows are valid there."""
        self.pending = None


@dataclass
class PreparedCaptions:
    words_bf16: torch.Tensor      # (n_tiles*128, 1024) bf16
    row_meta: torch.Tensor        # (n_tiles*128, 4) i32, device
    row_wnorm: torch.Tensor       # (n_tiles*128,) f32
    n_tiles: int
    n_cap: int
    sum_len: int
    meta_host: np.ndarray = None  # the same row metadata on the host (the ground-truth item planner reads it)
    plan_key: tuple = None        # identifies the packing (memoised plans)
    gq_frag: torch.Tensor = None  # (n_tiles, 32, 128) f32 word Gram in mma fragment order (fused i2t), built on first use


TC_MAX_WORDS = 128           # longest caption the fused tcgen05 t2i kernel scores (one 128-row word tile)
GENERIC_MAX_WORDS = 100      # longest caption whose phase-2 tile fits in shared memory (2*144*(n+1) + (n+1)^2 floats)


def tc_shapes(images, captions):
    """Shapes the tcgen05 main loop is built for: 36 regions x embed 1024."""
    return (images.dim() == 3 and captions.dim() == 3 and images.size(1) == capi.REGIONS
            and images.size(2) == capi.EMBED and captions.size(2) == capi.EMBED)


def tc_supported(images, captions, raw_feature_norm, cap_lens=None):
    """Shapes / modes the FUSED tcgen05 t2i kernel is built for."""
    return tc_shapes(images, captions) and raw_feature_norm in ("clipped_l2norm", "l2norm")


_PLAN_CACHE = {}       # (device, digest of the lengths) -> (row_meta on the device, n_tiles); a validation set is re-planned every epoch otherwise


def plan_words_device(lengths: np.ndarray, device):
    """plan_words + the upload of the row metadata, memoised on the lengths array (LRU of 16 plans per process)."""
    import hashlib
    lengths = np.ascontiguousarray(lengths, dtype=np.int32)
    key = (str(device), len(lengths), hashlib.blake2b(lengths.tobytes(), digest_size=16).digest())
    hit = _PLAN_CACHE.pop(key, None)
    if hit is None:
        meta_host, n_tiles = plan_words(lengths)
        hit = (torch.from_numpy(meta_host).to(device), n_tiles, meta_host, key)
    _PLAN_CACHE[key] = hit                      # most recently used last
    while len(_PLAN_CACHE) > 16:
        _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
    return hit


_GT_ITEMS_CACHE = {}


def gt_items(pc: "PreparedCaptions", cap_offset, caps_per_img, n_img):
    """(items int32 (n_items, 4) on the device, n_items): the (leader's word tile, peer's word tile, image tile, 0) items
    of the ground-truth pre-pass for this packing (csrc/plan.cpp, itr_scan_plan_gt_items), memoised next to the plan."""
    key = (pc.plan_key, int(cap_offset), int(caps_per_img), int(n_img), str(pc.row_meta.device))
    hit = _GT_ITEMS_CACHE.pop(key, None) if pc.plan_key is not None else None
    if hit is None:
        meta = pc.meta_host if pc.meta_host is not None else pc.row_meta.cpu().numpy()
        meta = np.ascontiguousarray(meta, dtype=np.int32)
        L = capi.lib()
        n = capi.C.c_int(0)
        check(L.itr_scan_plan_gt_items(meta.ctypes.data, pc.n_tiles, int(cap_offset), int(caps_per_img), int(n_img), None, 0,
                                       capi.C.byref(n)))
        items = np.empty((max(n.value, 1), 4), dtype=np.int32)
        check(L.itr_scan_plan_gt_items(meta.ctypes.data, pc.n_tiles, int(cap_offset), int(caps_per_img), int(n_img),
                                       items.ctypes.data, n.value, capi.C.byref(n)))
        hit = (torch.from_numpy(items).to(pc.row_meta.device), n.value)
    if pc.plan_key is not None:
        _GT_ITEMS_CACHE[key] = hit
        while len(_GT_ITEMS_CACHE) > 32:
            _GT_ITEMS_CACHE.pop(next(iter(_GT_ITEMS_CACHE)))
    return hit


def plan_words(lengths: np.ndarray):
    """Host-side bin packing (csrc/plan.cpp).  Returns (row_meta i32 (n_tiles*128, 4), n_tiles)."""
    lengths = np.ascontiguousarray(lengths, dtype=np.int32)
    L = capi.lib()
    n_tiles = L.itr_scan_plan_max_tiles(lengths.ctypes.data, len(lengths))
    if n_tiles < 0:
        check(-n_tiles)
    meta = np.empty((max(n_tiles, 1) * capi.TILE_WORDS, 4), dtype=np.int32)
    got = capi.C.c_int(0)
    check(L.itr_scan_plan_words(lengths.ctypes.data, len(lengths), meta.ctypes.data, capi.C.byref(got)))
    assert got.value == n_tiles
    return meta[: n_tiles * capi.TILE_WORDS], n_tiles


def prepare_images(images, out=None, gram=None) -> PreparedImages:
    images = _cuda_f32(images, "images")
    n_img = images.size(0)
    if out is None:
        out = torch.empty(n_img, capi.REGIONS, capi.EMBED, device=images.device, dtype=torch.bfloat16)
    if gram is None:
        gram = torch.empty(n_img, capi.GRAM_BYTES, device=images.device, dtype=torch.uint8)
    assert out.is_contiguous() and gram.is_contiguous() and out.shape[0] == n_img and gram.shape[0] == n_img
    with torch.cuda.device(images.device):
        check(capi.lib().itr_scan_prep_images_bf16(ptr(images), n_img, images.size(1), images.size(2), ptr(out), ptr(gram),
                                                   stream_ptr()))
    return PreparedImages(out, gram, n_img)


STREAMED_IMAGES_MIN_BYTES = 128 << 20
STREAMED_SHARD_MIN_BYTES = 48 << 20       # multi-GPU: a rank's own slice of the images
STREAMED_IMAGE_CHUNKS = 8
_UPLOAD_STREAMS = {}


def _upload_stream(dev):
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _UPLOAD_STREAMS:
        _UPLOAD_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _UPLOAD_STREAMS[key]


def prepare_images_streamed(images, device, chunks=None) -> PreparedImages:
    """prepare_images for a large HOST image array (pinned or pageable float32): the array is uploaded in `chunks` row
    ranges on an upload stream, each range cast / Grammed as soon as it has landed, and the result's row_ranges() hand the
    ranges out in upload order with the event to wait for -- the score kernels start on the first range while the others
    are still on the PCIe bus (at COCO-5K size the 737 MB of images are 14 ms of transfer nothing else could hide).
    Small arrays, and arrays already on the device, take prepare_images directly."""
    dev = torch.device(device)
    t = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images))
    if t.dtype != torch.float32:
        t = t.float()
    nbytes = t.numel() * 4
    if t.is_cuda or nbytes < STREAMED_IMAGES_MIN_BYTES or t.dim() != 3:
        if not t.is_cuda and not t.is_pinned() and nbytes >= STAGED_UPLOAD_MIN_BYTES:
            return prepare_images(upload_pageable(t.contiguous(), dev))
        return prepare_images(t.to(dev, non_blocking=True))
    t = t.contiguous()
    n_img = t.size(0)
    k = int(chunks or STREAMED_IMAGE_CHUNKS)
    per = max(capi.TILE_IMAGES, -(-n_img // k) // capi.TILE_IMAGES * capi.TILE_IMAGES)       # whole image tiles per range
    per += capi.TILE_IMAGES if per * k < n_img else 0
    img_all = torch.empty(n_img, capi.REGIONS, capi.EMBED, device=dev, dtype=torch.bfloat16)
    gram_all = torch.empty(n_img, capi.GRAM_BYTES, device=dev, dtype=torch.uint8)
    main, up = torch.cuda.current_stream(dev), _upload_stream(dev)
    up.wait_stream(main)
    pending = []
    with torch.cuda.stream(up):
        for lo in range(0, n_img, per):
            hi = min(lo + per, n_img)
            stage = t[lo:hi].to(dev, non_blocking=True) if t.is_pinned() else upload_pageable(t[lo:hi], dev)
            prepare_images(stage, out=img_all[lo:hi], gram=gram_all[lo:hi])
            ev = torch.cuda.Event()
            ev.record(up)
            pending.append((lo, hi, ev))
    for x in (img_all, gram_all):
        x.record_stream(up)
    return PreparedImages(img_all, gram_all, n_img, pending=pending)


_SYM = {}              # (group name, device) -> symmetric-memory workspace of the image gather
_GATHER_STREAMS = {}


def _gather_stream(dev):
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _GATHER_STREAMS:
        _GATHER_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _GATHER_STREAMS[key]


def _sym_workspace(group, nbytes, dev):
    """A symmetric-memory buffer of >= nbytes on every rank of `group` (torch.distributed._symmetric_memory): each rank
    can map its peers' buffers and pull from them with plain device-to-device copies, which run on the copy engines
    over NVLink and leave every SM to the score kernel.  None when symmetric memory is unavailable (-> NCCL)."""
    import os
    import torch.distributed as dist
    if os.environ.get("ITR_B200_GATHER", "").lower() == "nccl":
        return None
    try:
        import torch.distributed._symmetric_memory as symm
        pg = group if group is not None else dist.group.WORLD
        if dist.get_backend(pg) != "nccl":
            return None
        name = pg.group_name
        key = (name, str(dev))
        ent = _SYM.get(key)
        if ent is None or ent["nbytes"] < nbytes:
            size = max(nbytes, 1 << 20)
            buf = symm.empty(size, dtype=torch.uint8, device=dev)
            hdl = symm.rendezvous(buf, name)
            ent = {"buf": buf, "hdl": hdl, "nbytes": size, "done": None}
            _SYM[key] = ent
        return ent
    except Exception as exc:      # noqa: BLE001  (no fabric / IPC support on this box: fall back to the NCCL all-gather)
        _SYM[("failed", str(dev))] = repr(exc)
        os.environ["ITR_B200_GATHER"] = "nccl"
        return None


def image_shard_bounds(n_img, world):
    per = (n_img + world - 1) // world
    return per, [(min(r * per, n_img), min((r + 1) * per, n_img)) for r in range(world)]


def prepare_images_sharded(images, group=None, device=None) -> PreparedImages:
    """Multi-GPU form of prepare_images: every rank casts / Grams only its slice of the images (and, for host
    input, uploads only that slice); the other slices are then PULLED from the peers' symmetric-memory buffers by the
    copy engines on a side stream (no SMs, no NCCL kernel), so the caller can score its own slice meanwhile:
    the result's `local_rows` are valid at once, the rest after `wait_gathered()` (every consumer in this module
    waits by itself).  Without symmetric memory the slices are all-gathered with NCCL before returning.
    `images` is the FULL (n_img, 36, 1024) array on every rank: a host numpy array / tensor or a CUDA tensor."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        t = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images))
        d_ = torch.device(device if device is not None else "cuda")
        if not t.is_cuda and not t.is_pinned() and t.dtype == torch.float32 and t.numel() * 4 >= STAGED_UPLOAD_MIN_BYTES:
            return prepare_images(upload_pageable(t.contiguous(), d_))
        return prepare_images(t.to(d_, non_blocking=True))
    rank = dist.get_rank(group)
    n_img = len(images)
    per, bounds = image_shard_bounds(n_img, world)
    lo, hi = bounds[rank]
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    sl = images[lo:hi]
    if not isinstance(sl, torch.Tensor):
        sl = torch.from_numpy(np.ascontiguousarray(sl))
    if sl.dtype != torch.float32:
        sl = sl.float()
    img_all = torch.empty(world * per, capi.REGIONS, capi.EMBED, device=dev, dtype=torch.bfloat16)
    gram_all = torch.empty(world * per, capi.GRAM_BYTES, device=dev, dtype=torch.uint8)
    img_bytes, gram_bytes = per * capi.REGIONS * capi.EMBED * 2, per * capi.GRAM_BYTES
    ws = _sym_workspace(group, img_bytes + gram_bytes, dev)
    # a large HOST slice is uploaded in chunks on the upload stream (symmetric-memory path): this rank scores its own rows
    # chunk by chunk as they land, the peers pull the slice once it is complete
    streamed = ws is not None and not sl.is_cuda and hi > lo and sl.numel() * 4 >= STREAMED_SHARD_MIN_BYTES
    if not streamed:
        if not sl.is_cuda and not sl.is_pinned() and sl.numel() * 4 >= STAGED_UPLOAD_MIN_BYTES:
            sl = upload_pageable(sl.contiguous(), dev)
        else:
            sl = sl.to(dev, non_blocking=True)
    if ws is not None:
        main = torch.cuda.current_stream(dev)
        if ws["done"] is not None:
            main.wait_event(ws["done"])                    # the previous gather's last pull from this buffer is over
        mine = ws["buf"]
        sym_img = mine[:img_bytes].view(torch.bfloat16).view(per, capi.REGIONS, capi.EMBED)
        sym_gram = mine[img_bytes: img_bytes + gram_bytes].view(per, capi.GRAM_BYTES)
        pending = None
        side = _gather_stream(dev)
        if streamed:
            sl = sl.contiguous()
            up = _upload_stream(dev)
            up.wait_stream(main)
            n_loc = hi - lo
            step_rows = max(capi.TILE_IMAGES, -(-n_loc // STREAMED_IMAGE_CHUNKS) // capi.TILE_IMAGES * capi.TILE_IMAGES)
            step_rows += capi.TILE_IMAGES if step_rows * STREAMED_IMAGE_CHUNKS < n_loc else 0
            pending = []
            with torch.cuda.stream(up):
                for a in range(0, n_loc, step_rows):
                    b = min(a + step_rows, n_loc)
                    stage = sl[a:b].to(dev, non_blocking=True) if sl.is_pinned() else upload_pageable(sl[a:b], dev)
                    prepare_images(stage, out=sym_img[a:b], gram=sym_gram[a:b])
                    img_all[lo + a: lo + b].copy_(sym_img[a:b])
                    gram_all[lo + a: lo + b].copy_(sym_gram[a:b])
                    ev = torch.cuda.Event()
                    ev.record(up)
                    pending.append((lo + a, lo + b, ev))
            for x in (img_all, gram_all):
                x.record_stream(up)
            side.wait_stream(main)
            side.wait_event(pending[-1][2])
        else:
            if hi > lo:
                prepare_images(sl, out=sym_img[: hi - lo], gram=sym_gram[: hi - lo])
                img_all[lo:hi].copy_(sym_img[: hi - lo])
                gram_all[lo:hi].copy_(sym_gram[: hi - lo])
            side.wait_stream(main)
        with torch.cuda.stream(side):
            hdl = ws["hdl"]
            hdl.barrier()                                   # every rank's slice is written
            all_bytes = img_all.view(torch.uint8).view(world, img_bytes)
            for step in range(1, world):
                r = (rank - step) % world
                n_r = bounds[r][1] - bounds[r][0]
                if n_r <= 0:
                    continue
                src = hdl.get_buffer(r, (img_bytes + gram_bytes,), torch.uint8)
                all_bytes[r, : n_r * capi.REGIONS * capi.EMBED * 2].copy_(src[: n_r * capi.REGIONS * capi.EMBED * 2])
                gram_all[r * per: r * per + n_r].view(-1).copy_(src[img_bytes: img_bytes + n_r * capi.GRAM_BYTES])
            hdl.barrier()                                   # nobody rewrites its slice while a peer still reads it
            done = torch.cuda.Event()
            done.record(side)
        ws["done"] = done
        for t in (img_all, gram_all):
            t.record_stream(side)
        return PreparedImages(img_all[:n_img], gram_all[:n_img], n_img, local_rows=(lo, hi), gathered=done, pending=pending)
    img_loc = img_all[rank * per:(rank + 1) * per]          # all_gather_into_tensor gathers in place
    gram_loc = gram_all[rank * per:(rank + 1) * per]
    if hi > lo:
        prepare_images(sl, out=img_loc[: hi - lo], gram=gram_loc[: hi - lo])
    if hi - lo < per:
        img_loc[hi - lo:].zero_()
        gram_loc[hi - lo:].zero_()
    dist.all_gather_into_tensor(img_all, img_loc, group=group)
    dist.all_gather_into_tensor(gram_all, gram_loc, group=group)
    return PreparedImages(img_all[:n_img], gram_all[:n_img], n_img)


def all_ranks_agree(flag: bool, group, device) -> bool:
    """True iff `flag` is true on every rank of `group` (one all-reduce(MIN) of an int; no-op without a process group)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bool(flag)
    backend = dist.get_backend(group)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device if backend == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


def prepare_captions(captions, cap_lens, device=None) -> PreparedCaptions:
    """captions: CUDA f32 tensor, or a PINNED host f32 tensor (read in place over PCIe: only the
    true words of each caption cross the bus, not the zero padding)."""
    if not isinstance(captions, torch.Tensor):
        raise TypeError("captions must be a torch tensor")
    if captions.is_cuda:
        captions = _cuda_f32(captions, "captions")
        device = captions.device
    else:
        if not captions.is_pinned() or captions.dtype != torch.float32 or not captions.is_contiguous():
            raise RuntimeError("host captions must be a pinned, contiguous float32 tensor")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n_cap, lmax, d = captions.shape
    ln = lengths_to_numpy(cap_lens, n_cap)
    if n_cap and ln.max() > lmax:
        raise ValueError("caption length {} exceeds the padded width {}".format(int(ln.max()), lmax))
    meta, n_tiles, meta_host, plan_key = plan_words_device(ln, device)
    rows = n_tiles * capi.TILE_WORDS
    words = torch.empty(rows, capi.EMBED, device=device, dtype=torch.bfloat16)
    wnorm = torch.empty(rows, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        check(capi.lib().itr_scan_pack_words_bf16(ptr(captions), n_cap, lmax, d, ptr(meta), n_tiles, ptr(words), ptr(wnorm),
                                                  stream_ptr()))
    return PreparedCaptions(words, meta, wnorm, n_tiles, n_cap, int(ln.sum()), meta_host, plan_key)


def scan_t2i_scores_bf16(pi: PreparedImages, pc: PreparedCaptions, raw_feature_norm, agg_func, lambda_softmax,
                         lambda_lse, out=None):
    norm, agg = capi.norm_code(raw_feature_norm), capi.agg_code(agg_func)
    dev = pi.images_bf16.device
    if out is None:
        out = torch.empty(pi.n_img, pc.n_cap, device=dev, dtype=torch.float32)
    assert out.is_cuda and out.dtype == torch.float32 and out.stride(1) == 1 and out.shape == (pi.n_img, pc.n_cap)
    with torch.cuda.device(dev):
        # multi-GPU: this rank's own image rows first, the others once the copy engines have delivered them
        for lo, hi, token in pi.row_ranges():
            pi.wait_rows(token)
            o = out[lo:hi]
            check(capi.lib().itr_scan_t2i_scores_bf16(ptr(pi.images_bf16[lo:hi]), ptr(pi.gram_pack[lo:hi]), hi - lo,
                                                      ptr(pc.words_bf16), ptr(pc.row_meta), ptr(pc.row_wnorm), pc.n_tiles, norm,
                                                      agg, float(lambda_softmax), float(lambda_lse), ptr(o),
                                                      out.stride(0) if out.numel() else max(pc.n_cap, 1), stream_ptr()))
        pi.ranges_consumed()
    return out


def scan_t2i_gt_thresholds(pi: PreparedImages, pc: PreparedCaptions, raw_feature_norm, agg_func, lambda_softmax, lambda_lse,
                           cap_offset=0, caps_per_img=5):
    """Ground-truth pre-pass of the fused evaluation (itr_scan_t2i_gt_thresholds_bf16).  Returns (thr_row (n_img,),
    thr_col (n_cap,)): thr_col[c] = score of local caption c with its own image, thr_row[i] = best score of image i with
    its captions among THESE captions (-inf if none) -- all-reduce(MAX) it across caption shards before counting."""
    norm, agg = capi.norm_code(raw_feature_norm), capi.agg_code(agg_func)
    dev = pi.images_bf16.device
    pi.wait_gathered()
    thr_row = torch.empty(pi.n_img, device=dev, dtype=torch.float32)
    thr_col = torch.empty(pc.n_cap, device=dev, dtype=torch.float32)
    items, n_items = gt_items(pc, cap_offset, caps_per_img, pi.n_img)
    with torch.cuda.device(dev):
        check(capi.lib().itr_scan_t2i_gt_thresholds_bf16(ptr(pi.images_bf16), ptr(pi.gram_pack), pi.n_img, ptr(pc.words_bf16),
                                                         ptr(pc.row_meta), ptr(pc.row_wnorm), pc.n_tiles, pc.n_cap, ptr(items),
                                                         n_items, norm, agg, float(lambda_softmax), float(lambda_lse),
                                                         int(cap_offset), int(caps_per_img), ptr(thr_col), ptr(thr_row),
                                                         stream_ptr()))
    return thr_row, thr_col


def scan_t2i_count(pi: PreparedImages, pc: PreparedCaptions, raw_feature_norm, agg_func, lambda_softmax, lambda_lse,
                   thr_row, thr_col, cap_offset=0, out=None):
    """Counting pass of the fused evaluation (itr_scan_t2i_count_bf16): the score kernel compares every score with its
    row / column threshold as it is produced.  Returns (cnt_row i32 (n_img,), cnt_col i32 (n_cap,), best_row u64-as-i64,
    best_col u64-as-i64) in the formats of rank_count; the score matrix is written only if `out` is given.
    Multi-GPU: this rank's own image rows are counted while the other ranks' rows are still arriving."""
    norm, agg = capi.norm_code(raw_feature_norm), capi.agg_code(agg_func)
    dev = pi.images_bf16.device
    cnt_row = torch.empty(pi.n_img, device=dev, dtype=torch.int32)
    cnt_col = torch.empty(pc.n_cap, device=dev, dtype=torch.int32)
    best_row = torch.empty(pi.n_img, device=dev, dtype=torch.int64)
    best_col = torch.empty(pc.n_cap, device=dev, dtype=torch.int64)
    if out is not None:
        assert out.is_cuda and out.dtype == torch.float32 and out.stride(1) == 1 and out.shape == (pi.n_img, pc.n_cap)
    with torch.cuda.device(dev):
        for k, (lo, hi, token) in enumerate(pi.row_ranges()):
            pi.wait_rows(token)
            o = out[lo:hi] if out is not None else None
            check(capi.lib().itr_scan_t2i_count_bf16(ptr(pi.images_bf16[lo:hi]), ptr(pi.gram_pack[lo:hi]), hi - lo,
                                                     ptr(pc.words_bf16), ptr(pc.row_meta), ptr(pc.row_wnorm), pc.n_tiles,
                                                     pc.n_cap, norm, agg, float(lambda_softmax), float(lambda_lse),
                                                     int(cap_offset), ptr(thr_col), ptr(thr_row[lo:hi]), ptr(o),
                                                     out.stride(0) if out is not None else 0, ptr(cnt_row[lo:hi]), ptr(cnt_col),
                                                     ptr(best_row[lo:hi]), ptr(best_col), int(lo), int(k > 0), stream_ptr()))
        pi.ranges_consumed()
    return cnt_row, cnt_col, best_row, best_col


def host_caption_chunks(lens, fractions=(1.0 / 16, 3.0 / 16, 3.0 / 4), multiple=5, min_words=2048):
    """[(c0, c1)] caption ranges for the pipelined host path: a small first chunk (its PCIe gather is the only one
    nobody hides), then growing ones.  Boundaries on multiples of `multiple`; small inputs stay in one piece."""
    n_cap = len(lens)
    total = int(np.sum(lens))
    if n_cap == 0 or total < 4 * min_words:
        return [(0, n_cap)]
    csum = np.cumsum(lens)
    cuts, acc = [], 0.0
    for f in fractions[:-1]:
        acc += f
        c = int(np.searchsorted(csum, acc * total)) + 1
        c = min(n_cap, (c + multiple - 1) // multiple * multiple)
        if c > (cuts[-1] if cuts else 0) and c < n_cap:
            cuts.append(c)
    edges = [0] + cuts + [n_cap]
    return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def scores_to_host_f64(block, host_block):
    """Device float32 score block -> the matching block of a MAPPED page-locked float64 host matrix
    (itr_scores_to_host_f64) on the current stream.  host_block: a CPU float64 tensor view (stride(1) == 1) of pinned memory."""
    assert block.is_cuda and block.dtype == torch.float32 and block.stride(1) == 1
    assert not host_block.is_cuda and host_block.dtype == torch.float64 and host_block.stride(1) == 1 and host_block.shape == block.shape
    with torch.cuda.device(block.device):
        check(capi.lib().itr_scores_to_host_f64(ptr(block), block.stride(0), block.size(0), block.size(1), ptr(host_block),
                                                host_block.stride(0), stream_ptr()))


def scan_t2i_scores_from_host(pi: PreparedImages, captions, cap_lens, raw_feature_norm, agg_func, lambda_softmax, lambda_lse,
                              device=None, chunks=None, on_block=None):
    """Fused t2i scores for captions living in PINNED host memory: the gather of chunk k+1 over PCIe
    (itr_scan_pack_words_bf16 on a side stream) runs under the score kernel of chunk k, so only the first, small
    chunk's transfer is exposed.  Returns the (n_img, n_cap) score matrix on the current stream."""
    n_cap = captions.size(0)
    ln = lengths_to_numpy(cap_lens, n_cap)
    dev = pi.images_bf16.device
    out = torch.empty(pi.n_img, n_cap, device=dev, dtype=torch.float32)
    # images still arriving in chunks (prepare_images_streamed): the first caption chunk is scored range by range as they
    # land, so it is made large enough to keep the SMs busy for the whole upload, and the later caption chunks leave the
    # PCIe bus to the images until those are all in
    images_landed = pi.pending[-1][2] if pi.pending else None
    if chunks is None:
        if on_block is not None:
            # the caller ships every finished block to the host while the next is scored: small blocks at the end, so that
            # little is left to ship when the last kernel is done
            chunks = host_caption_chunks(ln, fractions=(3.0 / 16, 5.0 / 16, 1.0 / 4, 1.0 / 8, 1.0 / 8))
        elif images_landed is not None:
            chunks = host_caption_chunks(ln, fractions=(3.0 / 16, 5.0 / 16, 1.0 / 2))
        else:
            chunks = host_caption_chunks(ln)
    if len(chunks) == 1:
        pc = prepare_captions(captions, ln, device=dev)
        scan_t2i_scores_bf16(pi, pc, raw_feature_norm, agg_func, lambda_softmax, lambda_lse, out=out)
        if on_block is not None:
            on_block(out, 0, n_cap)
        return out
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(main)
    for k, (c0, c1) in enumerate(chunks):
        with torch.cuda.stream(side):
            if k > 0 and images_landed is not None:
                side.wait_event(images_landed)
            pc = prepare_captions(captions[c0:c1], ln[c0:c1], device=dev)
            ready = side.record_event()
        main.wait_event(ready)
        for t in (pc.words_bf16, pc.row_meta, pc.row_wnorm):
            t.record_stream(main)
        scan_t2i_scores_bf16(pi, pc, raw_feature_norm, agg_func, lambda_softmax, lambda_lse, out=out[:, c0:c1])
        if on_block is not None:
            on_block(out, c0, c1)
    return out


_SIDE_STREAMS = {}


def _side_stream(dev):
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def caption_rows(pc: PreparedCaptions, lengths):
    """cap_row0[c] = packed row index of caption c's first word (a caption's words are consecutive packed rows)."""
    meta = pc.row_meta
    first = (meta[:, 0] >= 0) & (meta[:, 1] == 0)
    rows = torch.nonzero(first, as_tuple=False).flatten().to(torch.int32)
    cap_row0 = torch.empty(pc.n_cap, device=meta.device, dtype=torch.int32)
    cap_row0[meta[rows.long(), 0].long()] = rows
    return cap_row0


def scan_scores_tc_generic(images, captions, cap_lens, cross_attn, raw_feature_norm, agg_func, lambda_softmax, lambda_lse,
                           pi: PreparedImages = None, pc: PreparedCaptions = None, max_affinity_bytes=6 << 30):
    """Two-phase tensor-core path for everything the fused t2i kernel does not cover (i2t, softmax / clipped /
    no_norm feature norms): tcgen05 affinities dumped per image chunk (itr_scan_affinity_bf16), then the reference's
    epilogue in fp32 (itr_scan_epilogue_f32).  Inputs are rounded to bf16 for the D-wide contraction only."""
    norm, agg = capi.norm_code(raw_feature_norm), capi.agg_code(agg_func)
    if cross_attn not in ("t2i", "i2t"):
        raise ValueError("unknown cross_attn: {}".format(cross_attn))
    if pi is None:
        pi = prepare_images(images)
    pi.wait_gathered()
    ln = lengths_to_numpy(cap_lens, len(cap_lens))
    if len(ln) and (ln.min() < 1 or ln.max() > GENERIC_MAX_WORDS):
        raise ValueError("the two-phase tensor-core path scores captions of 1..{} words, got {}..{}; use the float32 mode "
                         "(itr_b200_precision='fp32')".format(GENERIC_MAX_WORDS, int(ln.min()), int(ln.max())))
    if pc is None:
        pc = prepare_captions(captions, ln)
    dev = pi.images_bf16.device
    L = capi.lib()
    n_img, n_cap = pi.n_img, pc.n_cap
    out = torch.empty(n_img, n_cap, device=dev, dtype=torch.float32)
    if n_img == 0 or n_cap == 0:
        return out
    cap_row0 = caption_rows(pc, ln)
    lens_dev = torch.from_numpy(ln).to(dev)
    # length classes: the phase-2 kernels are sized by the longest caption of a launch
    classes = []
    for lo_len, hi_len in ((1, 8), (9, 12), (13, 16), (17, 24), (25, 32), (33, 64), (65, GENERIC_MAX_WORDS)):
        ids = np.nonzero((ln >= lo_len) & (ln <= hi_len))[0].astype(np.int32)
        if len(ids):
            classes.append((torch.from_numpy(ids).to(dev), len(ids), int(ln[ids].max())))
    region_norm = pi.gram_pack[:, 4608:].contiguous().view(torch.float32).sqrt().contiguous()      # |v_k| from the Gram diagonal
    cap_gram = gram_off = None
    with torch.cuda.device(dev):
        if cross_attn == "i2t":
            off = np.zeros(n_cap + 1, dtype=np.int64)
            np.cumsum(ln.astype(np.int64) ** 2, out=off[1:])
            gram_off = torch.from_numpy(off[:-1].copy()).to(dev)
            cap_gram = torch.empty(int(off[-1]), device=dev, dtype=torch.float32)
            check(L.itr_scan_caption_gram_f32(ptr(pc.words_bf16), ptr(cap_row0), ptr(lens_dev), ptr(gram_off), n_cap, capi.EMBED,
                                              ptr(cap_gram), stream_ptr()))
        per_img = pc.n_tiles * capi.TILE_WORDS * capi.REGIONS * 4
        chunk = max(capi.TILE_IMAGES, min(n_img, int(max_affinity_bytes // per_img)) // capi.TILE_IMAGES * capi.TILE_IMAGES)
        aff = torch.empty(pc.n_tiles * chunk * capi.TILE_WORDS * capi.REGIONS, device=dev, dtype=torch.float32)
        for i0 in range(0, n_img, chunk):
            i1 = min(i0 + chunk, n_img)
            check(L.itr_scan_affinity_bf16(ptr(pi.images_bf16[i0:i1]), i1 - i0, ptr(pc.words_bf16), pc.n_tiles, ptr(aff), stream_ptr()))
            rgram = None
            if cross_attn == "t2i":
                rgram = torch.empty(i1 - i0, capi.REGIONS, capi.REGIONS, device=dev, dtype=torch.float32)
                rounded = pi.images_bf16[i0:i1].float()
                check(L.itr_region_gram_f32(ptr(rounded), i1 - i0, capi.REGIONS, capi.EMBED, ptr(rgram), stream_ptr()))
            for ids, n_ids, max_len in classes:
                check(L.itr_scan_epilogue_f32(ptr(aff), i1 - i0, ptr(cap_row0), ptr(lens_dev), ptr(ids), n_ids, max_len,
                                              ptr(pc.row_wnorm), ptr(region_norm[i0:i1]), ptr(rgram), ptr(cap_gram), ptr(gram_off),
                                              capi.T2I if cross_attn == "t2i" else capi.I2T, norm, agg, float(lambda_softmax),
                                              float(lambda_lse), ptr(out[i0:i1]), out.stride(0), stream_ptr()))
    return out


I2T_FUSED_MAX_WORDS = 32     # longest caption the fused i2t kernel scores (a caption must fit one 32-lane quarter)


def caption_gram_frag(pc: PreparedCaptions):
    """Word Gram of the packed tiles in mma fragment order (itr_scan_caption_gram_frag_bf16), cached on the PreparedCaptions."""
    if pc.gq_frag is None:
        out = torch.empty(max(pc.n_tiles, 1), 32, capi.TILE_WORDS, device=pc.words_bf16.device, dtype=torch.float32)
        with torch.cuda.device(out.device):
            check(capi.lib().itr_scan_caption_gram_frag_bf16(ptr(pc.words_bf16), ptr(pc.row_meta), pc.n_tiles, ptr(out), stream_ptr()))
        pc.gq_frag = out
    return pc.gq_frag


def scan_i2t_scores_bf16(pi: PreparedImages, pc: PreparedCaptions, raw_feature_norm, agg_func, lambda_softmax, lambda_lse,
                         out=None):
    """Fused tensor-core i2t scores (itr_scan_i2t_scores_bf16) of every caption of up to 32 words; the columns of longer
    captions are left untouched (see scan_i2t_scores_tc)."""
    norm, agg = capi.norm_code(raw_feature_norm), capi.agg_code(agg_func)
    dev = pi.images_bf16.device
    pi.wait_gathered()
    if out is None:
        out = torch.empty(pi.n_img, pc.n_cap, device=dev, dtype=torch.float32)
    assert out.is_cuda and out.dtype == torch.float32 and out.stride(1) == 1 and out.shape == (pi.n_img, pc.n_cap)
    if pi.n_img == 0 or pc.n_cap == 0:
        return out
    region_norm = pi.gram_pack[:, 4608:].contiguous().view(torch.float32).sqrt().contiguous()      # |v_k| from the Gram diagonal
    gq = caption_gram_frag(pc)
    with torch.cuda.device(dev):
        check(capi.lib().itr_scan_i2t_scores_bf16(ptr(pi.images_bf16), ptr(region_norm), pi.n_img, ptr(pc.words_bf16),
                                                  ptr(pc.row_meta), ptr(gq), pc.n_tiles, norm, agg, float(lambda_softmax),
                                                  float(lambda_lse), ptr(out), out.stride(0), stream_ptr()))
    return out


def scan_i2t_scores_tc(images, captions, cap_lens, raw_feature_norm, agg_func, lambda_softmax, lambda_lse):
    """i2t scores on the tensor cores for raw_feature_norm in {clipped_l2norm, l2norm}: the fused kernel for captions of up
    to 32 words; the (rare) longer ones through the two-phase path (<= 100 words) or the float32 kernel."""
    ln = lengths_to_numpy(cap_lens, len(cap_lens))
    pi = prepare_images(images)
    pc = prepare_captions(captions, ln)
    out = scan_i2t_scores_bf16(pi, pc, raw_feature_norm, agg_func, lambda_softmax, lambda_lse)
    long_ids = np.nonzero(ln > I2T_FUSED_MAX_WORDS)[0]
    if len(long_ids):
        idx = torch.from_numpy(long_ids).to(out.device)
        caps_long, ln_long = captions[idx], ln[long_ids]
        if int(ln_long.max()) <= GENERIC_MAX_WORDS:
            sub = scan_scores_tc_generic(images, caps_long, ln_long, "i2t", raw_feature_norm, agg_func, lambda_softmax,
                                         lambda_lse, pi=pi)
        else:
            sub = scan_scores_f32(images, caps_long, ln_long, "i2t", raw_feature_norm, agg_func, lambda_softmax, lambda_lse)
        out[:, idx] = sub
    return out


def scan_t2i_affinity_debug(pi: PreparedImages, pc: PreparedCaptions, word_tile, image_tile):
    out = torch.empty(capi.TILE_WORDS, capi.TILE_IMAGES * capi.REGIONS, device=pi.images_bf16.device, dtype=torch.float32)
    with torch.cuda.device(out.device):
        check(capi.lib().itr_scan_t2i_affinity_debug(ptr(pi.images_bf16), pi.n_img, ptr(pc.words_bf16), pc.n_tiles,
                                                     int(word_tile), int(image_tile), ptr(out), stream_ptr()))
    return out


# ------------------------------------------------------------------------------ hinge
def hinge(scores, margin, max_violation, need_grad=True):
    scores = _cuda_f32(scores, "scores")
    if scores.dim() != 2 or scores.size(0) != scores.size(1):
        raise ValueError("the hinge loss needs a square score matrix, got {}".format(tuple(scores.shape)))
    n = scores.size(0)
    loss = torch.empty((), device=scores.device, dtype=torch.float32)
    ds = torch.empty_like(scores) if need_grad else None
    with torch.cuda.device(scores.device):
        check(capi.lib().itr_hinge_fwd_bwd_f32(ptr(scores), scores.stride(0), n, float(margin), int(bool(max_violation)),
                                               ptr(loss), ptr(ds), n, stream_ptr()))
    return loss, ds


def cosine_hinge(im, s, margin, max_violation, need_grad=True):
    im, s = _cuda_f32(im, "im"), _cuda_f32(s, "s")
    if im.shape != s.shape or im.dim() != 2:
        raise ValueError("im and s must both be (batch, d), got {} and {}".format(tuple(im.shape), tuple(s.shape)))
    n, d = im.shape
    ws = torch.empty(max(int(capi.lib().itr_cosine_hinge_workspace_f32(n, d)), 2 * n * n), device=im.device, dtype=torch.float32)
    loss = torch.empty((), device=im.device, dtype=torch.float32)
    d_im = torch.empty_like(im) if need_grad else None
    d_s = torch.empty_like(s) if need_grad else None
    with torch.cuda.device(im.device):
        check(capi.lib().itr_cosine_hinge_fwd_bwd_f32(ptr(im), ptr(s), n, d, float(margin), int(bool(max_violation)),
                                                      ptr(ws), ptr(loss), ptr(d_im), ptr(d_s), stream_ptr()))
    return loss, d_im, d_s


# ------------------------------------------------------------------------------ ranking
def rank_thresholds(scores, cap_offset=0, caps_per_img=5):
    n_img, n_cap = scores.shape
    thr_row = torch.empty(n_img, device=scores.device, dtype=torch.float32)
    thr_col = torch.empty(n_cap, device=scores.device, dtype=torch.float32)
    with torch.cuda.device(scores.device):
        check(capi.lib().itr_rank_thresholds_f32(ptr(scores), scores.stride(0), n_img, n_cap, int(cap_offset),
                                                 int(caps_per_img), ptr(thr_row), ptr(thr_col), stream_ptr()))
    return thr_row, thr_col


def rank_count(scores, thr_row, thr_col, cap_offset=0):
    n_img, n_cap = scores.shape
    dev = scores.device
    cnt_row = torch.empty(n_img, device=dev, dtype=torch.int32)
    cnt_col = torch.empty(n_cap, device=dev, dtype=torch.int32)
    best_row = torch.empty(n_img, device=dev, dtype=torch.int64)
    best_col = torch.empty(n_cap, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        check(capi.lib().itr_rank_count_f32(ptr(scores), scores.stride(0), n_img, n_cap, int(cap_offset), ptr(thr_row),
                                            ptr(thr_col), ptr(cnt_row), ptr(cnt_col), ptr(best_row), ptr(best_col),
                                            stream_ptr()))
    return cnt_row, cnt_col, best_row, best_col


def unpack_best_index(best):
    """low 32 bits of the packed arg-max key hold ~index."""
    return (~best) & 0xFFFFFFFF


def rank_f64(sims, caps_per_img=5):
    """sims: CUDA float64 (n_img, n_cap).  Returns (rank_row, rank_col, top1_row, top1_col) int32."""
    assert sims.is_cuda and sims.dtype == torch.float64 and sims.stride(1) == 1
    n_img, n_cap = sims.shape
    outs = [torch.empty(n, device=sims.device, dtype=torch.int32) for n in (n_img, n_cap, n_img, n_cap)]
    with torch.cuda.device(sims.device):
        check(capi.lib().itr_rank_f64(ptr(sims), sims.stride(0), n_img, n_cap, int(caps_per_img), *[ptr(o) for o in outs],
                                      stream_ptr()))
    return tuple(outs)

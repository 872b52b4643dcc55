"""Drop-ins for the hot-path symbols of ``itr.metricmodule.evaluation``:
``cal_sims`` (:124-153), ``i2t`` (:156-189), ``t2i`` (:192-222), ``cal_recall`` (:225-259),
plus the fused ``cal_sims_and_recall`` that ranks on the device and never ships the
matrix to the host unless asked.

Differences from the reference, all deliberate:
  D1  every caption is scored with ITS OWN length.  The reference hands the un-sliced
      ``lengths`` array to every caption block, so caption c is scored with
      ``lengths[c % shard_size]``; pass ``compat_unsliced_lengths=True`` to reproduce that.
  --  ``shard_size`` is accepted and ignored by the fused similarity functions (the whole
      matrix is produced in one launch); it still drives the blocked loop used for callables
      this package does not accelerate (``model.sim_enc``, ``model.mvm``).
  --  rank = number of scores strictly greater than the ground-truth score.  Identical to the
      reference's argsort position unless another score ties exactly with the ground truth,
      where numpy's unstable sort leaves the reference's own answer unspecified.
"""
from __future__ import annotations

import os
import threading
import time
import weakref

import numpy as np
import torch

from . import objectives, ops

CAPS_PER_IMG = 5   # hard-coded upstream, evaluation.py:173,208


# ----------------------------------------------------------------------------------------- host hand-off
class _PinnedPool:
    """Page-locked host buffers for the matrices ``cal_sims`` hands back, recycled when the array (and every view of
    it) has been garbage-collected.  Bounded: beyond ``ITR_B200_PINNED_POOL_MB`` (default 4096) the caller gets
    pageable memory, as with the reference."""

    def __init__(self):
        self.free, self.lock, self.bytes = [], threading.Lock(), 0
        self.cap = int(os.environ.get("ITR_B200_PINNED_POOL_MB", "4096")) << 20

    def take(self, nbytes):
        """A pinned uint8 tensor of at least nbytes, or None when the pool is exhausted / pinning fails."""
        with self.lock:
            fits = [i for i, t in enumerate(self.free) if nbytes <= t.numel() <= max(2 * nbytes, 1 << 20)]
            if fits:
                return self.free.pop(min(fits, key=lambda i: self.free[i].numel()))      # by index: `in` / remove() compare tensors elementwise
            if self.bytes + nbytes > self.cap:
                while self.free and self.bytes + nbytes > self.cap:
                    i = max(range(len(self.free)), key=lambda j: self.free[j].numel())
                    self.bytes -= self.free.pop(i).numel()
                if self.bytes + nbytes > self.cap:
                    return None
            self.bytes += nbytes
        try:
            return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        except RuntimeError:
            with self.lock:
                self.bytes -= nbytes
            return None

    def give_back(self, t):
        with self.lock:
            self.free.append(t)


_POOL = _PinnedPool()


def _release_pinned(t):
    _POOL.give_back(t)


class _SimsEntry:
    __slots__ = ("ref", "dev", "ranks")

    def __init__(self, ref, dev):
        self.ref, self.dev, self.ranks = ref, dev, None


_SIMS = {}      # id(host array cal_sims returned) -> _SimsEntry (device matrix it was copied from)


def _remember(arr, dev_matrix):
    key = id(arr)
    _SIMS[key] = _SimsEntry(weakref.ref(arr, lambda _r, k=key: _SIMS.pop(k, None)), dev_matrix)


def _device_twin(sims):
    """The cache entry of the device matrix behind a host array ``cal_sims`` returned, or None.  The host array is
    handed out read-only; if somebody made it writeable again its contents may have changed, so the twin is dropped."""
    if not isinstance(sims, np.ndarray):
        return None
    e = _SIMS.get(id(sims))
    if e is None or e.ref() is not sims:
        return None
    if sims.flags.writeable:
        _SIMS.pop(id(sims), None)
        return None
    return e


def _to_host_f64(d):
    """CUDA f32 (n_img, n_cap) -> host float64 ndarray: converted on the device, one DMA into a recycled pinned
    buffer (pageable when the pool is exhausted).  The result is marked read-only (see _device_twin)."""
    n = d.numel()
    if n == 0:
        return np.zeros(tuple(d.shape), dtype=np.float64)
    d64 = d.double()
    buf = _POOL.take(n * 8)
    if buf is None:
        out = d64.cpu().numpy()
    else:
        host = buf[: n * 8].view(torch.float64).view(d.shape)
        host.copy_(d64, non_blocking=True)
        torch.cuda.current_stream(d.device).synchronize()
        root = buf.numpy()
        weakref.finalize(root, _release_pinned, buf)
        out = root[: n * 8].view(np.float64).reshape(tuple(d.shape))
        del root
    out.setflags(write=False)
    return out


_D2H_STREAMS = {}


class _HostMatrixWriter:
    """The float64 host matrix of ``cal_sims``, filled block by block while the score kernels are still running: every
    finished column block is converted and written straight into a page-locked buffer by a small kernel on its own
    stream (itr_scores_to_host_f64), so only the last, small block's transfer is left when the last kernel ends.  Used
    when the similarity function reports its blocks (tensor-core SCAN t2i); anything else takes _to_host_f64."""

    def __init__(self, n_img, n_cap):
        self.shape, self.buf, self.host, self.done, self.stream = (n_img, n_cap), None, None, 0, None

    def block(self, matrix, c0, c1):
        if os.environ.get("ITR_B200_HOST_MATRIX", "") == "oneshot":      # A/B switch: convert and copy the whole matrix at the end
            self.done = -1
            return
        if self.done != c0 or tuple(matrix.shape) != self.shape or matrix.dtype != torch.float32:
            self.done = -1                               # not the contiguous left-to-right sequence this writer expects
            return
        if self.buf is None:
            n = matrix.numel()
            self.buf = _POOL.take(n * 8) if n else None
            if self.buf is None:
                self.done = -1
                return
            self.host = self.buf[: n * 8].view(torch.float64).view(self.shape)
            dev = matrix.device
            key = dev.index if dev.index is not None else torch.cuda.current_device()
            if key not in _D2H_STREAMS:
                _D2H_STREAMS[key] = torch.cuda.Stream(device=dev)
            self.stream = _D2H_STREAMS[key]
        ready = torch.cuda.current_stream(matrix.device).record_event()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            ops.scores_to_host_f64(matrix[:, c0:c1], self.host[:, c0:c1])
        self.done = c1

    def finish(self, matrix):
        """The finished read-only ndarray, or None when the blocks did not cover the matrix (the caller converts it whole)."""
        if self.buf is None or self.done != self.shape[1] or self.done < 0:
            if self.buf is not None:
                if self.stream is not None:
                    self.stream.synchronize()
                _release_pinned(self.buf)
            return None
        matrix.record_stream(self.stream)
        self.stream.synchronize()
        n = self.shape[0] * self.shape[1]
        root = self.buf.numpy()
        weakref.finalize(root, _release_pinned, self.buf)
        out = root[: n * 8].view(np.float64).reshape(self.shape)
        del root
        out.setflags(write=False)
        return out


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("itr_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device(x, dev):
    """host numpy / tensor -> CUDA f32 tensor.  Pinned memory copies asynchronously in one DMA; large PAGEABLE arrays
    (the reference de-duplicates the images with a list comprehension, evaluation.py:289, so what reaches cal_sims
    is a fresh pageable copy) go through ops.upload_pageable: chunks staged into two pinned buffers by a
    multi-threaded host copy that overlaps the DMA of the previous chunk."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if x.dtype != torch.float32:
        x = x.float()
    if not x.is_cuda and not x.is_pinned() and x.numel() * 4 >= ops.STAGED_UPLOAD_MIN_BYTES:
        return ops.upload_pageable(x.contiguous(), dev)
    return x.to(dev, non_blocking=True)


def _effective_lengths(lengths, n_cap, shard_size, compat_unsliced_lengths):
    if lengths is None:
        return None
    ln = ops.lengths_to_numpy(lengths, n_cap) if not compat_unsliced_lengths else np.asarray(lengths, dtype=np.int32)
    if compat_unsliced_lengths:
        ln = ln[np.arange(n_cap) % shard_size]     # what evaluation.py:149 + Objectives.py:340 amount to
    return ln


def _scan_t2i_from(pi, caps, ln, norm, config, dev, on_block=None):
    """Fused t2i scores from prepared images and captions that are either on the device or in pinned host memory
    (gathered in place over PCIe, pipelined against the score kernel).  on_block(matrix, c0, c1) is called after the
    launch that completes columns [c0, c1) (cal_sims ships them to the host meanwhile)."""
    args = (norm, config["agg_func"], config["lambda_softmax"], config.get("lambda_lse", 6.0))
    if not caps.is_cuda:
        return ops.scan_t2i_scores_from_host(pi, caps, ln, *args, device=dev, on_block=on_block)
    out = ops.scan_t2i_scores_bf16(pi, ops.prepare_captions(caps, ln, device=dev), *args)
    if on_block is not None:
        on_block(out, 0, out.size(1))
    return out


def _sim_function(model):
    config = model.config
    if config["name"] in ["CAMERA"]:
        return model.mvm
    return model.sim_enc if getattr(model, "sim_enc", None) is not None else model.criterion.sim


def _tc_t2i_inputs(model, img_embs, cap_embs, ln, image_group, dev):
    """Inputs of the fused tcgen05 SCAN t2i kernel, or None when this evaluation does not run on it.
    Returns (prepared images, captions tensor (CUDA or pinned host), raw_feature_norm)."""
    config = model.config
    cal_fun = _sim_function(model)
    norm = config.get("raw_feature_norm")
    # decided from rank-invariant inputs (config, image shape) ...
    tc_t2i = (cal_fun is objectives.xattn_score_t2i and objectives._precision(config) == "bf16"
              and norm in ("clipped_l2norm", "l2norm") and ln is not None
              and getattr(img_embs, "ndim", 0) == 3 and tuple(img_embs.shape[1:]) == (36, 1024)
              and getattr(cap_embs, "ndim", 0) == 3 and cap_embs.shape[2] == 1024)
    if not tc_t2i:
        return None
    # ... and from the caption lengths, which are rank-LOCAL when the captions are sharded: the sharded path runs
    # collectives (image all-gather), so every rank must take the same branch -- agree on one flag first (an empty
    # local block is fine: nothing to score, but the rank still joins the all-gather).
    lens_ok = ln.size == 0 or (int(np.min(ln)) >= 1 and int(np.max(ln)) <= ops.TC_MAX_WORDS)
    if image_group is not None:
        lens_ok = ops.all_ranks_agree(lens_ok, image_group, dev)
    if not lens_ok:
        return None
    caps = cap_embs if isinstance(cap_embs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cap_embs))
    if caps.dtype != torch.float32:
        caps = caps.float()
    if not caps.is_cuda and not caps.is_pinned():
        caps = _to_device(caps, dev)          # pageable host memory: staged copy of the padded array
    if image_group is not None:
        pi = ops.prepare_images_sharded(img_embs, image_group, dev)
    else:
        pi = ops.prepare_images_streamed(img_embs, dev)      # large host arrays: uploaded in chunks, scored as they land
    return pi, caps, norm


def device_sims(model, img_embs, cap_embs, lengths=None, shard_size=128, compat_unsliced_lengths=False,
                image_group=None, on_block=None):
    """The score matrix as a CUDA float32 tensor (n_img, n_cap); inputs host or device.
    image_group: torch.distributed group over which the image preparation is sharded and all-gathered
    (tensor-core SCAN path only; every rank passes the full image array and its own captions)."""
    dev = _device()
    config = model.config
    cal_fun = _sim_function(model)
    n_img, n_cap = len(img_embs), len(cap_embs)
    ln = _effective_lengths(lengths, n_cap, shard_size, compat_unsliced_lengths)

    fused = cal_fun in (objectives.cosine_sim, objectives.xattn_score_t2i, objectives.xattn_score_i2t)
    with torch.no_grad():
        if fused:
            tc = _tc_t2i_inputs(model, img_embs, cap_embs, ln, image_group, dev)
            if tc is not None:
                pi, caps, norm = tc
                if n_cap == 0:
                    return torch.empty(n_img, 0, device=dev, dtype=torch.float32)
                return _scan_t2i_from(pi, caps, ln, norm, config, dev, on_block)
            img = _to_device(img_embs, dev)
            if cal_fun is objectives.cosine_sim:
                return ops.cosine_scores(img, _to_device(cap_embs, dev))
            return cal_fun(img, _to_device(cap_embs, dev), ln, config)
        # anything else (learned similarity heads, CAMERA): the reference's blocked loop, with sliced lengths
        out = torch.empty(n_img, n_cap, device=dev, dtype=torch.float32)
        for i0 in range(0, n_img, shard_size):
            i1 = min(i0 + shard_size, n_img)
            img_block = _to_device(img_embs[i0:i1], dev)
            for c0 in range(0, n_cap, shard_size):
                c1 = min(c0 + shard_size, n_cap)
                cap_block = _to_device(cap_embs[c0:c1], dev)
                out[i0:i1, c0:c1] = cal_fun(img_block, cap_block, None if ln is None else ln[c0:c1], config)
        return out


def fused_ranks(model, img_embs, cap_embs, lengths=None, shard_size=128, compat_unsliced_lengths=False, image_group=None,
                start=0, n_cap_total=None, caps_per_img=CAPS_PER_IMG, return_block=False):
    """(i2t_ranks, i2t_top1, t2i_ranks, t2i_top1[, score block]) of this rank's captions against all images.

    SCAN t2i on the tensor-core path ranks INSIDE the score kernel (ground-truth pre-pass + counting epilogue,
    include/itr_b200.h "fused evaluation"): no (n_img, n_cap) matrix exists unless ``return_block``.  Captions that
    stream from pinned host memory in several chunks keep the pipelined matrix path (the thresholds of ALL captions
    would have to exist before the first chunk is counted, which would expose the whole PCIe gather), and so does
    every other similarity function."""
    from . import sharding
    dev = _device()
    n_img, n_cap = len(img_embs), len(cap_embs)
    n_cap_total = n_cap if n_cap_total is None else n_cap_total
    ln = _effective_lengths(lengths, n_cap, shard_size, compat_unsliced_lengths)
    config = model.config
    with torch.no_grad():
        tc = _tc_t2i_inputs(model, img_embs, cap_embs, ln, image_group, dev) if _sim_function(model) is objectives.xattn_score_t2i else None
        if tc is not None and n_cap > 0 and (tc[1].is_cuda or len(ops.host_caption_chunks(ln)) == 1):
            pi, caps, norm = tc
            pc = ops.prepare_captions(caps, ln, device=dev)
            block = torch.empty(n_img, n_cap, device=dev, dtype=torch.float32) if return_block else None
            stats = sharding.FusedScanStats(pi, pc, norm, config["agg_func"], config["lambda_softmax"],
                                            config.get("lambda_lse", 6.0), scores_out=block)
            out = sharding.sharded_ranks(stats.block(), start, n_cap_total, image_group, caps_per_img, stats)
            return out + ((block,) if return_block else ())
        if tc is not None:
            pi, caps, norm = tc
            block = _scan_t2i_from(pi, caps, ln, norm, config, dev) if n_cap else torch.empty(n_img, 0, device=dev)
        else:
            block = device_sims(model, img_embs, cap_embs, lengths, shard_size, compat_unsliced_lengths, image_group)
    out = sharding.sharded_ranks(block, start, n_cap_total, image_group, caps_per_img)
    return out + ((block,) if return_block else ())


class _ValLogger(dict):
    """Stand-in for the reference's LogCollector (evaluation.py:43-72) when that class is not importable."""

    def update(self, k, v, n=0):
        self[k] = v

    def tb_log(self, *a, **k):
        pass


def _host_zeros(shape, pin):
    """Zeroed float32 host tensor, page-locked if asked for and if the host grants it."""
    if pin:
        try:
            return torch.zeros(shape, dtype=torch.float32, pin_memory=True)
        except RuntimeError:
            pass
    return torch.zeros(shape, dtype=torch.float32)


def encode_data(model, data_loader, islength=False):
    """evaluation.py:75-121 with the device->host hand-off changed: embeddings are
    collected in PINNED host buffers and returned as numpy views of them, so every caller keeps working
    (`img_embs[::5]`, `numpy.array([...])`, slicing into folds) while `cal_sims` can gather just the real words of
    each caption straight from that memory over PCIe instead of re-uploading the zero-padded array block by block.
    Captions are always laid out at the dataset's maximum length (fixes the reference's first-batch sizing, defect D8)."""
    try:
        from itr.metricmodule.evaluation import LogCollector as _LC     # the reference's own collector, if importable
        val_logger = _LC()
    except Exception:
        val_logger = _ValLogger()
    model.val_start()
    # page-locked result buffers let cal_sims gather / stream them over PCIe; ITR_B200_PIN_EMBEDDINGS=0 keeps them pageable
    # (as the reference's numpy arrays are), and an allocation the host refuses falls back to pageable memory
    pin = torch.cuda.is_available() and os.environ.get("ITR_B200_PIN_EMBEDDINGS", "1") != "0"
    n = len(data_loader.dataset)
    max_n_word = 0
    if islength:
        for batch in data_loader:
            max_n_word = max(max_n_word, int(batch[4][0]))
    img_embs = cap_embs = cap_lens = None
    for batch_data in data_loader:
        model.logger = val_logger
        images, boxes, imgs_wh, captions, lengths, ids, captions_mask, captions_type_ids = batch_data
        with torch.no_grad():
            emd = model.forward_emb(images=images, boxes=boxes, imgs_wh=imgs_wh, captions=captions, lengths=lengths,
                                    ids=ids, captions_mask=captions_mask, captions_type_ids=captions_type_ids)
        img_emb, cap_emb = emd[0].detach().float(), emd[1].detach().float()
        if img_embs is None:
            cap_size = [n] + list(cap_emb.shape[1:])
            if islength:
                cap_size[1] = max_n_word
            img_embs = _host_zeros([n] + list(img_emb.shape[1:]), pin)
            cap_embs = _host_zeros(cap_size, pin)
            cap_lens = np.zeros(n, dtype=np.int32)
        if cap_emb.dim() == 3 and cap_emb.size(1) > cap_embs.size(1):          # a later, longer batch (defect D8)
            grown = _host_zeros([n, cap_emb.size(1)] + list(cap_embs.shape[2:]), pin)
            grown[:, : cap_embs.size(1)] = cap_embs
            cap_embs = grown
        idx = torch.as_tensor(np.asarray(ids), dtype=torch.long)
        img_embs[idx] = img_emb.cpu()
        if cap_emb.dim() == 3:
            cap_embs[idx, : cap_emb.size(1)] = cap_emb.cpu()
        else:
            cap_embs[idx] = cap_emb.cpu()
        cap_lens[np.asarray(ids)] = np.asarray(lengths)
    return img_embs.numpy(), cap_embs.numpy(), cap_lens


def cal_sims(model, img_embs, cap_embs, lengths=None, shard_size=128, compat_unsliced_lengths=False):
    """evaluation.py:124-153: host numpy in, host float64 (n_img, n_cap) out.

    The returned array is a genuine float64 ndarray (filled by one DMA from the device, read-only) and the device
    matrix it came from is remembered behind it: ``cal_recall`` / ``i2t`` / ``t2i`` called with that same array
    (evaluation.py:290-291, utils.py:158-167) rank from the device copy, both directions in one pass, instead of
    uploading the matrix again.  Anything derived from it (a copy, an average of two matrices) is ranked from the
    host values as before."""
    t0 = time.time()
    ship = _HostMatrixWriter(len(img_embs), len(cap_embs))
    d = device_sims(model, img_embs, cap_embs, lengths, shard_size, compat_unsliced_lengths, on_block=ship.block)
    out = ship.finish(d)
    if out is None:
        out = _to_host_f64(d)
    if out.size:
        _remember(out, d)
    print("Calculate similarity matrix elapses: {:.3f}s".format(time.time() - t0))
    return out


# ----------------------------------------------------------------------------------------- ranking
def _metrics(ranks):
    n = len(ranks)
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / n
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / n
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / n
    medr = np.floor(np.median(ranks)) + 1
    meanr = ranks.mean() + 1
    return (r1, r5, r10, medr, meanr)


def device_ranks(sims_dev, caps_per_img=CAPS_PER_IMG):
    """(i2t_ranks, i2t_top1, t2i_ranks, t2i_top1) as int64 CUDA tensors from a CUDA matrix (f32 or f64)."""
    if sims_dev.dtype == torch.float64:
        rr, rc, tr, tc = ops.rank_f64(sims_dev.contiguous(), caps_per_img)
        return rr.long(), tr.long(), rc.long(), tc.long()
    s = sims_dev if (sims_dev.dtype == torch.float32 and sims_dev.stride(1) == 1) else sims_dev.float().contiguous()
    thr_row, thr_col = ops.rank_thresholds(s, 0, caps_per_img)
    cnt_row, cnt_col, best_row, best_col = ops.rank_count(s, thr_row, thr_col, 0)
    return cnt_row.long(), ops.unpack_best_index(best_row), cnt_col.long(), ops.unpack_best_index(best_col)


def _rank_host(sims):
    twin = _device_twin(sims)
    if twin is not None:
        if twin.ranks is None:
            twin.ranks = [x.cpu().numpy().astype(np.float64) for x in device_ranks(twin.dev)]
        return [r.copy() for r in twin.ranks]
    dev = _device()
    if isinstance(sims, torch.Tensor):
        s = sims.to(dev)
    else:
        s = torch.from_numpy(np.ascontiguousarray(sims)).to(dev)
    if s.dtype not in (torch.float32, torch.float64):
        s = s.double()
    return [x.cpu().numpy().astype(np.float64) for x in device_ranks(s)]


def i2t(sims, return_ranks=False):
    """evaluation.py:156-189.  sims (N, 5N) -> (r1, r5, r10, medr, meanr)[, (ranks, top1)]."""
    ranks, top1, _, _ = _rank_host(sims)
    m = _metrics(ranks)
    return (m, (ranks, top1)) if return_ranks else m


def t2i(sims, return_ranks=False):
    """evaluation.py:192-222."""
    _, _, ranks, top1 = _rank_host(sims)
    m = _metrics(ranks)
    return (m, (ranks, top1)) if return_ranks else m


def _recall_dict(r, rt, ri, rti, verbose=True):
    ar = (r[0] + r[1] + r[2]) / 3
    ari = (ri[0] + ri[1] + ri[2]) / 3
    rsum = r[0] + r[1] + r[2] + ri[0] + ri[1] + ri[2]
    if verbose:
        print("rsum: %.1f" % rsum)
        print("Average i2t Recall: %.1f" % ar)
        print("Image to text: r1 %.1f; r5 %.1f; r10 %.1f; medr %.1f; meanr %.1f" % r)
        print("Average t2i Recall: %.1f" % ari)
        print("Text to image: r1 %.1f; r5 %.1f; r10 %.1f; medr %.1f; meanr %.1f" % ri)
    return {
        "result": [list(r) + list(ri) + [ar, ari, rsum]], "rsum": rsum,
        "i2t_ave_r": ar, "i2t_r1": r[0], "i2t_r5": r[1], "i2t_r10": r[2], "i2t_medr": r[3], "i2t_meanr": r[4],
        "i2t_ranks": rt[0], "i2t_top1": rt[1],
        "t2i_ave_r": ari, "t2i_r1": ri[0], "t2i_r5": ri[1], "t2i_r10": ri[2], "t2i_medr": ri[3], "t2i_meanr": ri[4],
        "t2i_ranks": rti[0], "t2i_top1": rti[1],
    }


def cal_recall(sims, verbose=True):
    """evaluation.py:225-259: same dict keys; one upload, both directions ranked in one pass."""
    a, b, c, d = _rank_host(sims)
    return _recall_dict(_metrics(a), (a, b), _metrics(c), (c, d), verbose)


def cal_sims_and_recall(model, img_embs, cap_embs, lengths=None, shard_size=128, return_sims=False, verbose=False,
                        compat_unsliced_lengths=False):
    """Fused evaluation: scores and both rankings stay on the device; only the (N,) / (5N,) rank and
    top-1 vectors come back (plus the matrix when ``return_sims``)."""
    out = fused_ranks(model, img_embs, cap_embs, lengths, shard_size, compat_unsliced_lengths, return_block=return_sims)
    a, b, c, d = [x.cpu().numpy().astype(np.float64) for x in out[:4]]
    res = _recall_dict(_metrics(a), (a, b), _metrics(c), (c, d), verbose)
    if return_sims:
        res["sims"] = out[4]
    return res


def cal_sims_and_recall_ensemble(models, img_embs_list, cap_embs_list, lengths_list=None, shard_size=128,
                                 return_sims=False, verbose=False):
    """Two-model (or n-model) ensemble of ``evalrank_ensemble`` (evaluation.py:378-401): the models' score matrices are
    averaged and ranked on the device.  The reference averages float64 matrices holding float32 values
    (``(sims + sims_2) / 2``), which is exact; so is the float64 accumulation here (a float32 average could round and
    create ties or rank flips the reference does not have).  Nothing but the rank vectors comes back."""
    if not (len(models) == len(img_embs_list) == len(cap_embs_list)) or not models:
        raise ValueError("models, img_embs_list and cap_embs_list must be equally long and non-empty")
    lengths_list = lengths_list if lengths_list is not None else [None] * len(models)
    sims = None
    for model, img, cap, ln in zip(models, img_embs_list, cap_embs_list, lengths_list):
        s = device_sims(model, img, cap, ln, shard_size)
        if len(models) > 1:
            s = s.double()
        sims = s if sims is None else sims.add_(s)
    if len(models) > 1:
        sims = sims.div_(float(len(models)))
    a, b, c, d = [x.cpu().numpy().astype(np.float64) for x in device_ranks(sims)]
    res = _recall_dict(_metrics(a), (a, b), _metrics(c), (c, d), verbose)
    if return_sims:
        res["sims"] = sims
    return res

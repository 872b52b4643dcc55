"""Caption-sharded evaluation across the GPUs of one box.

Every (image, caption) pair is independent, so the score matrix shards by caption blocks with
no data-path collective: rank r owns a contiguous block of captions that starts on a multiple
of ``caps_per_img`` (an image's ground-truth captions stay together), every rank holds all
images.  Ranking needs one small exchange:

  t2i  fully local -- a rank owns whole columns, so the rank and top-1 of its captions are exact;
       the per-caption vectors are concatenated.
  i2t  the threshold of image i is its best ground-truth score, which lives on exactly one rank:
       all-reduce(MAX) of a (n_img,) float vector; then every rank counts the local captions that
       beat it; counts are summed and the packed (score, caption) arg-max keys maxed.
  Everything after the threshold all-reduce travels in ONE all-gather of a packed int64 vector per rank.

All payloads are KBs, so the exchange is latency-bound; counts are integers, so the result is
bit-identical for every world size.  ``torch.distributed`` (NCCL over NVLink on the box, gloo in
the CPU tests) is plumbing only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import ops

_SIGN = torch.iinfo(torch.int64).min


def shard_bounds(n_cap: int, world_size: int, caps_per_img: int = 5):
    """[(start, end)] caption ranges, contiguous, starting on multiples of caps_per_img, sizes
    differing by at most one image's worth."""
    groups = (n_cap + caps_per_img - 1) // caps_per_img
    base, extra = divmod(groups, world_size)
    out, g = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((min(g * caps_per_img, n_cap), min((g + n) * caps_per_img, n_cap)))
        g += n
    return out


class CudaStats:
    """Local block statistics from the CUDA rank kernels."""

    @staticmethod
    def thresholds(block, cap_offset, caps_per_img):
        return ops.rank_thresholds(block, cap_offset, caps_per_img)

    @staticmethod
    def count(block, thr_row, thr_col, cap_offset):
        cnt_row, cnt_col, best_row, best_col = ops.rank_count(block, thr_row, thr_col, cap_offset)
        return cnt_row, cnt_col, best_row ^ _SIGN, ops.unpack_best_index(best_col)


class _BlockShape:
    """What sharded_ranks needs to know about a score block that is never materialised."""

    def __init__(self, n_img, n_local, device):
        self.shape, self.device = (n_img, n_local), device


class FusedScanStats:
    """Block statistics straight from the fused SCAN t2i kernel (no score matrix): thresholds = the ground-truth
    pre-pass on < 1 % of the items, count = the full pass with the comparisons in its epilogue."""

    def __init__(self, pi, pc, raw_feature_norm, agg_func, lambda_softmax, lambda_lse, scores_out=None):
        self.pi, self.pc, self.args, self.out = pi, pc, (raw_feature_norm, agg_func, lambda_softmax, lambda_lse), scores_out

    def block(self):
        return _BlockShape(self.pi.n_img, self.pc.n_cap, self.pi.images_bf16.device)

    def thresholds(self, block, cap_offset, caps_per_img):
        pi, n_local = self.pi, self.pc.n_cap
        if pi.gathered is not None and pi.local_rows is not None:
            # multi-GPU: the ground-truth images of this rank's captions are (normally) the images it prepared itself,
            # so the pre-pass does not have to wait for the other ranks' images
            lo, hi = pi.local_rows
            g_lo, g_hi = cap_offset // caps_per_img, min(-(-(cap_offset + n_local) // caps_per_img), pi.n_img)
            if n_local > 0 and lo <= g_lo and g_hi <= hi and cap_offset - lo * caps_per_img >= 0:
                pi.wait_local()               # this rank's own rows may still be arriving in chunks from the host
                tr, tc = ops.scan_t2i_gt_thresholds(pi.rows(lo, hi), self.pc, *self.args,
                                                    cap_offset=cap_offset - lo * caps_per_img, caps_per_img=caps_per_img)
                thr_row = torch.full((pi.n_img,), float("-inf"), device=tr.device)
                thr_row[lo:hi] = tr
                return thr_row, tc
        return ops.scan_t2i_gt_thresholds(pi, self.pc, *self.args, cap_offset=cap_offset, caps_per_img=caps_per_img)

    def count(self, block, thr_row, thr_col, cap_offset):
        cnt_row, cnt_col, best_row, best_col = ops.scan_t2i_count(self.pi, self.pc, *self.args, thr_row, thr_col,
                                                                  cap_offset=cap_offset, out=self.out)
        return cnt_row, cnt_col, best_row ^ _SIGN, ops.unpack_best_index(best_col)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def _all_reduce(t, op, group):
    if _world(group) > 1:
        dist.all_reduce(t, op=op, group=group)
    return t


def sharded_ranks(block, start, n_cap_total, group=None, caps_per_img=5, stats=CudaStats):
    """block: this rank's (n_img, n_local) scores, first column = global caption ``start``.
    Returns (i2t_ranks, i2t_top1, t2i_ranks, t2i_top1) as int64 tensors, identical on every rank.

    Two collectives: all-reduce(MAX) of the i2t thresholds, then ONE all-gather of every rank's packed int64
    statistics [i2t counts | i2t arg-max keys | t2i ranks of its captions | t2i top-1 of its captions]; the merge
    (sum of counts, max of keys, concatenation of the caption slices) is local.  best_row keys handed over by
    ``stats.count`` are signed-comparable int64 (unsigned key ^ 2^63) whose low 32 bits hold ~(global caption index)."""
    n_img, n_local = block.shape
    dev = block.device
    world = _world(group)
    if n_local > 0:
        thr_row, thr_col = stats.thresholds(block, start, caps_per_img)
    else:
        thr_row = torch.full((n_img,), float("-inf"), device=dev)
        thr_col = torch.empty(0, device=dev)
    thr_row = _all_reduce(thr_row.contiguous(), dist.ReduceOp.MAX, group)
    bounds = shard_bounds(n_cap_total, world, caps_per_img) if world > 1 else [(start, start + n_local)]
    width = max(hi - lo for lo, hi in bounds)
    mine = torch.zeros(2 * n_img + 2 * width, dtype=torch.int64, device=dev)
    mine[n_img: 2 * n_img] = _SIGN
    if n_local > 0:
        cnt_row, cnt_col, best_row, best_col_idx = stats.count(block, thr_row, thr_col, start)
        mine[:n_img] = cnt_row
        mine[n_img: 2 * n_img] = best_row
        mine[2 * n_img: 2 * n_img + n_local] = cnt_col
        mine[2 * n_img + width: 2 * n_img + width + n_local] = best_col_idx
    if world > 1:
        if bounds[dist.get_rank(group)] != (start, start + n_local):
            raise ValueError("sharded_ranks: this rank's block ({}, {}) is not shard_bounds()[rank] = {}".format(
                start, start + n_local, bounds[dist.get_rank(group)]))
        every = torch.empty(world * mine.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, mine, group=group)
        every = every.view(world, -1)
    else:
        every = mine.view(1, -1)
    i2t_ranks = every[:, :n_img].sum(dim=0)
    best = every[:, n_img: 2 * n_img].max(dim=0).values
    i2t_top1 = (~best) & 0xFFFFFFFF
    t2i_ranks = torch.cat([every[r, 2 * n_img: 2 * n_img + (hi - lo)] for r, (lo, hi) in enumerate(bounds)])
    t2i_top1 = torch.cat([every[r, 2 * n_img + width: 2 * n_img + width + (hi - lo)] for r, (lo, hi) in enumerate(bounds)])
    return i2t_ranks, i2t_top1, t2i_ranks, t2i_top1


def sharded_scan_eval(images, captions_local, lengths_local, start, n_cap_total, config, group=None,
                      caps_per_img=5, return_block=False):
    """SCAN evaluation of this rank's caption block + the exchange.  images: all images (host numpy /
    pinned or CUDA tensor); captions_local / lengths_local: this rank's block.  Returns the
    reference's ``cal_recall`` dict (evaluation.py:225-259), identical on every rank."""
    from . import evaluation as ev

    class _M:                      # the two attributes device_sims reads from the model
        sim_enc = None

    m = _M()
    m.config = config
    from .objectives import ContrastiveLoss
    m.criterion = ContrastiveLoss(config, margin=config.get("margin", 0), measure=config.get("measure", "cosine"),
                                  max_violation=config.get("max_violation", False))
    grp = group
    if grp is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        grp = dist.group.WORLD
    out = ev.fused_ranks(m, images, captions_local, lengths_local, image_group=grp, start=start, n_cap_total=n_cap_total,
                         caps_per_img=caps_per_img, return_block=return_block)
    a, b, c, d = [x.cpu().numpy().astype(np.float64) for x in out[:4]]
    block = out[4] if return_block else None
    res = ev._recall_dict(ev._metrics(a), (a, b), ev._metrics(c), (c, d), verbose=False)
    if return_block:
        res["sims_block"] = block
    return res

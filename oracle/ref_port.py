"""float32 torch-CPU port of the reference hot path (TEST INFRASTRUCTURE).

Purpose: the timed CPU baseline (``cpu_baseline.kind == "port"``) and the
``bench.py --impl reference`` arm on the GPU box, where /root/reference does
not exist.  It keeps the reference's cost structure on purpose -- one Python
iteration per caption, the caption materialised once per image, two batched
matmuls, a D-wide cosine -- so its timing stands in for the reference's own
CPU path (checked against the real thing in tests/test_oracle_golden.py and,
in the authoring container, tests/test_oracle_vs_reference.py).

Reference lines restated: itr/modalmodule/Objectives.py:10-21, 329-476,
482-517; itr/modalmodule/utils.py:11-15; itr/metricmodule/evaluation.py:156-222.
"""
from __future__ import annotations

import numpy as np
import torch


def _unit_l2(x, dim, eps=1e-8):
    # utils.py:11-15 (eps outside the sqrt)
    return x / (x.pow(2).sum(dim=dim, keepdim=True).sqrt() + eps)


def _cos(a, b, dim, eps=1e-8):
    # Objectives.py:10-15
    num = (a * b).sum(dim)
    den = (a.norm(2, dim) * b.norm(2, dim)).clamp(min=eps)
    return num / den


def _attention(query, context, norm_mode, smooth):
    # Objectives.py:421-476
    attn = torch.bmm(context, query.transpose(1, 2))           # (B, sL, qL)
    if norm_mode == "softmax":
        attn = torch.softmax(attn, dim=2)
    elif norm_mode == "l2norm":
        attn = _unit_l2(attn, 2)
    elif norm_mode == "clipped_l2norm":
        attn = _unit_l2(torch.nn.functional.leaky_relu(attn, 0.1), 2)
    elif norm_mode == "clipped":
        attn = torch.nn.functional.leaky_relu(attn, 0.1)
    elif norm_mode != "no_norm":
        raise ValueError("unknown first norm type: {}".format(norm_mode))
    attn = attn.transpose(1, 2).contiguous()                   # (B, qL, sL)
    attn = torch.softmax(attn * smooth, dim=2)
    attn_t = attn.transpose(1, 2).contiguous()                 # (B, sL, qL)
    ctx = torch.bmm(context.transpose(1, 2), attn_t)           # (B, D, qL)
    return ctx.transpose(1, 2)


def _agg(row_sim, agg_func, lambda_lse):
    if agg_func == "LogSumExp":
        return torch.log(torch.exp(row_sim * lambda_lse).sum(dim=1, keepdim=True)) / lambda_lse
    if agg_func == "Max":
        return row_sim.max(dim=1, keepdim=True)[0]
    if agg_func == "Sum":
        return row_sim.sum(dim=1, keepdim=True)
    if agg_func == "Mean":
        return row_sim.mean(dim=1, keepdim=True)
    raise ValueError("unknown aggfunc: {}".format(agg_func))


@torch.no_grad()
def scan_scores(images, captions, cap_lens, cross_attn, raw_feature_norm="clipped_l2norm",
                agg_func="LogSumExp", lambda_softmax=9.0, lambda_lse=6.0):
    """Objectives.py:329-372 (t2i) / 376-417 (i2t).  torch CPU tensors in, (B, C) out."""
    n_img = images.size(0)
    cols = []
    for c in range(captions.size(0)):
        n = int(cap_lens[c])
        cap = captions[c, :n, :].unsqueeze(0).contiguous().repeat(n_img, 1, 1)
        if cross_attn == "t2i":
            ctx = _attention(cap, images, raw_feature_norm, lambda_softmax).contiguous()
            row = _cos(cap, ctx, 2)
        elif cross_attn == "i2t":
            ctx = _attention(images, cap, raw_feature_norm, lambda_softmax)
            row = _cos(images, ctx, 2)
        else:
            raise ValueError("unknown cross_attn: {}".format(cross_attn))
        cols.append(_agg(row, agg_func, lambda_lse))
    return torch.cat(cols, 1)


@torch.no_grad()
def cosine_scores(im, s):
    """Objectives.py:18-21."""
    return im.mm(s.t())


def hinge(scores, margin=0.0, max_violation=False):
    """Objectives.py:492-517 (TripletLoss; same maths as ContrastiveLoss :93-115)."""
    n = scores.size(0)
    diag = scores.diag().view(n, 1)
    cost_s = (margin + scores - diag.expand_as(scores)).clamp(min=0)
    cost_im = (margin + scores - diag.t().expand_as(scores)).clamp(min=0)
    eye = torch.eye(n, dtype=torch.bool)
    cost_s = cost_s.masked_fill(eye, 0)
    cost_im = cost_im.masked_fill(eye, 0)
    if max_violation:
        cost_s = cost_s.max(1)[0]
        cost_im = cost_im.max(0)[0]
    return cost_s.sum() + cost_im.sum()


def hinge_any(scores, margin=0.0, max_violation=False):
    """``hinge`` for scores on any device (the eye mask follows the scores)."""
    n = scores.size(0)
    diag = scores.diag().view(n, 1)
    eye = torch.eye(n, dtype=torch.bool, device=scores.device)
    cost_s = (margin + scores - diag.expand_as(scores)).clamp(min=0).masked_fill(eye, 0)
    cost_im = (margin + scores - diag.t().expand_as(scores)).clamp(min=0).masked_fill(eye, 0)
    if max_violation:
        cost_s = cost_s.max(1)[0]
        cost_im = cost_im.max(0)[0]
    return cost_s.sum() + cost_im.sum()


def i2t_ranks(sims, caps_per_img=5):
    """evaluation.py:156-189, one argsort per image row (numpy, single thread)."""
    n = sims.shape[0]
    ranks = np.zeros(n)
    top1 = np.zeros(n)
    for i in range(n):
        order = np.argsort(sims[i])[::-1]
        where = np.empty(order.size, dtype=np.int64)
        where[order] = np.arange(order.size)
        ranks[i] = where[caps_per_img * i: caps_per_img * (i + 1)].min()
        top1[i] = order[0]
    return ranks, top1


def t2i_ranks(sims, caps_per_img=5):
    """evaluation.py:192-222, one argsort per caption column."""
    st = sims.T
    m = st.shape[0]
    ranks = np.zeros(m)
    top1 = np.zeros(m)
    for c in range(m):
        order = np.argsort(st[c])[::-1]
        ranks[c] = np.where(order == c // caps_per_img)[0][0]
        top1[c] = order[0]
    return ranks, top1

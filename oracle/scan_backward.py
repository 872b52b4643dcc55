"""Gradient oracle for the SCAN cross-attention scores (TEST INFRASTRUCTURE -- only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this).

The reference trains SCAN by back-propagating ``ContrastiveLoss`` through ``xattn_score_t2i`` /
``xattn_score_i2t`` with torch autograd (itr/modalmodule/Models.py:219-222,
itr/modalmodule/Objectives.py:76-115, 329-476).  Two restatements live here:

``autograd_grads``   autograd through the float64 torch port of the forward (oracle/ref_port.py) --
                     the same computation graph the reference differentiates.
``coefficient_form`` the closed form the CUDA backward implements, in float64 numpy.  With
                     s = source index (context rows), q = query index and, per (image, caption) pair,
                         a[s][q]   raw affinity  c_s . q_q
                         xh[s][q]  raw_feature_norm over q        (Objectives.py:436-455)
                         al[q][s]  softmax over s of lambda * xh  (Objectives.py:459-462)
                         x_q = sum_s al[q][s] c_s,  r_q = cos(q_q, x_q)   (:469-474, :10-15)
                     the gradient of  sum_pairs dS * agg_q(r_q)  is
                         dQuery   = MQ   . Context + diag(t) Query
                         dContext = MQ^T . Query   + MC . Context
                     with MQ (Q x S), t (Q), MC (S x S) built from a, the context Gram and dS only --
                     nothing D-wide.  t2i: query = words, context = regions; i2t: the reverse.

Pinned by tests/golden/scan_grad.npz (gradients produced by the reference itself, see
oracle/make_golden.py) in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ref_port

EPS = 1e-8


def autograd_grads(images, captions, cap_lens, d_scores, cross_attn, raw_feature_norm="clipped_l2norm",
                   agg_func="LogSumExp", lambda_softmax=9.0, lambda_lse=6.0):
    """d(sum(scores * d_scores))/d(images, captions) by autograd through the float64 port."""
    im = torch.as_tensor(np.asarray(images), dtype=torch.float64).clone().requires_grad_(True)
    cap = torch.as_tensor(np.asarray(captions), dtype=torch.float64).clone().requires_grad_(True)
    with torch.enable_grad():
        scores = ref_port.scan_scores.__wrapped__(im, cap, cap_lens, cross_attn, raw_feature_norm, agg_func,
                                                  lambda_softmax, lambda_lse)
        (scores * torch.as_tensor(np.asarray(d_scores), dtype=torch.float64)).sum().backward()
    return scores.detach().numpy(), im.grad.numpy(), cap.grad.numpy()


def _leaky_grad(a):
    return np.where(a > 0, 1.0, 0.1)


def pair_coefficients(a, gram_ctx, q_norm, d_score, raw_feature_norm, agg_func, lambda_softmax, lambda_lse):
    """One (image, caption) pair.  a: (S, Q) raw affinities; gram_ctx: (S, S) context Gram;
    q_norm: (Q,) |query_q|.  Returns score, MQ (Q, S), t (Q,), MC (S, S)."""
    S, Q = a.shape
    # ---- forward ------------------------------------------------------------------------------
    if raw_feature_norm in ("clipped_l2norm", "clipped"):
        l = np.where(a > 0, a, 0.1 * a)
    else:
        l = a
    if raw_feature_norm in ("clipped_l2norm", "l2norm"):
        rs = np.sqrt((l * l).sum(axis=1, keepdims=True))             # per source row, over q
        nrm = rs + EPS
        xh = l / nrm
    elif raw_feature_norm == "softmax":
        e = np.exp(a - a.max(axis=1, keepdims=True))
        xh = e / e.sum(axis=1, keepdims=True)
    elif raw_feature_norm in ("clipped", "no_norm"):
        xh = l
    else:
        raise ValueError("unknown first norm type: {}".format(raw_feature_norm))
    z = lambda_softmax * xh                                           # (S, Q)
    e = np.exp(z - z.max(axis=0, keepdims=True))
    al = (e / e.sum(axis=0, keepdims=True)).T                          # (Q, S)
    P = (al * a.T).sum(axis=1)                                         # query . ctx
    ga = al @ gram_ctx                                                 # (Q, S)  (G alpha)
    qf = (ga * al).sum(axis=1)                                         # |ctx|^2
    w2 = np.sqrt(np.maximum(qf, 0.0))
    den = q_norm * w2
    clamped = den < EPS
    r = P / np.maximum(den, EPS)
    # ---- aggregation and its gradient ---------------------------------------------------------
    if agg_func == "LogSumExp":
        ex = np.exp(lambda_lse * r)
        score = np.log(ex.sum()) / lambda_lse
        g = ex / ex.sum()
    elif agg_func == "Mean":
        score, g = r.mean(), np.full(Q, 1.0 / Q)
    elif agg_func == "Sum":
        score, g = r.sum(), np.ones(Q)
    elif agg_func == "Max":
        score, g = r.max(), np.zeros(Q)
        g[int(np.argmax(r))] = 1.0
    else:
        raise ValueError("unknown aggfunc: {}".format(agg_func))
    g = g * d_score
    # ---- cosine backward: coefficients on query_q and ctx_q -----------------------------------
    safe_den = np.where(clamped, 1.0, den)
    p = np.where(clamped, g / EPS, g / safe_den)                       # d r / d(query . ctx)
    u = np.where(clamped, 0.0, -g * r / np.where(qf > 0, qf, 1.0))     # coefficient of ctx in d r / d ctx
    t = np.where(clamped, 0.0, -g * r / np.where(q_norm > 0, q_norm * q_norm, 1.0))
    d_al = p[:, None] * a.T + u[:, None] * ga                          # (Q, S)  = context_s . d ctx_q
    dot = (al * d_al).sum(axis=1, keepdims=True)
    d_xh = (lambda_softmax * al * (d_al - dot)).T                      # (S, Q)
    # ---- raw_feature_norm backward (rows over q) ----------------------------------------------
    if raw_feature_norm in ("clipped_l2norm", "l2norm"):
        proj = (d_xh * l).sum(axis=1, keepdims=True)
        d_l = d_xh / nrm - l * proj / (nrm * nrm * np.where(rs > 0, rs, 1.0))
    elif raw_feature_norm == "softmax":
        d_l = xh * (d_xh - (xh * d_xh).sum(axis=1, keepdims=True))
    else:
        d_l = d_xh
    d_a = d_l * _leaky_grad(a) if raw_feature_norm in ("clipped_l2norm", "clipped") else d_l
    mq = p[:, None] * al + d_a.T                                       # (Q, S)
    mc = (al * u[:, None]).T @ al                                      # (S, S)
    return score, mq, t, mc


def coefficient_form(images, captions, cap_lens, d_scores, cross_attn, raw_feature_norm="clipped_l2norm",
                     agg_func="LogSumExp", lambda_softmax=9.0, lambda_lse=6.0):
    """Scores and both gradients through the closed form, float64 numpy."""
    V = np.asarray(images, np.float64)
    W = np.asarray(captions, np.float64)
    dS = np.asarray(d_scores, np.float64)
    n_img, n_cap = V.shape[0], W.shape[0]
    scores = np.zeros((n_img, n_cap))
    dV, dW = np.zeros_like(V), np.zeros_like(W)
    for c in range(n_cap):
        n = int(cap_lens[c])
        w = W[c, :n]
        gw = w @ w.T
        for i in range(n_img):
            v = V[i]
            a_rw = v @ w.T                                             # (R, n)
            if cross_attn == "t2i":
                s, mq, t, mc = pair_coefficients(a_rw, v @ v.T, np.linalg.norm(w, axis=1), dS[i, c], raw_feature_norm,
                                                 agg_func, lambda_softmax, lambda_lse)
                dW[c, :n] += mq @ v + t[:, None] * w
                dV[i] += mq.T @ w + mc @ v
            elif cross_attn == "i2t":
                s, mq, t, mc = pair_coefficients(a_rw.T, gw, np.linalg.norm(v, axis=1), dS[i, c], raw_feature_norm,
                                                 agg_func, lambda_softmax, lambda_lse)
                dV[i] += mq @ w + t[:, None] * v
                dW[c, :n] += mq.T @ v + mc @ w
            else:
                raise ValueError("unknown cross_attn: {}".format(cross_attn))
            scores[i, c] = s
    return scores, dV, dW

"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

The reference has no tests or golden vectors of its own,
so these files ARE the pin: inputs (stored exactly, as bf16 bit patterns so the
same numbers can be fed to reduced-precision kernels) and the outputs of the
reference's own functions on them:
  Objectives.xattn_score_t2i / xattn_score_i2t   (itr/modalmodule/Objectives.py:329-417)
  Objectives.cosine_sim                           (:18-21)
  Objectives.TripletLoss (+ torch autograd)       (:482-517; CPU-safe twin of ContrastiveLoss, defect D2)
  evaluation.i2t / t2i                            (itr/metricmodule/evaluation.py:156-222)
  torch autograd through xattn_score_* (+ TripletLoss)   -> scan_grad.npz, the pin of the training backward

  Objectives.func_attention / cosine_similarity / pdist / pdist_cos   -> helpers.npz

    python oracle/make_golden.py --only scan_grad        regenerates that file alone (also: aux_sims, helpers)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

NORMS = ("clipped_l2norm", "l2norm", "softmax", "clipped", "no_norm")
AGGS = ("LogSumExp", "Mean", "Max", "Sum")


def bf16_bits(x: torch.Tensor) -> np.ndarray:
    return x.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def from_bits(b: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(b.view(np.int16).copy()).view(torch.bfloat16).to(torch.float32)


def scan_case(O, name, n_img, lens, seed, d=1024, r=36):
    g = torch.Generator().manual_seed(seed)
    lens = np.asarray(lens, dtype=np.int32)
    lmax = int(lens.max())
    bank = torch.nn.functional.normalize(torch.randn(64, d, generator=g), dim=-1)
    concept = torch.randint(0, 64, (n_img, r), generator=g)
    images = torch.nn.functional.normalize(bank[concept] + 0.6 / d ** 0.5 * torch.randn(n_img, r, d, generator=g), dim=-1)
    n_cap = len(lens)
    caps = torch.zeros(n_cap, lmax, d)
    for c in range(n_cap):
        owner = c % n_img
        pick = torch.randint(0, r, (int(lens[c]),), generator=g)
        w = torch.nn.functional.normalize(bank[concept[owner, pick]] + 0.8 / d ** 0.5 * torch.randn(int(lens[c]), d, generator=g), dim=-1)
        caps[c, :lens[c]] = (0.5 + 1.5 * torch.rand(int(lens[c]), 1, generator=g)) * w
    img_bits, cap_bits = bf16_bits(images), bf16_bits(caps)
    images, caps = from_bits(img_bits), from_bits(cap_bits)       # what everybody consumes
    out = {"img_bits": img_bits, "cap_bits": cap_bits, "lens": lens}
    for direction, fn, lam_sm in (("t2i", O.xattn_score_t2i, 9.0), ("i2t", O.xattn_score_i2t, 4.0)):
        for norm in NORMS:
            for agg in AGGS:
                cfg = dict(raw_feature_norm=norm, agg_func=agg, lambda_lse=6.0, lambda_softmax=lam_sm)
                with torch.no_grad():
                    s32 = fn(images, caps, lens.tolist(), cfg).numpy()
                    s64 = fn(images.double(), caps.double(), lens.tolist(), cfg).numpy()
                out["{}|{}|{}|f32".format(direction, norm, agg)] = s32
                out["{}|{}|{}|f64".format(direction, norm, agg)] = s64
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if "|" not in k})


GRAD_FULL = (("clipped_l2norm", "LogSumExp"), ("l2norm", "Mean"), ("softmax", "Max"), ("clipped", "Sum"), ("no_norm", "LogSumExp"))


def scan_grad_case(O, name="scan_grad", n=6, lens=(9, 3, 14, 6, 11, 2), seed=105, d=128, r=36, n_probe=4):
    """Gradients of the reference's own SCAN scorers under torch autograd (float64).  Every mode combination is
    pinned through ``n_probe`` random projections of both gradients; five combinations per direction and the
    hinge-loss cases keep the full gradients (stored as float32)."""
    g = torch.Generator().manual_seed(seed)
    lens = np.asarray(lens, dtype=np.int32)
    lmax = int(lens.max())
    images = torch.nn.functional.normalize(torch.randn(n, r, d, generator=g), dim=-1)
    caps = torch.zeros(len(lens), lmax, d)
    for c in range(len(lens)):
        pick = torch.randint(0, r, (int(lens[c]),), generator=g)
        w = torch.nn.functional.normalize(images[c % n, pick] + 1.2 / d ** 0.5 * torch.randn(int(lens[c]), d, generator=g), dim=-1)
        caps[c, :lens[c]] = (0.5 + 1.5 * torch.rand(int(lens[c]), 1, generator=g)) * w
    img_bits, cap_bits = bf16_bits(images), bf16_bits(caps)
    images, caps = from_bits(img_bits).double(), from_bits(cap_bits).double()
    d_scores = torch.randn(n, len(lens), generator=g, dtype=torch.float64)
    probe_im = torch.randn(n_probe, *images.shape, generator=g, dtype=torch.float64)
    probe_cap = torch.randn(n_probe, *caps.shape, generator=g, dtype=torch.float64)
    out = {"img_bits": img_bits, "cap_bits": cap_bits, "lens": lens, "d_scores": d_scores.numpy(),
           "probe_im": probe_im.float().numpy(), "probe_cap": probe_cap.float().numpy()}
    probe_im, probe_cap = probe_im.float().double(), probe_cap.float().double()      # what the test will read back
    for direction, fn, lam_sm in (("t2i", O.xattn_score_t2i, 9.0), ("i2t", O.xattn_score_i2t, 4.0)):
        for norm in NORMS:
            for agg in AGGS:
                cfg = dict(raw_feature_norm=norm, agg_func=agg, lambda_lse=6.0, lambda_softmax=lam_sm)
                a, b = images.clone().requires_grad_(True), caps.clone().requires_grad_(True)
                (fn(a, b, lens.tolist(), cfg) * d_scores).sum().backward()
                key = "{}|{}|{}".format(direction, norm, agg)
                out[key + "|proj_im"] = (probe_im * a.grad).flatten(1).sum(1).numpy()
                out[key + "|proj_cap"] = (probe_cap * b.grad).flatten(1).sum(1).numpy()
                if (norm, agg) in GRAD_FULL:
                    out[key + "|d_im"] = a.grad.float().numpy()
                    out[key + "|d_cap"] = b.grad.float().numpy()
        for mv in (False, True):
            cfg = dict(raw_feature_norm="clipped_l2norm", agg_func="LogSumExp", lambda_lse=6.0, lambda_softmax=lam_sm)
            a, b = images.clone().requires_grad_(True), caps.clone().requires_grad_(True)
            loss = O.TripletLoss(margin=0.2, max_violation=mv)(fn(a, b, lens.tolist(), cfg))
            loss.backward()
            key = "{}|hinge|mv{}".format(direction, int(mv))
            out[key + "|loss"] = np.float64(loss.item())
            out[key + "|d_im"] = a.grad.float().numpy()
            out[key + "|d_cap"] = b.grad.float().numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, len(out), "arrays")


def aux_sims_case(O, name="aux_sims", seed=106):
    """order_sim (Objectives.py:24-30) and CAMERA MultiViewMatching (Fusionmodule.py:670-692): scores and the
    gradients torch autograd gives for sum(scores * d_scores), float64, from the reference's own code."""
    import importlib
    F = importlib.import_module("itr.modalmodule.Fusionmodule")
    g = torch.Generator().manual_seed(seed)
    out = {}
    n_img, n_cap, d = 19, 33, 96
    im = torch.randn(n_img, d, generator=g).abs()
    s = torch.randn(n_cap, d, generator=g).abs()
    s[3] = im[5] * 0.5                                          # a pair with distance exactly 0 (s <= im everywhere)
    im_bits, s_bits = bf16_bits(im), bf16_bits(s)
    im, s = from_bits(im_bits).double(), from_bits(s_bits).double()
    ds = torch.randn(n_img, n_cap, generator=g, dtype=torch.float64)
    a, b = im.clone().requires_grad_(True), s.clone().requires_grad_(True)
    sc = O.order_sim(a, b)
    mask = sc.detach() != 0                                     # autograd's sqrt'(0) is inf: pin the finite pairs only
    (sc * ds * mask).sum().backward()
    out.update({"order|im_bits": im_bits, "order|s_bits": s_bits, "order|d_scores": (ds * mask).numpy(), "order|scores": sc.detach().numpy(),
                "order|d_im": torch.nan_to_num(a.grad, nan=0.0).numpy(), "order|d_s": torch.nan_to_num(b.grad, nan=0.0).numpy()})
    mvm = F.MultiViewMatching()
    for tag, n_i, n_c in (("square", 12, 12), ("rect", 7, 20)):
        imgs = torch.nn.functional.normalize(torch.randn(n_i, 12, d, generator=g), dim=-1)
        caps = torch.nn.functional.normalize(torch.randn(n_c, d, generator=g), dim=-1)
        i_bits, c_bits = bf16_bits(imgs), bf16_bits(caps)
        imgs, caps = from_bits(i_bits).double(), from_bits(c_bits).double()
        ds = torch.randn(n_i, n_c, generator=g, dtype=torch.float64)
        a, b = imgs.clone().requires_grad_(True), caps.clone().requires_grad_(True)
        sc = mvm(a, b)
        (sc * ds).sum().backward()
        out.update({"mvm|%s|img_bits" % tag: i_bits, "mvm|%s|cap_bits" % tag: c_bits, "mvm|%s|d_scores" % tag: ds.numpy(),
                    "mvm|%s|scores" % tag: sc.detach().numpy(), "mvm|%s|d_imgs" % tag: a.grad.numpy(), "mvm|%s|d_caps" % tag: b.grad.numpy()})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, len(out), "arrays")


def helpers_case(O, name="helpers", seed=107):
    """The small exported helpers, from the reference's own code: func_attention (Objectives.py:421-476) for every
    raw_feature_norm in both roles, cosine_similarity (:10-15), SAEM's pdist / pdist_cos (:296-323)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    b, n_q, n_ctx, d = 3, 7, 36, 64
    query = torch.randn(b, n_q, d, generator=g)
    context = torch.nn.functional.normalize(torch.randn(b, n_ctx, d, generator=g), dim=-1)
    q_bits, c_bits = bf16_bits(query), bf16_bits(context)
    query, context = from_bits(q_bits).double(), from_bits(c_bits).double()
    out.update({"fa|query_bits": q_bits, "fa|context_bits": c_bits})
    for norm in NORMS:
        for smooth in (9.0, 4.0):
            with torch.no_grad():
                w, a = O.func_attention(query, context, dict(raw_feature_norm=norm), smooth=smooth)
            out["fa|{}|{}|weighted".format(norm, smooth)] = w.numpy()
            out["fa|{}|{}|attn".format(norm, smooth)] = a.numpy()
    x1 = torch.randn(5, 9, d, generator=g)
    x2 = torch.randn(5, 9, d, generator=g)
    x2[2, 4] = 0.0                                               # a zero row: the clamp(min=eps) branch
    x1_bits, x2_bits = bf16_bits(x1), bf16_bits(x2)
    x1, x2 = from_bits(x1_bits).double(), from_bits(x2_bits).double()
    out.update({"cs|x1_bits": x1_bits, "cs|x2_bits": x2_bits,
                "cs|dim2": O.cosine_similarity(x1, x2, dim=2).numpy(),
                "cs|dim1": O.cosine_similarity(x1, x2, dim=1).numpy()})
    a = torch.randn(17, 96, generator=g)
    bb = torch.randn(23, 96, generator=g)
    bb[5] = 0.0                                                  # pdist_cos zeroes the NaNs of 0/0
    a_bits, b_bits = bf16_bits(a), bf16_bits(bb)
    a, bb = from_bits(a_bits).double(), from_bits(b_bits).double()
    out.update({"pd|x1_bits": a_bits, "pd|x2_bits": b_bits, "pd|pdist": O.pdist(a, bb).numpy(),
                "pd|pdist_cos": O.pdist_cos(a, bb).numpy()})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, len(out), "arrays")


def main():
    O, E = ref_loader.load()
    torch.set_num_threads(8)
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    if "--only" in sys.argv:
        only = sys.argv[sys.argv.index("--only") + 1]
        if only == "scan_grad":
            scan_grad_case(O)
        elif only == "aux_sims":
            aux_sims_case(O)
        elif only == "helpers":
            helpers_case(O)
        else:
            raise SystemExit("--only supports: scan_grad, aux_sims, helpers")
        return

    scan_case(O, "scan_small", n_img=8, lens=[16, 3, 12, 9, 5, 14, 7, 11, 16, 4, 13, 8], seed=101)
    scan_case(O, "scan_long", n_img=5, lens=[72, 40, 33], seed=102)
    scan_grad_case(O)
    aux_sims_case(O)
    helpers_case(O)

    # cosine + hinge
    g = torch.Generator().manual_seed(103)
    im = torch.nn.functional.normalize(torch.randn(24, 1024, generator=g), dim=-1)
    s = torch.nn.functional.normalize(im + 0.9 / 32 * torch.randn(24, 1024, generator=g), dim=-1)
    im_bits, s_bits = bf16_bits(im), bf16_bits(s)
    im, s = from_bits(im_bits), from_bits(s_bits)
    out = {"im_bits": im_bits, "s_bits": s_bits,
           "cosine|f32": O.cosine_sim(im, s).numpy(), "cosine|f64": O.cosine_sim(im.double(), s.double()).numpy()}
    for margin in (0.0, 0.2):
        for mv in (False, True):
            sc = O.cosine_sim(im.double(), s.double()).clone().requires_grad_(True)
            loss = O.TripletLoss(margin=margin, max_violation=mv)(sc)
            loss.backward()
            key = "hinge|m{}|mv{}".format(margin, int(mv))
            out[key + "|loss"] = np.float64(loss.item())
            out[key + "|dscores"] = sc.grad.numpy()
            # gradient w.r.t. the embeddings through the cosine similarity
            a = im.double().clone().requires_grad_(True)
            b = s.double().clone().requires_grad_(True)
            O.TripletLoss(margin=margin, max_violation=mv)(O.cosine_sim(a, b)).backward()
            out[key + "|d_im"] = a.grad.numpy()
            out[key + "|d_s"] = b.grad.numpy()
    np.savez_compressed(os.path.join(gold, "vse_hinge.npz"), **out)
    print("vse_hinge", len(out))

    # ranking
    rng = np.random.default_rng(104)
    n = 40
    sims = rng.standard_normal((n, 5 * n))
    sims[np.arange(n).repeat(5), np.arange(5 * n)] += 1.5          # make the ground truth competitive
    (r, (ranks, top1)) = E.i2t(sims, return_ranks=True)
    (ri, (ranks_i, top1_i)) = E.t2i(sims, return_ranks=True)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        rd = E.cal_recall(sims)
    np.savez_compressed(os.path.join(gold, "ranking.npz"), sims=sims,
                        i2t_metrics=np.asarray(r, dtype=np.float64), i2t_ranks=ranks, i2t_top1=top1,
                        t2i_metrics=np.asarray(ri, dtype=np.float64), t2i_ranks=ranks_i, t2i_top1=top1_i,
                        result=np.asarray(rd["result"], dtype=np.float64), rsum=np.float64(rd["rsum"]))
    print("ranking", r, ri)


if __name__ == "__main__":
    main()

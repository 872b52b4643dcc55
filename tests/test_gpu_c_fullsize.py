"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle cannot
finish these sizes in test time): agreement of the two independent CUDA paths on samples, oracle on a
sub-block, run-to-run determinism, shard-count independence of every rank, and planted-structure recall."""
import numpy as np
import pytest
import torch

import itr_b200
from itr_b200 import evaluation as ev, objectives as ob, ops, sharding
from oracle import scan_oracle as so

pytestmark = pytest.mark.gpu


def cfg(**kw):
    base = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp",
                lambda_lse=6.0, lambda_softmax=9.0, margin=0.2, max_violation=True, measure="cosine")
    base.update(kw)
    return base


def virtual_shard_ranks(scores, world):
    """Ranks computed shard by shard and merged exactly like the multi-GPU exchange (sum of counts, max of
    thresholds / keys), on one device."""
    n_img, n_cap = scores.shape
    bounds = sharding.shard_bounds(n_cap, world)
    thr = torch.full((n_img,), float("-inf"), device=scores.device)
    for lo, hi in bounds:
        t, _ = ops.rank_thresholds(scores[:, lo:hi], lo, 5)
        thr = torch.maximum(thr, t)
    i2t = torch.zeros(n_img, dtype=torch.int64, device=scores.device)
    t2i = torch.zeros(n_cap, dtype=torch.int64, device=scores.device)
    for lo, hi in bounds:
        blk = scores[:, lo:hi]
        _, tc = ops.rank_thresholds(blk, lo, 5)
        cr, cc, _, _ = ops.rank_count(blk, thr, tc, lo)
        i2t += cr.long()
        t2i[lo:hi] = cc.long()
    return i2t, t2i


def test_config3_f30k_shape_full():
    """SCAN t2i LSE, Flickr30K-shaped 1000 x 5000 (72 707 words)."""
    sh = itr_b200.synth.F30K_SHAPE
    img, cap, lens = itr_b200.synth.scan_inputs(device="cuda", round_to="bf16", **sh)
    assert int(lens.sum()) == 72707
    a = ob.xattn_score_t2i(img, cap, lens, cfg())
    b = ob.xattn_score_t2i(img, cap, lens, cfg())
    assert torch.equal(a, b)                                       # deterministic
    assert torch.isfinite(a).all()
    f32 = ob.xattn_score_t2i(img, cap[:400], lens[:400], cfg(itr_b200_precision="fp32"))
    rel = ((a[:, :400] - f32).abs() / f32.abs().clamp_min(1e-6)).max().item()
    assert rel < 1e-3, rel
    want = so.scan_scores(img[:40].cpu().numpy(), cap[4000:4012].cpu().numpy(), lens[4000:4012], "t2i",
                          "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    np.testing.assert_allclose(a[:40, 4000:4012].cpu().numpy(), want, rtol=1e-3, atol=1e-6)
    i2t, _, t2i, _ = ev.device_ranks(a)
    for world in (2, 8):
        vi, vt = virtual_shard_ranks(a, world)
        assert torch.equal(vi, i2t) and torch.equal(vt, t2i)
    # the synthetic captions are planted on their image: recall must be high and ranks in range
    assert (t2i < 1).float().mean().item() > 0.9 and (i2t < 10).float().mean().item() > 0.95
    assert int(i2t.max()) < 5000 and int(t2i.max()) < 1000


def test_config5_coco5k_shape_full():
    """SCAN t2i LSE COCO-5K: 5000 x 25000 (312 906 words), plus the 8-way shard merge of the ranks."""
    sh = itr_b200.synth.COCO5K_SHAPE
    img, cap, lens = itr_b200.synth.scan_inputs(device="cuda", **sh)
    assert int(lens.sum()) == 312906 and cap.shape == (25000, 72, 1024)
    pi, pc = ops.prepare_images(img), ops.prepare_captions(cap, lens)
    assert pc.n_tiles % 1 == 0 and pc.sum_len == 312906
    a = ops.scan_t2i_scores_bf16(pi, pc, "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    assert torch.isfinite(a).all() and a.shape == (5000, 25000)
    # a column block recomputed on its own (different packing, different tile schedule) is bit-identical
    lo, hi = sharding.shard_bounds(25000, 8)[5]
    pc_blk = ops.prepare_captions(cap[lo:hi], lens[lo:hi])
    blk = ops.scan_t2i_scores_bf16(pi, pc_blk, "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    assert torch.equal(blk, a[:, lo:hi])
    # fp32 CUDA-core path on a sample of captions incl. the planted long ones (every 997th)
    cols = torch.tensor([0, 997, 1994, 12345, 24999], device="cuda")
    f32 = ops.scan_scores_f32(img[:256], cap[cols], lens[cols.cpu().numpy()], "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    rel = ((a[:256][:, cols] - f32).abs() / f32.abs().clamp_min(1e-6)).max().item()
    assert rel < 1.5e-3, rel          # inputs here are NOT pre-rounded: bf16 input quantisation included
    # the oracle (float64, the reference's algorithm) on a column block: first 32 images x captions incl. two planted
    # long ones (72 and 59 words), fed the bf16-rounded values the kernel consumes
    ocols = np.array([0, 997, 12345, 24999])
    rnd = lambda t: t.to(torch.bfloat16).float().cpu().numpy()
    want = so.scan_scores(rnd(img[:32]), rnd(cap[torch.from_numpy(ocols).cuda()]), lens[ocols], "t2i", "clipped_l2norm",
                          "LogSumExp", 9.0, 6.0)
    np.testing.assert_allclose(a[:32][:, torch.from_numpy(ocols).cuda()].cpu().numpy(), want, rtol=1e-3, atol=1e-6)
    assert lens[0] == 72 and lens[997] == 71
    i2t, _, t2i, _ = ev.device_ranks(a)
    vi, vt = virtual_shard_ranks(a, 8)
    assert torch.equal(vi, i2t) and torch.equal(vt, t2i)
    assert (t2i < 1).float().mean().item() > 0.9 and (i2t < 10).float().mean().item() > 0.9


def _ranks_match_up_to_ties(got, want_ranks, scores64, thr, axis, tol):
    """ranks equal except where another reference score sits within `tol` of the ground-truth score (north_star)."""
    got = np.asarray(got, dtype=np.int64)
    want_ranks = np.asarray(want_ranks, dtype=np.int64)
    bad = np.nonzero(got != want_ranks)[0]
    for q in bad:
        line = scores64[q] if axis == 1 else scores64[:, q]
        near = int(np.count_nonzero(np.abs(line - thr[q]) <= tol)) - 1
        assert abs(int(got[q]) - int(want_ranks[q])) <= near, (q, got[q], want_ranks[q], near)
    return len(bad)


def test_config1_vse_cosine_recall_full():
    """BASELINE config 1 at its stated size: VSE++ cosine scores + i2t/t2i Recall@K, 1000 x 5000 x 1024."""
    im, s = itr_b200.synth.vse_inputs(1000, 5000, 1, device="cuda")
    scores = ops.cosine_scores(im, s)
    want = so.cosine_scores(im.cpu().numpy(), s.cpu().numpy())                 # float64
    np.testing.assert_allclose(scores.cpu().numpy(), want, rtol=1e-5, atol=2e-6)
    # the rank kernels are exact on the matrix they are given ...
    i2t, top_i, t2i, top_c = [x.cpu().numpy() for x in ev.device_ranks(scores)]
    s32 = scores.double().cpu().numpy()
    ri, rt, tied_i, tied_c = so.strict_ranks(s32)
    np.testing.assert_array_equal(i2t, ri)
    np.testing.assert_array_equal(t2i, rt)
    np.testing.assert_array_equal(top_i[~tied_i], s32.argmax(axis=1)[~tied_i])
    np.testing.assert_array_equal(top_c[~tied_c], s32.argmax(axis=0)[~tied_c])
    # ... and agree with the reference's argsort ranking of the float64 matrix except at within-tolerance ties
    rd = so.recall_dict(want)
    gt = 5 * np.arange(1000)[:, None] + np.arange(5)[None, :]
    thr_i = np.take_along_axis(want, gt, axis=1).max(axis=1)
    thr_c = want[np.arange(5000) // 5, np.arange(5000)]
    n_bad = _ranks_match_up_to_ties(i2t, rd["i2t_ranks"], want, thr_i, 1, 4e-6)
    n_bad += _ranks_match_up_to_ties(t2i, rd["t2i_ranks"], want, thr_c, 0, 4e-6)
    assert n_bad <= 30, n_bad
    # through the drop-in entry points: host numpy in, the reference's dict out
    m = type("M", (), {"sim_enc": None})()
    m.config = cfg(name="VSE++")
    m.criterion = ob.ContrastiveLoss(m.config, margin=0.2, measure="cosine", max_violation=True)
    res = ev.cal_sims_and_recall(m, im.cpu().numpy(), s.cpu().numpy())
    np.testing.assert_array_equal(res["i2t_ranks"], i2t)
    np.testing.assert_array_equal(res["t2i_ranks"], t2i)
    assert abs(res["rsum"] - rd["rsum"]) <= 0.3


def test_config4_coco_5fold_i2t_mean_full():
    """BASELINE config 4 at its stated size: SCAN i2t Mean (lambda_softmax 4), 5 folds of 1000 images x 5000 captions
    with the COCO-shaped lengths, tensor-core mode against the float64 oracle on a block per fold (incl. the fold's
    longest caption) and against the float32 mode on every rank."""
    lens_all = itr_b200.synth.caption_lengths(25000, 10.5, 14)
    c4 = cfg(cross_attn="i2t", agg_func="Mean", lambda_softmax=4.0)
    rnd = lambda t: t.to(torch.bfloat16).float()
    for fold in range(5):
        lengths = lens_all[fold * 5000:(fold + 1) * 5000]
        img, cap, ln = itr_b200.synth.scan_inputs(1000, 5000, 10.5, 14 + fold, device="cuda", lengths=lengths, round_to="bf16")
        a = ob.xattn_score_i2t(img, cap, ln, c4)
        assert a.shape == (1000, 5000) and torch.isfinite(a).all()
        assert torch.equal(a, ob.xattn_score_i2t(img, cap, ln, c4))               # deterministic
        longest = int(np.argmax(ln))
        ocols = np.unique(np.concatenate([np.arange(15), [longest, 4999]]))
        assert ln[longest] >= 44
        oc = torch.from_numpy(ocols).cuda()
        want = so.scan_scores(img[:32].cpu().numpy(), cap[oc].cpu().numpy(), ln[ocols], "i2t", "clipped_l2norm", "Mean", 4.0, 6.0)
        np.testing.assert_allclose(a[:32][:, oc].cpu().numpy(), want, rtol=1e-3, atol=1e-6)
        if fold in (0, 3):
            # float32 CUDA-core mode on the whole fold: scores within the bf16 contract, recall identical up to near-ties
            f32 = ob.xattn_score_i2t(img, cap, ln, dict(c4, itr_b200_precision="fp32"))
            # Mean over 36 signed r_k cancels: a handful of the 5e6 scores are 100x below the typical |score| (0.05) and carry
            # the absolute error of the r_k (measured 2.8e-5; the 99.99th percentile of the relative error is 1.7e-4), so the
            # 1e-3 relative bound is taken with an absolute floor of 1e-3 x the typical score
            err = (a - f32).abs()
            assert bool((err <= 1e-3 * f32.abs() + 1e-3 * f32.abs().mean()).all()), (err.max().item(), f32.abs().mean().item())
            assert torch.quantile((err / f32.abs().clamp_min(1e-6)).flatten()[::5], 0.9999).item() < 5e-4
            ra, rb = ev.device_ranks(a), ev.device_ranks(f32)
            rsum = lambda r: sum(100.0 * (r[0] < k).float().mean().item() + 100.0 * (r[2] < k).float().mean().item() for k in (1, 5, 10))
            assert abs(rsum(ra) - rsum(rb)) <= 0.2, (rsum(ra), rsum(rb))
            assert (ra[0] != rb[0]).float().mean().item() < 0.01 and (ra[2] != rb[2]).float().mean().item() < 0.01
        i2t, _, t2i, _ = ev.device_ranks(a)
        assert int(i2t.max()) < 5000 and int(t2i.max()) < 1000
        assert (t2i < 10).float().mean().item() > 0.9


def test_config2_hinge_batch128_timing_sanity():
    im, s = itr_b200.synth.vse_inputs(128, 640, 2, device="cuda")
    s = s[::5].contiguous()
    a, b = im.clone().requires_grad_(True), s.clone().requires_grad_(True)
    crit = ob.ContrastiveLoss(cfg(name="VSE++"), margin=0.2, measure="cosine", max_violation=True)
    l1 = crit(a, b); l1.backward()
    g1 = a.grad.clone()
    a.grad = None; b.grad = None
    l2 = crit(a, b); l2.backward()
    assert l1.item() == l2.item() and torch.equal(g1, a.grad)     # deterministic loss and gradients
    want, ds = so.hinge_loss(so.cosine_scores(im.cpu().numpy(), s.cpu().numpy()), 0.2, True)
    np.testing.assert_allclose(l1.item(), want, rtol=1e-4)

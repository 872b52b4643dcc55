"""Host-side mirror of the reference interface: dispatch, error behaviour, install()."""
import types

import numpy as np
import pytest
import torch

import itr_b200
from itr_b200 import evaluation as ev, objectives as ob


def cfg(**kw):
    base = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp",
                lambda_lse=6.0, lambda_softmax=9.0, margin=0.2, max_violation=True, measure="cosine")
    base.update(kw)
    return base


def test_contrastive_loss_dispatch_matches_reference():
    # Objectives.py:45-74
    assert ob.ContrastiveLoss(cfg(name="VSE++"), measure="cosine").sim is ob.cosine_sim
    assert ob.ContrastiveLoss(cfg(name="VSE++"), measure="order").sim is ob.order_sim
    assert ob.ContrastiveLoss(cfg(), measure="cosine").sim is ob.xattn_score_t2i
    assert ob.ContrastiveLoss(cfg(cross_attn="i2t"), measure="cosine").sim is ob.xattn_score_i2t
    x = object()
    assert ob.ContrastiveLoss(cfg(name="SGRAF"), measure="cosine").sim(x, 1, 2, 3) is x
    with pytest.raises(ValueError):
        ob.ContrastiveLoss(cfg(), measure="euclid")
    with pytest.raises(ValueError):
        ob.ContrastiveLoss(cfg(cross_attn="both"), measure="cosine")
    c = ob.ContrastiveLoss(cfg(), margin=0.2, measure="cosine", max_violation=True)
    assert (c.margin, c.max_violation) == (0.2, True) and isinstance(c, torch.nn.Module)
    t = ob.TripletLoss(margin=0.1, max_violation=True)
    assert (t.margin, t.max_violation) == (0.1, True)


def test_no_cpu_fallback():
    im = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ob.cosine_sim(im, im)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ob.TripletLoss()(torch.randn(4, 4))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ev.i2t(np.zeros((2, 10)))


def test_scan_config_errors_raise_valueerror_before_any_launch():
    im, cap = torch.zeros(2, 36, 16), torch.zeros(2, 4, 16)
    with pytest.raises(ValueError, match="unknown aggfunc"):
        ob.xattn_score_t2i(im, cap, [4, 4], cfg(agg_func="Median"))
    with pytest.raises(ValueError):
        ob.xattn_score_i2t(im, cap, [4, 4], cfg(raw_feature_norm="l1norm"))
    with pytest.raises(ValueError):
        ob._precision(cfg(itr_b200_precision="fp8"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):          # the autograd path is native too
        ob.xattn_score_t2i(im.requires_grad_(), cap, [4, 4], cfg())


def test_effective_lengths_defect_d1():
    ln = np.array([5, 6, 7, 8, 9, 10, 11], dtype=np.int32)
    np.testing.assert_array_equal(ev._effective_lengths(ln, 7, 3, False), ln)
    # evaluation.py:149 passes the un-sliced array; Objectives.py:340 indexes it block-locally
    np.testing.assert_array_equal(ev._effective_lengths(ln, 7, 3, True), [5, 6, 7, 5, 6, 7, 5])


def test_metrics_formula():
    ranks = np.array([0, 0, 3, 7, 12, 100], dtype=np.float64)
    r1, r5, r10, medr, meanr = ev._metrics(ranks)
    assert (r1, r5, r10) == (100 * 2 / 6, 100 * 3 / 6, 100 * 4 / 6)
    assert medr == np.floor(np.median(ranks)) + 1 and meanr == ranks.mean() + 1


def test_install_and_uninstall_on_standin_modules():
    O = types.ModuleType("Objectives"); E = types.ModuleType("evaluation")
    O.cosine_sim = "orig_cos"; O.ContrastiveLoss = "orig_loss"; O.l1norm_d = "keep"; O.pdist = "orig_pdist"
    E.cal_sims = "orig_cal"; E.i2t = "orig_i2t"; E.encode_data = "keep"
    itr_b200.install(O, E)
    assert O.cosine_sim is ob.cosine_sim and O.ContrastiveLoss is ob.ContrastiveLoss and O.xattn_score_t2i is ob.xattn_score_t2i
    assert E.cal_sims is ev.cal_sims and E.i2t is ev.i2t and E.cal_recall is ev.cal_recall and E.encode_data is ev.encode_data
    assert O.l1norm_d == "keep" and O.pdist is ob.pdist and E._itr_b200_orig["encode_data"] == "keep"
    itr_b200.install(O, E)       # idempotent: originals are not overwritten by the patched ones
    assert O._itr_b200_orig["cosine_sim"] == "orig_cos"
    itr_b200.uninstall(O, E)
    assert O.cosine_sim == "orig_cos" and E.cal_sims == "orig_cal" and E.encode_data == "keep" and not hasattr(E, "_itr_b200_orig")


def test_synth_shapes_and_length_sums():
    assert itr_b200.synth.caption_lengths(5000, 12.4, 30).sum() == 72707
    ln = itr_b200.synth.caption_lengths(25000, 10.5, 14)
    assert ln.sum() == 312906 and ln.max() == 72 and ln.min() == 3
    img, cap, ln = itr_b200.synth.scan_inputs(6, 30, 10.5, 1, d=64, round_to="bf16")
    assert img.shape == (6, 36, 64) and cap.shape == (30, int(ln.max()), 64)
    assert torch.allclose(img.norm(dim=-1), torch.ones(6, 36), atol=2e-2)
    for c in range(30):
        assert (cap[c, ln[c]:] == 0).all() and (cap[c, : ln[c]].abs().sum(-1) > 0).all()
    assert torch.equal(img, img.to(torch.bfloat16).float())
    im, s = itr_b200.synth.vse_inputs(10, 50, 3, d=32, raw_dim=48)
    assert im.shape == (10, 32) and s.shape == (50, 32)

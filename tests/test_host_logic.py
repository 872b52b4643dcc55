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


def test_host_caption_chunks_cover_the_captions_in_order():
    """ops.host_caption_chunks: contiguous ranges on multiples of five that cover every caption once, whatever the
    fractions (pipelined host path, block-shipped cal_sims matrix); small inputs stay in one piece."""
    from itr_b200 import ops
    lens = itr_b200.synth.caption_lengths(4000, 10.5, 3)
    for fr in ((1.0 / 16, 3.0 / 16, 3.0 / 4), (3.0 / 16, 5.0 / 16, 1.0 / 2), (3.0 / 16, 5.0 / 16, 1.0 / 4, 1.0 / 8, 1.0 / 8)):
        chunks = ops.host_caption_chunks(lens, fractions=fr)
        assert len(chunks) == len(fr)
        assert chunks[0][0] == 0 and chunks[-1][1] == len(lens)
        assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:])) and all(c1 > c0 and c0 % 5 == 0 for c0, c1 in chunks)
        words = [int(lens[c0:c1].sum()) for c0, c1 in chunks]
        assert abs(words[0] / float(lens.sum()) - fr[0]) < 0.01
    assert ops.host_caption_chunks(lens[:100]) == [(0, 100)]
    assert ops.host_caption_chunks(lens[:0]) == [(0, 0)]


def test_prepared_images_row_ranges_order():
    """PreparedImages.row_ranges: uploaded chunks first (in upload order, each with its event), then the rows that wait for
    the other ranks' slices; once consumed, one range over everything."""
    import torch
    from itr_b200 import ops
    img = torch.zeros(40, 36, 1024, dtype=torch.bfloat16)
    gram = torch.zeros(40, 8, dtype=torch.uint8)
    pi = ops.PreparedImages(img, gram, 40)
    assert pi.row_ranges() == [(0, 40, False)]
    pi = ops.PreparedImages(img, gram, 40, local_rows=(8, 16), gathered="gather-event")
    assert pi.row_ranges() == [(8, 16, False), (0, 8, True), (16, 40, True)]
    pi = ops.PreparedImages(img, gram, 40, local_rows=(8, 16), gathered="gather-event", pending=[(8, 12, "e0"), (12, 16, "e1")])
    assert pi.row_ranges() == [(8, 12, "e0"), (12, 16, "e1"), (0, 8, True), (16, 40, True)]
    pi.ranges_consumed()
    assert pi.pending is None and pi.row_ranges() == [(8, 16, False), (0, 8, True), (16, 40, True)]
    pi = ops.PreparedImages(img, gram, 40, pending=[(0, 20, "e0"), (20, 40, "e1")])
    assert [r[:2] for r in pi.row_ranges()] == [(0, 20), (20, 40)]
    pi.ranges_consumed()
    assert pi.row_ranges() == [(0, 40, False)]

"""The oracle restatement against the committed outputs of the reference's own
functions (tests/golden, made by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import bits_to_f32, load_golden
from oracle import ref_port, scan_oracle as so

NORMS = so.RAW_FEATURE_NORMS
AGGS = so.AGG_FUNCS


@pytest.mark.parametrize("case", ["scan_small", "scan_long"])
@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
def test_scan_oracle_matches_reference(case, direction, lam_sm):
    g = load_golden(case)
    img, cap, lens = bits_to_f32(g["img_bits"]), bits_to_f32(g["cap_bits"]), g["lens"]
    for norm in NORMS:
        for agg in AGGS:
            got = so.scan_scores(img, cap, lens, direction, norm, agg, lam_sm, 6.0)
            ref64 = g["{}|{}|{}|f64".format(direction, norm, agg)]
            ref32 = g["{}|{}|{}|f32".format(direction, norm, agg)]
            # float64 restatement vs the reference run in float64: round-off only
            np.testing.assert_allclose(got, ref64, rtol=1e-11, atol=1e-13, err_msg=f"{norm}/{agg}")
            # and the reference's production float32 path stays within 1e-5 of it
            np.testing.assert_allclose(got, ref32, rtol=2e-5, atol=2e-6, err_msg=f"{norm}/{agg} f32")


@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
def test_ref_port_matches_reference_f32(direction, lam_sm):
    g = load_golden("scan_small")
    img = torch.from_numpy(bits_to_f32(g["img_bits"]))
    cap = torch.from_numpy(bits_to_f32(g["cap_bits"]))
    lens = g["lens"]
    for norm in NORMS:
        for agg in AGGS:
            got = ref_port.scan_scores(img, cap, lens, direction, norm, agg, lam_sm, 6.0).numpy()
            np.testing.assert_allclose(got, g["{}|{}|{}|f32".format(direction, norm, agg)], rtol=1e-5, atol=1e-6)


def test_unknown_modes_raise():
    img = np.zeros((2, 36, 8)); cap = np.ones((2, 4, 8)); lens = [4, 3]
    with pytest.raises(ValueError):
        so.scan_scores(img, cap, lens, "t2i", "l1norm")          # NameError upstream (defect D4) -> ValueError here
    with pytest.raises(ValueError):
        so.scan_scores(img, cap, lens, "t2i", "clipped_l2norm", "Median")
    with pytest.raises(ValueError):
        so.scan_scores(img, cap, lens, "sideways")


def test_cosine_and_hinge():
    g = load_golden("vse_hinge")
    im, s = bits_to_f32(g["im_bits"]), bits_to_f32(g["s_bits"])
    sc = so.cosine_scores(im, s)
    np.testing.assert_allclose(sc, g["cosine|f64"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(sc, g["cosine|f32"], rtol=1e-5, atol=1e-6)
    for margin in (0.0, 0.2):
        for mv in (False, True):
            key = "hinge|m{}|mv{}".format(margin, int(mv))
            loss, ds = so.hinge_loss(sc, margin, mv)
            np.testing.assert_allclose(loss, g[key + "|loss"], rtol=1e-12)
            np.testing.assert_array_equal(ds, g[key + "|dscores"])
            np.testing.assert_allclose(ds @ s.astype(np.float64), g[key + "|d_im"], rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(ds.T @ im.astype(np.float64), g[key + "|d_s"], rtol=1e-10, atol=1e-12)
            t = ref_port.hinge(torch.from_numpy(sc), margin, mv)
            np.testing.assert_allclose(float(t), g[key + "|loss"], rtol=1e-12)


def test_ranking():
    g = load_golden("ranking")
    sims = g["sims"]
    m, ranks, top1 = so.rank_i2t(sims)
    np.testing.assert_array_equal(ranks, g["i2t_ranks"]); np.testing.assert_array_equal(top1, g["i2t_top1"])
    np.testing.assert_allclose(m, g["i2t_metrics"])
    mi, ranks_i, top1_i = so.rank_t2i(sims)
    np.testing.assert_array_equal(ranks_i, g["t2i_ranks"]); np.testing.assert_array_equal(top1_i, g["t2i_top1"])
    np.testing.assert_allclose(mi, g["t2i_metrics"])
    rd = so.recall_dict(sims)
    np.testing.assert_allclose(rd["result"], g["result"]); np.testing.assert_allclose(rd["rsum"], g["rsum"])
    # tie-free definition the kernels implement == the reference's positions (continuous scores: no ties)
    a, b, ta, tb = so.strict_ranks(sims)
    assert not ta.any() and not tb.any()
    np.testing.assert_array_equal(a, ranks); np.testing.assert_array_equal(b, ranks_i)
    pr, pt = ref_port.i2t_ranks(sims)
    np.testing.assert_array_equal(pr, ranks); np.testing.assert_array_equal(pt, top1)
    pr, pt = ref_port.t2i_ranks(sims)
    np.testing.assert_array_equal(pr, ranks_i); np.testing.assert_array_equal(pt, top1_i)


# ----------------------------------------------------------------------------- training backward (row f3)
GRAD_FULL = (("clipped_l2norm", "LogSumExp"), ("l2norm", "Mean"), ("softmax", "Max"), ("clipped", "Sum"), ("no_norm", "LogSumExp"))


def _grad_inputs(g):
    return (bits_to_f32(g["img_bits"]).astype(np.float64), bits_to_f32(g["cap_bits"]).astype(np.float64), g["lens"],
            g["d_scores"], g["probe_im"].astype(np.float64), g["probe_cap"].astype(np.float64))


@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
def test_gradient_oracles_match_reference_autograd(direction, lam_sm):
    """Both gradient restatements (autograd through the port, and the closed form the CUDA kernels implement)
    against gradients the reference's own xattn_score_* produced under torch autograd in float64."""
    from oracle import scan_backward as sb
    g = load_golden("scan_grad")
    img, cap, lens, ds, p_im, p_cap = _grad_inputs(g)
    for norm in NORMS:
        for agg in AGGS:
            key = "{}|{}|{}".format(direction, norm, agg)
            for fn in (sb.autograd_grads, sb.coefficient_form):
                _, d_im, d_cap = fn(img, cap, lens, ds, direction, norm, agg, lam_sm, 6.0)
                scale_i, scale_c = np.abs(g[key + "|proj_im"]).max(), np.abs(g[key + "|proj_cap"]).max()
                np.testing.assert_allclose((p_im * d_im).reshape(len(p_im), -1).sum(1), g[key + "|proj_im"], rtol=1e-9,
                                           atol=1e-10 * scale_i, err_msg=key + " " + fn.__name__)
                np.testing.assert_allclose((p_cap * d_cap).reshape(len(p_cap), -1).sum(1), g[key + "|proj_cap"], rtol=1e-9,
                                           atol=1e-10 * scale_c, err_msg=key + " " + fn.__name__)
                if (norm, agg) in GRAD_FULL:      # stored as float32
                    np.testing.assert_allclose(d_im, g[key + "|d_im"], rtol=1e-6, atol=1e-7 * np.abs(d_im).max())
                    np.testing.assert_allclose(d_cap, g[key + "|d_cap"], rtol=1e-6, atol=1e-7 * np.abs(d_cap).max())


@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
@pytest.mark.parametrize("mv", [False, True])
def test_scan_hinge_gradient_oracle(direction, lam_sm, mv):
    """TripletLoss over the SCAN scores, back-propagated to the embeddings (Models.py:219-222)."""
    from oracle import scan_backward as sb
    g = load_golden("scan_grad")
    img, cap, lens = _grad_inputs(g)[:3]
    scores = so.scan_scores(img, cap, lens, direction, "clipped_l2norm", "LogSumExp", lam_sm, 6.0)
    loss, d_scores = so.hinge_loss(scores, 0.2, mv)
    key = "{}|hinge|mv{}".format(direction, int(mv))
    np.testing.assert_allclose(loss, g[key + "|loss"], rtol=1e-11)
    _, d_im, d_cap = sb.coefficient_form(img, cap, lens, d_scores, direction, "clipped_l2norm", "LogSumExp", lam_sm, 6.0)
    np.testing.assert_allclose(d_im, g[key + "|d_im"], rtol=1e-6, atol=1e-7 * np.abs(d_im).max())
    np.testing.assert_allclose(d_cap, g[key + "|d_cap"], rtol=1e-6, atol=1e-7 * np.abs(d_cap).max())


# ----------------------------------------------------------------------------- order_sim / MultiViewMatching (row f4)
def test_order_and_multiview_oracles_match_reference():
    g = load_golden("aux_sims")
    im, s = bits_to_f32(g["order|im_bits"]), bits_to_f32(g["order|s_bits"])
    np.testing.assert_allclose(so.order_scores(im, s), g["order|scores"], rtol=1e-12, atol=1e-14)
    assert (g["order|scores"] == 0).sum() == 1                     # the planted zero-distance pair
    d_im, d_s = so.order_grads(im, s, g["order|d_scores"])
    np.testing.assert_allclose(d_im, g["order|d_im"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(d_s, g["order|d_s"], rtol=1e-10, atol=1e-12)
    for tag in ("square", "rect"):
        imgs, caps = bits_to_f32(g["mvm|%s|img_bits" % tag]), bits_to_f32(g["mvm|%s|cap_bits" % tag])
        np.testing.assert_allclose(so.multiview_scores(imgs, caps), g["mvm|%s|scores" % tag], rtol=1e-12, atol=1e-14)
        d_imgs, d_caps = so.multiview_grads(imgs, caps, g["mvm|%s|d_scores" % tag])
        np.testing.assert_allclose(d_imgs, g["mvm|%s|d_imgs" % tag], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(d_caps, g["mvm|%s|d_caps" % tag], rtol=1e-10, atol=1e-12)

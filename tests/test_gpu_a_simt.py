"""GPU parity of the float32 kernels (1e-5 mode) against the oracle and the golden fixtures,
called through the Python drop-ins, i.e. through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import bits_to_f32, load_golden
import itr_b200
from itr_b200 import evaluation as ev, objectives as ob, ops
from oracle import scan_oracle as so

pytestmark = pytest.mark.gpu
RTOL32, ATOL32 = 1e-5, 2e-6      # "1e-5 in an fp32 mode" (BASELINE.json north_star)


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def cfg(**kw):
    base = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp",
                lambda_lse=6.0, lambda_softmax=9.0, margin=0.2, max_violation=True, measure="cosine",
                itr_b200_precision="fp32")
    base.update(kw)
    return base


class FakeModel:
    sim_enc = None

    def __init__(self, config):
        self.config = config
        self.criterion = ob.ContrastiveLoss(config, margin=config["margin"], measure=config["measure"],
                                            max_violation=config["max_violation"])


def test_library_sees_sm100():
    assert ops.capi.lib().itr_device_supported(0) == 1


def test_cosine_golden_and_odd_shapes():
    g = load_golden("vse_hinge")
    im, s = bits_to_f32(g["im_bits"]), bits_to_f32(g["s_bits"])
    got = ob.cosine_sim(dev(im), dev(s)).cpu().numpy()
    np.testing.assert_allclose(got, g["cosine|f64"], rtol=RTOL32, atol=ATOL32)
    rng = np.random.default_rng(0)
    for n_img, n_cap, d in [(1, 1, 1), (37, 53, 100), (130, 65, 1024), (64, 64, 17)]:
        a = rng.standard_normal((n_img, d)).astype(np.float32)
        b = rng.standard_normal((n_cap, d)).astype(np.float32)
        ref = so.cosine_scores(a, b)
        got = ob.cosine_sim(dev(a), dev(b)).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=RTOL32, atol=1e-5 * np.abs(ref).max())


@pytest.mark.parametrize("case", ["scan_small", "scan_long"])
@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
def test_scan_f32_golden_all_modes(case, direction, lam_sm):
    g = load_golden(case)
    img, cap, lens = dev(bits_to_f32(g["img_bits"])), dev(bits_to_f32(g["cap_bits"])), g["lens"]
    fn = ob.xattn_score_t2i if direction == "t2i" else ob.xattn_score_i2t
    for norm in so.RAW_FEATURE_NORMS:
        for agg in so.AGG_FUNCS:
            c = cfg(cross_attn=direction, raw_feature_norm=norm, agg_func=agg, lambda_softmax=lam_sm)
            got = fn(img, cap, lens.tolist(), c).cpu().numpy()
            want = g["{}|{}|{}|f64".format(direction, norm, agg)]
            np.testing.assert_allclose(got, want, rtol=RTOL32, atol=ATOL32, err_msg="{} {} {}".format(direction, norm, agg))


def test_scan_f32_vs_oracle_ragged_and_padding_invariance():
    img, cap, lens = itr_b200.synth.scan_inputs(13, 27, 10.5, 5)
    lens = lens.copy(); lens[:4] = [1, 2, 80 if cap.size(1) >= 80 else cap.size(1), 3]
    lens = np.minimum(lens, cap.size(1))
    for direction, lam in (("t2i", 9.0), ("i2t", 4.0)):
        want = so.scan_scores(img.numpy(), cap.numpy(), lens, direction, "clipped_l2norm", "LogSumExp", lam, 6.0)
        fn = ob.xattn_score_t2i if direction == "t2i" else ob.xattn_score_i2t
        c = cfg(cross_attn=direction, lambda_softmax=lam)
        got = fn(img.cuda(), cap.cuda(), lens, c).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=RTOL32, atol=ATOL32)
        # words beyond len must not matter (the reference slices them away, Objectives.py:341)
        dirty = cap.clone()
        for k, n in enumerate(lens):
            dirty[k, n:] = 7.0
        got2 = fn(img.cuda(), dirty.cuda(), torch.from_numpy(lens), c).cpu().numpy()
        np.testing.assert_array_equal(got, got2)


def test_scan_aggregation_ordering_property():
    img, cap, lens = itr_b200.synth.scan_inputs(20, 40, 10.5, 9)
    out = {}
    for agg in so.AGG_FUNCS:
        out[agg] = ob.xattn_score_t2i(img.cuda(), cap.cuda(), lens, cfg(agg_func=agg)).cpu().numpy()
    n = lens[None, :].astype(np.float64)
    assert (out["LogSumExp"] >= out["Max"] - 1e-6).all()           # lse_lambda >= max
    assert (out["Max"] >= out["Mean"] - 1e-6).all()
    np.testing.assert_allclose(out["Sum"], out["Mean"] * n, rtol=1e-5, atol=1e-6)
    assert (out["LogSumExp"] <= out["Max"] + np.log(n) / 6.0 + 1e-6).all()


def test_hinge_golden_and_autograd():
    g = load_golden("vse_hinge")
    im, s = dev(bits_to_f32(g["im_bits"])), dev(bits_to_f32(g["s_bits"]))
    for margin in (0.0, 0.2):
        for mv in (False, True):
            key = "hinge|m{}|mv{}".format(margin, int(mv))
            # TripletLoss on a ready matrix
            sc = dev(g["cosine|f64"].astype(np.float32)).requires_grad_(True)
            loss = ob.TripletLoss(margin=margin, max_violation=mv)(sc)
            loss.backward()
            np.testing.assert_allclose(loss.item(), g[key + "|loss"], rtol=2e-5)
            np.testing.assert_array_equal(sc.grad.cpu().numpy(), g[key + "|dscores"])
            # ContrastiveLoss (VSE++): fused scores + hinge + embedding gradients
            a, b = im.clone().requires_grad_(True), s.clone().requires_grad_(True)
            crit = ob.ContrastiveLoss(cfg(name="VSE++"), margin=margin, measure="cosine", max_violation=mv)
            l2 = crit(a, b)
            (2.0 * l2).backward()
            np.testing.assert_allclose(l2.item(), g[key + "|loss"], rtol=2e-5)
            np.testing.assert_allclose(a.grad.cpu().numpy(), 2.0 * g[key + "|d_im"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(b.grad.cpu().numpy(), 2.0 * g[key + "|d_s"], rtol=1e-5, atol=1e-6)
            # SGRAF calling style: the matrix arrives as `im`
            sc2 = sc.detach().clone().requires_grad_(True)
            l3 = ob.ContrastiveLoss(cfg(name="SGRAF"), margin=margin, measure="cosine", max_violation=mv)(sc2)
            l3.backward()
            np.testing.assert_array_equal(sc2.grad.cpu().numpy(), g[key + "|dscores"])


def test_hinge_batch128_vs_oracle_and_properties():
    im, s = itr_b200.synth.vse_inputs(128, 640, 2)
    s = s[::5].contiguous()
    sc = so.cosine_scores(im.numpy(), s.numpy())
    for mv in (True, False):
        want, dwant = so.hinge_loss(sc, 0.2, mv)
        a = im.cuda().requires_grad_(True); b = s.cuda().requires_grad_(True)
        loss = ob.ContrastiveLoss(cfg(name="VSE++"), margin=0.2, measure="cosine", max_violation=mv)(a, b)
        loss.backward()
        np.testing.assert_allclose(loss.item(), want, rtol=1e-4)
        np.testing.assert_allclose(a.grad.cpu().numpy(), dwant @ s.numpy().astype(np.float64), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(b.grad.cpu().numpy(), dwant.T @ im.numpy().astype(np.float64), rtol=1e-4, atol=1e-5)
    # margin satisfied everywhere -> zero loss, zero gradient
    big = torch.eye(16, device="cuda") * 5.0
    big.requires_grad_(True)
    l = ob.TripletLoss(margin=0.2, max_violation=True)(big)
    l.backward()
    assert l.item() == 0.0 and big.grad.abs().sum().item() == 0.0
    # single pair: no negatives
    one = torch.ones(1, 1, device="cuda", requires_grad=True)
    assert ob.TripletLoss(margin=0.2, max_violation=True)(one).item() == 0.0


def test_ranking_golden_f64_and_f32():
    g = load_golden("ranking")
    sims = g["sims"]
    (m, (ranks, top1)) = ev.i2t(sims, return_ranks=True)
    np.testing.assert_array_equal(ranks, g["i2t_ranks"]); np.testing.assert_array_equal(top1, g["i2t_top1"])
    np.testing.assert_allclose(m, g["i2t_metrics"])
    (mi, (ranks_i, top1_i)) = ev.t2i(sims, return_ranks=True)
    np.testing.assert_array_equal(ranks_i, g["t2i_ranks"]); np.testing.assert_array_equal(top1_i, g["t2i_top1"])
    np.testing.assert_allclose(mi, g["t2i_metrics"])
    rd = ev.cal_recall(sims, verbose=False)
    np.testing.assert_allclose(rd["result"], g["result"]); np.testing.assert_allclose(rd["rsum"], g["rsum"])
    assert sorted(rd) == sorted(so.recall_dict(sims))
    assert rd["i2t_ranks"].dtype == np.float64 and rd["t2i_top1"].dtype == np.float64
    assert ev.i2t(sims) == m                       # return_ranks=False form
    # float32 device path on a non-square-block shape, incl. exact ties (rank = strictly greater count)
    rng = np.random.default_rng(1)
    s32 = rng.standard_normal((301, 1505)).astype(np.float32)
    s32[5, 100] = s32[5, 25]                      # tie with a ground-truth score
    a, b, c, d = [x.cpu().numpy() for x in ev.device_ranks(dev(s32))]
    ia, ic, _, _ = so.strict_ranks(s32)
    np.testing.assert_array_equal(a, ia); np.testing.assert_array_equal(c, ic)
    np.testing.assert_array_equal(b, s32.argmax(1)); np.testing.assert_array_equal(d, s32.argmax(0))
    # and where nothing ties, it equals the reference's argsort positions
    ref_i, ref_c = so.rank_i2t(s32.astype(np.float64))[1], so.rank_t2i(s32.astype(np.float64))[1]
    ok_i = np.ones(301, bool); ok_i[5] = False
    np.testing.assert_array_equal(a[ok_i], ref_i[ok_i]); np.testing.assert_array_equal(c, ref_c)


def test_cal_sims_dropin_vse_and_scan_fp32():
    im, s = itr_b200.synth.vse_inputs(40, 200, 4)
    model = FakeModel(cfg(name="VSE++"))
    sims = ev.cal_sims(model, im.numpy(), s.numpy(), shard_size=64)
    assert sims.dtype == np.float64 and sims.shape == (40, 200)
    np.testing.assert_allclose(sims, so.cosine_scores(im.numpy(), s.numpy()), rtol=RTOL32, atol=ATOL32)
    res = ev.cal_sims_and_recall(model, im.numpy(), s.numpy())
    want = so.recall_dict(sims)
    for k in ("i2t_ranks", "t2i_ranks", "i2t_top1", "t2i_top1"):
        np.testing.assert_array_equal(res[k], want[k])
    np.testing.assert_allclose(res["result"], want["result"])
    # SCAN i2t Mean (config 4 style) through cal_sims, fp32 mode, plus the defect-D1 compat switch
    img, cap, lens = itr_b200.synth.scan_inputs(12, 60, 10.5, 11)
    c = cfg(cross_attn="i2t", agg_func="Mean", lambda_softmax=4.0)
    model = FakeModel(c)
    got = ev.cal_sims(model, img.numpy(), cap.numpy(), lens, shard_size=25)
    want = so.scan_scores(img.numpy(), cap.numpy(), lens, "i2t", "clipped_l2norm", "Mean", 4.0, 6.0)
    np.testing.assert_allclose(got, want, rtol=RTOL32, atol=ATOL32)
    compat = ev.cal_sims(model, img.numpy(), cap.numpy(), lens, shard_size=25, compat_unsliced_lengths=True)
    lens_d1 = np.minimum(lens[np.arange(60) % 25], cap.size(1))
    want_d1 = so.scan_scores(img.numpy(), cap.numpy(), lens_d1, "i2t", "clipped_l2norm", "Mean", 4.0, 6.0)
    np.testing.assert_allclose(compat, want_d1, rtol=RTOL32, atol=ATOL32)


def test_encode_data_pinned_handoff_and_validate_step_flow():
    """encode_data drop-in: pinned numpy views out, the reference's own validate_step post-processing
    (utils.py:152-167: list-comprehension dedupe, cal_sims, i2t, t2i) runs on them unchanged."""
    img, cap, lens = itr_b200.synth.scan_inputs(8, 40, 10.5, 3, round_to="bf16")
    img5 = img.repeat_interleave(5, dim=0)                       # the loader yields one image copy per caption
    order = np.argsort(-lens, kind="stable")                      # batches arrive sorted by length, ids scattered

    class DS:
        def __len__(self):
            return 40

    class Loader:
        dataset = DS()

        def __iter__(self):
            for s in range(0, 40, 16):
                ids = order[s:s + 16]
                l = lens[ids]
                yield (img5[ids].cuda(), None, None, cap[ids][:, : int(l.max())].cuda(), l.tolist(), ids.tolist(), None, None)

    class Model:
        sim_enc = None
        config = cfg(itr_b200_precision="bf16")
        criterion = ob.ContrastiveLoss(config, margin=0.2, measure="cosine", max_violation=True)

        def val_start(self):
            pass

        def forward_emb(self, images, captions, lengths, **kw):
            return images, captions, lengths

    m = Model()
    img_embs, cap_embs, cap_lens = ev.encode_data(m, Loader(), islength=True)
    assert isinstance(img_embs, np.ndarray) and img_embs.shape == (40, 36, 1024) and cap_embs.shape == (40, int(lens.max()), 1024)
    assert torch.from_numpy(cap_embs).is_pinned()
    np.testing.assert_array_equal(cap_lens, lens)
    np.testing.assert_array_equal(cap_embs, cap.numpy())
    img_embs = np.array([img_embs[i] for i in range(0, len(img_embs), 5)])      # utils.py:155
    sims = ev.cal_sims(m, img_embs, cap_embs, lengths=cap_lens, shard_size=100)
    want = so.scan_scores(img.numpy(), cap.numpy(), lens, "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    np.testing.assert_allclose(sims, want, rtol=1e-3, atol=1e-6)
    r = ev.i2t(sims); ri = ev.t2i(sims)
    assert len(r) == 5 and len(ri) == 5
    # first-batch sizing (islength=False) no longer breaks when a later batch is longer (defect D8)
    class Rev(Loader):
        def __iter__(self):
            return iter(list(Loader.__iter__(self))[::-1])
    _, cap2, _ = ev.encode_data(m, Rev(), islength=False)
    np.testing.assert_array_equal(cap2, cap.numpy())


# ----------------------------------------------------------------------------- order_sim / MultiViewMatching (row f4)
def test_order_sim_forward_backward_golden():
    g = load_golden("aux_sims")
    im = torch.from_numpy(bits_to_f32(g["order|im_bits"])).cuda().requires_grad_(True)
    s = torch.from_numpy(bits_to_f32(g["order|s_bits"])).cuda().requires_grad_(True)
    sc = ob.order_sim(im, s)
    np.testing.assert_allclose(sc.detach().cpu().numpy(), g["order|scores"], rtol=2e-6, atol=1e-6)
    (sc * torch.from_numpy(g["order|d_scores"]).float().cuda()).sum().backward()
    np.testing.assert_allclose(im.grad.cpu().numpy(), g["order|d_im"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(s.grad.cpu().numpy(), g["order|d_s"], rtol=1e-4, atol=2e-5)
    # ContrastiveLoss(measure="order") dispatches to it (Objectives.py:45-46)
    crit = ob.ContrastiveLoss(dict(name="VSE++"), margin=0.05, measure="order", max_violation=True)
    n = 19
    a, b = im.detach()[:n].clone().requires_grad_(True), s.detach()[:n].clone().requires_grad_(True)
    loss = crit(a, b)
    loss.backward()
    want_loss, d_sc = so.hinge_loss(so.order_scores(a.detach().cpu().numpy(), b.detach().cpu().numpy()), 0.05, True)
    np.testing.assert_allclose(loss.item(), want_loss, rtol=1e-5)
    want_a, want_b = so.order_grads(a.detach().cpu().numpy(), b.detach().cpu().numpy(), d_sc)
    np.testing.assert_allclose(a.grad.cpu().numpy(), want_a, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(b.grad.cpu().numpy(), want_b, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("tag", ["square", "rect"])
def test_multiview_matching_forward_backward_golden(tag):
    g = load_golden("aux_sims")
    imgs = torch.from_numpy(bits_to_f32(g["mvm|%s|img_bits" % tag])).cuda().requires_grad_(True)
    caps = torch.from_numpy(bits_to_f32(g["mvm|%s|cap_bits" % tag])).cuda().requires_grad_(True)
    mvm = itr_b200.MultiViewMatching()
    with torch.no_grad():
        np.testing.assert_allclose(mvm(imgs, caps).cpu().numpy(), g["mvm|%s|scores" % tag], rtol=2e-6, atol=1e-6)
        # chunked over images (small workspace): same numbers
        np.testing.assert_allclose(ops.multiview_scores(imgs, caps, max_workspace_bytes=4 * 12 * caps.size(0) * 2).cpu().numpy(),
                                   g["mvm|%s|scores" % tag], rtol=2e-6, atol=1e-6)
    sc = mvm(imgs, caps)
    (sc * torch.from_numpy(g["mvm|%s|d_scores" % tag]).float().cuda()).sum().backward()
    np.testing.assert_allclose(imgs.grad.cpu().numpy(), g["mvm|%s|d_imgs" % tag], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(caps.grad.cpu().numpy(), g["mvm|%s|d_caps" % tag], rtol=1e-4, atol=2e-6)
    # CAMERA's loss: TripletLoss over the multi-view scores (Models.py:628)
    if tag == "square":
        loss = ob.TripletLoss(margin=0.2, max_violation=True)(mvm(imgs.detach(), caps.detach()))
        want, _ = so.hinge_loss(g["mvm|square|scores"], 0.2, True)
        np.testing.assert_allclose(loss.item(), want, rtol=1e-5)


def test_empty_inputs_return_empty_matrices():
    """Zero images or zero captions: every scorer returns an empty matrix of the right shape without launching."""
    z_im, z_cap = torch.zeros(0, 36, 1024, device="cuda"), torch.zeros(0, 7, 1024, device="cuda")
    im, cap = torch.rand(3, 36, 1024, device="cuda"), torch.rand(2, 7, 1024, device="cuda")
    for c in (cfg(), cfg(itr_b200_precision="bf16"), cfg(cross_attn="i2t", itr_b200_precision="bf16")):
        fn = ob.xattn_score_t2i if c["cross_attn"] == "t2i" else ob.xattn_score_i2t
        assert tuple(fn(z_im, cap, [7, 3], c).shape) == (0, 2)
        assert tuple(fn(im, z_cap, [], c).shape) == (3, 0)
    assert tuple(ob.cosine_sim(torch.zeros(0, 64, device="cuda"), torch.rand(5, 64, device="cuda")).shape) == (0, 5)
    assert tuple(ob.cosine_sim(torch.rand(5, 64, device="cuda"), torch.zeros(0, 64, device="cuda")).shape) == (5, 0)
    assert tuple(ob.order_sim(torch.zeros(0, 64, device="cuda"), torch.rand(5, 64, device="cuda")).shape) == (0, 5)
    assert tuple(itr_b200.MultiViewMatching()(torch.zeros(0, 12, 64, device="cuda"), torch.rand(5, 64, device="cuda")).shape) == (0, 5)
    d_im, d_cap = ops.scan_backward_f32(z_im, cap, [7, 3], torch.zeros(0, 2, device="cuda"), "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    assert tuple(d_im.shape) == (0, 36, 1024) and not d_cap.any()


def test_ensemble_of_two_models_matches_averaged_matrices():
    """evalrank_ensemble (evaluation.py:378-401): (sims_1 + sims_2) / 2 then cal_recall -- here averaged and ranked on
    the device; one VSE++ model and one SCAN model on the same 40 x 200 split."""
    import types
    rng = np.random.default_rng(3)
    n, d = 40, 64
    im = rng.standard_normal((n, d)); im /= np.linalg.norm(im, axis=1, keepdims=True)
    s = np.repeat(im, 5, axis=0) + 0.7 * rng.standard_normal((5 * n, d)); s /= np.linalg.norm(s, axis=1, keepdims=True)
    img2, cap2, ln2 = itr_b200.synth.scan_inputs(n, 5 * n, 10.5, 9)
    m1 = types.SimpleNamespace(config=dict(name="VSE++"), sim_enc=None)
    m1.criterion = ob.ContrastiveLoss(m1.config, 0.2, "cosine", True)
    m2 = types.SimpleNamespace(config=cfg(), sim_enc=None)
    m2.criterion = ob.ContrastiveLoss(m2.config, 0.2, "cosine", True)
    res = ev.cal_sims_and_recall_ensemble([m1, m2], [im.astype(np.float32), img2.numpy()], [s.astype(np.float32), cap2.numpy()],
                                          [None, ln2], return_sims=True)
    s1 = ev.cal_sims(m1, im.astype(np.float32), s.astype(np.float32))
    s2 = ev.cal_sims(m2, img2.numpy(), cap2.numpy(), ln2)
    want = so.recall_dict((s1 + s2) / 2)                 # the reference's own average: float64 matrices of float32 values
    assert res["sims"].dtype == torch.float64
    np.testing.assert_array_equal(res["sims"].cpu().numpy(), (s1 + s2) / 2)
    assert res["rsum"] == pytest.approx(want["rsum"]) and res["i2t_r1"] == pytest.approx(want["i2t_r1"])
    np.testing.assert_array_equal(res["t2i_ranks"], want["t2i_ranks"])
    np.testing.assert_array_equal(res["i2t_ranks"], want["i2t_ranks"])


def test_saem_pdist_measures():
    """SAEM's pdist / pdist_cos (Objectives.py:296-323) over the native GEMM, with autograd, against plain torch."""
    g = torch.Generator().manual_seed(8)
    x1 = torch.randn(17, 96, generator=g).cuda().requires_grad_(True)
    x2 = torch.randn(23, 96, generator=g).cuda().requires_grad_(True)
    a, b = x1.detach().clone().requires_grad_(True), x2.detach().clone().requires_grad_(True)
    w = torch.randn(17, 23, generator=g).cuda()
    for ours, ref in ((ob.pdist, lambda p, q: torch.sqrt((p * p).sum(1).view(-1, 1) - 2 * p.mm(q.t()) + (q * q).sum(1).view(1, -1) + 1e-4)),
                      (ob.pdist_cos, lambda p, q: (p / p.norm(dim=1)[:, None]).mm((q / q.norm(dim=1)[:, None]).t()))):
        for t in (x1, x2, a, b):
            t.grad = None
        got, want = ours(x1, x2), ref(a, b)
        np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().cpu().numpy(), rtol=2e-5, atol=2e-6)
        (got * w).sum().backward(); (want * w).sum().backward()
        np.testing.assert_allclose(x1.grad.cpu().numpy(), a.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(x2.grad.cpu().numpy(), b.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    z = torch.zeros(2, 96, device="cuda")
    assert torch.equal(ob.pdist_cos(z, x2.detach()), torch.zeros(2, 23, device="cuda"))     # the reference zeroes 0/0
    crit = ob.ContrastiveLoss(dict(name="SAEM"), margin=0.2, measure="cosine", max_violation=True)
    assert crit.sim is ob.pdist_cos


@pytest.mark.parametrize("n,d", [(1, 4), (7, 36), (33, 100), (128, 1024), (200, 2048), (264, 64), (300, 128)])
def test_fused_vse_step_shapes(n, d):
    """The one-launch VSE++ step (csrc/vse_step.cu: batches up to 264) against the float64 oracle, for ragged tile /
    chunk shapes, both hinge modes, with and without gradients; 300 exercises the multi-launch path."""
    g = torch.Generator().manual_seed(n * 1000 + d)
    im = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
    s = torch.nn.functional.normalize(im + 0.5 / d ** 0.5 * torch.randn(n, d, generator=g), dim=-1)
    sc = so.cosine_scores(im.numpy(), s.numpy())
    for mv in (True, False):
        want, dwant = so.hinge_loss(sc, 0.2, mv)
        loss, d_im, d_s = ops.cosine_hinge(im.cuda(), s.cuda(), 0.2, mv)
        np.testing.assert_allclose(loss.item(), want, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(d_im.cpu().numpy(), dwant @ s.numpy().astype(np.float64), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(d_s.cpu().numpy(), dwant.T @ im.numpy().astype(np.float64), rtol=1e-4, atol=2e-5)
        loss2, a2, b2 = ops.cosine_hinge(im.cuda(), s.cuda(), 0.2, mv)             # bit-reproducible
        assert loss2.item() == loss.item() and torch.equal(a2, d_im) and torch.equal(b2, d_s)
        loss3, none_a, none_b = ops.cosine_hinge(im.cuda(), s.cuda(), 0.2, mv, need_grad=False)
        assert loss3.item() == loss.item() and none_a is None and none_b is None

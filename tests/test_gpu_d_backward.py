"""SCAN training backward: the native closed-form kernel chain behind
torch.autograd against gradients the reference produced under autograd (tests/golden/scan_grad.npz) and against
the float64 gradient oracle on seeded inputs.  Everything goes through the C ABI (itr_scan_backward_f32)."""
import numpy as np
import pytest
import torch

from conftest import bits_to_f32, load_golden
import itr_b200
from itr_b200 import objectives as ob, ops
from oracle import scan_backward as sb, scan_oracle as so

pytestmark = pytest.mark.gpu

NORMS = so.RAW_FEATURE_NORMS
AGGS = so.AGG_FUNCS
GRAD_FULL = (("clipped_l2norm", "LogSumExp"), ("l2norm", "Mean"), ("softmax", "Max"), ("clipped", "Sum"), ("no_norm", "LogSumExp"))
# float32 kernels against float64 gradients: relative to the largest entry of each gradient tensor
GRAD_TOL = 2e-4


def cfg(**kw):
    c = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp", lambda_lse=6.0,
             lambda_softmax=9.0, margin=0.2, max_violation=False, measure="cosine")
    c.update(kw)
    return c


def close(got, want, tol=GRAD_TOL, msg=""):
    got = got.detach().cpu().double().numpy() if isinstance(got, torch.Tensor) else np.asarray(got, np.float64)
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want).max() / scale
    assert err < tol, "{}: max error {:.2e} of the gradient scale".format(msg, err)


@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
def test_backward_matches_reference_autograd_golden(direction, lam_sm):
    g = load_golden("scan_grad")
    img = torch.from_numpy(bits_to_f32(g["img_bits"])).cuda()
    cap = torch.from_numpy(bits_to_f32(g["cap_bits"])).cuda()
    lens, ds = g["lens"], torch.from_numpy(g["d_scores"]).float().cuda()
    p_im, p_cap = g["probe_im"].astype(np.float64), g["probe_cap"].astype(np.float64)
    for norm in NORMS:
        for agg in AGGS:
            key = "{}|{}|{}".format(direction, norm, agg)
            d_im, d_cap = ops.scan_backward_f32(img, cap, lens, ds, direction, norm, agg, lam_sm, 6.0)
            d_im, d_cap = d_im.cpu().double().numpy(), d_cap.cpu().double().numpy()
            # random projections pin every mode combination
            pi, pc = (p_im * d_im).reshape(len(p_im), -1).sum(1), (p_cap * d_cap).reshape(len(p_cap), -1).sum(1)
            # (a projection of an error e onto a unit-variance probe is ~ |e|_2; float32 kernels: |e|_2 ~ 1e-6 |grad|_2)
            n_i, n_c = np.sqrt((d_im ** 2).sum()), np.sqrt((d_cap ** 2).sum())
            assert np.abs(pi - g[key + "|proj_im"]).max() < 3e-4 * n_i, key
            assert np.abs(pc - g[key + "|proj_cap"]).max() < 3e-4 * n_c, key
            if (norm, agg) in GRAD_FULL:
                close(d_im, g[key + "|d_im"].astype(np.float64), msg=key + " d_images")
                close(d_cap, g[key + "|d_cap"].astype(np.float64), msg=key + " d_captions")
            assert not d_cap[np.arange(cap.size(1))[None, :] >= lens[:, None]].any(), "padding rows must get zero gradient"


@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
@pytest.mark.parametrize("mv", [False, True])
def test_contrastive_loss_scan_backward_golden(direction, lam_sm, mv):
    """The call the reference's training step makes (Models.py:219-222): criterion(img_emb, cap_emb, cap_len).backward()."""
    g = load_golden("scan_grad")
    img = torch.from_numpy(bits_to_f32(g["img_bits"])).cuda().requires_grad_(True)
    cap = torch.from_numpy(bits_to_f32(g["cap_bits"])).cuda().requires_grad_(True)
    crit = ob.ContrastiveLoss(cfg(cross_attn=direction, lambda_softmax=lam_sm), margin=0.2, measure="cosine", max_violation=mv)
    loss = crit(img, cap, g["lens"].tolist())
    loss.backward()
    key = "{}|hinge|mv{}".format(direction, int(mv))
    np.testing.assert_allclose(loss.item(), g[key + "|loss"], rtol=2e-5)
    close(img.grad, g[key + "|d_im"].astype(np.float64), msg=key + " d_images")
    close(cap.grad, g[key + "|d_cap"].astype(np.float64), msg=key + " d_captions")


@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
def test_backward_seeded_batch_vs_oracle(direction, lam_sm):
    """A ragged batch at embed 1024 with a partial image group, lengths from 1 to 96, gradient through autograd."""
    rng = np.random.default_rng(7)
    n_img, d = 10, 1024
    lens = np.array([1, 96, 17, 33, 2, 64, 12, 9, 81, 5, 27], dtype=np.int32)
    V = rng.standard_normal((n_img, 36, d)); V /= np.linalg.norm(V, axis=-1, keepdims=True)
    W = np.zeros((len(lens), 96, d))
    for c, n in enumerate(lens):
        W[c, :n] = rng.standard_normal((n, d)) / d ** 0.5 + 0.6 * V[c % n_img, rng.integers(0, 36, n)]
    V, W = V.astype(np.float32), W.astype(np.float32)
    dS = rng.standard_normal((n_img, len(lens))).astype(np.float32)
    _, want_im, want_cap = sb.autograd_grads(V, W, lens, dS, direction, "clipped_l2norm", "LogSumExp", lam_sm, 6.0)
    img, cap = torch.from_numpy(V).cuda().requires_grad_(True), torch.from_numpy(W).cuda().requires_grad_(True)
    fn = ob.xattn_score_t2i if direction == "t2i" else ob.xattn_score_i2t
    scores = fn(img, cap, lens, cfg(cross_attn=direction, lambda_softmax=lam_sm))
    (scores * torch.from_numpy(dS).cuda()).sum().backward()
    close(img.grad, want_im, msg="d_images")
    close(cap.grad, want_cap, msg="d_captions")
    # image chunking (small workspace cap) accumulates the caption gradient across chunks: same numbers
    d_im2, d_cap2 = ops.scan_backward_f32(img.detach(), cap.detach(), lens, torch.from_numpy(dS).cuda(), direction, "clipped_l2norm",
                                          "LogSumExp", lam_sm, 6.0, max_workspace_bytes=1 << 20)
    close(d_im2, want_im, msg="chunked d_images")
    close(d_cap2, want_cap, msg="chunked d_captions")
    # only one side needs a gradient
    img3 = torch.from_numpy(V).cuda().requires_grad_(True)
    fn(img3, cap.detach(), lens, cfg(cross_attn=direction, lambda_softmax=lam_sm)).sum().backward()
    assert img3.grad is not None and torch.isfinite(img3.grad).all()


def test_training_batch_128_step():
    """The SCAN training shape (128 x 128, embed 1024, max_violation hinge): loss and gradients against the float64
    oracle on a caption / image subset (the full float64 autograd on the CPU would take minutes)."""
    rng = np.random.default_rng(11)
    n, d = 128, 1024
    lens = np.clip(rng.poisson(11, n) + 2, 3, 40).astype(np.int32)
    V = rng.standard_normal((n, 36, d)); V /= np.linalg.norm(V, axis=-1, keepdims=True)
    W = np.zeros((n, int(lens.max()), d))
    for c, m in enumerate(lens):
        W[c, :m] = rng.standard_normal((m, d)) / d ** 0.5 + 0.6 * V[c, rng.integers(0, 36, m)]
    V, W = V.astype(np.float32), W.astype(np.float32)
    img, cap = torch.from_numpy(V).cuda().requires_grad_(True), torch.from_numpy(W).cuda().requires_grad_(True)
    crit = ob.ContrastiveLoss(cfg(max_violation=True), margin=0.2, measure="cosine", max_violation=True)
    loss = crit(img, cap, lens.tolist())
    loss.backward()
    scores = ops.scan_scores_f32(img.detach(), cap.detach(), lens, "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    want_loss, d_scores = so.hinge_loss(scores.cpu().double().numpy(), 0.2, True)
    np.testing.assert_allclose(loss.item(), want_loss, rtol=1e-5)
    # the oracle on the rows / columns the max-violation hinge touched for the first 6 images and captions
    sel = np.arange(6)
    cols = np.unique(np.nonzero(d_scores[sel])[1])
    rows = np.unique(np.concatenate([sel, np.nonzero(d_scores[:, sel])[0]]))
    _, part_im, _ = sb.coefficient_form(V[sel], W[cols], lens[cols], d_scores[np.ix_(sel, cols)], "t2i")
    close(img.grad[:6], part_im, msg="d_images[:6]")
    _, _, part_cap = sb.coefficient_form(V[rows], W[sel], lens[sel], d_scores[np.ix_(rows, sel)], "t2i")
    close(cap.grad[:6], part_cap, msg="d_captions[:6]")


def test_backward_argument_errors():
    img, cap = torch.zeros(4, 36, 64, device="cuda"), torch.zeros(3, 5, 64, device="cuda")
    with pytest.raises(ValueError):
        ops.scan_backward_f32(img, cap, [5, 5, 5], torch.zeros(4, 2, device="cuda"), "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    with pytest.raises(ValueError):
        ops.scan_backward_f32(img, cap, [5, 9, 5], torch.zeros(4, 3, device="cuda"), "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    # the kernels are bounded by the batch's true longest caption (96 words), not by the padded width
    with pytest.raises(ValueError):
        ops.scan_backward_f32(img, torch.zeros(3, 100, 64, device="cuda"), [5, 100, 5], torch.zeros(4, 3, device="cuda"), "i2t",
                              "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    wide = torch.randn(3, 100, 64, device="cuda")
    wide[:, 5:] = 0
    g = torch.randn(4, 3, device="cuda")
    im = torch.nn.functional.normalize(torch.randn(4, 36, 64, device="cuda"), dim=-1)
    d_im_w, d_cap_w = ops.scan_backward_f32(im, wide, [5, 5, 5], g, "i2t", "clipped_l2norm", "LogSumExp", 4.0, 6.0)
    d_im_n, d_cap_n = ops.scan_backward_f32(im, wide[:, :5].contiguous(), [5, 5, 5], g, "i2t", "clipped_l2norm", "LogSumExp", 4.0, 6.0)
    assert d_cap_w.shape == (3, 100, 64) and torch.equal(d_cap_w[:, :5], d_cap_n) and not d_cap_w[:, 5:].any()
    assert torch.equal(d_im_w, d_im_n)
    assert torch.equal(ops.scan_scores_f32(im, wide, [5, 5, 5], "i2t", "clipped_l2norm", "LogSumExp", 4.0, 6.0),
                       ops.scan_scores_f32(im, wide[:, :5].contiguous(), [5, 5, 5], "i2t", "clipped_l2norm", "LogSumExp", 4.0, 6.0))
    with pytest.raises(ValueError):
        ops.scan_backward_f32(img, cap, [5, 5, 5], torch.zeros(4, 3, device="cuda"), "t2i", "l1norm", "LogSumExp", 9.0, 6.0)


@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
@pytest.mark.parametrize("n_regions,d", [(20, 300), (7, 64), (36, 2048)])
def test_other_region_counts_and_embed_sizes(direction, lam_sm, n_regions, d):
    """Grid features / other encoders: any region count up to 36 and any embedding size (multiple of 4 for the
    backward) run the float32 kernels, forward and backward."""
    rng = np.random.default_rng(n_regions * 1000 + d)
    n_img = 6
    lens = np.array([5, 11, 2, 23, 8], dtype=np.int32)
    V = rng.standard_normal((n_img, n_regions, d)); V /= np.linalg.norm(V, axis=-1, keepdims=True)
    W = np.zeros((len(lens), int(lens.max()), d))
    for c, n in enumerate(lens):
        W[c, :n] = rng.standard_normal((n, d)) / d ** 0.5 + 0.6 * V[c % n_img, rng.integers(0, n_regions, n)]
    V, W = V.astype(np.float32), W.astype(np.float32)
    dS = rng.standard_normal((n_img, len(lens))).astype(np.float32)
    want, want_im, want_cap = sb.autograd_grads(V, W, lens, dS, direction, "clipped_l2norm", "LogSumExp", lam_sm, 6.0)
    img, cap = torch.from_numpy(V).cuda().requires_grad_(True), torch.from_numpy(W).cuda().requires_grad_(True)
    fn = ob.xattn_score_t2i if direction == "t2i" else ob.xattn_score_i2t
    scores = fn(img, cap, lens, cfg(cross_attn=direction, lambda_softmax=lam_sm))
    np.testing.assert_allclose(scores.detach().cpu().numpy(), want, rtol=2e-5, atol=2e-6)
    with torch.no_grad():       # the no-grad call takes the same float32 kernel (the tensor-core path is 36 x 1024 only)
        np.testing.assert_allclose(fn(img, cap, lens, cfg(cross_attn=direction, lambda_softmax=lam_sm)).cpu().numpy(), want,
                                   rtol=2e-5, atol=2e-6)
    (scores * torch.from_numpy(dS).cuda()).sum().backward()
    close(img.grad, want_im, msg="d_images")
    close(cap.grad, want_cap, msg="d_captions")


def test_training_forward_on_tensor_cores_is_opt_in():
    """itr_b200_train_precision="bf16": the scores under autograd come from the fused tcgen05 kernel (within 1e-3 of the
    float32 ones), the gradients are the same float32 closed form."""
    img, cap, ln = itr_b200.synth.scan_inputs(12, 12, 10.5, 23, device="cuda", lengths=np.array([5, 9, 14, 3, 22, 7, 11, 30, 4, 16, 8, 12]))
    out = {}
    for mode in ("fp32", "bf16"):
        a, b = img.clone().requires_grad_(True), cap.clone().requires_grad_(True)
        crit = ob.ContrastiveLoss(cfg(itr_b200_train_precision=mode, max_violation=False), margin=0.2, measure="cosine", max_violation=False)
        loss = crit(a, b, ln.tolist())
        loss.backward()
        out[mode] = (loss.item(), a.grad.clone(), b.grad.clone())
    assert abs(out["bf16"][0] - out["fp32"][0]) <= 2e-3 * abs(out["fp32"][0])
    assert out["bf16"][0] != out["fp32"][0]                       # it really took the other forward
    for k in (1, 2):                                              # sum hinge: same active set up to near-ties -> close gradients
        scale = out["fp32"][k].abs().max().item()
        assert (out["bf16"][k] - out["fp32"][k]).abs().max().item() <= 0.05 * scale
    with pytest.raises(ValueError):
        ob.xattn_score_t2i(img.clone().requires_grad_(True), cap, ln, cfg(itr_b200_train_precision="fp8"))

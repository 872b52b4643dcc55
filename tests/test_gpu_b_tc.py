"""GPU parity of the tcgen05 SCAN t2i kernel (bf16 inputs, fp32 accumulate) against the oracle fed
the SAME bf16-rounded values: scores within 1e-3 relative, ranks exact except where reference
scores tie within that tolerance (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from conftest import bits_to_f32, load_golden
import itr_b200
from itr_b200 import evaluation as ev, objectives as ob, ops, sharding
from oracle import scan_oracle as so

pytestmark = pytest.mark.gpu
RTOL_TC = 1e-3       # "1e-3 relative for bf16/TF32 inputs"
ATOL_TC = 1e-6


def cfg(**kw):
    base = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp",
                lambda_lse=6.0, lambda_softmax=9.0, margin=0.2, max_violation=True, measure="cosine",
                itr_b200_precision="bf16")
    base.update(kw)
    return base


class FakeModel:
    sim_enc = None

    def __init__(self, config):
        self.config = config
        self.criterion = ob.ContrastiveLoss(config, margin=0.2, measure="cosine", max_violation=True)


def test_affinity_tile_matches_matmul():
    """The raw tensor-core contraction of one (word tile, image tile) pair: TMA swizzle, UMMA
    descriptors and the TMEM lane/column mapping all have to be right for this to hold."""
    img, cap, lens = itr_b200.synth.scan_inputs(11, 60, 10.5, 21, round_to="bf16")
    pi = ops.prepare_images(img.cuda())
    pc = ops.prepare_captions(cap.cuda(), lens)
    meta = pc.row_meta.cpu().numpy().reshape(pc.n_tiles, 128, 4)
    words = pc.words_bf16.float().cpu().numpy().reshape(pc.n_tiles, 128, 1024)
    np.testing.assert_array_equal(pi.images_bf16.float().cpu().numpy(), img.numpy())
    for wt in range(pc.n_tiles):
        # packed rows are the right words, padding rows are zero, norms are the rounded rows' norms
        for row in range(128):
            c, j = meta[wt, row, 0], meta[wt, row, 1]
            want = cap[c, j].numpy() if c >= 0 else np.zeros(1024, np.float32)
            np.testing.assert_array_equal(words[wt, row], want)
        np.testing.assert_allclose(pc.row_wnorm.cpu().numpy().reshape(-1, 128)[wt],
                                   np.linalg.norm(words[wt].astype(np.float64), axis=1), rtol=1e-6)
        for it in range(3):
            got = ops.scan_t2i_affinity_debug(pi, pc, wt, it).cpu().numpy()
            v = np.zeros((144, 1024), np.float64)
            rows = img.numpy().reshape(-1, 1024)[it * 144:(it + 1) * 144]
            v[: len(rows)] = rows
            want = words[wt].astype(np.float64) @ v.T
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
    # Gram pack of the rounded regions: fp16 off-diagonal block in UMMA core-matrix order + fp32 diagonal
    G = np.einsum("ikd,ild->ikl", img.numpy().astype(np.float64), img.numpy().astype(np.float64))
    pack = pi.gram_pack.cpu().numpy()
    assert pack.shape == (11, 4752)
    for i in (0, 10):
        g16 = pack[i, :4608].copy().view(np.float16).astype(np.float64)
        gd = pack[i, 4608:].copy().view(np.float32)
        np.testing.assert_allclose(gd, np.diag(G[i]), rtol=2e-6)        # fp32 accumulation of 1024 products
        dense = np.zeros((48, 48))
        for n_ in range(48):
            for k_ in range(48):
                dense[n_, k_] = g16[(n_ // 8) * 384 + (k_ // 8) * 64 + (n_ % 8) * 8 + (k_ % 8)]
        want = G[i] - np.eye(36)          # unit diagonal stays in fp32 on the CUDA cores (sum_k e_k^2)
        np.testing.assert_allclose(dense[:36, :36], want, rtol=1e-3, atol=1e-6)      # fp16 rounding
        # output column 36 is the all-ones vector over k < 36 (softmax denominator); the rest is zero padding
        assert (dense[36, :36] == 1).all() and (dense[36, 36:] == 0).all()
        assert (dense[37:] == 0).all() and (dense[:36, 36:] == 0).all()


@pytest.mark.parametrize("case", ["scan_small", "scan_long"])
def test_tc_golden(case):
    g = load_golden(case)
    img = torch.from_numpy(bits_to_f32(g["img_bits"])).cuda()
    cap = torch.from_numpy(bits_to_f32(g["cap_bits"])).cuda()
    lens = g["lens"]
    for norm in ("clipped_l2norm", "l2norm"):
        for agg in so.AGG_FUNCS:
            got = ob.xattn_score_t2i(img, cap, lens, cfg(raw_feature_norm=norm, agg_func=agg)).cpu().numpy()
            want = g["t2i|{}|{}|f64".format(norm, agg)]
            np.testing.assert_allclose(got, want, rtol=RTOL_TC, atol=ATOL_TC, err_msg="{} {}".format(norm, agg))


def test_tc_vs_oracle_ragged_lengths_and_image_tail():
    # 13 images: the last image tile is partial; lengths cover 1, 32, 33 (long tile), 72, 128
    img, cap, lens = itr_b200.synth.scan_inputs(13, 45, 10.5, 33, round_to="bf16")
    lmax = cap.size(1)
    extra = torch.zeros(45, 128 - lmax, 1024)
    cap = torch.cat([cap, extra], 1)
    g = torch.Generator().manual_seed(1)
    for c, n in [(0, 1), (1, 32), (2, 33), (3, 72), (4, 128), (5, 2)]:
        cap[c, :n] = (torch.randn(n, 1024, generator=g) / 32).to(torch.bfloat16).float()
        cap[c, n:] = 0
        lens[c] = n
    for agg, lam_lse in (("LogSumExp", 6.0), ("Mean", 6.0), ("Max", 6.0), ("Sum", 6.0), ("LogSumExp", 20.0)):
        want = so.scan_scores(img.numpy(), cap.numpy(), lens, "t2i", "clipped_l2norm", agg, 9.0, lam_lse)
        got = ob.xattn_score_t2i(img.cuda(), cap.cuda(), lens, cfg(agg_func=agg, lambda_lse=lam_lse)).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=RTOL_TC, atol=ATOL_TC, err_msg=agg)
    # padding content must not matter
    dirty = cap.clone()
    for c, n in enumerate(lens):
        dirty[c, n:] = 3.0
    a = ob.xattn_score_t2i(img.cuda(), cap.cuda(), lens, cfg()).cpu().numpy()
    b = ob.xattn_score_t2i(img.cuda(), dirty.cuda(), lens, cfg()).cpu().numpy()
    np.testing.assert_array_equal(a, b)


def test_tc_vs_fp32_kernel_and_ranks_medium():
    """1000 x 500 block (the CPU-baseline sub-block of BASELINE.md section 4): the two CUDA paths agree to
    1e-3, the tensor-core path agrees with the float64 oracle on a sample of columns, and ranks are
    equal wherever the reference scores do not tie within the tolerance."""
    img, cap, lens = itr_b200.synth.scan_inputs(1000, 500, 12.4, 30, device="cuda", round_to="bf16")
    tc = ob.xattn_score_t2i(img, cap, lens, cfg())
    f32 = ob.xattn_score_t2i(img, cap, lens, cfg(itr_b200_precision="fp32"))
    rel = ((tc - f32).abs() / f32.abs().clamp_min(1e-6)).max().item()
    assert rel < RTOL_TC, rel
    cols = np.arange(0, 500, 37)
    want = so.scan_scores(img[:96].cpu().numpy(), cap[cols].cpu().numpy(), lens[cols], "t2i", "clipped_l2norm",
                          "LogSumExp", 9.0, 6.0)
    np.testing.assert_allclose(tc[:96][:, cols].cpu().numpy(), want, rtol=RTOL_TC, atol=ATOL_TC)
    # ranks: captions 0..499 belong to images 0..99
    a_tc = [x.cpu().numpy() for x in ev.device_ranks(tc[:100])]
    ref = f32[:100].double().cpu().numpy()
    ri, rc, _, _ = so.strict_ranks(ref)
    # a query is "tie-exempt" if some competitor is within 1e-3 relative of its threshold
    thr_i = np.take_along_axis(ref, 5 * np.arange(100)[:, None] + np.arange(5)[None], 1).max(1)
    near_i = (np.abs(ref - thr_i[:, None]) <= 2e-3 * np.abs(thr_i[:, None])).sum(1) > 1
    thr_c = ref[np.arange(500) // 5, np.arange(500)]
    near_c = (np.abs(ref - thr_c[None]) <= 2e-3 * np.abs(thr_c[None])).sum(0) > 1
    np.testing.assert_array_equal(a_tc[0][~near_i], ri[~near_i])
    np.testing.assert_array_equal(a_tc[2][~near_c], rc[~near_c])
    assert (~near_i).sum() > 50 and (~near_c).sum() > 250


def test_tc_equivariance_and_cal_sims_fused_path():
    img, cap, lens = itr_b200.synth.scan_inputs(40, 200, 10.5, 8, round_to="bf16")
    model = FakeModel(cfg())
    base = ev.device_sims(model, img.numpy(), cap.numpy(), lens)
    # permuting captions / images permutes the matrix (packing must not leak into results)
    pc = torch.randperm(200, generator=torch.Generator().manual_seed(0))
    pim = torch.randperm(40, generator=torch.Generator().manual_seed(1))
    perm = ev.device_sims(model, img[pim].numpy(), cap[pc].numpy(), lens[pc.numpy()])
    torch.testing.assert_close(perm, base[pim.cuda()][:, pc.cuda()], rtol=1e-5, atol=1e-6)
    # pinned host captions are gathered in place and give the same numbers
    pinned = cap.pin_memory()
    torch.testing.assert_close(ev.device_sims(model, img, pinned, lens), base, rtol=0, atol=0)
    # host API: float64 out, same keys as the reference's cal_recall
    sims = ev.cal_sims(model, img.numpy(), cap.numpy(), lens, shard_size=64)
    assert sims.dtype == np.float64
    res = ev.cal_sims_and_recall(model, img.numpy(), cap.numpy(), lens)
    want = so.recall_dict(sims)
    for k in ("i2t_ranks", "t2i_ranks", "i2t_top1", "t2i_top1"):
        np.testing.assert_array_equal(res[k], want[k])
    # single-process sharded path == plain path
    res2 = sharding.sharded_scan_eval(img, cap, lens, 0, 200, cfg())
    for k in ("i2t_ranks", "t2i_ranks", "i2t_top1", "t2i_top1", "rsum"):
        np.testing.assert_array_equal(res2[k], res[k])
    # two "virtual ranks" on one GPU: column blocks ranked separately and merged by hand
    lo, hi = sharding.shard_bounds(200, 2)[1]
    blk = sharding.sharded_scan_eval(img, cap[lo:hi], lens[lo:hi], lo, 200, cfg(), return_block=True)["sims_block"]
    torch.testing.assert_close(blk, base[:, lo:hi], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", ["scan_small", "scan_long"])
@pytest.mark.parametrize("direction,lam_sm", [("t2i", 9.0), ("i2t", 4.0)])
def test_tc_generic_two_phase_golden(case, direction, lam_sm):
    """tcgen05 affinities + fp32 epilogue kernel: both directions, all five norm modes, all four aggregations."""
    g = load_golden(case)
    img = torch.from_numpy(bits_to_f32(g["img_bits"])).cuda()
    cap = torch.from_numpy(bits_to_f32(g["cap_bits"])).cuda()
    lens = g["lens"]
    for norm in so.RAW_FEATURE_NORMS:
        for agg in so.AGG_FUNCS:
            got = ops.scan_scores_tc_generic(img, cap, lens, direction, norm, agg, lam_sm, 6.0).cpu().numpy()
            want = g["{}|{}|{}|f64".format(direction, norm, agg)]
            np.testing.assert_allclose(got, want, rtol=RTOL_TC, atol=ATOL_TC, err_msg="{} {} {}".format(direction, norm, agg))
    # the public entry points dispatch here for i2t (bf16 mode is the default)
    fn = ob.xattn_score_i2t if direction == "i2t" else ob.xattn_score_t2i
    got = fn(img, cap, lens, cfg(cross_attn=direction, raw_feature_norm="softmax", agg_func="Mean", lambda_softmax=lam_sm)).cpu().numpy()
    np.testing.assert_allclose(got, g["{}|softmax|Mean|f64".format(direction)], rtol=RTOL_TC, atol=ATOL_TC)


def test_tc_generic_i2t_medium_chunked_vs_fp32_kernel():
    """Config-4 style block (i2t, Mean, lambda 4): chunked image loop == one chunk == fp32 CUDA-core kernel (1e-3)."""
    img, cap, lens = itr_b200.synth.scan_inputs(300, 400, 10.5, 14, device="cuda", round_to="bf16")
    a = ops.scan_scores_tc_generic(img, cap, lens, "i2t", "clipped_l2norm", "Mean", 4.0, 6.0)
    b = ops.scan_scores_tc_generic(img, cap, lens, "i2t", "clipped_l2norm", "Mean", 4.0, 6.0, max_affinity_bytes=64 << 20)
    assert torch.equal(a, b)
    f32 = ops.scan_scores_f32(img, cap, lens, "i2t", "clipped_l2norm", "Mean", 4.0, 6.0)
    rel = ((a - f32).abs() / f32.abs().clamp_min(1e-6)).max().item()
    assert rel < 1e-3, rel
    want = so.scan_scores(img[:16].cpu().numpy(), cap[:24].cpu().numpy(), lens[:24], "i2t", "clipped_l2norm", "Mean", 4.0, 6.0)
    np.testing.assert_allclose(a[:16, :24].cpu().numpy(), want, rtol=RTOL_TC, atol=ATOL_TC)


@pytest.mark.parametrize("n_img,lens", [(1, [1]), (1, [5, 3]), (3, [2, 40]), (5, [32] * 9), (7, [1] * 130 + [128, 33]),
                                        (9, list(range(1, 33)) * 5)])
def test_tc_edge_shapes(n_img, lens):
    """Degenerate and schedule-stressing shapes: single image / caption / word, exactly full quarters, more single-word
    captions than one tile holds, many tiles in several bands, a lone long caption."""
    lens = np.asarray(lens, dtype=np.int32)
    g = torch.Generator().manual_seed(int(lens.sum()) + n_img)
    img = torch.nn.functional.normalize(torch.randn(n_img, 36, 1024, generator=g), dim=-1).to(torch.bfloat16).float()
    cap = torch.zeros(len(lens), int(lens.max()), 1024)
    for c, n in enumerate(lens):
        cap[c, :n] = (torch.randn(int(n), 1024, generator=g) / 32 + 0.3 * img[c % n_img, torch.randint(0, 36, (int(n),), generator=g)]).to(torch.bfloat16).float()
    sel = np.unique(np.linspace(0, len(lens) - 1, 12).astype(int))
    want = so.scan_scores(img.numpy(), cap[sel].numpy(), lens[sel], "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    got = ob.xattn_score_t2i(img.cuda(), cap.cuda(), lens, cfg()).cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got[:, sel], want, rtol=RTOL_TC, atol=ATOL_TC)
    i2t_cfg = cfg(cross_attn="i2t", agg_func="Mean", lambda_softmax=4.0)
    if int(lens.max()) > ops.GENERIC_MAX_WORDS:           # documented limit of the two-phase path: loud, not silent
        with pytest.raises(ValueError):
            ob.xattn_score_i2t(img.cuda(), cap.cuda(), lens, i2t_cfg)
        return
    got_i2t = ob.xattn_score_i2t(img.cuda(), cap.cuda(), lens, i2t_cfg).cpu().numpy()
    want_i2t = so.scan_scores(img.numpy(), cap[sel].numpy(), lens[sel], "i2t", "clipped_l2norm", "Mean", 4.0, 6.0)
    np.testing.assert_allclose(got_i2t[:, sel], want_i2t, rtol=RTOL_TC, atol=ATOL_TC)


def test_tc_pipelined_host_captions_match_device_path():
    """Captions in pinned host memory: chunked PCIe gather on a side stream under the score kernel gives the same
    matrix as the one-shot device path, bit for bit (same kernel, same packing within a chunk boundary or not)."""
    rng = np.random.default_rng(5)
    lens = np.clip(rng.poisson(9, 400) + 1, 1, 30).astype(np.int32)
    img, cap, ln = itr_b200.synth.scan_inputs(13, 400, 10.5, 21, device="cuda", lengths=lens)
    pi = ops.prepare_images(img)
    want = ops.scan_t2i_scores_bf16(pi, ops.prepare_captions(cap, ln), "clipped_l2norm", "LogSumExp", 9.0, 6.0)
    host = cap.cpu().pin_memory()
    assert ops.host_caption_chunks(ln) == [(0, 400)]
    chunks = ops.host_caption_chunks(ln, min_words=64)
    assert len(chunks) == 3 and chunks[0][0] == 0 and chunks[-1][1] == 400 and all(a % 5 == 0 for a, _ in chunks)
    for ch in (None, chunks, [(0, 5), (5, 395), (395, 400)]):
        got = ops.scan_t2i_scores_from_host(pi, host, ln, "clipped_l2norm", "LogSumExp", 9.0, 6.0, chunks=ch)
        torch.cuda.synchronize()
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=2e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------------
# fused evaluation: ranking inside the score kernel (ground-truth pre-pass + counting epilogue), no score matrix
def _fused_vs_matrix(n_img, n_cap, lam, seed, lengths=None, agg="LogSumExp", shards=1):
    img, cap, lens = itr_b200.synth.scan_inputs(n_img, n_cap, lam, seed, device="cuda", lengths=lengths)
    pi = ops.prepare_images(img)
    args = ("clipped_l2norm", agg, 9.0, 6.0)
    pc_all = ops.prepare_captions(cap, lens)
    scores = ops.scan_t2i_scores_bf16(pi, pc_all, *args)
    want = [x.cpu().numpy() for x in ev.device_ranks(scores)]
    # per shard: thresholds from the pre-pass, merged like the multi-GPU exchange (max of thresholds, sum of counts,
    # max of keys), then the counting pass
    bounds = sharding.shard_bounds(n_cap, shards)
    pcs = [ops.prepare_captions(cap[lo:hi], lens[lo:hi]) if hi > lo else None for lo, hi in bounds]
    thr_row = torch.full((n_img,), float("-inf"), device="cuda")
    thr_cols = []
    for (lo, hi), pc in zip(bounds, pcs):
        if pc is None:
            thr_cols.append(None)
            continue
        tr, tc_ = ops.scan_t2i_gt_thresholds(pi, pc, *args, cap_offset=lo, caps_per_img=5)
        thr_row = torch.maximum(thr_row, tr)
        thr_cols.append(tc_)
        # the thresholds ARE the matrix entries, bit for bit
        gt_img = (torch.arange(lo, hi, device="cuda") // 5)
        has = gt_img < n_img
        assert torch.equal(tc_[has], scores[gt_img[has], torch.arange(lo, hi, device="cuda")[has]])
        assert torch.isnan(tc_[~has]).all()
    i2t = torch.zeros(n_img, dtype=torch.int64, device="cuda")
    best = torch.zeros(n_img, dtype=torch.int64, device="cuda")
    t2i, t2i_top = [], []
    for (lo, hi), pc, tc_ in zip(bounds, pcs, thr_cols):
        if pc is None:
            continue
        out = torch.full((n_img, hi - lo), float("nan"), device="cuda")
        cr, cc, br, bc = ops.scan_t2i_count(pi, pc, *args, thr_row, tc_, cap_offset=lo, out=out)
        assert torch.equal(out, scores[:, lo:hi])                       # the optional matrix is the same matrix
        cr2, cc2, br2, bc2 = ops.scan_t2i_count(pi, pc, *args, thr_row, tc_, cap_offset=lo)      # and without it
        assert torch.equal(cr, cr2) and torch.equal(cc, cc2) and torch.equal(br, br2) and torch.equal(bc, bc2)
        i2t += cr.long()
        best = torch.maximum(best ^ sharding._SIGN, br ^ sharding._SIGN) ^ sharding._SIGN
        t2i.append(cc.long()); t2i_top.append(ops.unpack_best_index(bc))
    np.testing.assert_array_equal(i2t.cpu().numpy(), want[0])
    np.testing.assert_array_equal(torch.cat(t2i).cpu().numpy(), want[2])
    np.testing.assert_array_equal(ops.unpack_best_index(best).cpu().numpy(), want[1])
    has_gt = (np.arange(n_cap) // 5) < n_img                             # top-1 of a caption without an image is undefined
    np.testing.assert_array_equal(torch.cat(t2i_top).cpu().numpy()[has_gt], want[3][has_gt])


@pytest.mark.parametrize("n_img,n_cap,shards", [(10, 50, 1), (37, 185, 1), (37, 185, 3), (64, 333, 2), (5, 25, 4), (130, 650, 1)])
def test_fused_ranking_matches_matrix_ranking(n_img, n_cap, shards):
    _fused_vs_matrix(n_img, n_cap, 10.5, 31 + n_img, shards=shards)


def test_fused_ranking_long_captions_and_other_aggregations():
    lens = np.array([72, 40, 33, 128, 3, 5, 17, 32, 31, 1] * 4, dtype=np.int32)
    for agg in ("LogSumExp", "Mean", "Max", "Sum"):
        _fused_vs_matrix(8, 40, 10.5, 77, lengths=lens, agg=agg, shards=2)


def test_cal_sims_and_recall_never_builds_the_matrix():
    """cal_sims_and_recall(return_sims=False) on the tensor-core SCAN t2i path: same dict as the matrix path, and no
    (n_img, n_cap) allocation: its peak device memory is lower than the matrix path's by the size of the matrix."""
    n_img, n_cap = 1000, 5000
    img, cap, lens = itr_b200.synth.scan_inputs(n_img, n_cap, 10.5, 8, device="cuda")
    model = FakeModel(cfg())

    def peak_of(**kw):
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        res = ev.cal_sims_and_recall(model, img, cap, lens, **kw)
        torch.cuda.synchronize()
        return res, torch.cuda.max_memory_allocated() - base

    peak_of()                                   # warm the plan caches so both measured runs allocate the same operands
    got, peak_fused = peak_of()
    want, peak_matrix = peak_of(return_sims=True)
    assert want.pop("sims").shape == (n_img, n_cap)
    for k in want:
        if isinstance(want[k], np.ndarray):
            np.testing.assert_array_equal(got[k], want[k])
        else:
            assert got[k] == want[k], k
    matrix_bytes = n_img * n_cap * 4
    assert peak_matrix - peak_fused >= 0.9 * matrix_bytes, (peak_matrix, peak_fused, matrix_bytes)


@pytest.mark.parametrize("agg", ["LogSumExp", "Max"])
def test_fused_i2t_many_tiny_captions(agg):
    """Captions of 1..3 words pack 11 to 32 to a 32-row quarter: the fused i2t kernel then walks several caption n-tiles
    of eight per quarter (its `ct` loop) -- against the two-phase path and the float64 oracle."""
    n_img, n_cap = 21, 300
    lens = (np.arange(n_cap) % 3 + 1).astype(np.int32)
    lens[:40] = 1                                     # a few quarters of 32 single-word captions
    img, cap, lens = itr_b200.synth.scan_inputs(n_img, n_cap, 2.0, 77, device="cuda", lengths=lens, round_to="bf16")
    got = ops.scan_i2t_scores_tc(img, cap, lens, "l2norm", agg, 4.0, 6.0)
    two = ops.scan_scores_tc_generic(img, cap, lens, "i2t", "l2norm", agg, 4.0, 6.0)
    assert torch.isfinite(got).all()
    assert ((got - two).abs() / two.abs().clamp_min(1e-6)).max().item() < 2e-3
    want = so.scan_scores(img.cpu().numpy(), cap.cpu().numpy(), lens, "i2t", "l2norm", agg, 4.0, 6.0)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL_TC, atol=ATOL_TC)
    assert torch.equal(got, ops.scan_i2t_scores_tc(img, cap, lens, "l2norm", agg, 4.0, 6.0))


def test_streamed_host_images_give_the_same_results(monkeypatch):
    """Large host image arrays are uploaded in chunks and scored range by range as they land
    (ops.prepare_images_streamed): pinned or pageable source, matrix or fused ranking, device or host captions -- the
    results must be those of one upload + one launch, bit for bit."""
    monkeypatch.setattr(ops, "STREAMED_IMAGES_MIN_BYTES", 0)
    n_img, n_cap = 70, 350
    img, cap, lens = itr_b200.synth.scan_inputs(n_img, n_cap, 10.5, 43, device="cuda")
    args = ("clipped_l2norm", "LogSumExp", 9.0, 6.0)
    pc = ops.prepare_captions(cap, lens)
    want = ops.scan_t2i_scores_bf16(ops.prepare_images(img), pc, *args)
    want_ranks = sharding.sharded_ranks(want, 0, n_cap, None, 5)
    pinned = torch.empty(img.shape, dtype=torch.float32, pin_memory=True).copy_(img)
    pageable = img.cpu().numpy().copy()
    cap_pinned = torch.empty(cap.shape, dtype=torch.float32, pin_memory=True).copy_(cap)
    for src in (pinned, pageable):
        for chunks in (3, 8):
            pi = ops.prepare_images_streamed(src, "cuda", chunks=chunks)
            ranges = pi.row_ranges()
            assert len(ranges) >= 3 and ranges[0][0] == 0 and ranges[-1][1] == n_img and all(lo % 4 == 0 for lo, _, _ in ranges)
            assert torch.equal(ops.scan_t2i_scores_bf16(pi, pc, *args), want)
            assert pi.pending is None and pi.row_ranges() == [(0, n_img, False)]       # every range has been waited for
            # captions streaming from pinned host memory in three chunks at the same time
            pi = ops.prepare_images_streamed(src, "cuda", chunks=chunks)
            got = ops.scan_t2i_scores_from_host(pi, cap_pinned, lens, *args, chunks=[(0, 60), (60, 200), (200, n_cap)])
            assert torch.equal(got, want)
            # fused ranking: pre-pass on all rows, counting range by range with column accumulation
            stats = sharding.FusedScanStats(ops.prepare_images_streamed(src, "cuda", chunks=chunks), pc, *args)
            got_ranks = sharding.sharded_ranks(stats.block(), 0, n_cap, None, 5, stats)
            assert all(torch.equal(a, b) for a, b in zip(got_ranks, want_ranks))
    # small arrays and device arrays are not streamed
    monkeypatch.setattr(ops, "STREAMED_IMAGES_MIN_BYTES", 1 << 40)
    assert ops.prepare_images_streamed(pinned, "cuda").pending is None
    assert ops.prepare_images_streamed(img, "cuda", chunks=4).pending is None


def test_image_row_ranges_give_the_same_results():
    """The multi-GPU path scores its own image rows before the other ranks' rows have arrived (PreparedImages.row_ranges):
    launching the kernels per row range -- scores, ground-truth pre-pass on the local rows only, counting with column
    accumulation -- must change nothing."""
    n_img, n_cap = 50, 250
    img, cap, lens = itr_b200.synth.scan_inputs(n_img, n_cap, 10.5, 41, device="cuda")
    args = ("clipped_l2norm", "LogSumExp", 9.0, 6.0)
    pc = ops.prepare_captions(cap, lens)
    whole = ops.prepare_images(img)
    want = ops.scan_t2i_scores_bf16(whole, pc, *args)
    want_ranks = sharding.sharded_ranks(want, 0, n_cap, None, 5)

    def split(lo, hi):
        pi = ops.prepare_images(img)
        ev_ = torch.cuda.Event()
        ev_.record()
        pi.local_rows, pi.gathered = (lo, hi), ev_
        return pi

    for lo, hi in ((0, 10), (13, 31), (40, 50), (0, 50)):
        pi = split(lo, hi)
        assert [r[:2] for r in pi.row_ranges()][0] == (lo, hi)
        assert torch.equal(ops.scan_t2i_scores_bf16(pi, pc, *args), want)
        assert pi.gathered is None or (lo, hi) == (0, n_img)        # waited for, unless there was nothing to wait for
        # captions [5 lo, 5 hi) are this "rank's" shard: their ground-truth images are the local rows
        c0, c1 = 5 * lo, 5 * hi
        pc_loc = ops.prepare_captions(cap[c0:c1], lens[c0:c1])
        stats = sharding.FusedScanStats(split(lo, hi), pc_loc, *args)
        tr, tc_ = stats.thresholds(stats.block(), c0, 5)
        assert torch.equal(tc_, want[torch.arange(c0, c1, device="cuda") // 5, torch.arange(c0, c1, device="cuda")])
        assert torch.isinf(tr[:lo]).all() and torch.isinf(tr[hi:]).all()
        assert torch.equal(tr[lo:hi], want[lo:hi, c0:c1].view(hi - lo, hi - lo, 5)[torch.arange(hi - lo), torch.arange(hi - lo)].max(dim=1).values)
        stats = sharding.FusedScanStats(split(lo, hi), pc, *args)
        got = sharding.sharded_ranks(stats.block(), 0, n_cap, None, 5, stats)
        for a, b in zip(got, want_ranks):
            assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------------------------
# fused i2t kernel (csrc/scan_i2t_tc2.cu)
@pytest.mark.parametrize("case", ["scan_small", "scan_long"])
def test_fused_i2t_golden(case):
    """xattn_score_i2t on the fused tensor-core kernel against the reference's own outputs: both l2 feature norms, all
    four aggregations (scan_long: every caption is longer than 32 words and takes the two-phase path)."""
    g = load_golden(case)
    img = torch.from_numpy(bits_to_f32(g["img_bits"])).cuda()
    cap = torch.from_numpy(bits_to_f32(g["cap_bits"])).cuda()
    lens = g["lens"]
    for norm in ("clipped_l2norm", "l2norm"):
        for agg in so.AGG_FUNCS:
            got = ob.xattn_score_i2t(img, cap, lens, cfg(cross_attn="i2t", raw_feature_norm=norm, agg_func=agg, lambda_softmax=4.0))
            want = g["i2t|{}|{}|f64".format(norm, agg)]
            np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL_TC, atol=ATOL_TC, err_msg="{} {}".format(norm, agg))
    if case == "scan_small":
        pi, pc = ops.prepare_images(img), ops.prepare_captions(cap, lens)
        a = ops.scan_i2t_scores_bf16(pi, pc, "clipped_l2norm", "Mean", 4.0, 6.0)
        np.testing.assert_allclose(a.cpu().numpy(), g["i2t|clipped_l2norm|Mean|f64"], rtol=RTOL_TC, atol=ATOL_TC)
        assert torch.equal(a, ops.scan_i2t_scores_bf16(pi, pc, "clipped_l2norm", "Mean", 4.0, 6.0))      # deterministic


@pytest.mark.parametrize("n_img,n_cap,agg", [(37, 185, "Mean"), (130, 333, "LogSumExp"), (5, 64, "Max"), (64, 200, "Sum")])
def test_fused_i2t_matches_two_phase_and_oracle(n_img, n_cap, agg):
    lens = itr_b200.synth.caption_lengths(n_cap, 10.5, 7 + n_img)
    lens[::13] = 32                                  # exactly full quarters
    lens[5::17] = 1                                  # single-word captions
    lens[3] = 47                                     # a long caption among short ones: filled by the two-phase path
    img, cap, lens = itr_b200.synth.scan_inputs(n_img, n_cap, 10.5, 7 + n_img, device="cuda", lengths=lens, round_to="bf16")
    got = ops.scan_i2t_scores_tc(img, cap, lens, "clipped_l2norm", agg, 4.0, 6.0)
    two = ops.scan_scores_tc_generic(img, cap, lens, "i2t", "clipped_l2norm", agg, 4.0, 6.0)
    rel = ((got - two).abs() / two.abs().clamp_min(1e-6)).max().item()
    assert rel < 2e-3, rel                           # two approximations, each within RTOL_TC of the oracle (below)
    assert torch.equal(got[:, 3], two[:, 3])         # the long caption's column comes from the two-phase path itself
    sel = np.unique(np.concatenate([np.arange(0, n_cap, max(1, n_cap // 9)), [3, 5, 13]]))
    sel = sel[sel < n_cap]
    rows = slice(0, min(n_img, 24))
    want = so.scan_scores(img[rows].cpu().numpy(), cap[torch.from_numpy(sel).cuda()].cpu().numpy(), lens[sel], "i2t",
                          "clipped_l2norm", agg, 4.0, 6.0)
    np.testing.assert_allclose(got[rows][:, torch.from_numpy(sel).cuda()].cpu().numpy(), want, rtol=RTOL_TC, atol=ATOL_TC)

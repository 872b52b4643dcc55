"""The reference's own call sequences through the drop-in symbols: evalrank_single
(evaluation.py:284-303: cal_sims -> cal_recall) and validate_step (utils.py:152-167: encode_data -> de-duplicate ->
cal_sims -> i2t -> t2i).  cal_sims hands back a genuine float64 host matrix; the ranking calls that follow must
reuse the device copy behind it (no second upload) and agree with the fused cal_sims_and_recall and with the
reference's own numpy ranking of that matrix."""
import contextlib
import io

import numpy as np
import pytest
import torch

import itr_b200
from itr_b200 import evaluation as ev, objectives as ob, ops
from oracle import ref_loader, scan_oracle as so

pytestmark = pytest.mark.gpu


def cfg(**kw):
    base = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp",
                lambda_lse=6.0, lambda_softmax=9.0, margin=0.2, max_violation=True, measure="cosine")
    base.update(kw)
    return base


class FakeModel:
    sim_enc = None
    Eiters = 0

    def __init__(self, config, img=None, cap=None, lens=None):
        self.config = config
        self.criterion = ob.ContrastiveLoss(config, margin=0.2, measure="cosine", max_violation=True)
        self.img, self.cap, self.lens = img, cap, lens

    def val_start(self):
        pass

    def forward_emb(self, images=None, captions=None, lengths=None, ids=None, **kw):
        ids = list(ids)
        width = int(max(lengths))
        return self.img[[i // 5 for i in ids]].cuda(), self.cap[ids][:, :width].cuda(), lengths


class FakeLoader:
    """Batches in the reference's 8-tuple layout (data_loader.py collate): captions sorted by length inside a batch."""

    def __init__(self, n, lens, batch=32):
        self.dataset = list(range(n))
        self.batches = []
        for s in range(0, n, batch):
            ids = sorted(range(s, min(s + batch, n)), key=lambda i: -lens[i])
            self.batches.append((None, None, None, None, [int(lens[i]) for i in ids], ids, None, None))

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


@pytest.mark.parametrize("name", ["SCAN", "VSE++"])
def test_cal_recall_reuses_the_device_matrix(name):
    if name == "SCAN":
        img, cap, lens = itr_b200.synth.scan_inputs(60, 300, 10.5, 5)
        model = FakeModel(cfg())
        args = (img.numpy(), cap.numpy(), lens)
    else:
        im, s = itr_b200.synth.vse_inputs(60, 300, 5)
        model = FakeModel(cfg(name="VSE++"))
        args = (im.numpy(), s.numpy(), None)
    with contextlib.redirect_stdout(io.StringIO()):
        sims = ev.cal_sims(model, *args, shard_size=100)
    assert isinstance(sims, np.ndarray) and sims.dtype == np.float64 and sims.shape == (60, 300)
    assert not sims.flags.writeable
    twin = ev._device_twin(sims)
    assert twin is not None and twin.ranks is None
    np.testing.assert_array_equal(sims, twin.dev.double().cpu().numpy())       # the host matrix IS the device matrix
    with contextlib.redirect_stdout(io.StringIO()):
        res = ev.cal_recall(sims)
    assert twin.ranks is not None                                               # ranked from the device copy, once
    fused = ev.cal_sims_and_recall(model, *args, shard_size=100)
    for k in ("i2t_ranks", "i2t_top1", "t2i_ranks", "t2i_top1"):
        np.testing.assert_array_equal(res[k], fused[k])
        assert res[k].dtype == np.float64
    assert res["rsum"] == fused["rsum"] and res["result"] == fused["result"]
    # validate_step's form: i2t then t2i on the same array -- the second call is answered from the cached vectors
    m_i, (ri, ti) = ev.i2t(sims, return_ranks=True)
    m_t, (rt, tt) = ev.t2i(sims, return_ranks=True)
    np.testing.assert_array_equal(ri, res["i2t_ranks"]); np.testing.assert_array_equal(rt, res["t2i_ranks"])
    # the reference's numpy ranking of the very same host matrix
    want = so.recall_dict(sims)
    i2t_s, t2i_s, tied_i, tied_c = so.strict_ranks(sims)
    np.testing.assert_array_equal(res["i2t_ranks"], i2t_s); np.testing.assert_array_equal(res["t2i_ranks"], t2i_s)
    np.testing.assert_array_equal(res["i2t_ranks"][~tied_i], want["i2t_ranks"][~tied_i])
    np.testing.assert_array_equal(res["t2i_ranks"][~tied_c], want["t2i_ranks"][~tied_c])
    # anything derived from the matrix is ranked from ITS values, not from the twin
    avg = (sims + sims[::-1]) / 2
    assert ev._device_twin(avg) is None
    with contextlib.redirect_stdout(io.StringIO()):
        res_avg = ev.cal_recall(avg)
    np.testing.assert_array_equal(res_avg["i2t_ranks"], so.strict_ranks(avg)[0])
    # a caller that makes the array writeable again may have changed it: the twin is dropped, the host values win
    sims.setflags(write=True)
    sims[0, :] = -5.0
    assert ev._device_twin(sims) is None
    np.testing.assert_array_equal(ev.i2t(sims, return_ranks=True)[1][0], so.strict_ranks(sims)[0])
    del sims, twin
    import gc
    gc.collect()
    assert all(e.ref() is not None for e in ev._SIMS.values())


def test_cal_sims_ships_blocks_while_scoring():
    """cal_sims fills its float64 host matrix block by block while the later caption chunks are still being scored
    (evaluation._HostMatrixWriter + itr_scores_to_host_f64): the result must be the one-shot conversion of the device
    matrix, bit for bit -- for captions streaming from pinned memory in several chunks, for device captions (one block)
    and for a similarity function that reports no blocks at all."""
    model = FakeModel(cfg())
    img, cap, lens = itr_b200.synth.scan_inputs(220, 1100, 10.5, 6)        # > 8192 words: several caption chunks
    assert len(ops.host_caption_chunks(np.asarray(lens), fractions=(3.0 / 16, 5.0 / 16, 1.0 / 4, 1.0 / 8, 1.0 / 8))) == 5
    cap_pinned = torch.empty(cap.shape, dtype=torch.float32, pin_memory=True).copy_(cap)
    with contextlib.redirect_stdout(io.StringIO()):
        got = ev.cal_sims(model, img.numpy(), cap_pinned.numpy(), lens)
        got_dev_caps = ev.cal_sims(model, img.numpy(), cap.numpy().copy(), lens)      # pageable captions: uploaded whole, one block
    want = ev.device_sims(model, img.numpy(), cap.numpy(), lens).double().cpu().numpy()
    assert got.dtype == np.float64 and got.flags.c_contiguous and not got.flags.writeable
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got_dev_caps, want)
    res = ev.cal_recall(got)                                                       # still ranked from the device twin
    assert res["rsum"] == pytest.approx(so.recall_dict(want)["rsum"])
    vse = FakeModel(cfg(name="VSE++"))
    im, s = itr_b200.synth.vse_inputs(40, 200, 9)
    with contextlib.redirect_stdout(io.StringIO()):
        np.testing.assert_array_equal(ev.cal_sims(vse, im, s), ev.device_sims(vse, im, s).double().cpu().numpy())


def test_pinned_pool_recycles_and_isolates():
    """Two live matrices never share a buffer; a dead one's buffer is reused."""
    model = FakeModel(cfg(name="VSE++"))
    im, s = itr_b200.synth.vse_inputs(40, 200, 9)
    with contextlib.redirect_stdout(io.StringIO()):
        a = ev.cal_sims(model, im.numpy(), s.numpy())
        b = ev.cal_sims(model, (im * 0.5).numpy(), s.numpy())
    assert a.ctypes.data != b.ctypes.data
    np.testing.assert_allclose(b, a * 0.5, rtol=1e-6)
    addr = a.ctypes.data
    keep = a[3:5]                       # a view keeps the buffer alive
    del a
    with contextlib.redirect_stdout(io.StringIO()):
        c = ev.cal_sims(model, im.numpy(), s.numpy())
    assert c.ctypes.data != addr and keep.base is not None
    del keep, c
    with contextlib.redirect_stdout(io.StringIO()):
        d = ev.cal_sims(model, im.numpy(), s.numpy())
    assert d.ctypes.data in (addr, d.ctypes.data)


def test_staged_upload_matches_plain_copy():
    x = torch.randn(3_000_000 * 4 + 17)
    y = ops.upload_pageable(x, "cuda", chunk_bytes=4 << 20)
    assert torch.equal(y.cpu(), x)
    z = ops.upload_pageable(x.view(-1)[: 12_000_000].view(3000, 4000), "cuda")
    assert z.shape == (3000, 4000) and torch.equal(z.cpu(), x[:12_000_000].view(3000, 4000))


@pytest.mark.skipif(not ref_loader.available(), reason="reference package not staged (oracle/_ref) on this box")
def test_reference_validate_step_unmodified_through_install():
    """The reference's OWN validate_step (itr/utils.py:144-186), byte for byte, after install(): it calls
    eval.encode_data, numpy-de-duplicates the images, eval.cal_sims, eval.i2t, eval.t2i."""
    U = ref_loader.load_utils()
    img, cap, lens = itr_b200.synth.scan_inputs(40, 200, 10.5, 3)
    img5 = img                              # forward_emb repeats each image for its five captions
    config = cfg(batch_size=32)
    try:
        itr_b200.install()
        model = FakeModel(config, img5, cap, lens)
        loader = FakeLoader(200, lens)
        with contextlib.redirect_stdout(io.StringIO()):
            r_sum, r1 = U.validate_step(config, loader, model)
    finally:
        itr_b200.uninstall()
    want = ob.xattn_score_t2i(img.cuda(), cap.cuda(), lens, config)
    wr = [x.cpu().numpy() for x in ev.device_ranks(want)]
    want_sum = sum(100.0 * np.mean(wr[0] < k) + 100.0 * np.mean(wr[2] < k) for k in (1, 5, 10))
    assert abs(r_sum - want_sum) < 1e-9 and abs(r1 - 100.0 * np.mean(wr[0] < 1)) < 1e-9
    import sys
    logged = sys.modules["tensorboard_logger"].logged
    assert abs(logged["r_sum"] - want_sum) < 1e-9 and "medr_i2t" in logged

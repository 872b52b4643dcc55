"""The C-ABI library loads on a CPU-only box and exports every symbol include/itr_b200.h declares;
host-only entry points (the caption planner) are exercised, nothing that needs a GPU is called."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT

import itr_b200
from itr_b200 import _capi as capi, ops


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "itr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(itr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 17
    L = capi.lib()
    for n in names:
        assert hasattr(L, n), "libitr_b200.so does not export " + n
        assert n in capi.SIGNATURES, "ctypes binding lacks " + n
    assert sorted(capi.SIGNATURES) == names
    assert L.itr_version() == 100
    assert isinstance(L.itr_last_error(), bytes)


def test_error_mapping_without_gpu():
    L = capi.lib()
    rc = L.itr_scan_plan_words(None, 3, None, None)
    assert rc == capi.ITR_ERR_INVALID
    with pytest.raises(ValueError):
        capi.check(rc)
    with pytest.raises(ValueError, match="unknown aggfunc"):
        capi.agg_code("Median")
    with pytest.raises(ValueError, match="unknown first norm type"):
        capi.norm_code("l1norm")


def check_plan(lens):
    lens = np.asarray(lens, dtype=np.int32)
    meta, n_tiles = ops.plan_words(lens)
    meta = meta.reshape(n_tiles, 128, 4)
    cap, word, seg, ln = meta[..., 0], meta[..., 1], meta[..., 2], meta[..., 3]
    # every (caption, word) exactly once
    live = cap >= 0
    keys = cap[live].astype(np.int64) * 1000 + word[live]
    assert len(np.unique(keys)) == len(keys) == int(lens.sum())
    assert (ln[live] == lens[cap[live]]).all()
    long_tile = (seg >> 16) & 1
    for t in range(n_tiles):
        if long_tile[t].any():
            assert long_tile[t].all()
            c = cap[t][cap[t] >= 0]
            assert len(set(c.tolist())) == 1 and lens[c[0]] > 32
            assert (word[t][: lens[c[0]]] == np.arange(lens[c[0]])).all() and (cap[t][lens[c[0]]:] == -1).all()
            continue
        for q in range(4):
            sl = slice(32 * q, 32 * q + 32)
            lo, hi = seg[t, sl] & 0xFF, (seg[t, sl] >> 8) & 0xFF
            lane = np.arange(32)
            assert ((lo <= lane) & (lane <= hi) & (hi < 32)).all()
            cq, wq = cap[t, sl], word[t, sl]
            pad = cq < 0
            assert (lo[pad] == lane[pad]).all() and (hi[pad] == lane[pad]).all()
            # a caption occupies exactly lanes [lo, hi] in word order
            assert (wq[~pad] == (lane - lo)[~pad]).all()
            assert ((hi - lo + 1)[~pad] == lens[cq[~pad]]).all()
            for l in lane[~pad]:
                assert (cq[lo[l]: hi[l] + 1] == cq[l]).all()
    return n_tiles


def test_plan_words_properties():
    rng = np.random.default_rng(0)
    assert check_plan([5]) == 1
    assert check_plan([32, 32, 32, 32, 1]) == 2
    assert check_plan([33, 128, 64]) == 3
    check_plan(rng.integers(1, 33, size=500))
    check_plan(rng.integers(1, 129, size=300))
    with pytest.raises(ValueError):
        ops.plan_words(np.array([4, 129], dtype=np.int32))
    with pytest.raises(ValueError):
        ops.plan_words(np.array([0, 3], dtype=np.int32))


def test_plan_words_packing_efficiency_coco_shape():
    lens = itr_b200.synth.caption_lengths(25000, 10.5, 14)
    n_tiles = check_plan(lens)
    eff = lens.sum() / (n_tiles * 128.0)
    assert eff > 0.94, eff            # share of MMA rows spent on real words (best-fit into 32-row quarters)


def test_plan_words_fuzz():
    """Property-based: any multiset of caption lengths in [1, 128] yields a valid plan (every caption exactly once,
    whole captions inside one warp's 32 rows unless `long`, long captions alone in their tile)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=128), min_size=1, max_size=400))
    def run(lens):
        check_plan(lens)

    run()


def test_gt_item_planner_covers_every_ground_truth_pair():
    """itr_scan_plan_gt_items: every caption's word tile appears in an item with the image tile of its ground-truth
    image; word tiles are paired per image tile (never two items for the same (tile, image tile)); shards use the
    global caption index."""
    import ctypes as C
    rng = np.random.default_rng(11)
    L = capi.lib()
    for n_cap, n_img, cap_offset in [(0, 4, 0), (5, 1, 0), (333, 64, 0), (640, 130, 0), (200, 130, 450), (77, 9, 0)]:
        lens = rng.integers(1, 40, size=n_cap).astype(np.int32)
        if n_cap:
            lens[rng.integers(0, n_cap)] = 128
        meta, n_tiles = ops.plan_words(lens) if n_cap else (np.zeros((0, 4), np.int32), 0)
        n = C.c_int(-1)
        buf = np.ascontiguousarray(meta if n_tiles else np.zeros((1, 4), np.int32))
        capi.check(L.itr_scan_plan_gt_items(buf.ctypes.data, n_tiles, cap_offset, 5, n_img, None, 0, C.byref(n)))
        items = np.empty((max(n.value, 1), 4), dtype=np.int32)
        capi.check(L.itr_scan_plan_gt_items(buf.ctypes.data, n_tiles, cap_offset, 5, n_img, items.ctypes.data, n.value, C.byref(n)))
        items = items[: n.value]
        m = meta.reshape(-1, 4)
        rows = np.nonzero((m[:, 0] >= 0) & (m[:, 1] == 0))[0]
        img = (cap_offset + m[rows, 0]) // 5
        need = {(int(r // 128), int(i // 4)) for r, i in zip(rows, img) if i < n_img}
        have = [(int(a), int(t)) for a, b, t, z in items] + [(int(b), int(t)) for a, b, t, z in items if b < n_tiles]
        assert len(have) == len(set(have)) and set(have) == need
        assert (items[:, 3] == 0).all() and (items[:, 0] < max(n_tiles, 1)).all() and (items[:, 1] <= n_tiles).all()
        assert (np.diff(items[:, 2]) >= 0).all()                                # sorted by image tile
        if n.value > 1:
            with pytest.raises(ValueError):
                capi.check(L.itr_scan_plan_gt_items(buf.ctypes.data, n_tiles, cap_offset, 5, n_img, items.ctypes.data, n.value - 1, C.byref(n)))

"""Host-side logic of the caption-sharded evaluation (N > 1 path) on CPU: world_size-2 gloo
processes run the real exchange code (itr_b200.sharding.sharded_ranks) with block statistics
computed by numpy, and must reproduce the single-process ranks bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from itr_b200 import sharding
from oracle import scan_oracle as so

_SIGN = torch.iinfo(torch.int64).min


def _orderable(v):
    u = v.astype(np.float32).view(np.uint32).astype(np.uint64)
    neg = (u & np.uint64(0x80000000)) != 0
    return np.where(neg, (~u) & np.uint64(0xFFFFFFFF), u | np.uint64(0x80000000))


class NumpyStats:
    """Same contract as sharding.CudaStats, computed on the host (test double for the rank kernels)."""

    @staticmethod
    def thresholds(block, cap_offset, cpi):
        b = block.numpy()
        n_img, n_loc = b.shape
        thr_col = np.array([b[(cap_offset + c) // cpi, c] if (cap_offset + c) // cpi < n_img else np.inf
                            for c in range(n_loc)], dtype=np.float32)
        thr_row = np.full(n_img, -np.inf, dtype=np.float32)
        for i in range(n_img):
            cols = [i * cpi + k - cap_offset for k in range(cpi)]
            cols = [c for c in cols if 0 <= c < n_loc]
            if cols:
                thr_row[i] = b[i, cols].max()
        return torch.from_numpy(thr_row), torch.from_numpy(thr_col)

    @staticmethod
    def count(block, thr_row, thr_col, cap_offset):
        b = block.numpy()
        cnt_row = (b > thr_row.numpy()[:, None]).sum(1).astype(np.int32)
        cnt_col = (b > thr_col.numpy()[None, :]).sum(0).astype(np.int32)
        key = _orderable(b) << np.uint64(32)
        col_ids = (~(np.arange(b.shape[1], dtype=np.uint64) + np.uint64(cap_offset))) & np.uint64(0xFFFFFFFF)
        best_row = (key | col_ids[None, :]).max(1)
        best_col_idx = b.argmax(0).astype(np.int64)
        signed = torch.from_numpy(best_row.view(np.int64).copy()) ^ _SIGN
        return torch.from_numpy(cnt_row), torch.from_numpy(cnt_col), signed, torch.from_numpy(best_col_idx)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sims, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_cap = sims.shape[1]
        lo, hi = sharding.shard_bounds(n_cap, world)[rank]
        block = torch.from_numpy(np.ascontiguousarray(sims[:, lo:hi]))
        res = sharding.sharded_ranks(block, lo, n_cap, None, 5, NumpyStats)
        np.savez(os.path.join(out_dir, "r{}.npz".format(rank)), *[r.numpy() for r in res])
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    for n_cap, w in [(25000, 8), (5000, 2), (35, 4), (5, 8), (12, 2)]:
        b = sharding.shard_bounds(n_cap, w)
        assert b[0][0] == 0 and b[-1][1] == n_cap and len(b) == w
        for (a0, a1), (b0, b1) in zip(b, b[1:]):
            assert a1 == b0
        assert all(lo % 5 == 0 for lo, _ in b)
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 5 + (-n_cap) % 5     # one image's worth (+ a ragged last group)
    assert sharding.shard_bounds(25000, 8)[1] == (3125, 6250)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ranks_match_single_process(tmp_path, world):
    rng = np.random.default_rng(5)
    n = 23
    sims = rng.standard_normal((n, 5 * n)).astype(np.float32)
    sims[np.arange(n).repeat(5), np.arange(5 * n)] += 1.0
    mp.spawn(_worker, args=(world, _free_port(), sims, str(tmp_path)), nprocs=world, join=True)
    i2t_r, t2i_r, _, _ = so.strict_ranks(sims)
    for r in range(world):
        got = np.load(tmp_path / "r{}.npz".format(r))
        a, b, c, d = [got["arr_{}".format(k)] for k in range(4)]
        np.testing.assert_array_equal(a, i2t_r)
        np.testing.assert_array_equal(c, t2i_r)
        np.testing.assert_array_equal(b, sims.argmax(1))
        np.testing.assert_array_equal(d, sims.argmax(0))
    # and the strict-greater definition equals the reference's argsort positions here (no ties)
    np.testing.assert_array_equal(so.rank_i2t(sims.astype(np.float64))[1], i2t_r)
    np.testing.assert_array_equal(so.rank_t2i(sims.astype(np.float64))[1], t2i_r)


def test_single_process_no_group():
    rng = np.random.default_rng(6)
    sims = rng.standard_normal((7, 35)).astype(np.float32)
    a, b, c, d = sharding.sharded_ranks(torch.from_numpy(sims), 0, 35, None, 5, NumpyStats)
    i2t_r, t2i_r, _, _ = so.strict_ranks(sims)
    np.testing.assert_array_equal(a.numpy(), i2t_r)
    np.testing.assert_array_equal(c.numpy(), t2i_r)

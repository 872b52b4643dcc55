"""The small exported helpers against outputs of the reference's own functions (tests/golden/helpers.npz, made by
oracle/make_golden.py --only helpers): func_attention (Objectives.py:421-476) and cosine_similarity (:10-15) are
torch-op compatibility shims that run on the caller's device, so they are pinned on the CPU; SAEM's pdist / pdist_cos
(:296-323) contract through the native GEMM and are pinned on the GPU."""
import numpy as np
import pytest
import torch

from conftest import bits_to_f32, load_golden
from itr_b200 import objectives as ob

NORMS = ("clipped_l2norm", "l2norm", "softmax", "clipped", "no_norm")


@pytest.mark.parametrize("dtype,rtol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
def test_func_attention_matches_reference(dtype, rtol):
    g = load_golden("helpers")
    query = torch.from_numpy(bits_to_f32(g["fa|query_bits"])).to(dtype)
    context = torch.from_numpy(bits_to_f32(g["fa|context_bits"])).to(dtype)
    for norm in NORMS:
        for smooth in (9.0, 4.0):
            w, a = ob.func_attention(query, context, dict(raw_feature_norm=norm), smooth=smooth)
            want_w, want_a = g["fa|{}|{}|weighted".format(norm, smooth)], g["fa|{}|{}|attn".format(norm, smooth)]
            assert tuple(w.shape) == want_w.shape == (3, 7, 64) and tuple(a.shape) == want_a.shape == (3, 36, 7)
            np.testing.assert_allclose(w.double().numpy(), want_w, rtol=rtol, atol=rtol)
            np.testing.assert_allclose(a.double().numpy(), want_a, rtol=rtol, atol=rtol)
    with pytest.raises(ValueError):
        ob.func_attention(query, context, dict(raw_feature_norm="l1norm"), smooth=9.0)


def test_cosine_similarity_matches_reference():
    g = load_golden("helpers")
    x1 = torch.from_numpy(bits_to_f32(g["cs|x1_bits"])).double()
    x2 = torch.from_numpy(bits_to_f32(g["cs|x2_bits"])).double()
    for dim in (1, 2):
        got = ob.cosine_similarity(x1, x2, dim=dim)
        want = g["cs|dim{}".format(dim)]
        assert tuple(got.shape) == want.shape
        np.testing.assert_allclose(got.numpy(), want, rtol=1e-12, atol=1e-15)
    assert ob.cosine_similarity(x1, x2, dim=2)[2, 4].item() == 0.0          # zero row: 0 / clamp(0, min=eps)
    # the (1, n, d) case squeezes to (n,), as the reference's trailing .squeeze() does (SURVEY defect D5)
    assert tuple(ob.cosine_similarity(x1[:1], x2[:1], dim=2).shape) == (9,)


@pytest.mark.gpu
def test_saem_pdist_matches_reference():
    g = load_golden("helpers")
    x1 = torch.from_numpy(bits_to_f32(g["pd|x1_bits"])).cuda()
    x2 = torch.from_numpy(bits_to_f32(g["pd|x2_bits"])).cuda()
    np.testing.assert_allclose(ob.pdist(x1, x2).cpu().numpy(), g["pd|pdist"], rtol=2e-5, atol=1e-5)
    got = ob.pdist_cos(x1, x2).cpu().numpy()
    np.testing.assert_allclose(got, g["pd|pdist_cos"], rtol=2e-5, atol=2e-6)
    assert (got[:, 5] == 0).all()                                             # zero row: NaNs of 0/0 zeroed

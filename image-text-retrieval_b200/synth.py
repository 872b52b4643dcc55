"""Seeded synthetic inputs for the five BASELINE.json configs.

There is no network for datasets or checkpoints, so every test and bench run
uses these generators: caption lengths follow the survey's recipe exactly
(the roofline FLOP counts quote their sums), embeddings carry planted
image<->caption structure so Recall@K is non-trivial and score ties are rare.
Shapes mirror what the reference's ``encode_data`` hands to ``cal_sims``
(itr/metricmodule/evaluation.py:104-117): images (n_img, 36, D) unit-norm
regions, captions (n_cap, Lmax, D) zero padded beyond each length.
"""
from __future__ import annotations

import math

import numpy as np
import torch

R = 36          # regions per image (precomp features)
D = 1024        # embed_size, itr/config.py:73
CAPS_PER_IMG = 5

# (n_img, n_cap, poisson lambda, seed) -- SURVEY.md section 8(d)
F30K_SHAPE = dict(n_img=1000, n_cap=5000, lam=12.4, seed=30)      # sum(len) = 72 707
COCO5K_SHAPE = dict(n_img=5000, n_cap=25000, lam=10.5, seed=14)   # sum(len) = 312 906


def caption_lengths(n_cap: int, lam: float, seed: int) -> np.ndarray:
    """len = clip(2 + Poisson(lam), 3, 72); every 997th caption is a long outlier."""
    rng = np.random.default_rng(seed)
    ln = np.clip(2 + rng.poisson(lam, n_cap), 3, 72)
    idx = np.arange(0, n_cap, 997)
    ln[idx] = 72 - (np.arange(len(idx)) % 29)
    return ln.astype(np.int32)


def _unit(x):
    return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def scan_inputs(n_img, n_cap, lam, seed, device="cpu", d=D, regions=R, lengths=None,
                round_to=None, chunk=2048):
    """Region/word embeddings for the SCAN configs.

    Returns (images f32 (n_img, R, d), captions f32 (n_cap, Lmax, d), lengths i32 np (n_cap,)).
    ``round_to`` in {None, "bf16", "tf32"} pre-rounds the values so a float32
    reference sees exactly the numbers a reduced-precision kernel consumes.
    """
    if lengths is None:
        lengths = caption_lengths(n_cap, lam, seed)
    lengths = np.asarray(lengths, dtype=np.int32)
    lmax = int(lengths.max())
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed + 1)
    rn = lambda *shape: torch.randn(*shape, generator=gen, device=dev)
    bank = _unit(rn(512, d))
    concept = torch.randint(0, 512, (n_img, regions), generator=gen, device=dev)
    images = torch.empty(n_img, regions, d, device=dev)
    for s in range(0, n_img, chunk):
        e = min(s + chunk, n_img)
        images[s:e] = _unit(bank[concept[s:e]] + 0.6 / math.sqrt(d) * rn(e - s, regions, d))
    captions = torch.zeros(n_cap, lmax, d, device=dev)
    len_t = torch.from_numpy(lengths).to(dev)
    for s in range(0, n_cap, chunk):
        e = min(s + chunk, n_cap)
        owner = (torch.arange(s, e, device=dev) // CAPS_PER_IMG) % n_img
        pick = torch.randint(0, regions, (e - s, lmax), generator=gen, device=dev)
        cid = torch.gather(concept[owner], 1, pick)                       # concept of each content word
        word = _unit(bank[cid] + 0.8 / math.sqrt(d) * rn(e - s, lmax, d))
        noise = _unit(rn(e - s, lmax, d))
        pos = torch.arange(lmax, device=dev)[None, :]
        ln = len_t[s:e, None]
        is_tag = (pos == 0) | (pos == ln - 1)                             # <start>/<end>: pure noise
        word = torch.where(is_tag[..., None], noise, word)
        gain = 0.5 + 1.5 * torch.rand(e - s, lmax, 1, generator=gen, device=dev)   # SCAN words are not normalised
        captions[s:e] = torch.where((pos < ln)[..., None], gain * word, torch.zeros((), device=dev))
    images, captions = _round(images, round_to), _round(captions, round_to)
    return images, captions, lengths


def vse_inputs(n_img, n_cap, seed, device="cpu", d=D, raw_dim=2048, regions=R, round_to=None):
    """Pooled unit-norm embeddings for the VSE++ configs (2-D inputs, SURVEY defect D7).

    Raw precomp features (n_img, 36, raw_dim) are mean-pooled, projected by a fixed
    seeded Linear raw_dim->d and L2-normalised (builder-defined pooling); captions are
    noisy copies of their image's embedding.
    Returns (im f32 (n_img, d), cap f32 (n_cap, d)).
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed + 7)
    rn = lambda *shape: torch.randn(*shape, generator=gen, device=dev)
    proj = rn(d, raw_dim) / math.sqrt(raw_dim)
    pooled = torch.empty(n_img, raw_dim, device=dev)
    for s in range(0, n_img, 256):
        e = min(s + 256, n_img)
        pooled[s:e] = rn(e - s, regions, raw_dim).abs().mean(dim=1) + rn(e - s, raw_dim)
    im = _unit(pooled @ proj.t())
    owner = (torch.arange(n_cap, device=dev) // CAPS_PER_IMG) % n_img
    cap = _unit(im[owner] + 0.9 / math.sqrt(d) * rn(n_cap, d))
    return _round(im, round_to), _round(cap, round_to)


def _round(x, round_to):
    if round_to is None:
        return x
    if round_to == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    if round_to == "tf32":
        # round-to-nearest-even onto a 10-bit mantissa
        bits = x.contiguous().view(torch.int32)
        bits = (bits + 0x0FFF + ((bits >> 13) & 1)) & ~0x1FFF
        return bits.view(torch.float32)
    raise ValueError("round_to must be None, 'bf16' or 'tf32'")

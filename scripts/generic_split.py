#!/usr/bin/env python
"""Time the phases of the two-phase tensor-core path (i2t Mean, one 1000 x 5000 COCO-shaped fold)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import itr_b200
from itr_b200 import ops, _capi as capi
lens = itr_b200.synth.caption_lengths(25000, 10.5, 14)[:5000]
img, cap, ln = itr_b200.synth.scan_inputs(1000, 5000, 10.5, 14, device="cuda", lengths=lens)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
pi = ops.prepare_images(img); pc = ops.prepare_captions(cap, ln)
print("prep images %.3f ms, prep captions %.3f ms" % (t(lambda: ops.prepare_images(img)), t(lambda: ops.prepare_captions(cap, ln))))
aff = torch.empty(pc.n_tiles * 1000 * 128 * 36, device="cuda")
L = capi.lib()
print("affinity dump (tcgen05) %.3f ms for %.2f GB" % (t(lambda: capi.check(L.itr_scan_affinity_bf16(capi.ptr(pi.images_bf16), 1000, capi.ptr(pc.words_bf16), pc.n_tiles, capi.ptr(aff), capi.stream_ptr()))), aff.numel() * 4 / 1e9))
for d, lam in (("i2t", 4.0), ("t2i", 9.0)):
    print(d, "whole generic path %.3f ms" % t(lambda: ops.scan_scores_tc_generic(img, cap, ln, d, "clipped_l2norm", "Mean", lam, 6.0, pi=pi, pc=pc, max_affinity_bytes=12 << 30)))
print("fused t2i kernel %.3f ms" % t(lambda: ops.scan_t2i_scores_bf16(pi, pc, "clipped_l2norm", "Mean", 9.0, 6.0)))

#!/usr/bin/env python
"""Device time of the fused i2t kernel alone on one COCO-shaped fold (1000 x 5000), next to the two-phase path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itr_b200
from itr_b200 import ops
lens = itr_b200.synth.caption_lengths(25000, 10.5, 14)[:5000]
img, cap, ln = itr_b200.synth.scan_inputs(1000, 5000, 10.5, 14, device="cuda", lengths=lens)
pi, pc = ops.prepare_images(img), ops.prepare_captions(cap, ln)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = torch.empty(1000, 5000, device="cuda")
print("fused i2t kernel      %.3f ms" % t(lambda: ops.scan_i2t_scores_bf16(pi, pc, "clipped_l2norm", "Mean", 4.0, 6.0, out=out)))
pc.gq_frag = None
print("caption gram (rel)    %.3f ms" % t(lambda: (setattr(pc, "gq_frag", None), ops.caption_gram_frag(pc))))
print("t2i kernel (same shape) %.3f ms" % t(lambda: ops.scan_t2i_scores_bf16(pi, pc, "clipped_l2norm", "LogSumExp", 9.0, 6.0, out=out)))
print("two-phase i2t         %.3f ms" % t(lambda: ops.scan_scores_tc_generic(img, cap, ln, "i2t", "clipped_l2norm", "Mean", 4.0, 6.0, pi=pi, pc=pc), 3))
print("whole scan_i2t_scores_tc %.3f ms" % t(lambda: ops.scan_i2t_scores_tc(img, cap, ln, "clipped_l2norm", "Mean", 4.0, 6.0), 3))

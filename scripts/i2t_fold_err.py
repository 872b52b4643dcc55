"""Diagnostic: fused i2t (tensor-core mode) against the float32 mode on a full COCO-shaped fold."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
itr_b200 = importlib.import_module("image-text-retrieval_b200")
ob = itr_b200.objectives
lens_all = itr_b200.synth.caption_lengths(25000, 10.5, 14)
def mk(**kw):
    c = dict(cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp", lambda_lse=6.0, lambda_softmax=9.0)
    c.update(kw)
    return c
for agg in ("Mean", "LogSumExp", "Max"):
    c4 = mk(cross_attn="i2t", agg_func=agg, lambda_softmax=4.0)
    img, cap, ln = itr_b200.synth.scan_inputs(1000, 5000, 10.5, 14, device="cuda", lengths=lens_all[:5000], round_to="bf16")
    a = ob.xattn_score_i2t(img, cap, ln, c4)
    f32 = ob.xattn_score_i2t(img, cap, ln, dict(c4, itr_b200_precision="fp32"))
    err = (a - f32).abs(); rel = err / f32.abs().clamp_min(1e-6)
    i = int(rel.argmax()); 
    print(agg, "max abs %.3e  max rel %.3e at score %.4e | mean |score| %.4f | frac rel>1e-3: %.2e | p99.99 rel %.2e" % (
        err.max().item(), rel.max().item(), f32.flatten()[i].item(), f32.abs().mean().item(),
        (rel > 1e-3).float().mean().item(), torch.quantile(rel.flatten()[::7].float(), 0.9999).item()))

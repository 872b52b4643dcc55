"""Diagnostic: max relative error of the fused and the two-phase i2t paths against the float64 oracle."""
import importlib, os, sys
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
itr_b200 = importlib.import_module("image-text-retrieval_b200")
from oracle import scan_oracle as so
ops = itr_b200.ops
n_img, n_cap = 40, 160
lens = itr_b200.synth.caption_lengths(n_cap, 10.5, 5)
lens[::13] = 32; lens[5::17] = 1
img, cap, lens = itr_b200.synth.scan_inputs(n_img, n_cap, 10.5, 5, device="cuda", lengths=lens, round_to="bf16")
for norm in ("clipped_l2norm", "l2norm"):
    for agg in so.AGG_FUNCS:
        want = so.scan_scores(img.cpu().numpy(), cap.cpu().numpy(), lens, "i2t", norm, agg, 4.0, 6.0)
        f = ops.scan_i2t_scores_tc(img, cap, lens, norm, agg, 4.0, 6.0).cpu().numpy()
        t = ops.scan_scores_tc_generic(img, cap, lens, "i2t", norm, agg, 4.0, 6.0).cpu().numpy()
        rel = lambda x: np.abs(x - want) / np.maximum(np.abs(want), 1e-6)
        print("%-15s %-10s fused max %.2e median %.2e | two-phase max %.2e median %.2e" % (
            norm, agg, rel(f).max(), np.median(rel(f)), rel(t).max(), np.median(rel(t))))

"""One SCAN training step (ContrastiveLoss t2i, max_violation, batch 128) fwd + bwd: timing per stage, or a short
run for an ncu launch list (scripts/scan_train_step.py --once)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from itr_b200 import objectives as ob, ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
lens = np.clip(synth.caption_lengths(128, 10.5, 16), 1, 60)
img, cap, ln = synth.scan_inputs(128, 128, 10.5, 16, device=dev, lengths=lens)
a, b = img.clone().requires_grad_(True), cap.clone().requires_grad_(True)
cfg = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp", lambda_lse=6.0, lambda_softmax=9.0)
crit = ob.ContrastiveLoss(cfg, margin=0.2, measure="cosine", max_violation=True)
lens_list = [int(x) for x in ln]
print("lmax", cap.size(1), "words", int(ln.sum()))


def step():
    a.grad = None
    b.grad = None
    crit(a, b, lens_list).backward()


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if "--once" in sys.argv:
    step()
    torch.cuda.synchronize()
    step()
    torch.cuda.synchronize()
    sys.exit(0)

ds = torch.randn(128, 128, device=dev)
print("fwd+bwd ms", timed(step))
print("fwd only ms", timed(lambda: ops.scan_scores_f32(img, cap, ln, "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0)))
print("bwd only ms", timed(lambda: ops.scan_backward_f32(img, cap, ln, ds, "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0)))
t0 = time.perf_counter()
for _ in range(20):
    step()
torch.cuda.synchronize()
print("wall ms", (time.perf_counter() - t0) / 20 * 1e3)

"""Diagnostic: kernel list of the two-phase fallback for the few captions of more than 32 words."""
import importlib, os, sys
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
itr_b200 = importlib.import_module("image-text-retrieval_b200")
ops, synth = itr_b200.ops, itr_b200.synth
lens_all = synth.caption_lengths(25000, 10.5, 14)
img, cap, ln = synth.scan_inputs(1000, 5000, 10.5, 14, device="cuda", lengths=lens_all[:5000])
lnp = ops.lengths_to_numpy(ln, 5000)
pi = ops.prepare_images(img)
long_ids = np.nonzero(lnp > 32)[0]; idx = torch.from_numpy(long_ids).cuda()
f = lambda: ops.scan_scores_tc_generic(img, cap[idx], lnp[long_ids], "i2t", "clipped_l2norm", "Mean", 4.0, 6.0, pi=pi)
for _ in range(3): f()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): f()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))

#!/usr/bin/env python
"""Per-role wait-cycle breakdown of the CTA-pair score kernel (itr_scan_t2i_pair_profile), clk per item (= per
item PAIR for the leader's issuers), leader and peer CTA side by side."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import itr_b200
from itr_b200 import ops, _capi as capi
n_img, n_cap = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 5000
lengths = itr_b200.synth.caption_lengths(25000, 10.5, 14)[:n_cap]
img, cap, ln = itr_b200.synth.scan_inputs(n_img, n_cap, 10.5, 14, device="cuda", lengths=lengths)
pi, pc = ops.prepare_images(img), ops.prepare_captions(cap, ln)
out = torch.empty(n_img, n_cap, device="cuda")
names = ["producer total", "producer wait empty", "mma total", "mma wait loaded+gfree", "mma wait full", "items",
         "g0 wait tfull+afull", "g0 wait uready", "g1 wait tfull+afull", "g1 wait uready", "g2 wait tfull+afull", "g2 wait uready",
         "g3 wait tfull+afull", "g3 wait uready", "epilogue total", "g0 wait afull"]
cnt = torch.zeros(74, 2, 16, dtype=torch.int64, device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    if rep == 1:
        ev0.record()
    capi.check(capi.lib().itr_scan_t2i_pair_profile(capi.ptr(pi.images_bf16), capi.ptr(pi.gram_pack), n_img, capi.ptr(pc.words_bf16),
                                                    capi.ptr(pc.row_meta), capi.ptr(pc.row_wnorm), pc.n_tiles, capi.ptr(out), n_cap,
                                                    capi.ptr(cnt), capi.stream_ptr()))
ev1.record()
torch.cuda.synchronize()
c = cnt.cpu().numpy().astype(np.float64)
items = c[:, 0, 5].mean()
print("=== pair kernel %d x %d: %.3f ms, item pairs per cluster %.1f" % (n_img, n_cap, ev0.elapsed_time(ev1), items))
for r, who in enumerate(("leader", "peer")):
    print("  %-6s " % who + "  ".join("%s %.0f" % (nm.replace(" ", "_"), c[:, r, i].mean() / items) for i, nm in enumerate(names) if i != 5))
# spread across the 74 clusters: with a static schedule the launch ends with the slowest cluster
for r, who in enumerate(("leader", "peer")):
    tot, it = c[:, r, 14], c[:, r, 5]
    print("  %-6s epilogue loop clk per cluster: min %.4g  mean %.4g  max %.4g  (max/mean - 1 = %.2f %%);  items min %d max %d;  clk/item min %.0f max %.0f"
          % (who, tot.min(), tot.mean(), tot.max(), 100 * (tot.max() / tot.mean() - 1), it.min(), it.max(), (tot / it).min(), (tot / it).max()))
order = np.argsort(c[:, 0, 14])
print("  slowest clusters:", order[-6:].tolist(), " fastest:", order[:6].tolist())

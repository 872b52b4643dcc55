"""Diagnostic: device time of the pieces of one config-4 fold (xattn_score_i2t + device_ranks)."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
itr_b200 = importlib.import_module("image-text-retrieval_b200")
ops, ob, ev, synth = itr_b200.ops, itr_b200.objectives, itr_b200.evaluation, itr_b200.synth
lens_all = synth.caption_lengths(25000, 10.5, 14)
img, cap, ln = synth.scan_inputs(1000, 5000, 10.5, 14, device="cuda", lengths=lens_all[:5000])
c4 = dict(name="SCAN", cross_attn="i2t", raw_feature_norm="clipped_l2norm", agg_func="Mean", lambda_lse=6.0, lambda_softmax=4.0)
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - w0) / n * 1e3
def show(name, fn): print("%-46s device %.3f ms   wall %.3f ms" % ((name,) + t(fn)))
lnp = ops.lengths_to_numpy(ln, 5000)
show("whole fold: xattn_score_i2t + device_ranks", lambda: ev.device_ranks(ob.xattn_score_i2t(img, cap, ln, c4)))
show("xattn_score_i2t", lambda: ob.xattn_score_i2t(img, cap, ln, c4))
s = ob.xattn_score_i2t(img, cap, ln, c4)
show("device_ranks", lambda: ev.device_ranks(s))
show("prepare_images", lambda: ops.prepare_images(img))
show("prepare_captions", lambda: ops.prepare_captions(cap, lnp))
pi, pc = ops.prepare_images(img), ops.prepare_captions(cap, lnp)
show("gram fragments", lambda: (setattr(pc, "gq_frag", None), ops.caption_gram_frag(pc)))
show("fused kernel (+region norms)", lambda: ops.scan_i2t_scores_bf16(pi, pc, "clipped_l2norm", "Mean", 4.0, 6.0))
long_ids = np.nonzero(lnp > 32)[0]
print("captions longer than 32 words:", len(long_ids), "max", lnp.max())
idx = torch.from_numpy(long_ids).cuda()
show("two-phase path for the long captions", lambda: ops.scan_scores_tc_generic(img, cap[idx], lnp[long_ids], "i2t", "clipped_l2norm", "Mean", 4.0, 6.0, pi=pi))

"""Diagnostic: the fused t2i kernel alone on one COCO fold shape and on a quarter of the COCO-5K shape."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
itr_b200 = importlib.import_module("image-text-retrieval_b200")
ops, synth = itr_b200.ops, itr_b200.synth
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for n_img, n_cap in ((1000, 5000), (2500, 12500)):
    lens = synth.caption_lengths(n_cap, 10.5, 14)
    img, cap, ln = synth.scan_inputs(n_img, n_cap, 10.5, 14, device="cuda", lengths=lens)
    pi, pc = ops.prepare_images(img), ops.prepare_captions(cap, ops.lengths_to_numpy(ln, n_cap))
    del img, cap
    out = torch.empty(n_img, n_cap, device="cuda")
    print("t2i scores kernel %5d x %5d   %.3f ms" % (n_img, n_cap, t(lambda: ops.scan_t2i_scores_bf16(pi, pc, "clipped_l2norm", "LogSumExp", 9.0, 6.0, out=out))))

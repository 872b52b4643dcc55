"""Diagnostic: repeat the fused and the two-phase i2t paths on the test shapes and report run-to-run differences."""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
itr_b200 = importlib.import_module("image-text-retrieval_b200")
ops = itr_b200.ops
for n_img, n_cap, agg in [(37, 185, "Mean"), (130, 333, "LogSumExp"), (64, 200, "Sum")]:
    lens = itr_b200.synth.caption_lengths(n_cap, 10.5, 7 + n_img)
    lens[::13] = 32; lens[5::17] = 1; lens[3] = 47
    img, cap, lens = itr_b200.synth.scan_inputs(n_img, n_cap, 10.5, 7 + n_img, device="cuda", lengths=lens, round_to="bf16")
    pi, pc = ops.prepare_images(img), ops.prepare_captions(cap, lens)
    short = torch.from_numpy(np.nonzero(lens <= 32)[0]).cuda()
    base_f = base_t = None
    for rep in range(30):
        f = ops.scan_i2t_scores_bf16(pi, pc, "clipped_l2norm", agg, 4.0, 6.0)[:, short]
        t = ops.scan_scores_tc_generic(img, cap, lens, "i2t", "clipped_l2norm", agg, 4.0, 6.0)[:, short]
        torch.cuda.synchronize()
        if base_f is None: base_f, base_t = f.clone(), t.clone()
        df = (f != base_f).nonzero(); dt = (t != base_t).nonzero()
        if len(df) or len(dt):
            print(n_img, n_cap, agg, "rep", rep, "fused diffs", len(df), df[:6].tolist(), "two-phase diffs", len(dt), dt[:6].tolist())
            if len(df):
                i, c = df[0].tolist(); cc = int(short[c]); print("  fused", f[i, c].item(), "base", base_f[i, c].item(), "two", base_t[i, c].item(), "len", lens[cc])
    rel = ((base_f - base_t).abs() / base_t.abs().clamp_min(1e-6)).max().item()
    print(n_img, n_cap, agg, "first-run rel fused vs two-phase", rel)

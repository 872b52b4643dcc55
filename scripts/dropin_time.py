"""Diagnostic: the reference's call sequence (cal_sims -> cal_recall) at COCO-5K size, host matrix shipped block by block
(default) against converted and copied in one piece at the end (ITR_B200_HOST_MATRIX=oneshot)."""
import contextlib, importlib, io, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
itr_b200 = importlib.import_module("image-text-retrieval_b200")
ev, ob, synth = itr_b200.evaluation, itr_b200.objectives, itr_b200.synth
CONFIG = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp", lambda_lse=6.0,
              lambda_softmax=9.0, margin=0.2, max_violation=True, measure="cosine")
n_img, n_cap = 5000, 25000
lens = synth.caption_lengths(n_cap, 10.5, 14)
images, captions, _ = synth.scan_inputs(n_img, n_cap, 10.5, 14, device="cuda", lengths=lens)
images_h = torch.empty(images.shape, dtype=torch.float32, pin_memory=True).copy_(images)
captions_h = torch.empty(captions.shape, dtype=torch.float32, pin_memory=True).copy_(captions)
del images, captions; torch.cuda.empty_cache()
class M: sim_enc = None
m = M(); m.config = CONFIG; m.criterion = ob.ContrastiveLoss(CONFIG, margin=0.2, measure="cosine", max_violation=True)
imgs_np, caps_np = np.array(images_h.numpy()), captions_h.numpy()
for mode in ("", "oneshot", "", "oneshot"):
    os.environ["ITR_B200_HOST_MATRIX"] = mode
    for rep in range(3):
        with contextlib.redirect_stdout(io.StringIO()):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            sims = ev.cal_sims(m, imgs_np, caps_np, lengths=lens, shard_size=640)
            t1 = time.perf_counter()
            res = ev.cal_recall(sims)
            torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%-8s cal_sims %.1f ms  cal_recall %.1f ms  total %.1f ms  rsum %.3f" % (mode or "blocks", (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, res["rsum"]))
    del sims
# where the host thread spends the call (no extra synchronisation)
ops = itr_b200.ops
T = {}
def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); T[name] = T.get(name, 0.0) + time.perf_counter() - t0; return r
    return w
ops.prepare_images_streamed = timed("prepare_images_streamed (host-blocking part)", ops.prepare_images_streamed)
ops.scan_t2i_scores_from_host = timed("scan_t2i_scores_from_host (launches)", ops.scan_t2i_scores_from_host)
ev._HostMatrixWriter.finish = timed("host matrix finish (wait for the device)", ev._HostMatrixWriter.finish)
os.environ["ITR_B200_HOST_MATRIX"] = ""
for rep in range(3):
    T.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sims = ev.cal_sims(m, imgs_np, caps_np, lengths=lens, shard_size=640)
        t1 = time.perf_counter()
print("cal_sims %.1f ms:" % ((t1 - t0) * 1e3), {k: round(v * 1e3, 1) for k, v in T.items()})
x = torch.from_numpy(imgs_np[:625].reshape(-1)); pin = torch.empty(x.numel(), pin_memory=True)
t0 = time.perf_counter()
for _ in range(5): pin.copy_(x)
print("host copy pageable -> pinned, 92 MB: %.1f ms (%.1f GB/s)" % ((time.perf_counter() - t0) / 5 * 1e3, x.numel() * 4 / ((time.perf_counter() - t0) / 5) / 1e9))
# is the block path taken, and when do the blocks leave the device?
calls = []
orig = ops.scores_to_host_f64
def counted(block, host_block):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(block, host_block); e1.record(); calls.append((block.shape[1], e0, e1))
ops.scores_to_host_f64 = counted
start = torch.cuda.Event(enable_timing=True)
with contextlib.redirect_stdout(io.StringIO()):
    torch.cuda.synchronize(); start.record()
    sims = ev.cal_sims(m, imgs_np, caps_np, lengths=lens, shard_size=640)
torch.cuda.synchronize()
print("blocks shipped:", [(w, round(start.elapsed_time(a), 1), round(start.elapsed_time(b), 1)) for w, a, b in calls])

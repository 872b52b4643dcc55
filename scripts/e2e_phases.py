"""Diagnostic: where the end-to-end step (host buffers -> recall dict) spends its time at N ranks.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 scripts/e2e_phases.py"""
import importlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
itr_b200 = importlib.import_module("image-text-retrieval_b200")
ev, ops, sharding, synth = itr_b200.evaluation, itr_b200.ops, itr_b200.sharding, itr_b200.synth
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
CONFIG = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp", lambda_lse=6.0,
              lambda_softmax=9.0, margin=0.2, max_violation=True, measure="cosine")
n_img, n_cap = 5000, 25000
lens = synth.caption_lengths(n_cap, 10.5, 14)
lo, hi = sharding.shard_bounds(n_cap, world)[rank]
images, captions, ln = synth.scan_inputs(n_img, hi - lo, 10.5, 14, device=dev, lengths=lens[lo:hi])
images_h = torch.empty(images.shape, dtype=torch.float32, pin_memory=True).copy_(images)
captions_h = torch.empty(captions.shape, dtype=torch.float32, pin_memory=True).copy_(captions)
del captions, images
torch.cuda.empty_cache()
T = {}
def timed(name, fn):
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k)
        if SYNC: torch.cuda.synchronize()
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
        return r
    return w
ev._tc_t2i_inputs = timed("1 inputs (agree, image H2D+prep+gather)", ev._tc_t2i_inputs)
ev._scan_t2i_from = timed("2 scores from host captions", ev._scan_t2i_from)
sharding.sharded_ranks = timed("3 sharded_ranks", sharding.sharded_ranks)
ev._recall_dict = timed("5 recall dict", ev._recall_dict)
def step():
    return sharding.sharded_scan_eval(images_h, captions_h, ln, lo, n_cap, CONFIG, None)
for SYNC in (False, True, False):
    for _ in range(2): step()
    T.clear()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): step()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / 4
    if rank == 0:
        print("sync after each phase:", SYNC, " e2e step %.2f ms" % (tot * 1e3))
        for k in sorted(T): print("   %-45s %.2f ms" % (k, T[k] / 4 * 1e3))
        print("   %-45s %.2f ms" % ("(unaccounted: .cpu() x4, python)", (tot - sum(T.values()) / 4) * 1e3))
if world > 1: dist.destroy_process_group()

#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], the config the metric is quoted on; it fits one GPU):
SCAN t2i, clipped_l2norm, LogSumExp (lambda_lse 6, lambda_softmax 9), COCO-5K shape =
5000 images x 25000 captions (312 906 words), synthetic embeddings.
One step = one full evaluation pass: fp32 embeddings -> bf16 prep (cast, pack, Grams) ->
tcgen05 score kernel -> rank kernels (-> the tiny rank exchange when N > 1).  The captions are
sharded across ranks (strong scaling: the total problem is fixed).

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the
same metric through the public API (itr_b200.sharding.sharded_scan_eval, the multi-GPU form of
cal_sims + cal_recall) with pinned HOST buffers, copies inside the timed region; `e2e_dropin` (N = 1) is
the reference's own call sequence through the drop-in symbols -- cal_sims(numpy inputs) -> float64 host
matrix -> cal_recall(sims) -- and `device_breakdown_ms` splits the device step into its four stages.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = dict(name="SCAN", cross_attn="t2i", raw_feature_norm="clipped_l2norm", agg_func="LogSumExp",
              lambda_lse=6.0, lambda_softmax=9.0, margin=0.2, max_violation=True, measure="cosine")
METRIC = "image-caption pair scores/sec (SCAN t2i COCO-5K eval)"
R, D = 36, 1024
# the dominant kernel: the CTA-pair (tcgen05 cta_group::2) score kernel unless ITR_B200_SCORE_KERNEL=single selects round 1's
SCORE_KERNEL = "scan_t2i_tc_kernel" if os.environ.get("ITR_B200_SCORE_KERNEL", "").startswith("s") else "scan_t2i_tc2_kernel"


def measured_traffic(n_img, n_cap, world):
    """DRAM bytes per launch of the score kernel from the committed ncu capture (profiles/), or None."""
    for rnd in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True):
        path = os.path.join(ROOT, "profiles", rnd, "ncu_traffic.json")
        if os.path.exists(path):
            rec = json.load(open(path)).get(SCORE_KERNEL, {}).get("{}x{}@{}".format(n_img, n_cap, world))
            if rec:
                return rec["dram_bytes_read"] + rec["dram_bytes_write"]
    return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-gpu-eager-baseline", action="store_true",
                    help="skip the same-GPU baseline (the reference's op sequence in eager PyTorch on this GPU, ~0.5 s sample)")
    ap.add_argument("--n-img", type=int, default=5000, help="override for debugging only (invalidates the number)")
    ap.add_argument("--n-cap", type=int, default=25000)
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5, 6],
                    help="BASELINE.json config to measure; 5 (default) is the headline the driver runs; "
                         "6 = SCAN training step (fwd + bwd, batch 128), not a BASELINE config")
    ap.add_argument("--rank-mode", default="fused", choices=["fused", "matrix"],
                    help="device step: 'fused' ranks inside the score kernel (no score matrix), 'matrix' writes the matrix "
                         "and ranks it with the rank kernels (round 1); the other mode is timed briefly and reported too")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-caps", type=int, default=CPU_SAMPLE[1])
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]), src="measured (sustained)")
    return dict(hbm_gbs=6650.0, tflops=1590.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.file = index, None, None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.file,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [l.strip().split(", ") for l in open(self.file.name) if l.strip()]
        os.unlink(self.file.name)
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


_CPU_INPUTS = {}
CPU_SAMPLE = (1000, 700)      # images x captions of the COCO-5K-shaped workload scored per CPU step (both CPU legs): ~11 s on 16 cores


def _reference_fns():
    """(kind, scan_t2i(img, cap, lens) -> (n_img, n_cap) tensor, i2t, t2i): the reference's own functions when its
    package is importable (oracle/_ref on the GPU box, /root/reference in the authoring container), else the port."""
    from oracle import ref_loader, ref_port
    if ref_loader.available():
        try:
            O, E = ref_loader.load()
            cfg = dict(CONFIG)
            return ("reference", lambda img, cap, ln: O.xattn_score_t2i(img, cap, [int(x) for x in ln], cfg),
                    lambda s: E.i2t(s), lambda s: E.t2i(s))
        except Exception:      # noqa: BLE001  (a missing third-party import of the reference: fall back to the port)
            pass
    return ("port", lambda img, cap, ln: ref_port.scan_scores(img, cap, ln, "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0),
            ref_port.i2t_ranks, ref_port.t2i_ranks)


def cpu_reference_rate(n_img_s, n_cap_s, seed=14, lam=10.5, repeats=1):
    """The reference's CPU path (float32 torch, one Python iteration per caption, then numpy argsort ranking) on a
    bounded sample of the workload: first n_img_s images x first n_cap_s captions.
    Returns (pairs/s, seconds, cores, kind)."""
    from itr_b200 import synth
    if (n_img_s, n_cap_s) not in _CPU_INPUTS:
        lengths = synth.caption_lengths(25000, lam, seed)[:n_cap_s]
        _CPU_INPUTS[(n_img_s, n_cap_s)] = synth.scan_inputs(n_img_s, n_cap_s, lam, seed, device="cpu", lengths=lengths)
    img, cap, ln = _CPU_INPUTS[(n_img_s, n_cap_s)]
    kind, scan, i2t, t2i = _reference_fns()
    cores = torch.get_num_threads()
    with torch.no_grad():
        scan(img[:8], cap[:4], ln[:4])   # warm-up
        best = float("inf")
        for _ in range(repeats):
            t0 = time.perf_counter()
            sims = scan(img, cap, ln).double().numpy()
            # the reference then ranks on the host; time it on the sample's square part
            m = min(n_img_s, n_cap_s // 5)
            if m >= 1:
                i2t(sims[:m, : 5 * m]); t2i(sims[:m, : 5 * m])
            best = min(best, time.perf_counter() - t0)
    return n_img_s * n_cap_s / best, best, cores, kind


def _cpu_sample_text(n_img_s, n_cap_s, kind, secs=None):
    what = "the reference's own xattn_score_t2i + i2t/t2i (oracle/_ref, unmodified)" if kind == "reference" else \
        "the oracle's float32 torch port of the reference's per-caption loop + numpy ranking"
    return "first {} images x first {} captions of the COCO-5K-shaped workload per step, {} on the host cores{}".format(
        n_img_s, n_cap_s, what, "" if secs is None else ", {:.1f} s".format(secs))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on this box's host cores (the reference package staged under
    oracle/_ref when present, else the oracle port), same sample as the `cpu_baseline` leg of the b200 arm."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    n_img_s, n_cap_s = min(CPU_SAMPLE[0], args.n_img), min(CPU_SAMPLE[1], args.n_cap)
    # the whole --steps K --warmup W run has to end within a few minutes whatever K is: a 40-caption probe gives the rate,
    # and the per-step sample shrinks (never below 100 captions) if K + W full samples would take more than ~150 s
    probe_rate = cpu_reference_rate(n_img_s, min(40, n_cap_s))[0]
    budget_caps = int(probe_rate * 150.0 / max(1, args.warmup + args.steps) / n_img_s)
    n_cap_s = max(min(100, n_cap_s), min(n_cap_s, budget_caps // 20 * 20))
    times = []
    for i in range(args.warmup + args.steps):
        rate, secs, cores, kind = cpu_reference_rate(n_img_s, n_cap_s)
        if i >= args.warmup:
            times.append(secs)
    t = float(np.mean(times))
    value = n_img_s * n_cap_s / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SCAN t2i LSE COCO-5K shape (5000 img x 25000 caps); the CPU arm scores a bounded sample per step "
                                   "and reports the rate",
                       "n_img": args.n_img, "n_cap": args.n_cap},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind,
                             "sample": _cpu_sample_text(n_img_s, n_cap_s, kind)},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _time_cuda(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_side_config(args):
    """Configs 1-4 of BASELINE.json on one GPU (informational lines; the driver's contract is config 5)."""
    import itr_b200
    from itr_b200 import evaluation as ev, objectives as ob, ops, synth
    from oracle import ref_port
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    out = {"config": args.config, "n_gpus": 1, "data": "synthetic", "cores": cores}
    if args.config == 1:      # VSE++ cosine sim + i2t/t2i Recall@K, 1000 x 5000 x 1024
        im, s = synth.vse_inputs(1000, 5000, 1, device=dev)
        def dev_step():
            return ev.device_ranks(ops.cosine_scores(im, s))
        ms = _time_cuda(dev_step, 50)
        im_h, s_h = im.cpu().numpy(), s.cpu().numpy()
        class M: sim_enc = None
        m = M(); m.config = dict(CONFIG, name="VSE++"); m.criterion = ob.ContrastiveLoss(m.config, 0.2, "cosine", True)
        for _ in range(2):
            res = ev.cal_sims_and_recall(m, im_h, s_h)
        t0 = time.perf_counter()
        for _ in range(10):
            res = ev.cal_sims_and_recall(m, im_h, s_h)
        e2e = (time.perf_counter() - t0) / 10
        t0 = time.perf_counter()
        sims = ref_port.cosine_scores(torch.from_numpy(im_h), torch.from_numpy(s_h)).double().numpy()
        ref_port.i2t_ranks(sims); ref_port.t2i_ranks(sims)
        cpu = time.perf_counter() - t0
        out.update(workload="VSE++ cosine + i2t/t2i Recall@K, 1000 img x 5000 caps x 1024 (fp32 kernels)", device_ms=ms,
                   pairs_per_s=5e6 / (ms * 1e-3), e2e_ms=e2e * 1e3, cpu_port_ms=cpu * 1e3, rsum=res["rsum"])
    elif args.config == 2:    # ContrastiveLoss max_violation, batch 128, embed 1024, fwd + bwd
        im, s = synth.vse_inputs(128, 640, 2, device=dev)
        s = s[::5].contiguous()
        a, b = im.clone().requires_grad_(True), s.clone().requires_grad_(True)
        crit = ob.ContrastiveLoss(dict(CONFIG, name="VSE++"), margin=0.2, measure="cosine", max_violation=True)
        def ours():
            a.grad = None; b.grad = None
            crit(a, b).backward()
        def eager():          # the reference's op sequence in eager PyTorch on the same GPU
            a.grad = None; b.grad = None
            sc = a.mm(b.t())
            d = sc.diag().view(-1, 1)
            eye = torch.eye(128, device=dev) > .5
            cs = (0.2 + sc - d.expand_as(sc)).clamp(min=0).masked_fill_(eye, 0)
            ci = (0.2 + sc - d.t().expand_as(sc)).clamp(min=0).masked_fill_(eye, 0)
            (cs.max(1)[0].sum() + ci.max(0)[0].sum()).backward()
        us_ours, us_eager = _time_cuda(ours, 300) * 1e3, _time_cuda(eager, 300) * 1e3
        # the native call alone (one cooperative launch: scores + hinge + both gradients), without autograd bookkeeping,
        # and replayed from a CUDA graph (device time of the launch itself)
        im_c, s_c = im.contiguous(), s.contiguous()
        us_native = _time_cuda(lambda: ops.cosine_hinge(im_c, s_c, 0.2, True), 300) * 1e3
        us_graph = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                ops.cosine_hinge(im_c, s_c, 0.2, True)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    held = ops.cosine_hinge(im_c, s_c, 0.2, True)
            torch.cuda.current_stream().wait_stream(side)
            us_graph = _time_cuda(graph.replay, 300) * 1e3
        except Exception as exc:      # noqa: BLE001  (informational)
            us_graph = repr(exc)[:80]
        ac, bc = im.cpu().clone().requires_grad_(True), s.cpu().clone().requires_grad_(True)
        t0 = time.perf_counter()
        for _ in range(50):
            ac.grad = None; bc.grad = None
            ref_port.hinge(ac.mm(bc.t()), 0.2, True).backward()
        cpu = (time.perf_counter() - t0) / 50
        out.update(workload="VSE++ ContrastiveLoss max_violation fwd+bwd, batch 128 x 1024", us_per_call=us_ours,
                   us_per_call_native_no_autograd=us_native, us_per_call_cuda_graph_replay=us_graph,
                   us_per_call_eager_pytorch_same_gpu=us_eager, us_per_call_cpu_port=cpu * 1e6)
    elif args.config == 6:    # SCAN ContrastiveLoss t2i, max_violation, batch 128, embed 1024, fwd + bwd (row f3)
        lens_np = np.clip(synth.caption_lengths(128, 10.5, 16), 1, 60)
        img, cap, ln = synth.scan_inputs(128, 128, 10.5, 16, device=dev, lengths=lens_np)
        a, b = img.clone().requires_grad_(True), cap.clone().requires_grad_(True)
        cfg = dict(CONFIG, cross_attn="t2i", agg_func="LogSumExp", lambda_softmax=9.0)
        crit = ob.ContrastiveLoss(cfg, margin=0.2, measure="cosine", max_violation=True)
        lens_list = [int(x) for x in ln]
        def ours():
            a.grad = None; b.grad = None
            crit(a, b, lens_list).backward()
        def eager(x=a, y=b):  # the reference's per-caption op sequence under autograd, on the same GPU / on the CPU
            x.grad = None; y.grad = None
            ref_port.hinge_any(ref_port.scan_scores.__wrapped__(x, y, lens_list, "t2i", "clipped_l2norm", "LogSumExp", 9.0, 6.0),
                               0.2, True).backward()
        ms_ours, ms_eager = _time_cuda(ours, 30), _time_cuda(eager, 5, warmup=1)
        ac, bc = img.cpu().clone().requires_grad_(True), cap.cpu().clone().requires_grad_(True)
        t0 = time.perf_counter()
        eager(ac, bc)
        cpu = time.perf_counter() - t0
        out.update(workload="SCAN t2i ContrastiveLoss max_violation fwd+bwd, batch 128 x 128, 36 regions, embed 1024 (fp32 kernels)",
                   ms_per_step=ms_ours, ms_per_step_eager_pytorch_same_gpu=ms_eager, ms_per_step_cpu_port=cpu * 1e3)
    else:                     # SCAN 1000 x 5000 blocks
        if args.config == 3:
            shapes = [dict(synth.F30K_SHAPE)]
            direction, agg, lam = "t2i", "LogSumExp", 9.0
        else:
            shapes = [dict(n_img=1000, n_cap=5000, lam=10.5, seed=14, fold=f) for f in range(5)]
            direction, agg, lam = "i2t", "Mean", 4.0
        cfg = dict(CONFIG, cross_attn=direction, agg_func=agg, lambda_softmax=lam)
        tot_ms, tot_ms32, tot_ms2p, rsums = 0.0, 0.0, 0.0, []
        lens_all = synth.caption_lengths(25000, 10.5, 14)
        for sh in shapes:
            lengths = lens_all[sh["fold"] * 5000:(sh["fold"] + 1) * 5000] if "fold" in sh else None
            img, cap, ln = synth.scan_inputs(sh["n_img"], sh["n_cap"], sh["lam"], sh["seed"] + sh.get("fold", 0), device=dev, lengths=lengths)
            fn = ob.xattn_score_t2i if direction == "t2i" else ob.xattn_score_i2t
            def dev_step(c=cfg):
                return ev.device_ranks(fn(img, cap, ln, c))
            tot_ms += _time_cuda(dev_step, 3, warmup=1)
            if direction == "t2i":
                tot_ms32 += _time_cuda(lambda: dev_step(dict(cfg, itr_b200_precision="fp32")), 1, warmup=1)
            else:                                     # the round-1 route: affinity kernel + separate fp32 epilogue kernel
                os.environ["ITR_B200_I2T"] = "twophase"
                tot_ms2p += _time_cuda(dev_step, 2, warmup=1)
                del os.environ["ITR_B200_I2T"]
            r = [x.cpu().numpy().astype(np.float64) for x in dev_step()]
            rsums.append(sum(100.0 * np.mean(r[0] < k) + 100.0 * np.mean(r[2] < k) for k in (1, 5, 10)))
        n_s, c_s = 1000, max(20, args.cpu_sample_caps)
        t0 = time.perf_counter()
        ref_port.scan_scores(img[:n_s].cpu(), cap[:c_s].cpu(), ln[:c_s], direction, "clipped_l2norm", agg, lam, 6.0)
        cpu_rate = n_s * c_s / (time.perf_counter() - t0)
        pairs = sum(sh["n_img"] * sh["n_cap"] for sh in shapes)
        out.update(workload="SCAN {} {} ({} block(s) of 1000 img x 5000 caps){}".format(
                       direction, agg, len(shapes), ", fused tcgen05 bf16 kernel" if direction == "t2i" else
                       ", fused kernel: tcgen05 bf16 affinities + mma.sync tf32 per-caption contractions (captions > 32 words: two-phase path)"),
                   device_ms=tot_ms, pairs_per_s=pairs / (tot_ms * 1e-3), rsum=rsums,
                   cpu_port_pairs_per_s=cpu_rate, cpu_sample="1000 img x {} caps".format(c_s))
        if direction == "t2i":
            out.update(device_ms_fp32_mode=tot_ms32)
        else:
            # affinity flops of the five folds (2 x regions x words x 1024 each) over the whole fold time: the per-caption
            # contractions on mma.sync and everything else in the fold count as overhead
            pk = peaks()
            flop = 2.0 * 36 * 1024 * sum(sh["n_img"] * int(lens_all[sh["fold"] * 5000:(sh["fold"] + 1) * 5000].sum()) for sh in shapes)
            out.update(device_ms_two_phase=tot_ms2p,
                       roofline={"bound": "tensor", "achieved": flop / (tot_ms * 1e-3) / 1e12, "peak": pk["tflops"], "unit": "TFLOP/s",
                                 "frac": flop / (tot_ms * 1e-3) / 1e12 / pk["tflops"], "peak_source": pk["src"],
                                 "scope": "whole fold (prep, Gram fragments, fused kernel, two-phase path of the captions > 32 words, ranking)"})
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    if args.config != 5 and args.impl == "b200":
        run_side_config(args)
        return
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    import itr_b200
    from itr_b200 import evaluation as ev, ops, sharding, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ synthetic workload
    n_img, n_cap = args.n_img, args.n_cap
    lengths = synth.caption_lengths(n_cap, 10.5, 14)
    lo, hi = sharding.shard_bounds(n_cap, world)[rank]
    images, captions, _ = synth.scan_inputs(n_img, n_cap, 10.5, 14, device=dev, lengths=lengths)
    captions = captions[lo:hi].contiguous()
    ln_local = lengths[lo:hi]
    sum_words, sum_words_local = int(lengths.sum()), int(ln_local.sum())
    torch.cuda.synchronize()

    scores = torch.empty(n_img, hi - lo, device=dev, dtype=torch.float32)
    # per-step device breakdown: [start, images prepared (+ all-gathered), captions packed, scores done, ranks merged]
    marks = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]

    kern_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]

    class TimedStats(sharding.FusedScanStats):
        """FusedScanStats with CUDA events around the counting launch (the dominant kernel of the fused step)."""
        slot = None

        def count(self, block, thr_row, thr_col, cap_offset):
            if self.slot is not None:
                kern_ev[self.slot][0].record()
            res = super().count(block, thr_row, thr_col, cap_offset)
            if self.slot is not None:
                kern_ev[self.slot][1].record()
            return res

    def step(i=None, mode=None):
        mode = mode or args.rank_mode
        mark = (lambda k: marks[i][k].record()) if i is not None else (lambda k: None)
        mark(0)
        pi = ops.prepare_images_sharded(images, None, dev)      # each rank preps 1/N of the images + NCCL all-gather
        mark(1)
        pc = ops.prepare_captions(captions, ln_local)
        mark(2)
        if mode == "matrix":
            if i is not None:
                kern_ev[i][0].record()
            ops.scan_t2i_scores_bf16(pi, pc, CONFIG["raw_feature_norm"], CONFIG["agg_func"], CONFIG["lambda_softmax"],
                                     CONFIG["lambda_lse"], out=scores)
            if i is not None:
                kern_ev[i][1].record()
            mark(3)
            out = sharding.sharded_ranks(scores, lo, n_cap, None, 5)
        else:
            # ground-truth pre-pass (<1 % of the items) -> threshold all-reduce -> counting pass -> one all-gather
            stats = TimedStats(pi, pc, CONFIG["raw_feature_norm"], CONFIG["agg_func"], CONFIG["lambda_softmax"], CONFIG["lambda_lse"])
            stats.slot = i
            mark(3)
            out = sharding.sharded_ranks(stats.block(), lo, n_cap, None, 5, stats)
        mark(4)
        return out

    for _ in range(args.warmup):
        out = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        out = step(i)
    t1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    seg = [float(np.mean([m[k].elapsed_time(m[k + 1]) for m in marks])) for k in range(4)]
    kern_ms = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in kern_ev]))], device=dev)
    seg_ms = torch.tensor(seg, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(seg_ms, op=dist.ReduceOp.MAX)
    if args.rank_mode == "matrix":
        names = ["prep_images" + ("(+async gather)" if world > 1 else ""), "pack_captions", "score_kernel",
                 "rank_kernels" + ("+exchange" if world > 1 else "")]
    else:
        names = ["prep_images" + ("(+async gather)" if world > 1 else ""), "pack_captions", "host_setup",
                 "gt_prepass+score_count_kernel" + ("+exchange" if world > 1 else "")]
    breakdown = dict(zip(names, [round(x, 3) for x in seg_ms.tolist()]))
    breakdown["score_kernel_alone"] = round(kern_ms.item(), 3)
    # the other ranking mode, timed briefly on the same inputs
    other = "matrix" if args.rank_mode == "fused" else "fused"
    for _ in range(2):
        out_other = step(mode=other)
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for _ in range(3):
        out_other = step(mode=other)
    o1.record()
    barrier()
    other_ms = torch.tensor([o0.elapsed_time(o1) / 3], device=dev)
    if world > 1:
        dist.all_reduce(other_ms, op=dist.ReduceOp.MAX)
    same_ranks = all(torch.equal(a, b) for a, b in zip(out, out_other))
    ms_per_step = elapsed_ms.item() / args.steps
    value = n_img * n_cap / (ms_per_step * 1e-3)
    i2t_ranks, _, t2i_ranks, _ = out
    r1 = 100.0 * (i2t_ranks < 1).float().mean().item()
    r1_t = 100.0 * (t2i_ranks < 1).float().mean().item()

    # ------------------------------------------------------------------ end to end, host buffers
    images_h = torch.empty(images.shape, dtype=torch.float32, pin_memory=True).copy_(images)
    captions_h = torch.empty(captions.shape, dtype=torch.float32, pin_memory=True).copy_(captions)
    del captions
    torch.cuda.empty_cache()

    def e2e_step():
        return sharding.sharded_scan_eval(images_h, captions_h, ln_local, lo, n_cap, CONFIG, None)

    for _ in range(min(args.warmup, 2)):
        res = e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - w0) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_img * n_cap / e2e_s.item()

    # ------------------------------------------------------------------ the reference's own call sequence (N = 1)
    # evalrank_single (evaluation.py:284-291) / validate_step (utils.py:152-167): numpy inputs -- the image array is the
    # PAGEABLE copy the reference's de-duplication makes, the captions are what encode_data returned -- then
    # sims = cal_sims(...) (host float64 matrix) and cal_recall(sims) / i2t(sims) + t2i(sims).
    dropin = None
    if world == 1:
        class _M:
            sim_enc = None
        from itr_b200.objectives import ContrastiveLoss
        model = _M(); model.config = CONFIG
        model.criterion = ContrastiveLoss(CONFIG, margin=0.2, measure="cosine", max_violation=True)
        imgs_np = np.array(images_h.numpy())                       # pageable, like numpy.array([img_embs[i] ...])
        caps_np = captions_h.numpy()                               # view of the pinned buffer encode_data fills
        import contextlib, io

        def dropin_step(validate=False):
            with contextlib.redirect_stdout(io.StringIO()):
                sims = ev.cal_sims(model, imgs_np, caps_np, lengths=ln_local, shard_size=640)
                if validate:
                    return sims, (ev.i2t(sims), ev.t2i(sims))
                return sims, ev.cal_recall(sims)

        for _ in range(2):
            sims_h, res_d = dropin_step()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            sims_h, res_d = dropin_step()
        torch.cuda.synchronize()
        dropin_s = (time.perf_counter() - w0) / args.steps
        w0 = time.perf_counter()
        sims_h, (r_i, r_t) = dropin_step(validate=True)
        validate_s = time.perf_counter() - w0
        dropin = {"value": n_img * n_cap / dropin_s, "unit": "pairs/s", "ms_per_step": dropin_s * 1e3,
                  "ms_per_step_validate_sequence": validate_s * 1e3,
                  "vs_e2e": dropin_s / e2e_s.item(), "rsum": res_d["rsum"],
                  "sims": "{} {} host matrix returned by cal_sims".format(sims_h.dtype, tuple(sims_h.shape)),
                  "call": "cal_sims(model, pageable numpy images, numpy captions, lengths) -> float64 host matrix; cal_recall(sims)",
                  "h2d_bytes_per_step": n_img * R * D * 4 + sum_words_local * D * 4,
                  "d2h_bytes_per_step": n_img * n_cap * 8 + (n_img * 2 + n_cap * 2) * 8}
        del sims_h, imgs_np
    n_tiles = ops.plan_words(ln_local)[1]
    h2d = -(-n_img // world) * R * D * 4 + sum_words_local * D * 4 + n_tiles * 128 * 16
    d2h = (n_img * 2 + n_cap * 2) * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    f_alg = 2.0 * R * D * sum_words_local * n_img            # true words, affinity contraction only
    achieved = f_alg / (kern_ms.item() * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "SCAN t2i clipped_l2norm LogSumExp (lambda_lse 6, lambda_softmax 9), COCO-5K shape: "
                               "{} images x {} captions ({} words), captions sharded over {} GPU(s)".format(n_img, n_cap, sum_words, world),
                   "n_img": n_img, "n_cap": n_cap, "sum_words": sum_words, "parallelism": "caption-shard x{}".format(world),
                   "l2_policy": "inputs larger than L2 (bf16 operands {:.0f} MB per GPU{})".format(
                       (n_img * R * D * 2 + n_tiles * 128 * D * 2) / 1e6,
                       "" if args.rank_mode == "fused" else " + {:.0f} MB score block".format(n_img * (hi - lo) * 4 / 1e6)),
                   "step": "prep(cast,pack,gram" + (", image shards pulled from the peers' symmetric memory by the copy engines while the local shard is scored"
                                                         if world > 1 else "") + ") + " +
                           ("ground-truth pre-pass + tcgen05 scores with the ranking in the epilogue (no score matrix)" if args.rank_mode == "fused"
                            else "tcgen05 scores + rank kernels") + (" + rank exchange" if world > 1 else "")},
        "eval_wall_ms": {"device": ms_per_step, "e2e": e2e_s.item() * 1e3},
        "device_breakdown_ms": breakdown,
        "rank_mode": {"timed": args.rank_mode, "ms_per_step": ms_per_step, "other": other, "other_ms_per_step": other_ms.item(),
                      "identical_ranks": bool(same_ranks)},
        "recall_check": {"i2t_r1": r1, "t2i_r1": r1_t, "e2e_rsum": res["rsum"]},
        "roofline": {"bound": "tensor", "kernel": SCORE_KERNEL, "achieved": achieved, "peak": pk["tflops"],
                     "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": measured_traffic(n_img, n_cap, world),
                     "traffic_unit": "DRAM bytes per launch (ncu, profiles/*/ncu_traffic.json); algorithmic minimum {:.2f} GB".format(
                         (n_img * R * D * 2 + sum_words_local * D * 2 + n_img * (hi - lo) * 4) / 1e9),
                     "peak_source": pk["src"],
                     "kernel_ms": kern_ms.item(), "algorithmic_flop_per_launch": f_alg,
                     "kernel_share_of_step": kern_ms.item() / ms_per_step},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s.item() * 1e3},
        "gpu_launches": 5 * args.steps,      # fused: prep, pack, gt pre-pass, threshold un-key, score+count; matrix: prep, pack, scores, thresholds, counts
        "clocks": clocks,
    }
    if dropin is not None:
        line["e2e_dropin"] = dropin
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        n_s, c_s = min(CPU_SAMPLE[0], n_img), min(args.cpu_sample_caps, n_cap)
        rate, secs, cores, kind = cpu_reference_rate(n_s, c_s)
        line["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": kind,
                                "sample": _cpu_sample_text(n_s, c_s, kind, secs)}
    # second baseline: the reference's own op sequence -- one Python iteration per caption,
    # repeat / bmm / softmax / bmm / cosine in float32 -- in eager PyTorch on this same GPU, on a bounded sample of the
    # workload (all images x the first captions, ~0.5 s); informational, the driver's ratio uses the CPU arm
    if world == 1 and not args.no_gpu_eager_baseline:
        try:
            kind, scan, _, _ = _reference_fns()
            c_g = min(60, n_cap)
            caps_g = captions_h[:c_g].to(dev)
            with torch.no_grad():
                scan(images, caps_g[:8], lengths[:8])
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                scan(images, caps_g, lengths[:c_g])
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            line["gpu_eager_baseline"] = {"value": n_img * c_g / dt, "unit": "pairs/s", "kind": kind,
                                          "sample": "all {} images x first {} captions, the reference's xattn_score_t2i op sequence in "
                                                    "eager PyTorch fp32 on the same GPU, {:.2f} s".format(n_img, c_g, dt)}
        except Exception as exc:          # noqa: BLE001  (informational leg only)
            line["gpu_eager_baseline"] = {"unavailable": repr(exc)[:120]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

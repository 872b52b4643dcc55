#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py`: one block per step, a step
starting at each prep_images_kernel launch.  usage: scripts/launch_summary.py gpurun_out/launches_TAG.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip().replace("usecond", "us").replace("nsecond", "ns").replace("msecond", "ms"), 1e-6)
    launches.append((r[ki], v))
starts = [i for i, (k, _) in enumerate(launches) if "prep_images_kernel" in k]
# an end-to-end step prepares the images in several uploaded chunks (one prep launch each): a new step starts only at a
# prep launch that follows a score kernel (or is the first)
merged, seen_score = [], True
for a, b in zip(starts, starts[1:] + [len(launches)]):
    if seen_score:
        merged.append(a)
    seen_score = any("scan_t2i_tc" in k for k, _ in launches[a:b])
starts = merged
print("ncu --metrics gpu__time_duration.sum --clock-control none   command: python bench.py --steps 2 --warmup 1 --no-cpu-baseline")
print("(per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_share_of_step, not absolutes)")
print("{} launches captured, {} steps (a step starts at prep_images_kernel); launches before the first step are input generation".format(len(launches), len(starts)))
for n, s in enumerate(starts):
    e = starts[n + 1] if n + 1 < len(starts) else len(launches)
    step = launches[s:e]
    tot = sum(v for _, v in step)
    agg = OrderedDict()
    for k, v in step:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    pack = sum(v for k, v in step if "pack_words_kernel" in k)
    kind = "end-to-end step (captions gathered from pinned host memory by pack_words_kernel)" if pack > 5.0 else "device-resident step"
    print("\nstep {}: {}  ({} launches, {:.3f} ms of kernel time)".format(n, kind, len(step), tot))
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ours = "(ours)" if ("itr::" in k or "tc::" in k or "tc2::" in k) else "(torch plumbing)"
        print("  {:86s} x{:<3d} {:10.3f} ms {:6.2f}% {}".format(k[:86], c, v, 100 * v / tot, ours))

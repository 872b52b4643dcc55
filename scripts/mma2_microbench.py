#!/usr/bin/env python
"""tcgen05.mma cost model, CTA-pair form (cta_group::2, M = 256): cycles per MMA of ONE issuer for various N /
operand sources / issuer counts, next to the single-CTA numbers of scripts/mma_microbench.py."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from itr_b200 import _capi as capi
L = capi.lib()
iters = 4096
n_pairs = 74
print("%-4s %-5s %-6s %-6s %-8s %10s %10s %9s" % ("cg", "N", "n_acc", "A-src", "issuers", "clk/MMA", "agg clk", "N/2"))
def run2(n, n_acc, a_tmem, iss):
    cyc = torch.zeros(n_pairs, dtype=torch.int64, device="cuda")
    for _ in range(2):
        capi.check(L.itr_tc_mma2_microbench(n, n_acc, iters, a_tmem, iss, n_pairs, capi.ptr(cyc), capi.stream_ptr()))
    torch.cuda.synchronize()
    per = cyc.double().mean().item() / iters
    print("%-4d %-5d %-6d %-6s %-8d %10.1f %10.1f %9.1f" % (2, n, n_acc, "TMEM" if a_tmem else "SMEM", iss, per, per / iss, n / 2), flush=True)
def run1(n, n_acc, a_tmem, iss):
    cyc = torch.zeros(148, dtype=torch.int64, device="cuda")
    for _ in range(2):
        capi.check(L.itr_tc_mma_microbench(n, n_acc, iters, a_tmem, 1, iss, 148, capi.ptr(cyc), capi.stream_ptr()))
    torch.cuda.synchronize()
    per = cyc.double().mean().item() / iters
    print("%-4d %-5d %-6d %-6s %-8d %10.1f %10.1f %9.1f" % (1, n, n_acc, "TMEM" if a_tmem else "SMEM", iss, per, per / iss, n / 2), flush=True)
for n in (48, 144, 256):
    run1(n, 1, 0, 1)
for n, iss in ((48, 2), (144, 2), (144, 3), (224, 2)):
    run1(n, 1, 0, iss)
run1(48, 1, 1, 1); run1(48, 1, 1, 2)
for n in (32, 48, 96, 144, 160, 192, 224, 256):
    run2(n, 1, 0, 1)
run2(144, 2, 0, 1)
for n, iss in ((48, 2), (144, 2), (144, 3), (224, 2)):
    run2(n, 1, 0, iss)
for n in (48, 144):
    run2(n, 1, 1, 1)
run2(48, 1, 1, 2)

#!/usr/bin/env python
"""tcgen05.mma cost model on this GPU: cycles per M=128,K=16 MMA for various N / accumulator counts / operand
sources / number of concurrently issuing warps.  clk/MMA is per MMA of ONE issuer; agg = tensor-pipe clk per MMA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from itr_b200 import _capi as capi
L = capi.lib()
iters = 4096
print("%-5s %-6s %-6s %-5s %-8s %-5s %10s %10s %9s" % ("N", "n_acc", "A-src", "kadv", "issuers", "CTAs", "clk/MMA", "agg clk", "N/2"))
cases = []
for n in (16, 48, 64, 128, 144, 160, 192, 256):
    cases.append((n, 1, 0, 1, 1))
for n in (48, 144, 256):
    cases.append((n, 1, 1, 1, 1))
for n, iss in ((16, 2), (16, 4), (48, 2), (48, 4), (144, 2), (144, 3), (64, 4), (112, 4), (224, 2)):
    cases.append((n, 1, 0, 1, iss))
cases += [(144, 3, 0, 1, 1), (144, 1, 0, 0, 1), (48, 1, 1, 1, 2), (48, 1, 1, 1, 4)]
for n_ctas in (148,):
    for (n, n_acc, a_tmem, kadv, iss) in cases:
        cyc = torch.zeros(n_ctas, dtype=torch.int64, device="cuda")
        for _ in range(2):
            capi.check(L.itr_tc_mma_microbench(n, n_acc, iters, a_tmem, kadv, iss, n_ctas, capi.ptr(cyc), capi.stream_ptr()))
        torch.cuda.synchronize()
        per = cyc.double().mean().item() / iters
        print("%-5d %-6d %-6s %-5d %-8d %-5d %10.1f %10.1f %9.1f" % (n, n_acc, "TMEM" if a_tmem else "SMEM", kadv, iss, n_ctas, per, per / iss, n / 2))

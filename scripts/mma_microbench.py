#!/usr/bin/env python
"""tcgen05.mma cost model on this GPU: cycles per M=128,K=16 MMA for various N / accumulator counts / operand sources."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from itr_b200 import _capi as capi
L = capi.lib()
iters = 4096
print("%-6s %-6s %-7s %-5s %-6s %10s %10s" % ("N", "n_acc", "A-src", "kadv", "CTAs", "clk/MMA", "N/2 floor"))
for n_ctas in (1, 148):
    for a_tmem in (0, 1):
        for kadv in (1, 0):
            for n, n_acc in ((16, 1), (64, 1), (128, 1), (144, 1), (144, 2), (144, 3), (160, 1), (192, 1), (256, 1)):
                if a_tmem and kadv == 0 and n not in (144, 256):
                    continue
                cyc = torch.zeros(n_ctas, dtype=torch.int64, device="cuda")
                for _ in range(2):
                    capi.check(L.itr_tc_mma_microbench(n, n_acc, iters, a_tmem, kadv, n_ctas, capi.ptr(cyc), capi.stream_ptr()))
                torch.cuda.synchronize()
                print("%-6d %-6d %-7s %-5d %-6d %10.1f %10.1f" % (n, n_acc, "TMEM" if a_tmem else "SMEM", kadv, n_ctas,
                                                                  cyc.double().mean().item() / iters, n / 2))

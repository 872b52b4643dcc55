"""Diagnostic: throughput of the zero-copy float64 host write and of the one-shot conversion + DMA."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
itr_b200 = importlib.import_module("image-text-retrieval_b200")
ops, ev = itr_b200.ops, itr_b200.evaluation
d = torch.randn(5000, 25000, device="cuda")
host = torch.empty(5000, 25000, dtype=torch.float64, pin_memory=True)
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("zero-copy kernel, whole matrix        %.1f ms" % t(lambda: ops.scores_to_host_f64(d, host)))
print("zero-copy kernel, 1/8 column block    %.1f ms" % t(lambda: ops.scores_to_host_f64(d[:, :3125], host[:, :3125])))
print("d.double() + one DMA (old path)       %.1f ms" % t(lambda: host.copy_(d.double(), non_blocking=True)))
f32 = torch.empty(5000, 25000, dtype=torch.float32, pin_memory=True)
print("f32 DMA only                          %.1f ms" % t(lambda: f32.copy_(d, non_blocking=True)))
assert torch.equal(host, d.double().cpu())

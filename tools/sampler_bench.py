"""Sampler kernel timings (64 rows x 128 256 logits): greedy vs the stochastic paths.  python tools/sampler_bench.py"""
import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import torch, ollamamq_b200 as m
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
rows, V = 64, 128256
logits = torch.randn(rows, V, device='cuda')
out = torch.zeros(rows, dtype=torch.int32, device='cuda')
temp0 = torch.zeros(rows, device='cuda'); temp1 = torch.full((rows,), 0.8, device='cuda')
k = torch.full((rows,), 40, dtype=torch.int32, device='cuda'); p = torch.full((rows,), 0.9, device='cuda')
seed = torch.zeros(rows, dtype=torch.int64, device='cuda'); cnt = torch.zeros(rows, dtype=torch.int32, device='cuda')
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1000
print("argmax_kernel        %.1f us" % t(lambda: m.lib.mq_debug_argmax(P(logits), rows, V, V, P(out), None, None, None, None)))
print("sample_kernel greedy %.1f us" % t(lambda: m.lib.mq_debug_sample(P(logits), rows, V, V, P(out), P(temp0), P(k), P(p), P(seed), P(cnt))))
print("sample_kernel T=0.8 top_k=40 top_p=0.9 %.1f us" % t(lambda: m.lib.mq_debug_sample(P(logits), rows, V, V, P(out), P(temp1), P(k), P(p), P(seed), P(cnt))))
k0 = torch.zeros(rows, dtype=torch.int32, device='cuda'); p0 = torch.zeros(rows, device='cuda')
print("sample_kernel T=0.8 only %.1f us" % t(lambda: m.lib.mq_debug_sample(P(logits), rows, V, V, P(out), P(temp1), P(k0), P(p0), P(seed), P(cnt))))
sharp = logits * 10
print("sample_kernel T=0.8 top_p=0.9 alone, peaked row (nucleus inside the candidate set) %.1f us" % t(lambda: m.lib.mq_debug_sample(P(sharp), rows, V, V, P(out), P(temp1), P(k0), P(p), P(seed), P(cnt))))
print("sample_kernel T=0.8 top_p=0.9 alone, flat row (radix walk) %.1f us" % t(lambda: m.lib.mq_debug_sample(P(logits), rows, V, V, P(out), P(temp1), P(k0), P(p), P(seed), P(cnt))))

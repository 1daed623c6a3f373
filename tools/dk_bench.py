"""Micro-benchmark of the decode-chain GEMM (csrc/gemm_dk.cuh) against the plane-based split-K kernel, same shapes.
MQ_DK_DBG=1 prints CTA 0's phase stamps.   python tools/dk_bench.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ollamamq_b200 as m  # noqa: E402

P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
dev = torch.device("cuda:0")
# a pool of weight matrices larger than L2 so every timed launch streams from HBM
for (T, n_out, K, css, splits) in [(64, 4096, 4096, (1, 2, 4), 4), (64, 4096, 14336, (2, 4), 4), (16, 4096, 4096, (1, 2, 4), 4),
                                   (64, 6144, 4096, (2, 3), 3)]:
    W = (torch.randn(n_out, K, device=dev) * 0.03).bfloat16()
    X = torch.randn(64, K, device=dev).bfloat16()
    h = torch.zeros(T, n_out, device=dev)
    gamma = torch.ones(n_out, device=dev).bfloat16()
    xg = torch.zeros(T, n_out, device=dev, dtype=torch.bfloat16)
    ssq = torch.zeros((n_out + 127) // 128, 64, device=dev)
    ms = C.c_float()
    for cs in css:
        if T // cs > 32 or (T + cs - 1) // cs > 32:
            continue
        rc = m.lib.mq_debug_gemm_dk_resid(P(W), n_out, K, P(X), 64, T, cs, P(h), P(gamma), P(xg), P(ssq), 64, 50, C.byref(ms))
        print("dk_resid T=%d n_out=%d K=%d cs=%d: rc=%d %.2f us/launch (%.0f GB/s)" % (T, n_out, K, cs, rc, ms.value * 1e3, n_out * K * 2 / ms.value / 1e6), flush=True)
    out = torch.zeros(splits, T, n_out, device=dev)
    rc = m.lib.mq_debug_gemm(P(W), n_out, n_out, K, P(X), 64, T, 0, P(out), n_out, splits, T * n_out, 0, 0, 50, C.byref(ms))
    print("planes   T=%d n_out=%d K=%d splits=%d: rc=%d %.2f us/launch (%.0f GB/s)" % (T, n_out, K, splits, rc, ms.value * 1e3, n_out * K * 2 / ms.value / 1e6), flush=True)

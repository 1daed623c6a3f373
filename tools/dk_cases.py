"""Run decode-chain GEMM cases one per subprocess with a short timeout (a deadlocked cluster kernel must not eat the
GPU lease).   python tools/dk_cases.py            # all cases
              python tools/dk_cases.py T n_out K cs  # one case, in-process"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(16, 4096, 14336, 4), (16, 4096, 4096, 4), (16, 4096, 4096, 2), (16, 1024, 1024, 1), (64, 4096, 4096, 4),
         (32, 4096, 4096, 4), (32, 4096, 4096, 2), (48, 1024, 2048, 3), (5, 512, 512, 2), (33, 3584, 3584, 4)]


def one(T, n_out, K, cs):
    import torch
    import ollamamq_b200 as m
    P = lambda t: C.c_void_p(t.data_ptr())
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(1)
    W = (torch.randn(n_out, K, device=dev, generator=g) * 0.03).bfloat16()
    X = torch.randn(64, K, device=dev, generator=g).bfloat16()
    h0 = torch.randn(T, n_out, device=dev, generator=g)
    h = h0.clone()
    gamma = torch.ones(n_out, device=dev).bfloat16()
    xg = torch.zeros(T, n_out, device=dev, dtype=torch.bfloat16)
    ssq = torch.zeros((n_out + 127) // 128, 64, device=dev)
    rc = m.lib.mq_debug_gemm_dk_resid(P(W), n_out, K, P(X), 64, T, cs, P(h), P(gamma), P(xg), P(ssq), 64, 0, None)
    ref = h0 + X[:T].float() @ W.float().T
    err = ((h - ref).norm() / ref.norm()).item()
    print("rc=%d relerr=%.2e %s" % (rc, err, m.last_error() if rc else ""))


if len(sys.argv) == 5:
    one(*[int(x) for x in sys.argv[1:]])
else:
    for c in CASES:
        try:
            r = subprocess.run(["timeout", "-k", "5", "40", sys.executable, __file__] + [str(x) for x in c],
                               capture_output=True, text=True, env=dict(os.environ, MQ_DK_DBG="1"))
            print(c, "exit", r.returncode, (r.stdout.strip().splitlines() or [""])[-1], "|", (r.stderr.strip().splitlines() or [""])[-1][:300], flush=True)
        except Exception as e:
            print(c, "EXC", e, flush=True)

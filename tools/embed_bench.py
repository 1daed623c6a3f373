"""BASELINE configs[4] shape on the embedding worker: batches of `n_seq` x `seq_len` random token ids through the
blocking C-ABI call (host token arrays in, host fp32 embeddings out), bge-small geometry, random-init weights.

    python tools/embed_bench.py [n_seq] [seq_len] [tokens_per_pass] [reps]
Prints sequences/s, tokens/s and the achieved TFLOP/s against SURVEY.md 8(d)'s algorithmic work
(2 * P_mm * T + attention flops; 27.2 TFLOP for 1024 x 512)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import ollamamq_b200 as mq  # noqa: E402
from ollamamq_b200.models import BGE_SMALL  # noqa: E402

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seq_len = int(sys.argv[2]) if len(sys.argv) > 2 else 512
per_pass = int(sys.argv[3]) if len(sys.argv) > 3 else 75776   # 2 x 148 x 256: whole waves on 148 SMs for 512-token inputs
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cfg = BGE_SMALL
H, I, L = cfg["hidden"], cfg["ffn"], cfg["n_layers"]
p_mm = L * (3 * H * H + H * H + 2 * H * I)
flops = 2.0 * p_mm * n_seq * seq_len + L * 4.0 * seq_len * seq_len * H * n_seq      # bidirectional: full s^2
rng = np.random.default_rng(0)
seqs = [rng.integers(0, cfg["vocab"], seq_len).astype("int32").tolist() for _ in range(n_seq)]
with mq.Encoder(0, mq.encoder_cfg(cfg, max_seq=seq_len, max_tokens_per_pass=per_pass)) as e:
    e.init_random(0, 0.05)
    e.embed(seqs[:64])
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        out = e.embed(seqs)
        best = min(best, time.perf_counter() - t0)
    arr = np.asarray(seqs, dtype=np.int32)            # [n_seq, seq_len]: the same batch without per-token python work
    s0 = e.stats()
    best_np = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        out2 = e.embed(arr)
        best_np = min(best_np, time.perf_counter() - t0)
    st = e.stats()
    gpu_ms = (st["gpu_us"] - s0["gpu_us"]) / 1e3 / reps
    assert np.array_equal(out, out2)
    print("  same batch as one int32 array: %.1f ms per batch = %.0f sequences/s; device time of its passes (CUDA events) %.1f ms "
          "= %.1f TFLOP/s = %.1f %% of the sustained 1 457.8 TFLOP/s" %
          (best_np * 1e3, n_seq / best_np, gpu_ms, flops / (gpu_ms * 1e-3) / 1e12, 100 * flops / (gpu_ms * 1e-3) / 1457.8e12))
    print("bge-small %d x %d tokens, passes of %d: %.1f ms per batch = %.0f sequences/s, %.2f M tokens/s, %.1f TFLOP/s "
          "(algorithmic %.1f TFLOP; end to end through mq_encoder_embed incl. python list marshalling) | %d launches per pass"
          % (n_seq, seq_len, per_pass, best * 1e3, n_seq / best, n_seq * seq_len / best / 1e6, flops / best / 1e12,
             flops / 1e12, st["kernel_launches"] // max(1, st["passes"])))

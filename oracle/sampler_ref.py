"""sampler_ref.py - TEST INFRASTRUCTURE ONLY (imported by tests/; never by the product path).

numpy restatement of the sampling step the reference leaves to the external backend (Ollama "options":
temperature / top_k / top_p / seed; call site /root/reference/src/dispatcher.rs:287-290, nothing pinned there):

  keep   = {i : logit_i >= k-th largest logit}                     (top_k; ties at the threshold stay in)
  keep  &= smallest top set of `keep`, by descending logit, whose softmax(l / T) mass reaches top_p * mass(keep)
  token  = argmax_{i in keep} ( logit_i / T + g_i ),   g_i = -log(-log u_i)           (Gumbel-max = exact sampling)
  u_i    = (top 24 bits of splitmix64(seed ^ position * C1 ^ i * C2) + 0.5) / 2^24

which is what csrc/kernels.cu:sample_kernel computes (there with radix selection instead of sorting and 2^40 fixed-point
masses).  temperature <= 0 is greedy argmax, lowest index on ties.
"""
from __future__ import annotations

import numpy as np

M64 = (1 << 64) - 1
C1, C2 = 0xD1342543DE82EF95, 0xA24BAED4963EE407


def _mix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return x ^ (x >> np.uint64(31))


def gumbel(seed: int, counter: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        key = np.uint64(seed & M64) ^ np.uint64((counter * C1) & M64) ^ (idx * np.uint64(C2))
        r = _mix64(key)
    u = ((r >> np.uint64(40)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    return -np.log(-np.log(u, dtype=np.float32), dtype=np.float32)


def keep_mask(logits: np.ndarray, temperature: float, top_k: int = 0, top_p: float = 0.0) -> np.ndarray:
    l = logits.astype(np.float32)
    keep = np.ones(l.shape[0], dtype=bool)
    if 0 < top_k < l.shape[0]:
        thr = np.partition(l, -top_k)[-top_k]
        keep &= l >= thr
    if 0.0 < top_p < 1.0:
        w = np.where(keep, np.exp((l.astype(np.float64) - l.max()) / temperature), 0.0)
        order = np.argsort(-l, kind="stable")
        cum = np.cumsum(w[order])
        n_keep = int(np.searchsorted(cum, top_p * cum[-1], side="left")) + 1
        thr_p = l[order[min(n_keep, l.shape[0]) - 1]]
        keep &= l >= thr_p
    return keep


def sample(logits: np.ndarray, temperature: float, top_k: int = 0, top_p: float = 0.0, seed: int = 0,
           counter: int = 0):
    """Returns (token, perturbed scores with -inf outside the kept set)."""
    l = logits.astype(np.float32)
    if not temperature > 0:
        return int(np.argmax(l)), l
    keep = keep_mask(l, temperature, top_k, top_p)
    sc = l * np.float32(1.0 / temperature) + gumbel(seed, counter, l.shape[0])
    sc = np.where(keep, sc, -np.inf).astype(np.float32)
    return int(np.argmax(sc)), sc

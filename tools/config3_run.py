"""BASELINE configs[2] live: Qwen2.5-7B geometry (random init), 32 users whose request counts follow Zipf(1.1) and total 256,
2 VIP + 4 Boost users (the documented set EXTENSION of the reference's single slots), one B200 worker behind the
fair-share dispatcher.  Everything is submitted at t = 0; per priority class it reports the mean position in the
dispatch order, the mean time to the first token and the mean completion time - the live counterpart of
tests/test_dispatch.py::test_config3_multi_vip_boost_extension_priority_fairness.

    python tools/config3_run.py [capacity] [prompt_len] [gen_len]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import ollamamq_b200 as mq  # noqa: E402
from ollamamq_b200.models import QWEN25_7B  # noqa: E402

capacity = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prompt_len = int(sys.argv[2]) if len(sys.argv) > 2 else 256
gen_len = int(sys.argv[3]) if len(sys.argv) > 3 else 64


def zipf_counts(n_users=32, total=256, s=1.1, seed=0):
    w = 1.0 / np.arange(1, n_users + 1) ** s
    counts = np.maximum(1, np.floor(w / w.sum() * total).astype(int))
    rng = np.random.default_rng(seed)
    while counts.sum() < total:
        counts[rng.integers(0, n_users)] += 1
    while counts.sum() > total:
        counts[int(np.argmax(counts))] -= 1
    return counts.tolist()


users = ["user%02d" % i for i in range(32)]
counts = zipf_counts()
vips, boosts = ["user09", "user20"], ["user03", "user12", "user25", "user30"]
cls = {u: ("vip" if u in vips else "boost" if u in boosts else "other") for u in users}
cfg = QWEN25_7B
wk = mq.Worker(0, mq.model_cfg(cfg, max_batch=capacity, max_seq=prompt_len + gen_len + 16, max_prefill_tokens=4096,
                               use_graphs=1, use_pdl=1, model_name="qwen2.5-7b-random"))
wk.init_random(0, 0.02)
d = mq.Dispatcher([wk], capacity=capacity)
for u in vips:
    d.add_vip(u)
for u in boosts:
    d.add_boost(u)
rng = np.random.default_rng(1)
order = [u for u, c in zip(users, counts) for _ in range(c)]
rng.shuffle(order)
# warm-up (graph capture, attribute setup)
w0 = d.submit("warm", prompt_tokens=rng.integers(0, cfg["vocab"], prompt_len).astype("int32").tolist(), max_new_tokens=4)
w0.wait(300)
t0 = time.perf_counter()
streams = [(u, d.submit(u, prompt_tokens=rng.integers(0, cfg["vocab"], prompt_len).astype("int32").tolist(),
                        max_new_tokens=gen_len)) for u in order]
d.drain(600000)
wall = time.perf_counter() - t0
log = [x for x in d.log() if x[0] != "warm"]
rank = {}
for pos, (u, seq, be) in enumerate(log):
    rank.setdefault(u, []).append(pos)
print("# BASELINE configs[2] live on one B200: Qwen2.5-7B geometry, 32 users / 256 requests (Zipf 1.1), prompt %d, gen %d, "
      "capacity %d; 2 VIP + 4 Boost (extension); %.2f s wall, %.0f tokens/s" %
      (prompt_len, gen_len, capacity, wall, 256 * gen_len / wall))
print("# class   users  requests | mean dispatch position (0..255) | mean TTFT ms | mean completion ms")
for c in ("vip", "boost", "other"):
    us = [u for u in users if cls[u] == c]
    ss = [s for u, s in streams if cls[u] == c]
    assert all(s.rc == 0 for s in ss)
    pos = np.mean([p for u in us for p in rank[u]])
    ttft = np.mean([s.ttft for s in ss]) * 1e3
    done = np.mean([s.chunk_times[-1] - s.t_submit for s in ss]) * 1e3
    print("%-7s %6d %9d | %31.1f | %12.1f | %18.1f" % (c, len(us), len(ss), pos, ttft, done))
d.close()
wk.close()

"""Overlapped timeline of ONE live decode step (CUDA graph + PDL chain), from in-kernel %globaltimer stamps.

    MQ_TRACE=1 python tools/decode_timeline.py [users] [gen_len] [model] [prompt_len] > profiles/rNN_decode_timeline.txt

ncu serialises launches (cold caches, no overlap) and nsys is not in the image, so this is the only view of how
the kernels of a step actually overlap: per launch, the first CTA's start, the moment its dependency wait
(griddepcontrol.wait) returned, the first and the last CTA's end.  The stamps cost one atomic per CTA; the step
is ~1 % slower with them on.  Numbers printed per kernel type are means over the layers of the step.
"""
import ctypes as C
import os
import sys

os.environ["MQ_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import ollamamq_b200 as mq  # noqa: E402
from ollamamq_b200 import models  # noqa: E402

users = int(sys.argv[1]) if len(sys.argv) > 1 else 64
gen = int(sys.argv[2]) if len(sys.argv) > 2 else 64
model = sys.argv[3] if len(sys.argv) > 3 else "LLAMA3_8B"      # LLAMA3_8B | QWEN25_7B | PHI3_MINI
plen = int(sys.argv[4]) if len(sys.argv) > 4 else 512
cfg = getattr(models, model)
L = cfg["n_layers"]
P = [np.random.default_rng(u).integers(0, cfg["vocab"], plen).astype("int32").tolist() for u in range(users)]
wk = mq.Worker(0, mq.model_cfg(cfg, max_batch=max(64, users), max_seq=plen + 128 + 16, max_prefill_tokens=4736, use_graphs=1,
                               use_pdl=int(os.environ.get("MQ_PDL", "1"))))
wk.init_random(0, 0.02)
wk.set_timing(True)
d = mq.Dispatcher([wk], capacity=max(64, users))
for rep in range(2):
    if rep == 1:
        wk.reset_stats()
    ss = [d.submit("user%02d" % u, prompt_tokens=P[u], max_new_tokens=gen) for u in range(users)]
    for s in ss:
        s.wait(600)
        assert s.rc == 0, s.err
st = wk.stats()
SLOTS = 512
buf = (C.c_ulonglong * (SLOTS * 4))()
n = mq.lib.mq_debug_trace_read(wk._h, buf, SLOTS)
assert n == SLOTS, mq.last_error()
t = np.frombuffer(buf, dtype=np.uint64).reshape(SLOTS, 4).copy()
FF = np.uint64(0xFFFFFFFFFFFFFFFF)
t[:, 3] = ~t[:, 3]
names = ["norm1", "qkv", "rope", "attn", "o", "norm2", "gate_up", "down"]
ids = [(1 + 8 * l + k, l, names[k]) for l in range(L) for k in range(8)] + [(510, L, "final_norm"), (511, L, "lm_head")]
ids = [(i, l, nm) for i, l, nm in ids if t[i, 0] != FF]
t0 = min(int(t[i, 0]) for i, _, _ in ids)
t_end = max(int(t[i, 3]) for i, _, _ in ids)
print("# live decode step, %d users, ctx ~%d, %s: %.1f us from first kernel start to last kernel end "
      "(worker stats: %.3f ms/step over %d steps incl. argmax + launch)" %
      (users, plen + gen, model, (t_end - t0) / 1e3, st["decode_ms"] / max(1, st["decode_steps"]), st["decode_steps"]))
print("# kernels per layer: %d (%s)" % (sum(1 for _, l, _ in ids if l == 1), ", ".join(nm for _, l, nm in ids if l == 1)))
print("# per launch: start / dependency-wait-returned / first-CTA-end / last-CTA-end, us from the step's first stamp")
rows = []
prev_end = t0
for i, l, nm in ids:
    s0, w0, e0, e1 = (int(x) for x in t[i])
    rows.append((l, nm, (s0 - t0) / 1e3, (w0 - t0) / 1e3, (e0 - t0) / 1e3, (e1 - t0) / 1e3, (w0 - prev_end) / 1e3))
    prev_end = e1
print("# layer kernel     start   waited  first_end last_end | lead (start before prev end) | exposed (last_end - prev last_end)")
prev = 0.0
agg = {}
for l, nm, s0, w0, e0, e1, gap in rows:
    lead = prev - s0
    exposed = e1 - prev
    if l in (0, 1, L // 2, L - 1, L):
        print("%5d %-10s %8.1f %8.1f %8.1f %8.1f | %6.1f | %6.1f" % (l, nm, s0, w0, e0, e1, lead, exposed))
    if 1 <= l < L:  # steady-state layers only
        a = agg.setdefault(nm, [])
        a.append((e1 - s0, e1 - w0, exposed, lead, w0 - prev, e1 - e0))
    prev = e1
print("\n# steady state (layers 1..%d), means in us" % (L - 1))
print("# kernel      span(start->last_end)  busy(waited->last_end)  exposed(critical path)  lead-in  wait-after-prev-end  tail(first->last CTA end)")
tot = 0.0
for nm in names:
    if nm not in agg:  # e.g. no rope launch when RoPE is fused into the attention kernel
        continue
    a = np.array(agg[nm])
    m = a.mean(0)
    tot += m[2]
    print("%-10s %10.2f %22.2f %22.2f %12.2f %14.2f %16.2f" % (nm, m[0], m[1], m[2], m[3], m[4], m[5]))
print("# sum of exposed per layer: %.2f us  -> x %d layers = %.3f ms" % (tot, L, tot * L / 1e3))
d.close()
wk.close()

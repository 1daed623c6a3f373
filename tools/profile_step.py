"""Small, profiler-friendly run of the hot path for ncu (used under gpurun; see profiles/README).

    ncu --profile-from-start off ... python tools/profile_step.py [users] [gen_len]

Warm-up (graph capture, attribute setup) happens before cudaProfilerStart(); the profiled region is one pass of
the trace with `users` users x (512-token prompt + gen_len tokens) on Llama-3-8B random-init.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ollamamq_b200 as mq  # noqa: E402
from ollamamq_b200.models import LLAMA3_8B  # noqa: E402

users = int(sys.argv[1]) if len(sys.argv) > 1 else 64
gen = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pdl = int(os.environ.get("MQ_PDL", "1"))
P = [np.random.default_rng(u).integers(0, LLAMA3_8B["vocab"], 512).astype("int32").tolist() for u in range(users)]
wk = mq.Worker(0, mq.model_cfg(LLAMA3_8B, max_batch=64, max_seq=512 + 128 + 16, max_prefill_tokens=int(os.environ.get("MQ_PREFILL_TOKENS", "9472")), use_graphs=1,
                               use_pdl=pdl))
wk.init_random(0, 0.02)
wk.set_timing(True)
d = mq.Dispatcher([wk], capacity=64)


def step():
    ss = [d.submit("user%02d" % u, prompt_tokens=P[u], max_new_tokens=gen) for u in range(users)]
    for s in ss:
        s.wait(600)
        assert s.rc == 0, s.err


step()
torch.cuda.synchronize()
wk.reset_stats()
torch.cuda.profiler.start()
t0 = time.time()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
st = wk.stats()
print("profiled step: %.2fs wall, decode %.3f ms/step over %d steps, prefill %.1f ms over %d passes" %
      (time.time() - t0, st["decode_ms"] / max(1, st["decode_steps"]), st["decode_steps"], st["prefill_ms"],
       st["prefill_passes"]))
d.close()
wk.close()

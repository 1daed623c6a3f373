"""Run the headline trace once (or twice) and dump the generated tokens: python tools/trace_tokens.py out.npy [warm_runs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench_trace as BT  # noqa: E402
import ollamamq_b200 as mq  # noqa: E402
from ollamamq_b200.models import LLAMA3_8B  # noqa: E402

out = sys.argv[1]
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 0
P = BT.prompts()
wk = mq.Worker(0, mq.model_cfg(LLAMA3_8B, max_batch=64, max_seq=BT.PROMPT_LEN + BT.GEN_LEN + 16, max_prefill_tokens=4736, use_graphs=1,
                               use_pdl=1))
wk.init_random(seed=0, std=0.02)
if os.environ.get("TT_TIMING"):
    wk.set_timing(True)
if os.environ.get("TT_TORCH"):
    import torch
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
d = mq.Dispatcher([wk], capacity=64)
order = list(range(BT.USERS))
if os.environ.get("TT_ORDER") == "sched":   # the order bench.py submits in: the scheduler's own dispatch order of the t=0 trace
    sc = mq.Scheduler(1, capacity=BT.USERS)
    for u in range(BT.USERS):
        sc.enqueue("user%02d" % u)
    order = [int(x.user[4:]) for x in sc.drain()]
elif os.environ.get("TT_ORDER") == "rev":
    order = order[::-1]
for rep in range(warm + 1):
    sub = {u: d.submit("user%02d" % u, prompt_tokens=P[u], max_new_tokens=BT.GEN_LEN, stream=1) for u in order}
    ss = [sub[u] for u in range(BT.USERS)]
    toks = []
    for s in ss:
        s.wait(600)
        assert s.rc == 0
        toks.append(s.tokens())
    st = wk.stats()
    print("run %d: checksum %s, prefill passes so far %d, decode steps %d" % (rep, BT.token_checksum(toks), st["prefill_passes"], st["decode_steps"]))
np.save(out, np.array(toks, dtype=np.int32))
d.close()
wk.close()

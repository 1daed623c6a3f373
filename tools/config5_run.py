"""BASELINE configs[4] live: /api/embed batches (bge-small geometry, 512-token inputs) interleaved 1:4 with configs[1]-style
chat (Llama-3-8B geometry, 512-token prompt / 128 tokens) across N B200s behind ONE dispatcher: every backend is a
generation worker plus an embedding worker sharing one GPU on separate streams; the scheduler's least-connections pick
(dispatcher.rs:247-254) spreads chat and embed requests alike.

    python tools/config5_run.py [chat_users] [embed_requests] [seqs_per_embed_request] [n_gpus]
Reports chat tokens/s and embedding sequences/s alone and together.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import ollamamq_b200 as mq  # noqa: E402
from ollamamq_b200.models import BGE_SMALL, LLAMA3_8B  # noqa: E402

chat_users = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_embed = int(sys.argv[2]) if len(sys.argv) > 2 else 16          # 1 embed request per 4 chat requests
per_req = int(sys.argv[3]) if len(sys.argv) > 3 else 64          # 16 x 64 = 1024 sequences of 512 tokens
n_gpus = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rng = np.random.default_rng(0)
wks, encs = [], []
for g in range(n_gpus):
    wk = mq.Worker(g, mq.model_cfg(LLAMA3_8B, max_batch=64, max_seq=512 + 128 + 16, max_prefill_tokens=4736, use_graphs=1,
                                   use_pdl=1))
    wk.init_random(0, 0.02)
    enc = mq.Encoder(g, mq.encoder_cfg(BGE_SMALL, max_seq=512, max_tokens_per_pass=32768))
    enc.init_random(0, 0.05)
    wks.append(wk)
    encs.append(enc)
d = mq.Dispatcher(wks, capacity=chat_users + n_embed)
for g in range(n_gpus):
    d.attach_encoder(g, encs[g])
prompts = [rng.integers(0, LLAMA3_8B["vocab"], 512).astype("int32").tolist() for _ in range(chat_users)]
embed_bodies = [json.dumps({"model": "bge-small", "input": [rng.integers(1000, 30000, 510).tolist() for _ in range(per_req)]}).encode()
                for _ in range(n_embed)]


def run(chat: bool, embed: bool):
    t0 = time.perf_counter()
    cs, es = [], []
    for i in range(max(chat_users if chat else 0, 4 * n_embed if embed else 0)):
        if chat and i < chat_users:
            cs.append(d.submit("chat%02d" % i, prompt_tokens=prompts[i], max_new_tokens=128))
        if embed and i % 4 == 3 and i // 4 < n_embed:
            es.append(d.submit("embed%02d" % (i // 4), endpoint=6, body=embed_bodies[i // 4], path="/api/embed",
                               max_new_tokens=0))
    d.drain(600000)
    for s in cs + es:
        assert s.rc == 0, s.err
    t_chat = max((s.chunk_times[-1] for s in cs), default=t0) - t0
    t_emb = max((s.chunk_times[-1] for s in es), default=t0) - t0
    return t_chat, t_emb, (np.median([s.ttft for s in cs]) * 1e3 if cs else 0.0)


run(True, True)  # warm-up: graphs, attributes
a = run(True, False)
b = run(False, True)
c = run(True, True)
ntok, nseq = chat_users * 128, n_embed * per_req
print("# BASELINE configs[4] live on %d B200(s): %d chat users (Llama-3-8B geometry, 512 / 128) + %d /api/embed requests of %d x 512 "
      "tokens (bge-small geometry), 1 embed per 4 chat requests, one dispatcher over %d backend(s)" %
      (n_gpus, chat_users, n_embed, per_req, n_gpus))
print("chat alone     : %6.0f tokens/s, p50 TTFT %.0f ms" % (ntok / a[0], a[2]))
print("embed alone    : %6.0f sequences/s (JSON in, JSON out through the dispatcher)" % (nseq / b[1]))
print("both together  : %6.0f tokens/s, p50 TTFT %.0f ms  |  %6.0f sequences/s" % (ntok / c[0], c[2], nseq / c[1]))
print("per-backend processed:", [d.backend_stats(g)["processed_count"] for g in range(n_gpus)])
d.close()
for e in encs:
    e.close()
for w in wks:
    w.close()

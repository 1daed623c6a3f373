"""GPU: the serving engine (prefill passes, paged KV, CUDA-graphed decode, sampler) through the C ABI against
the model oracle (oracle/llama_ref.py, pinned to HF transformers by tests/golden/llama_tiny.json).

Floating point, bf16 kernels vs fp32 oracle.  Stated tolerance: per-position logits
max|delta| <= 2e-2 * max|logit|, and the engine's greedy token must be an oracle near-argmax: its oracle logit
within 2e-2 * max|logit| of the oracle maximum (exact greedy-token equality with an fp32 run is not claimed)."""
import json
import os
import time

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import ollamamq_b200 as mq  # noqa: E402
from oracle import llama_ref as R  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "llama_tiny.json")))
CFGS = {"tiny_llama": R.TINY_LLAMA, "tiny_qwen": R.TINY_QWEN, "tiny_phi3": R.TINY_PHI3}
MID = dict(vocab=2048, hidden=1024, ffn=2816, n_layers=3, n_q_heads=8, n_kv_heads=2, head_dim=128, qkv_bias=0,
           rope_theta=500000.0, rms_eps=1e-5)
TOL = 2e-2


def _open(cfg, w, **kw):
    args = dict(max_batch=16, max_seq=512, max_prefill_tokens=256, use_graphs=1, use_pdl=0)
    args.update(kw)
    wk = mq.Worker(0, mq.model_cfg(cfg, **args))
    wk.load_weights(w)
    return wk


@pytest.mark.parametrize("i", range(len(GOLD["cases"])))
def test_logits_match_hf_golden(i):
    c = GOLD["cases"][i]
    cfg = CFGS[c["model"]]
    w = R.make_weights(cfg, seed=c["seed"])
    with _open(cfg, w) as wk:
        got = wk.forward_logits(c["tokens"])[0]
    ref = np.array(c["last_logits"], dtype=np.float32)
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= TOL * scale, (np.abs(got - ref).max(), scale)


@pytest.mark.parametrize("pdl,graphs", [(0, 1), (1, 1), (1, 0)])
def test_all_position_logits_and_chunked_prefill(pdl, graphs):
    cfg = MID
    w = R.make_weights(cfg, seed=11, device="cuda")
    toks = torch.randint(0, cfg["vocab"], (300,), generator=torch.Generator().manual_seed(5)).tolist()
    ref = R.forward(w, cfg, toks, torch.float32).cpu().numpy()
    # max_prefill_tokens 128 < 300: three chunks, the later ones attend to cached context
    with _open(cfg, w, max_prefill_tokens=128, use_pdl=pdl, use_graphs=graphs) as wk:
        got = wk.forward_logits(toks, all_positions=True)
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max()
    assert err <= TOL * scale, (err, scale)
    agree = (got.argmax(-1) == ref.argmax(-1)).mean()
    assert agree >= 0.97, agree


def _check_greedy(w, cfg, prompt, gen):
    """Teacher-forced check of an engine generation against the fp32 oracle."""
    seq = list(prompt) + list(gen)
    ref = R.forward(w, cfg, seq[:-1], torch.float32)
    for j, tok in enumerate(gen):
        row = ref[len(prompt) - 1 + j]
        scale = row.abs().max().item()
        gap = (row.max() - row[tok]).item()
        assert gap <= TOL * scale, f"token {j}: engine picked {tok}, oracle gap {gap} (scale {scale})"


@pytest.mark.parametrize("pdl,graphs", [(0, 1), (0, 0), (1, 1)])
def test_generation_single_and_concurrent(pdl, graphs):
    cfg = MID
    w = R.make_weights(cfg, seed=21, device="cuda")
    g = torch.Generator().manual_seed(9)
    prompts = [torch.randint(0, cfg["vocab"], (n,), generator=g).tolist() for n in (5, 64, 17, 130, 1, 33, 200, 48)]
    with _open(cfg, w, max_batch=8, use_pdl=pdl, use_graphs=graphs) as wk:
        solo = wk.generate(prompts[1], 24)
        assert len(solo) == 24
        _check_greedy(w, cfg, prompts[1], solo)
        # 8 concurrent requests with different prompt and generation lengths (ragged batch, slots leaving early)
        lens = [24, 24, 7, 12, 30, 1, 9, 16]
        streams = [wk.submit(mq.Stream(), prompt_tokens=p, max_new_tokens=n) for p, n in zip(prompts, lens)]
        for s in streams:
            s.wait(120)
            assert s.rc == 0, s.err
        for p, n, s in zip(prompts, lens, streams):
            toks = s.tokens()
            assert len(toks) == n
            _check_greedy(w, cfg, p, toks)
            mq.lib.mq_req_release(s.handle)
        st = wk.stats()
        assert st["kernel_launches"] > 0 and st["decode_steps"] > 0
        if graphs:
            assert st["graph_launches"] > 0
        # more requests than slots: the rest wait for a free slot, everything still completes
        streams = [wk.submit(mq.Stream(), prompt_tokens=prompts[i % 8], max_new_tokens=5) for i in range(20)]
        for s in streams:
            s.wait(120)
            assert s.rc == 0 and len(s.tokens()) == 5
            mq.lib.mq_req_release(s.handle)


@pytest.mark.parametrize("cfg_name", ["mid", "qwen_bias_gqa7", "phi_d96", "d64"])
def test_decode_chain_and_plane_path_both_hold_parity(monkeypatch, cfg_name):
    """The decode chain (cluster split-K GEMMs with fused RoPE / residual epilogues, RMSNorm folded into the consumers:
    5 launches per layer) is the default; MQ_DECODE_CHAIN=0 selects the round-1 plane-based path (8 launches per layer).
    Both must hold parity with the fp32 oracle - GQA 4, GQA 7 + q/k/v bias, head_dim 96 MHA, head_dim 64 - at batch
    sizes that take the one-warp and the multi-warp forms of the attention kernel, and the chain must be deterministic."""
    cfg = {"mid": MID,
           "qwen_bias_gqa7": dict(MID, hidden=1024, n_q_heads=7, n_kv_heads=1, qkv_bias=1, rope_theta=1000000.0),
           "phi_d96": dict(MID, hidden=1536, n_q_heads=16, n_kv_heads=16, head_dim=96, rope_theta=10000.0),
           "d64": dict(MID, hidden=512, n_q_heads=8, n_kv_heads=4, head_dim=64, rope_theta=10000.0)}[cfg_name]
    w = R.make_weights(cfg, seed=29, device="cuda")
    g = torch.Generator().manual_seed(5)
    prompts = [torch.randint(0, cfg["vocab"], (n,), generator=g).tolist() for n in (9, 70, 33, 15, 16, 17, 130, 1)]

    def run(chain, order=None):
        monkeypatch.setenv("MQ_DECODE_CHAIN", "1" if chain else "0")
        with _open(cfg, w, max_batch=8, use_pdl=1, use_graphs=1) as wk:
            subs = {i: wk.submit(mq.Stream(), prompt_tokens=prompts[i], max_new_tokens=20) for i in (order or range(len(prompts)))}
            streams = [subs[i] for i in range(len(prompts))]
            toks = []
            for p, s in zip(prompts, streams):
                s.wait(120)
                assert s.rc == 0, s.err
                _check_greedy(w, cfg, p, s.tokens())
                toks.append(s.tokens())
                mq.lib.mq_req_release(s.handle)
            st = wk.stats()
            solo = wk.generate(prompts[1], 12)          # batch of one: 8 warps per attention CTA
            _check_greedy(w, cfg, prompts[1], solo)
            return toks, st["kernel_launches"], st["decode_steps"], st["prefill_passes"]

    def run_one_pass(chain, order=None):
        """Bit-equality between runs is a statement about the kernels, so the runs must batch alike: the eight prompts
        (291 tokens, 256 per pass) make TWO prefill passes when they arrive inside the worker's batching window.  A
        scheduling hiccup of the test process between two submits makes the worker start with a one- or two-prompt pass
        instead: <= 64 tokens take the split-K decode-width GEMMs, whose summation order differs from the prefill
        tiles', and a near-tie argmax can flip (seen once in ~10 runs of this test).  Retry until the run batched as
        designed."""
        for _ in range(5):
            r = run(chain, order)
            if r[3] == 2 and r[2] == 19:   # two prefill passes, then 19 decode steps of all eight slots in lockstep
                return r
        return None

    chain1, chain2 = run_one_pass(True), run_one_pass(True)
    rev = run_one_pass(True, order=list(range(len(prompts)))[::-1])
    toks_plane, n_plane, steps_plane, _ = run(False)
    if chain1 is None or (chain2 is None and rev is None):
        print("host too loaded: the worker never batched the trace as designed twice; bit-equality not checked (parity was)")
    else:
        toks_chain = chain1[0]
        if chain2 is not None:
            assert toks_chain == chain2[0]              # fixed-order reductions: bit-reproducible
        if rev is not None:
            assert toks_chain == rev[0]                 # ... and independent of the slot a sequence lands in
    # three launches per layer fewer in every decode step (compare runs with the same number of steps)
    ref = chain1 if chain1 is not None else run(True)
    if ref[2] == steps_plane:
        assert ref[1] < n_plane


def test_cancel_timeout_and_framing():
    cfg = MID
    w = R.make_weights(cfg, seed=31, device="cuda")
    with _open(cfg, w, max_batch=4, max_seq=2048) as wk:
        # client goes away after 3 chunks: worker cancels, on_done still fires
        n = [0]

        def on_chunk(b):
            n[0] += 1
            return n[0] < 3

        s = wk.submit(mq.Stream(on_chunk=on_chunk), prompt_tokens=[1, 2, 3], max_new_tokens=1500).wait(60)
        assert s.rc == -125 and 3 <= n[0] < 1500
        # whole-request timeout
        s = wk.submit(mq.Stream(), prompt_tokens=[1, 2, 3], max_new_tokens=1900, timeout_ms=30).wait(60)
        assert s.rc == -110
        # NDJSON framing on /api/chat, stream flag taken from the body
        body = json.dumps({"model": "m", "messages": [{"role": "user", "content": "Req 1"}], "stream": True,
                           "options": {"num_predict": 4}}).encode()
        s = wk.submit(mq.Stream(), endpoint=1, body=body, max_new_tokens=0, stream=-1).wait(60)
        assert s.rc == 0 and s.status == 200 and s.content_type == "application/x-ndjson"
        lines = [json.loads(x) for x in s.body.decode().strip().split("\n")]
        assert len(lines) == 5 and lines[-1]["done"] is True and lines[-1]["eval_count"] == 4
        assert all(not x["done"] and x["message"]["role"] == "assistant" for x in lines[:-1])
        # SSE on /v1/chat/completions, non-stream aggregate on /api/generate
        s = wk.submit(mq.Stream(), endpoint=2, body=body, max_new_tokens=3, stream=1).wait(60)
        assert s.content_type == "text/event-stream" and s.body.endswith(b"data: [DONE]\n\n")
        body2 = json.dumps({"model": "m", "prompt": "Req 2", "stream": False}).encode()
        s = wk.submit(mq.Stream(), endpoint=0, body=body2, max_new_tokens=6, stream=-1).wait(60)
        assert s.content_type == "application/json" and len(s.chunks) == 1
        obj = json.loads(s.body)
        assert obj["done"] is True and obj["eval_count"] == 6 and obj["response"]
        assert wk.healthy()


def test_dispatcher_over_gpu_worker_least_connections():
    """Dispatcher + one real worker with capacity 4: fair-share order holds and every stream completes."""
    cfg = MID
    w = R.make_weights(cfg, seed=41, device="cuda")
    with _open(cfg, w, max_batch=4) as wk:
        d = mq.Dispatcher([wk], capacity=4)
        try:
            streams = [d.submit(u, prompt_tokens=[3, 1, 4, 1, 5], max_new_tokens=6)
                       for u in ["alice", "bob", "alice", "carol", "bob", "alice"]]
            d.drain(120000)
            assert all(s.rc == 0 and len(s.tokens()) == 6 for s in streams)
            log = d.log()
            assert sorted(log) == sorted([("alice", 0, 0), ("alice", 1, 0), ("alice", 2, 0), ("bob", 0, 0),
                                          ("bob", 1, 0), ("carol", 0, 0)])
            assert d.user_stats("alice")["processed"] == 3
        finally:
            d.close()


def test_full_size_llama3_8b_logits_and_decode():
    """BASELINE configs[1] at FULL size: Llama-3-8B geometry, weights random-initialised on the device by the
    worker, read back through the C ABI so the fp32 oracle runs on the very same bf16 weights.
    Checks (a) last-position logits of a 512-token prompt, (b) 12 greedy tokens, teacher-forced."""
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~60 GB of HBM")
    cfg = R.LLAMA3_8B
    with mq.Worker(0, mq.model_cfg(cfg, max_batch=16, max_seq=640, max_prefill_tokens=1024, use_graphs=1,
                                   use_pdl=1)) as wk:
        wk.init_random(seed=5, std=0.02)
        w = {}
        for name, shape in R.tensor_shapes(cfg).items():
            w[name] = wk.read_tensor(name, torch.empty(shape, dtype=torch.bfloat16, device="cuda"))
        assert abs(float(w["layers.3.wqkv"].float().std()) - 0.02) < 2e-3       # N(0, 0.02^2) as documented
        prompt = np.random.default_rng(0).integers(0, cfg["vocab"], 512).astype("int32").tolist()
        got = wk.forward_logits(prompt)[0]
        ref = R.forward(w, cfg, prompt, torch.float32)[-1].cpu().numpy()
        cmp16 = R.forward(w, cfg, prompt, torch.bfloat16)[-1].float().cpu().numpy()   # same-precision comparator
        scale = np.abs(ref).max()
        err, err16 = np.abs(got - ref).max(), np.abs(cmp16 - ref).max()
        rel, rel16 = np.linalg.norm(got - ref) / np.linalg.norm(ref), np.linalg.norm(cmp16 - ref) / np.linalg.norm(ref)
        print("full-size: engine max|d| %.4f (%.2f%% of max|logit| %.3f), rel-L2 %.4f; bf16 torch comparator max|d| %.4f, "
              "rel-L2 %.4f" % (err, 100 * err / scale, scale, rel, err16, rel16))
        # 32 layers of bf16 tensor-core inputs: the stated bar at full depth is "no worse than 1.25x a plain bf16
        # torch forward of the same weights, and within 8e-2 * max|logit| of fp32"
        assert rel <= 1.25 * rel16 + 1e-3, (rel, rel16)
        assert err <= 8e-2 * scale, (err, scale)
        assert got.argmax() == ref.argmax() or (ref.max() - ref[got.argmax()]) <= 8e-2 * scale
        gen = wk.generate(prompt, 12)
        assert len(gen) == 12
        seq = prompt + gen
        r = R.forward(w, cfg, seq[:-1], torch.float32)
        for j, tok in enumerate(gen):
            row = r[len(prompt) - 1 + j]
            assert (row.max() - row[tok]).item() <= 8e-2 * row.abs().max().item(), j


def test_http_end_to_end_over_gpu_worker():
    """curl-style requests through the HTTP ingress -> dispatcher -> GPU worker."""
    import http.client
    cfg = MID
    w = R.make_weights(cfg, seed=51, device="cuda")
    with _open(cfg, w, max_batch=4) as wk:
        d = mq.Dispatcher([wk], capacity=4)
        try:
            port = d.serve_http(0)
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
            c.request("GET", "/health")
            r = c.getresponse()
            assert (r.status, r.read()) == (200, b"OK")
            body = json.dumps({"model": "m", "messages": [{"role": "user", "content": "Req 1"}], "stream": True,
                               "options": {"num_predict": 5}})
            c.request("POST", "/api/chat", body=body, headers={"X-User-ID": "alice", "Content-Type": "application/json"})
            r = c.getresponse()
            lines = [json.loads(x) for x in r.read().decode().strip().split("\n")]
            assert r.status == 200 and r.getheader("Content-Type") == "application/x-ndjson"
            assert len(lines) == 6 and lines[-1]["done"] and lines[-1]["eval_count"] == 5
            # OpenAI route, stream omitted -> one JSON document; raw token prompt through "prompt": [ids]
            c.request("POST", "/v1/completions", body=json.dumps({"model": "m", "prompt": [5, 6, 7], "max_tokens": 4}),
                      headers={"X-User-ID": "bob"})
            r = c.getresponse()
            obj = json.loads(r.read())
            assert r.status == 200 and obj["object"] == "text_completion" and obj["usage"]["completion_tokens"] == 4
            c.request("GET", "/api/tags", headers={"X-User-ID": "bob"})
            r = c.getresponse()
            assert r.status == 200 and "models" in json.loads(r.read())
            c.request("POST", "/api/embed", body="{}", headers={"X-User-ID": "bob"})
            r = c.getresponse()
            assert r.status == 501 and r.read()
            c.close()
            assert d.user_stats("alice")["processed"] == 1 and d.user_stats("bob")["processed"] == 3
        finally:
            d.close()


def test_qwen25_7b_geometry_logits():
    """BASELINE configs[2] geometry (Qwen2.5-7B: hidden 3584, 28 q / 4 kv heads -> GQA group 7, q/k/v bias,
    ffn 18944 = 148 weight tiles) with 4 layers: engine logits vs the fp32 oracle on identical weights."""
    cfg = dict(R.QWEN25_7B, n_layers=4, vocab=32768)
    w = R.make_weights(cfg, seed=61, device="cuda")
    toks = torch.randint(0, cfg["vocab"], (200,), generator=torch.Generator().manual_seed(6)).tolist()
    ref = R.forward(w, cfg, toks, torch.float32).cpu().numpy()
    with _open(cfg, w, max_batch=8, max_seq=512, max_prefill_tokens=256) as wk:
        got = wk.forward_logits(toks, all_positions=True)
        scale = np.abs(ref).max()
        err = np.abs(got - ref).max()
        assert err <= TOL * scale, (err, scale)
        gen = wk.generate(toks[:77], 16)
        _check_greedy(w, cfg, toks[:77], gen)


def test_one_dispatcher_over_two_gpu_workers():
    """The product topology of SURVEY.md 8(e): ONE scheduler thread driving one worker per GPU in-process;
    the backend pick (least connections, round-robin tie-break, dispatcher.rs:247-254) spreads the users."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg = MID
    ws = [R.make_weights(cfg, seed=71, device="cuda:%d" % i) for i in range(2)]
    wks = [mq.Worker(i, mq.model_cfg(cfg, max_batch=4, max_seq=512, max_prefill_tokens=256)) for i in range(2)]
    try:
        for wk, w in zip(wks, ws):
            wk.load_weights(w)
        d = mq.Dispatcher(wks, capacity=4)
        try:
            prompts = [[7, 8, 9, i] for i in range(8)]
            streams = [d.submit("user%d" % (i % 4), prompt_tokens=p, max_new_tokens=6) for i, p in enumerate(prompts)]
            d.drain(120000)
            for p, s in zip(prompts, streams):
                assert s.rc == 0 and len(s.tokens()) == 6
                _check_greedy(ws[0], cfg, p, s.tokens())
            backends = [b for _, _, b in d.log()]
            assert sorted(backends) == [0, 0, 0, 0, 1, 1, 1, 1]          # 8 tasks, two workers, 4 slots each
            assert all(d.backend_stats(b)["processed_count"] == 4 for b in range(2))
        finally:
            d.close()
    finally:
        for wk in wks:
            wk.close()


def test_config4_shape_256_users_short_prompts_sse():
    """BASELINE configs[3] load shape: 256 concurrent users, 32-token prompts, SSE on /v1/chat/completions (high-fanout
    TTFT).  Geometry is the mid-size head_dim-128 model (Phi-3's head_dim 96 is not instantiated): the point is the
    256-slot path - decode GEMMs with 256-token tiles, 256 x n_kv attention CTAs, ragged completion."""
    cfg = MID
    w = R.make_weights(cfg, seed=81, device="cuda")
    g = torch.Generator().manual_seed(8)
    prompts = [torch.randint(0, cfg["vocab"], (32,), generator=g).tolist() for _ in range(256)]
    with _open(cfg, w, max_batch=256, max_seq=128, max_prefill_tokens=2048) as wk:
        d = mq.Dispatcher([wk], capacity=256)
        try:
            streams = [d.submit("user%03d" % i, endpoint=2, prompt_tokens=p, max_new_tokens=8 + (i % 5), stream=1)
                       for i, p in enumerate(prompts)]
            d.drain(240000)
            for i, s in enumerate(streams):
                assert s.rc == 0 and s.content_type == "text/event-stream", (i, s.rc, s.err)
                events = [e for e in s.body.decode().split("\n\n") if e]
                assert events[-1] == "data: [DONE]" and len(events) == 8 + (i % 5) + 2
            # parity of a few of them (teacher-forced): token ids are embedded in the SSE deltas as " t<id>"
            import re
            for i in (0, 77, 255):
                toks = [int(x) for x in re.findall(r'"content":" t(\d+)"', streams[i].body.decode())]
                assert len(toks) == 8 + (i % 5)
                _check_greedy(w, cfg, prompts[i], toks)
            st = wk.stats()
            assert st["decode_steps"] > 0 and d.user_stats("user255")["processed"] == 1
        finally:
            d.close()


def test_full_size_phi3_mini_config4_logits_and_256_user_sse():
    """BASELINE configs[3] at FULL size: Phi-3-mini geometry (hidden 3072, 32 MHA heads of head_dim 96, vocab 32 064 =
    250.5 weight tiles, fused qkv / gate_up exactly as HF stores them), 256 users x 32-token prompts x 32 tokens over
    /v1/chat/completions SSE.  Weights are initialised on the device and read back so the fp32 oracle sees the same
    bf16 values."""
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~40 GB of HBM")
    cfg = R.PHI3_MINI
    rng = np.random.default_rng(4)
    prompts = [rng.integers(0, cfg["vocab"], 32).astype("int32").tolist() for _ in range(256)]
    with mq.Worker(0, mq.model_cfg(cfg, max_batch=256, max_seq=320, max_prefill_tokens=4096, use_graphs=1,
                                   use_pdl=1)) as wk:
        wk.init_random(seed=9, std=0.02)
        w = {}
        for name, shape in R.tensor_shapes(cfg).items():
            w[name] = wk.read_tensor(name, torch.empty(shape, dtype=torch.bfloat16, device="cuda"))
        long_prompt = rng.integers(0, cfg["vocab"], 300).astype("int32").tolist()
        got = wk.forward_logits(long_prompt)[0]
        ref = R.forward(w, cfg, long_prompt, torch.float32)[-1].cpu().numpy()
        cmp16 = R.forward(w, cfg, long_prompt, torch.bfloat16)[-1].float().cpu().numpy()
        scale = np.abs(ref).max()
        rel, rel16 = np.linalg.norm(got - ref) / np.linalg.norm(ref), np.linalg.norm(cmp16 - ref) / np.linalg.norm(ref)
        print("phi3-mini full size: engine rel-L2 %.4f, bf16 torch comparator rel-L2 %.4f, max|d| %.4f of max|logit| %.3f"
              % (rel, rel16, np.abs(got - ref).max(), scale))
        assert rel <= 1.25 * rel16 + 1e-3, (rel, rel16)          # same bar as the Llama-3-8B full-size test
        assert np.abs(got - ref).max() <= 8e-2 * scale
        d = mq.Dispatcher([wk], capacity=256)
        try:
            t0 = time.time()
            streams = [d.submit("user%03d" % i, endpoint=2, prompt_tokens=p, max_new_tokens=32, stream=1)
                       for i, p in enumerate(prompts)]
            d.drain(240000)
            dt = time.time() - t0
            import re
            for i, s in enumerate(streams):
                assert s.rc == 0 and s.content_type == "text/event-stream", (i, s.rc, s.err)
                events = [e for e in s.body.decode().split("\n\n") if e]
                assert events[-1] == "data: [DONE]" and len(events) == 32 + 2
            for i in (0, 131, 255):
                toks = [int(x) for x in re.findall(r'"content":" t(\d+)"', streams[i].body.decode())]
                seq = prompts[i] + toks
                r = R.forward(w, cfg, seq[:-1], torch.float32)
                for j, tok in enumerate(toks):
                    row = r[len(prompts[i]) - 1 + j]
                    assert (row.max() - row[tok]).item() <= 8e-2 * row.abs().max().item(), (i, j)
            st = wk.stats()
            print("config 4: 256 users x 32/32 in %.2f s wall (%.0f tok/s incl. python callbacks), decode steps %d"
                  % (dt, 256 * 32 / dt, st["decode_steps"]))
        finally:
            d.close()


def test_dispatcher_wide_timeout_like_the_reference_client_timeout():
    """`--timeout` of the reference (main.rs:31-33 -> reqwest client timeout, dispatcher.rs:165-167): a whole-request limit
    for every dispatched request that does not carry its own."""
    cfg = MID
    w = R.make_weights(cfg, seed=31, device="cuda")
    with _open(cfg, w, max_batch=4, max_seq=4096, max_prefill_tokens=256) as wk:
        d = mq.Dispatcher([wk], capacity=4)
        try:
            d.set_timeout(0.05)                       # 50 ms: far less than 3000 decode steps
            s = d.submit("alice", prompt_tokens=[1, 2, 3], max_new_tokens=3000)
            s.wait(60)
            # the worker ends the request with MQ_ERR_TIMEOUT after the head and some chunks went out; behind the
            # dispatcher that is the reference's mid-stream upstream error: the relay just ends (dispatcher.rs:300-312),
            # the client sees a truncated 200 body and the request counts as processed (:314-316)
            assert s.rc == 0 and s.status == 200 and 0 < len(s.tokens()) < 3000
            d.set_timeout(0)
            s = d.submit("alice", prompt_tokens=[1, 2, 3], max_new_tokens=40)
            s.wait(60)
            assert s.rc == 0 and len(s.tokens()) == 40
            assert d.user_stats("alice")["processed"] == 2 and d.user_stats("alice")["dropped"] == 0
        finally:
            d.close()


def test_soak_random_arrivals_cancels_timeouts_leave_nothing_behind():
    """A few seconds of mixed traffic through dispatcher + worker: ragged prompts (some longer than one prefill pass),
    more users than slots, random client disconnects and per-request timeouts.  Afterwards: every request ended exactly
    once, completed ones hold greedy parity, and the worker is EMPTY - all KV pages back in the pool, no slot in use,
    nothing queued (the leak / lost-wakeup check)."""
    import random
    import threading
    cfg = MID
    w = R.make_weights(cfg, seed=37, device="cuda")
    rnd = random.Random(3)
    with _open(cfg, w, max_batch=8, max_seq=512, max_prefill_tokens=128) as wk:
        total_pages = wk.occupancy()["total_pages"]
        assert wk.occupancy()["free_pages"] == total_pages
        d = mq.Dispatcher([wk], capacity=8)
        try:
            g = torch.Generator().manual_seed(12)
            streams, plan = [], []
            for i in range(120):
                n_prompt = rnd.choice([1, 3, 17, 64, 129, 300])
                n_new = rnd.choice([1, 2, 8, 25, 60])
                kind = rnd.choices(["ok", "cancel", "timeout"], weights=[6, 2, 1])[0]
                p = torch.randint(0, cfg["vocab"], (n_prompt,), generator=g).tolist()
                s = d.submit("user%02d" % rnd.randrange(20), prompt_tokens=p, max_new_tokens=n_new,
                             timeout_ms=3 if kind == "timeout" else 0)
                streams.append(s)
                plan.append((kind, p, n_new))
                if kind == "cancel":
                    threading.Timer(rnd.random() * 0.05, d.client_gone, args=(s.task_id,)).start()
                if i % 10 == 9:
                    time.sleep(rnd.random() * 0.03)
            d.drain(120000)
            done_ok = 0
            for s, (kind, p, n_new) in zip(streams, plan):
                s.wait(30)
                assert s.rc is not None
                toks = s.tokens()
                assert len(toks) <= n_new
                if s.rc == 0:
                    # a timeout that fires mid-stream ends the relay like the reference does: truncated body, rc 0
                    assert len(toks) == n_new or kind == "timeout", (kind, len(toks), n_new)
                    done_ok += 1
                    if done_ok % 7 == 0 and toks:
                        _check_greedy(w, cfg, p, toks)
                else:
                    assert kind in ("cancel", "timeout") and s.rc in (-125, -110), (kind, s.rc, s.err)
            assert done_ok >= 60
            time.sleep(0.2)
            oc = wk.occupancy()
            assert oc["free_pages"] == total_pages and oc["active_slots"] == 0 and oc["waiting"] == 0, oc
            assert oc["in_flight_gpu_passes"] == 0
            assert wk.healthy()
        finally:
            d.close()


def test_sampling_temperature_topk_seed_through_the_engine():
    """The backend's "sample" step: temperature / top_k / top_p / seed per request (struct fields or Ollama "options" in
    the body).  Same seed -> same tokens across runs and batch compositions; other seed -> other tokens; top_k = 1 is
    greedy; every sampled token lies in the oracle's top-k of the teacher-forced fp32 logits."""
    cfg = MID
    w = R.make_weights(cfg, seed=41, device="cuda")
    g = torch.Generator().manual_seed(6)
    prompt = torch.randint(0, cfg["vocab"], (40,), generator=g).tolist()
    other = torch.randint(0, cfg["vocab"], (23,), generator=g).tolist()
    with _open(cfg, w, max_batch=8, use_pdl=1, use_graphs=1) as wk:
        def run(**kw):
            s = wk.submit(mq.Stream(), prompt_tokens=prompt, max_new_tokens=24, **kw)
            s.wait(120)
            assert s.rc == 0, s.err
            t = s.tokens()
            mq.lib.mq_req_release(s.handle)
            return t
        greedy = run()
        assert run(temperature=0.9, top_k=1, seed=5) == greedy
        a = run(temperature=0.9, top_k=40, top_p=0.95, seed=1234)
        b = run(temperature=0.9, top_k=40, top_p=0.95, seed=1234)
        c = run(temperature=0.9, top_k=40, top_p=0.95, seed=99)
        assert a == b and a != c and a != greedy
        # same request inside a batch of others (different slot, different batch size): same stream of tokens
        noise = [wk.submit(mq.Stream(), prompt_tokens=other, max_new_tokens=30, temperature=1.0, seed=i) for i in range(5)]
        d = run(temperature=0.9, top_k=40, top_p=0.95, seed=1234)
        for s in noise:
            s.wait(120)
            mq.lib.mq_req_release(s.handle)
        assert d == a
        # every sampled token is one of the 40 most likely under the oracle's fp32 logits (teacher-forced)
        ref = R.forward(w, cfg, (prompt + a)[:-1], torch.float32)
        for j, tok in enumerate(a):
            row = ref[len(prompt) - 1 + j]
            kth = torch.topk(row, 40).values[-1]
            assert row[tok] >= kth - TOL * row.abs().max(), (j, tok)
        # Ollama-style options in a JSON body
        body = json.dumps({"model": "m", "prompt": "hello", "stream": False,
                           "options": {"temperature": 0.7, "top_k": 20, "seed": 3, "num_predict": 12}}).encode()
        s1 = wk.submit(mq.Stream(), endpoint=0, body=body, max_new_tokens=0, stream=-1); s1.wait(60)
        s2 = wk.submit(mq.Stream(), endpoint=0, body=body, max_new_tokens=0, stream=-1); s2.wait(60)
        s3 = wk.submit(mq.Stream(), endpoint=0, body=body.replace(b'"seed": 3', b'"seed": 4'), max_new_tokens=0, stream=-1)
        s3.wait(60)
        r1, r2, r3 = (json.loads(x.body)["response"] for x in (s1, s2, s3))
        assert r1 == r2 and r1 != r3


def test_eos_token_ends_generation_with_done_reason_stop():
    """cfg.eos_token_id: a request without ignore_eos ends when that token is drawn - it is not relayed, the final
    frame says "stop" instead of "length", and the slot is free again."""
    cfg = MID
    w = R.make_weights(cfg, seed=43, device="cuda")
    prompt = torch.randint(0, cfg["vocab"], (30,), generator=torch.Generator().manual_seed(8)).tolist()
    with _open(cfg, w, max_batch=4) as wk:
        full = wk.generate(prompt, 20)
    stop_at = 7
    eos = full[stop_at]
    assert eos not in full[:stop_at] and eos > 0
    with _open(dict(cfg, eos_token_id=eos), w, max_batch=4) as wk:
        s = wk.submit(mq.Stream(), prompt_tokens=prompt, max_new_tokens=20, ignore_eos=0)
        s.wait(60)
        assert s.rc == 0 and s.tokens() == full[:stop_at]
        mq.lib.mq_req_release(s.handle)
        s = wk.submit(mq.Stream(), prompt_tokens=prompt, max_new_tokens=20, ignore_eos=1)   # benchmark mode runs on
        s.wait(60)
        assert s.tokens() == full
        mq.lib.mq_req_release(s.handle)
        body = json.dumps({"model": "m", "prompt": prompt, "stream": False, "options": {"num_predict": 20}}).encode()
        s = wk.submit(mq.Stream(), endpoint=3, body=body, max_new_tokens=0, stream=-1, ignore_eos=0)
        s.wait(60)
        js = json.loads(s.body)
        assert js["choices"][0]["finish_reason"] == "stop" and js["usage"]["completion_tokens"] == stop_at
        oc = wk.occupancy()
        assert oc["active_slots"] == 0 and oc["free_pages"] == oc["total_pages"]


def test_health_path_on_a_real_worker_probe_failure_and_sticky_fault():
    """dispatcher.rs:171-193 over a GPU worker.  (1) The probe stops answering (mq_debug_worker_set_probe_fail): the
    prober flips is_online, new work is not dispatched to it, the request it has in flight still completes and is
    counted, and when the probe answers again nobody wakes the scheduler - the queued request moves at the next notify.
    (2) A sticky fault (what a CUDA error raises): requests on the worker end with an error value, the backend goes
    offline and stays there."""
    cfg = MID
    w = R.make_weights(cfg, seed=31, device="cuda")
    g = torch.Generator().manual_seed(2)
    prompts = [torch.randint(0, cfg["vocab"], (n,), generator=g).tolist() for n in (40, 33, 21, 60)]
    with _open(cfg, w, max_batch=4, use_pdl=1, use_graphs=1) as wk:
        d = mq.Dispatcher([wk], capacity=4)
        try:
            d.start_health(10)
            s1 = d.submit("alice", prompt_tokens=prompts[0], max_new_tokens=200)
            for _ in range(500):                                   # in flight
                if d.user_stats("alice")["processing"] == 1:
                    break
                time.sleep(0.002)
            mq.check(mq.lib.mq_debug_worker_set_probe_fail(wk._h, 1))
            for _ in range(500):
                if not d.backend_stats(0)["is_online"]:
                    break
                time.sleep(0.005)
            assert not d.backend_stats(0)["is_online"]
            s2 = d.submit("bob", prompt_tokens=prompts[1], max_new_tokens=8)
            s1.wait(120)
            assert s1.rc == 0 and len(s1.tokens()) == 200          # in-flight work completes ...
            _check_greedy(w, cfg, prompts[0], s1.tokens()[:16])
            d.wait_parked()
            assert d.user_stats("alice")["processed"] == 1         # ... and is accounted
            assert d.user_stats("bob")["queued"] == 1              # offline backend: not dispatched (:201-209)
            mq.check(mq.lib.mq_debug_worker_set_probe_fail(wk._h, 0))
            for _ in range(500):
                if d.backend_stats(0)["is_online"]:
                    break
                time.sleep(0.005)
            assert d.backend_stats(0)["is_online"]
            time.sleep(0.1)
            assert d.user_stats("bob")["queued"] == 1              # recovery does not notify run_worker
            s3 = d.submit("carol", prompt_tokens=prompts[2], max_new_tokens=8)   # the next notify does
            for s in (s2, s3):
                s.wait(120)
                assert s.rc == 0 and len(s.tokens()) == 8
            # (2) sticky fault with a request in flight
            s4 = d.submit("dave", prompt_tokens=prompts[3], max_new_tokens=400)
            for _ in range(500):
                if d.user_stats("dave")["processing"] == 1:
                    break
                time.sleep(0.002)
            mq.check(mq.lib.mq_debug_worker_inject_fault(wk._h, b"injected by the test"))
            s4.wait(60)
            assert len(s4.tokens()) < 400                          # cut short (a mid-stream upstream error ends the body, :310), no CPU fallback
            for _ in range(500):
                if not d.backend_stats(0)["is_online"]:
                    break
                time.sleep(0.005)
            assert not d.backend_stats(0)["is_online"] and not wk.healthy()
            s5 = d.submit("erin", prompt_tokens=prompts[1], max_new_tokens=4)
            d.wait_parked()
            assert d.user_stats("erin")["queued"] == 1             # nothing is dispatched to a faulted worker
        finally:
            d.close()

"""GPU: the embedding worker (BASELINE configs[4], /api/embed) against the fp32 BERT oracle (oracle/bert_ref.py, pinned to
HF BertModel by tests/golden/bert_tiny.json).  Tolerance: bf16 tensor-core GEMMs / bf16 activations vs fp32 -
per embedding `max|d| <= 3e-2 * max|e|` and cosine >= 0.999 (rows are unit vectors)."""
import json
import os
import urllib.request

import numpy as np
import pytest

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)
pytestmark = pytest.mark.gpu

import ollamamq_b200 as mq  # noqa: E402
from ollamamq_b200.models import LLAMA3_8B  # noqa: E402,F401
from oracle import bert_ref as B  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bert_tiny.json")))
TOL = 3e-2


def _close(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref).max()
    cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref)))
    assert err <= TOL * np.abs(ref).max() and cos >= 0.999, (err, np.abs(ref).max(), cos)


def _open(cfg, w, **kw):
    e = mq.Encoder(0, mq.encoder_cfg(cfg, **kw))
    e.load_weights(w)
    return e


@pytest.mark.parametrize("i", range(len(GOLD["cases"])))
def test_embedding_matches_hf_golden(i):
    c = GOLD["cases"][i]
    cfg = B.TINY_BERT
    w = B.make_weights(cfg, seed=c["seed"])
    with _open(cfg, w, max_seq=128, max_tokens_per_pass=512) as e:
        got = e.embed([c["tokens"]])[0]
    assert abs(float(np.linalg.norm(got)) - 1.0) < 1e-3
    _close(got, c["embedding"])


@pytest.mark.parametrize("pdl", [0, 1])
def test_ragged_batches_single_and_multi_pass(pdl):
    cfg = dict(B.TINY_BERT, hidden=256, n_heads=8, ffn=512, n_layers=3)
    w = B.make_weights(cfg, seed=5, device="cuda")
    g = torch.Generator().manual_seed(11)
    lens = [1, 2, 15, 16, 17, 63, 64, 65, 100, 128, 3, 77, 128, 31, 5, 90]
    seqs = [torch.randint(0, cfg["vocab"], (n,), generator=g).tolist() for n in lens]
    ref = B.embed(w, cfg, seqs).cpu().numpy()
    with _open(cfg, w, max_seq=128, max_tokens_per_pass=1024, use_pdl=pdl) as e:     # everything in one pass
        one = e.embed(seqs)
        assert e.stats()["passes"] == 1 and e.stats()["sequences"] == len(seqs)
    with _open(cfg, w, max_seq=128, max_tokens_per_pass=256, use_pdl=pdl) as e:      # several passes
        many = e.embed(seqs)
        assert e.stats()["passes"] >= 4
        solo = e.embed([seqs[8]])[0]
    for i in range(len(seqs)):
        _close(one[i], ref[i])
        _close(many[i], ref[i])
    # packing must not leak between sequences: a sequence alone gives the same row as inside a batch
    assert np.abs(solo - many[8]).max() <= 2e-3 and np.abs(one - many).max() <= 2e-3
    # inputs longer than max_seq are truncated, like a backend context limit
    with _open(cfg, w, max_seq=64, max_tokens_per_pass=256) as e:
        _close(e.embed([seqs[9]])[0], B.embed(w, cfg, [seqs[9][:64]])[0].cpu().numpy())


def test_full_size_bge_small_geometry():
    """BASELINE configs[4] geometry at full size (hidden 384, 12 layers, 12 heads of 32, ffn 1536, vocab 30 522):
    weights initialised on the device, read back so the oracle sees the same bf16 values; 512-token inputs."""
    cfg = B.BGE_SMALL
    with mq.Encoder(0, mq.encoder_cfg(cfg, max_seq=512, max_tokens_per_pass=32768)) as e:
        e.init_random(seed=3, std=0.05)
        w = {}
        for name, shape in B.tensor_shapes(cfg).items():
            w[name] = e.read_tensor(name, torch.empty(shape, dtype=torch.bfloat16, device="cuda"))
        # give the LayerNorms and biases non-trivial values too (init_random leaves them at 1 / 0)
        g = torch.Generator().manual_seed(4)
        for name in list(w):
            leaf = name.split(".")[-1]
            if leaf.endswith("ln_g"):
                w[name] = (1.0 + 0.1 * torch.randn(w[name].shape, generator=g)).to(torch.bfloat16).cuda()
            elif leaf.endswith("ln_b") or leaf.startswith("b"):
                w[name] = (0.1 * torch.randn(w[name].shape, generator=g)).to(torch.bfloat16).cuda()
        e.load_weights(w)
        rng = np.random.default_rng(2)
        seqs = [rng.integers(0, cfg["vocab"], n).astype("int32").tolist() for n in (512, 512, 300, 7, 512, 129)]
        got = e.embed(seqs)
        ref = B.embed(w, cfg, seqs).cpu().numpy()
        worst = 0.0
        for i in range(len(seqs)):
            _close(got[i], ref[i])
            worst = max(worst, float(np.abs(got[i] - ref[i]).max() / np.abs(ref[i]).max()))
        print("bge-small full size: worst max|d| / max|e| = %.4f over %d sequences" % (worst, len(seqs)))
        # a config-5 sized slice: 80 x 512 tokens = one 32768-token pass + one 8192-token pass
        batch = [rng.integers(0, cfg["vocab"], 512).astype("int32").tolist() for _ in range(80)]
        p0 = e.stats()["passes"]
        out = e.embed(batch)
        assert out.shape == (80, 384) and np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-3)
        for i in (0, 31, 63, 64, 79):
            _close(out[i], B.embed(w, cfg, [batch[i]])[0].cpu().numpy())
        assert e.stats()["passes"] - p0 == 2


def test_embed_routes_through_dispatcher_and_http():
    """The three embedding routes of main.rs:89-121 end to end: HTTP ingress -> fair-share dispatcher -> the backend's
    embedding worker; without an attached encoder the route answers 501 like any unimplemented backend path."""
    from oracle import llama_ref as R
    MID = dict(vocab=2048, hidden=1024, ffn=2816, n_layers=3, n_q_heads=8, n_kv_heads=2, head_dim=128, qkv_bias=0,
               rope_theta=500000.0, rms_eps=1e-5)
    cfg = B.TINY_BERT
    w = B.make_weights(cfg, seed=7, device="cuda")
    lw = R.make_weights(MID, seed=1, device="cuda")
    with mq.Worker(0, mq.model_cfg(MID, max_batch=4, max_seq=128, max_prefill_tokens=256)) as wk, \
            _open(cfg, w, max_seq=64, max_tokens_per_pass=256) as enc:
        wk.load_weights(lw)
        d = mq.Dispatcher([wk], capacity=4)
        try:
            port = d.serve_http()

            def post(path, body, user="alice"):
                rq = urllib.request.Request("http://127.0.0.1:%d%s" % (port, path), data=json.dumps(body).encode(),
                                            headers={"X-User-ID": user, "Content-Type": "application/json"})
                try:
                    with urllib.request.urlopen(rq, timeout=60) as r:
                        return r.status, r.headers.get("Content-Type"), r.read()
                except urllib.error.HTTPError as ex:
                    return ex.code, ex.headers.get("Content-Type"), ex.read()

            st, _, body = post("/api/embed", {"model": "m", "input": "hello"})
            assert st == 501 and b"not implemented" in body                  # no encoder attached yet
            d.attach_encoder(0, enc)
            st, ct, body = post("/api/embed", {"model": "bge", "input": ["hello world", "a"]})
            assert st == 200 and ct.startswith("application/json")
            js = json.loads(body)
            assert js["model"] == "bge" and len(js["embeddings"]) == 2 and len(js["embeddings"][0]) == cfg["hidden"]
            # byte-level tokenisation restated: [CLS]=1, bytes from 3, [SEP]=2 for this small vocabulary
            toks = [1] + [3 + (b % (cfg["vocab"] - 3)) for b in b"hello world"] + [2]
            _close(js["embeddings"][0], B.embed(w, cfg, [toks])[0].cpu().numpy())
            assert js["prompt_eval_count"] == len(toks) + 3
            st, _, body = post("/v1/embeddings", {"model": "bge", "input": [[5, 6, 7], [9]]})
            js = json.loads(body)
            assert st == 200 and js["object"] == "list" and [x["index"] for x in js["data"]] == [0, 1]
            _close(js["data"][0]["embedding"], B.embed(w, cfg, [[5, 6, 7]])[0].cpu().numpy())
            assert js["usage"]["prompt_tokens"] == 4
            st, _, body = post("/api/embeddings", {"model": "bge", "prompt": "hello world"})
            js = json.loads(body)
            assert st == 200 and len(js["embedding"]) == cfg["hidden"]
            _close(js["embedding"], B.embed(w, cfg, [toks])[0].cpu().numpy())
            # chat still works next to it on the same backend, and the embed requests were accounted to the user
            st, _, body = post("/api/chat", {"model": "m", "messages": [{"role": "user", "content": "hi"}],
                                             "stream": False, "options": {"num_predict": 3}})
            assert st == 200 and json.loads(body)["done"] is True
            assert d.user_stats("alice")["processed"] == 5
        finally:
            d.close()


# ---------------------------------------------------------------------------------------------------------------
# kernel level: the round-2 encoder kernels against plain torch fp32 references of the same op
# ---------------------------------------------------------------------------------------------------------------
import ctypes as C  # noqa: E402


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@pytest.mark.parametrize("T,n_out,K,epi,with_bias", [
    (32768, 1536, 384, 3, True),     # bge-small up projection: bias + erf-GELU, 6 k-blocks per tile
    (32768, 1152, 384, 4, True),     # QKV: ragged last 256-feature tile (1152 = 4.5 tiles)
    (4096, 384, 1536, 4, False),     # down projection: 384 features = 1.5 tiles, no bias (added by the LayerNorm kernel)
    (777, 384, 384, 4, True),        # ragged token tile
    (300, 1000, 512, 3, True),       # features not a multiple of 128
    (129, 256, 128, 3, True),        # smallest shape on the 2-CTA kernel
    (100, 384, 384, 3, True),        # T <= 128: the one-tile-per-CTA kernel, same epilogue semantics
])
def test_encoder_gemm_bias_gelu_epilogues(T, n_out, K, epi, with_bias):
    g = torch.Generator(device="cuda").manual_seed(T + n_out + K + epi)
    rows = (T + 255) // 256 * 256
    W = (torch.randn(n_out, K, device="cuda", generator=g) * 0.05).bfloat16()
    X = torch.randn(rows, K, device="cuda", generator=g).bfloat16()
    bias = (torch.randn(n_out, device="cuda", generator=g) * 0.5).bfloat16() if with_bias else None
    out = torch.full((T, n_out), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = mq.lib.mq_debug_gemm_bias(_p(W), n_out, K, _p(X), rows, T, epi, _p(bias), _p(out), n_out)
    assert rc == 0, mq.last_error()
    ref = X[:T].float() @ W.float().T
    if with_bias:
        ref = ref + bias.float()
    if epi == 3:
        ref = torch.nn.functional.gelu(ref)          # exact erf form, like BertIntermediate
    assert torch.isfinite(out.float()).all(), "unwritten / non-finite outputs"
    err = (out.float() - ref).abs().max().item()
    assert err <= 2 ** -7 * ref.abs().max().item() + 1e-3, (err, ref.abs().max().item())
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel < 5e-3, rel


@pytest.mark.parametrize("lens,H,heads", [
    ([512] * 6, 384, 12),                                   # bge-small
    ([1, 2, 63, 64, 65, 127, 128, 129, 255, 256, 257, 300, 511, 512, 7, 77], 384, 12),
    ([200, 33, 512, 5], 256, 8),
])
def test_encoder_attention_tcgen05_vs_torch(lens, H, heads):
    """Bidirectional softmax(Q K^T / sqrt(32)) V per sequence and head, packed q | k | v rows; rows of OTHER sequences and
    stale rows behind the last one are loaded by the 64-row TMA boxes and must not leak in."""
    g = torch.Generator(device="cuda").manual_seed(sum(lens) + H)
    T = sum(lens)
    rows = (T + 255) // 256 * 256 + 256
    qkv = torch.randn(rows, 3 * H, device="cuda", generator=g).bfloat16()     # rows >= T: finite garbage
    first = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    ln = np.asarray(lens, dtype=np.int32)
    out = torch.full((T, H), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = mq.lib.mq_debug_enc_attn(_p(qkv), rows, H, heads, first.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p),
                                  len(lens), _p(out))
    assert rc == 0, mq.last_error()
    d = H // heads
    worst = 0.0
    for s, n in enumerate(lens):
        blk = qkv[first[s]: first[s] + n].float()
        q, k, v = (blk[:, i * H:(i + 1) * H].reshape(n, heads, d).transpose(0, 1) for i in range(3))
        ref = torch.softmax(q @ k.transpose(1, 2) / d ** 0.5, dim=-1) @ v              # [heads, n, d]
        ref = ref.transpose(0, 1).reshape(n, H)
        got = out[first[s]: first[s] + n].float()
        assert torch.isfinite(got).all()
        worst = max(worst, ((got - ref).abs().max() / ref.abs().max()).item())
    assert worst < 2e-2, worst     # bf16 probabilities (2^-9 each) and a bf16 result


@pytest.mark.parametrize("T,n_out,K,with_bias,want_h32", [
    (32768, 384, 384, True, False),     # bge-small attention output
    (4096, 384, 1536, True, True),      # bge-small MLP output (+ the fp32 copy of the model's last LayerNorm)
    (777, 384, 384, False, False),      # ragged last token tile
    (300, 256, 512, True, True),        # one MMA per k-step
    (1000, 256, 256, True, False),
    (1, 384, 64, True, False),
])
def test_projection_residual_layernorm_in_one_kernel(T, n_out, K, with_bias, want_h32):
    """x <- LayerNorm(x + X W^T + bias) * gamma + beta (BertSelfOutput / BertOutput), tokens on the TMEM lanes; rows behind T
    of the (128-row padded) operand are garbage and must leave x untouched."""
    g = torch.Generator(device="cuda").manual_seed(T + n_out + K)
    rows = (T + 255) // 256 * 256
    W = (torch.randn(n_out, K, device="cuda", generator=g) * 0.05).bfloat16()
    X = torch.randn(rows, K, device="cuda", generator=g).bfloat16()
    x0 = torch.randn(rows, n_out, device="cuda", generator=g).bfloat16()
    bias = (torch.randn(n_out, device="cuda", generator=g) * 0.3).bfloat16() if with_bias else None
    gamma = (1 + 0.2 * torch.randn(n_out, device="cuda", generator=g)).bfloat16()
    beta = (0.2 * torch.randn(n_out, device="cuda", generator=g)).bfloat16()
    eps = 1e-12
    x = x0.clone()
    h32 = torch.full((T, n_out), float("nan"), device="cuda") if want_h32 else None
    rc = mq.lib.mq_debug_gemm_rowln(_p(W), n_out, K, _p(X), rows, T, _p(x), _p(bias), _p(gamma), _p(beta), eps, _p(h32))
    assert rc == 0, mq.last_error()
    pre = x0[:T].float() + X[:T].float() @ W.float().T + (bias.float() if with_bias else 0.0)
    ref = torch.nn.functional.layer_norm(pre, (n_out,), gamma.float(), beta.float(), eps)
    got = x[:T].float()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 2e-3
    assert ((got - ref).norm() / ref.norm()).item() < 5e-3
    assert torch.equal(x[T:], x0[T:]), "rows behind T were written"
    if want_h32:
        assert torch.isfinite(h32).all() and (h32 - ref).abs().max().item() <= 4e-3 * ref.abs().max().item() + 1e-3
        assert torch.equal(h32.bfloat16(), x[:T])

"""CPU: the request parsers and response framers of the worker (csrc/framing.cpp) through their host-only test ABI -
the JSON a client sends to the reference is opaque there (it is relayed verbatim, dispatcher.rs:287-290); here the
worker terminates it, so the parser has to survive anything the HTTP ingress lets through."""
import ctypes as C
import json

import numpy as np
from hypothesis import given, settings, strategies as st

import ollamamq_b200 as mq


def _call(fn, *args):
    need = fn(*args, None, 0)
    buf = C.create_string_buffer(int(need) + 16)
    assert fn(*args, buf, len(buf)) == need
    return buf.value.decode("utf-8", "replace")


def parse(endpoint, body: bytes, vocab=2048):
    return json.loads(_call(mq.lib.mq_debug_parse_body, endpoint, body, len(body), vocab))


def parse_embed(body: bytes, vocab=30522, max_len=512):
    return json.loads(_call(mq.lib.mq_debug_parse_embed, body, len(body), vocab, max_len))


def test_chat_and_generate_bodies():
    p = parse(1, json.dumps({"model": "llama3", "messages": [{"role": "system", "content": "a"}, {"role": "user", "content": [
        {"type": "text", "text": "bé"}, {"type": "image_url", "image_url": {"url": "x"}}]}], "stream": False,
        "options": {"num_predict": 7, "temperature": 0.7, "top_k": 40, "top_p": 0.9, "seed": 42}}).encode())
    assert p["ok"] and p["model"] == "llama3" and p["has_stream"] and not p["stream"] and p["num_predict"] == 7
    assert p["has_temperature"] and abs(p["temperature"] - 0.7) < 1e-9 and p["top_k"] == 40 and p["seed"] == 42
    assert abs(p["top_p"] - 0.9) < 1e-9
    assert bytes(p["tokens"]) == b"a\nb\xc3\xa9\n"                   # byte-level tokens of the concatenated contents
    p = parse(3, b'{"model":"m","prompt":[5,6,7],"max_tokens":3,"temperature":1.5,"seed":9,"unknown":{"a":[1,{"b":null}]}}')
    assert p["ok"] and p["tokens"] == [5, 6, 7] and p["num_predict"] == 3 and p["temperature"] == 1.5 and p["seed"] == 9
    assert not p["has_top_k"] and not p["has_stream"]
    p = parse(0, b'{"prompt":"hi \\"there\\"\\n","context":[1,2]}')
    assert p["ok"] and p["tokens"] == [1, 2]
    assert not parse(0, b'["not an object"]')["ok"] and not parse(0, b'')["ok"] and not parse(0, b'{"prompt": "unterminated')["ok"]


def test_embed_bodies_and_reply_shapes():
    e = parse_embed(b'{"model":"bge","input":["ab","c"]}')
    assert e["ok"] and e["model"] == "bge" and e["seqs"] == [[101, 1097, 1098, 102], [101, 1099, 102]]
    assert parse_embed(b'{"input":"ab"}')["seqs"] == [[101, 1097, 1098, 102]]
    assert parse_embed(b'{"prompt":"ab"}')["seqs"] == [[101, 1097, 1098, 102]]          # /api/embeddings
    assert parse_embed(b'{"input":[[5,6],[7]]}')["seqs"] == [[5, 6], [7]] and parse_embed(b'{"input":[5,6]}')["seqs"] == [[5, 6]]
    assert parse_embed(b'{"input":[]}')["seqs"] == [] and not parse_embed(b'nope')["ok"]
    long = parse_embed(json.dumps({"input": "x" * 1000}).encode(), max_len=16)["seqs"][0]
    assert len(long) == 16 and long[0] == 101                                             # truncated to the context
    small = parse_embed(b'{"input":"a"}', vocab=512)["seqs"][0]
    assert small == [1, 3 + ord("a"), 2]                                                  # small vocabularies: CLS 1, SEP 2
    emb = np.arange(6, dtype=np.float32).reshape(2, 3) / 8
    ptr = emb.ctypes.data_as(C.c_void_p)
    a = json.loads(_call(mq.lib.mq_debug_frame_embeddings, b"/api/embed", b'm"x', ptr, 2, 3, 9))
    assert a == {"model": 'm"x', "embeddings": emb.tolist(), "prompt_eval_count": 9}
    b = json.loads(_call(mq.lib.mq_debug_frame_embeddings, b"/v1/embeddings", b"m", ptr, 2, 3, 9))
    assert b["object"] == "list" and [d["index"] for d in b["data"]] == [0, 1] and b["data"][1]["embedding"] == emb[1].tolist()
    assert b["usage"] == {"prompt_tokens": 9, "total_tokens": 9}
    c = json.loads(_call(mq.lib.mq_debug_frame_embeddings, b"/api/embeddings", b"m", ptr, 2, 3, 9))
    assert c == {"embedding": emb[0].tolist()}


def test_final_frames_carry_the_stop_reason():
    f = lambda *a: _call(mq.lib.mq_debug_frame_final, *a)
    j = json.loads(f(1, 0, b"m", b" t1 t2", 5, 2, 1))
    assert j["done"] and j["done_reason"] == "stop" and j["message"]["content"] == " t1 t2" and j["eval_count"] == 2
    assert json.loads(f(0, 1, b"m", b"", 5, 2, 0))["done_reason"] == "length"
    sse = f(2, 1, b"m", b"", 5, 2, 1)
    assert sse.endswith("data: [DONE]\n\n") and '"finish_reason":"stop"' in sse
    assert json.loads(f(3, 0, b"m", b"x", 5, 2, 0))["choices"][0]["finish_reason"] == "length"


@settings(max_examples=300, deadline=None)
@given(st.binary(max_size=400))
def test_parsers_survive_arbitrary_bytes(blob):
    for ep in (0, 1, 2, 3):
        out = parse(ep, blob)
        assert isinstance(out["ok"], bool) and all(0 <= t < 2048 for t in out["tokens"])
    assert isinstance(parse_embed(blob)["ok"], bool)


_json = st.recursive(st.none() | st.booleans() | st.integers(-10**6, 10**6) | st.floats(allow_nan=False, allow_infinity=False) |
                     st.text(max_size=12), lambda c: st.lists(c, max_size=4) | st.dictionaries(st.text(max_size=6), c, max_size=4),
                     max_leaves=12)


@settings(max_examples=200, deadline=None)
@given(st.dictionaries(st.sampled_from(["model", "prompt", "messages", "stream", "options", "input", "seed", "top_p",
                                        "temperature", "max_tokens", "context", "x"]), _json, max_size=6))
def test_parsers_survive_well_formed_json_of_any_shape(obj):
    body = json.dumps(obj).encode()
    assert isinstance(parse(1, body)["ok"], bool)
    out = parse_embed(body)
    assert isinstance(out["ok"], bool) and all(isinstance(s, list) for s in out["seqs"])


def test_round1_advisor_inputs_fail_fast_instead_of_hanging_or_crashing():
    """The two inputs of ADVICE.md (round 1): a value position holding '}' used to spin the parser forever, and a body
    nested two million levels deep used to overflow the stack.  Both now return ok = false at once; so do near misses."""
    import time
    t0 = time.time()
    assert not parse(1, b'{"messages":[{"content":[}')["ok"]
    assert not parse(1, b'{"x":' + b"[" * 2_000_000)["ok"]
    assert not parse(1, b'{"x":' + b'{"a":' * 500_000)["ok"]
    assert not parse_embed(b'{"input":' + b"[" * 1_000_000)["ok"]
    for bad in (b'{"a":}', b'{"a":1 "b":2}', b'{"a":[1 2]}', b'{"a":[1,]}', b'{"a":1,}', b'{"a":1}x', b'{"a":"\\u12G4"}',
                b'{"a":"\\q"}', b'{"messages":[{"content":[{"text":}]}]}', b'{,}', b'{"a"}', b'{"a":nul'):
        assert not parse(0, bad)["ok"], bad
    assert time.time() - t0 < 5.0
    # depth 64 is the limit; 60 levels of a skipped value still parse
    assert parse(0, b'{"x":' + b"[" * 60 + b"]" * 60 + b',"prompt":"ok"}')["ok"]
    assert not parse(0, b'{"x":' + b"[" * 70 + b"]" * 70 + b"}")["ok"]


def test_numbers_are_clamped_before_they_are_narrowed():
    p = parse(0, b'{"prompt":"a","options":{"num_predict":1e300,"top_k":1e300,"seed":-5,"temperature":-3,"top_p":7}}')
    assert p["ok"] and p["num_predict"] == 1073741824 and p["top_k"] == 1073741824 and p["seed"] == 0
    assert p["temperature"] == 0 and p["top_p"] == 1
    p = parse(0, b'{"prompt":"a","options":{"num_predict":nan,"temperature":inf}}')     # not JSON numbers: skipped as bare words
    assert p["ok"] and p["num_predict"] == 0 and not p["has_temperature"]
    p = parse(0, b'{"prompt":[1e300,-1e300,3]}')
    assert p["ok"] and p["tokens"] == [2147483647, -2147483647, 3]   # clamped, not UB; mq_submit maps ids outside the vocabulary to 0


def test_unicode_escapes_and_long_model_names():
    p = parse(0, '{"prompt":"\\ud83d\\ude00é\\u00e9"}'.encode())
    assert p["ok"] and bytes(p["tokens"]) == "\U0001F600éé".encode()   # surrogate pair combined: valid UTF-8
    f = _call(mq.lib.mq_debug_frame_final, 1, 0, b"m" * 5000, b"x", 1, 1, 0)
    j = json.loads(f)                                                  # a 5 000-byte model name no longer truncates the frame
    assert j["done"] and len(j["model"]) == 256


def test_embedding_floats_round_trip_float32_exactly():
    """frame_embeddings writes 9 significant digits through integer arithmetic (no printf on the embedding worker's
    thread): every finite float32 must parse back to the same float32; NaN / inf have no JSON spelling and become 0."""
    rng = np.random.default_rng(7)
    special = np.array([0, 1, -1, 0.1, 0.5, 1e-5, 9.9999999e-6, 123456789.0, 999999999.0, 1e9, 1e-7, 3.4e38, -2.5e-3,
                        0.99999994, 1.0000001, 7.0, 1234.5678, 1e-45, 16777216.0, 99999.9921875, 0.001, 0.01, 10.0, 100.0,
                        99999999.0, 1e8], dtype=np.float32)
    rnd = (rng.standard_normal(50000) * 10.0 ** rng.integers(-7, 11, 50000)).astype(np.float32)
    unit = rng.standard_normal((8, 384)).astype(np.float32)
    unit /= np.linalg.norm(unit, axis=1, keepdims=True)
    v = np.concatenate([special, rnd, unit.reshape(-1)]).reshape(1, -1)
    buf = C.create_string_buffer(v.size * 30 + 256)
    n = mq.lib.mq_debug_frame_embeddings(b"/api/embeddings", b"m", v.ctypes.data_as(C.c_void_p), 1, v.shape[1], 1, buf, len(buf))
    assert 0 < n <= len(buf)
    got = np.array(json.loads(buf.value)["embedding"], dtype=np.float32)
    assert np.array_equal(got, v[0])
    bad = np.array([[np.nan, np.inf, -np.inf, 0.25]], dtype=np.float32)
    n = mq.lib.mq_debug_frame_embeddings(b"/api/embeddings", b"m", bad.ctypes.data_as(C.c_void_p), 1, 4, 1, buf, len(buf))
    assert json.loads(buf.value) == {"embedding": [0, 0, 0, 0.25]}

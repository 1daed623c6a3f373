"""llama_ref.py — TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product path).

Plain-torch restatement of the decoder-only transformer forward pass that the reference delegates to an
external Ollama / llama.cpp server (call site /root/reference/src/dispatcher.rs:287-290).  That dependency is
NOT in /root/reference, not in Cargo.lock and not pinned to any version, so the published
algorithm restated here is the Llama architecture as implemented by HuggingFace transformers 5.5.0
(models/llama/modeling_llama.py: LlamaRMSNorm, rotate_half / apply_rotary_pos_emb, LlamaAttention with
repeat_kv GQA, LlamaMLP SwiGLU; models/qwen2 adds a bias on q/k/v; models/phi3 is the same arithmetic with the
q/k/v and gate/up matrices stored fused - exactly the layout used here - and head_dim 96).

Pinned against: HF LlamaForCausalLM / Qwen2ForCausalLM / Phi3ForCausalLM outputs on seeded random-init weights
(tests/golden/llama_tiny.json, generated in the build container by tests/golden/make_llama_golden.py).
The reference itself pins nothing at this boundary.

Weight naming (same as the C ABI, include/ollamamq_b200.h): embed, final_norm, lm_head,
layers.<i>.{attn_norm, wqkv, bqkv, wo, mlp_norm, w_gate_up, w_down}; torch Linear [out, in] layout.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch


def tensor_shapes(cfg: dict) -> Dict[str, tuple]:
    H, I, V, L = cfg["hidden"], cfg["ffn"], cfg["vocab"], cfg["n_layers"]
    D, nq, nkv = cfg["head_dim"], cfg["n_q_heads"], cfg["n_kv_heads"]
    qkv = (nq + 2 * nkv) * D
    out = {"embed": (V, H), "final_norm": (H,), "lm_head": (V, H)}
    for l in range(L):
        p = f"layers.{l}."
        out[p + "attn_norm"] = (H,)
        out[p + "wqkv"] = (qkv, H)
        if cfg.get("qkv_bias"):
            out[p + "bqkv"] = (qkv,)
        out[p + "wo"] = (H, nq * D)
        out[p + "mlp_norm"] = (H,)
        out[p + "w_gate_up"] = (2 * I, H)
        out[p + "w_down"] = (H, I)
    return out


def make_weights(cfg: dict, seed: int = 0, std: float = 0.02, device="cpu") -> Dict[str, torch.Tensor]:
    """Seeded random-init bf16 weights (N(0, std^2); norm gains 1 + 0.1 N(0,1))."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = {}
    for name, shape in tensor_shapes(cfg).items():
        if name.endswith("norm"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bqkv"):
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            t = std * torch.randn(shape, generator=g)
        w[name] = t.to(torch.bfloat16).to(device)
    return w


def _rmsnorm(x, g, eps):
    # LlamaRMSNorm: variance in fp32, x * rsqrt(var + eps), times weight
    v = x.pow(2).mean(-1, keepdim=True)
    return x * torch.rsqrt(v + eps) * g


def _rope(x, pos, theta):
    # x [T, heads, D]; rotate_half convention of modeling_llama.apply_rotary_pos_emb
    D = x.shape[-1]
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.float32, device=x.device) / D))
    ang = pos.to(torch.float32)[:, None] * inv[None, :]
    emb = torch.cat([ang, ang], -1)
    cos, sin = emb.cos()[:, None, :].to(x.dtype), emb.sin()[:, None, :].to(x.dtype)
    x1, x2 = x[..., : D // 2], x[..., D // 2:]
    return x * cos + torch.cat([-x2, x1], -1) * sin


@torch.no_grad()
def forward(w: Dict[str, torch.Tensor], cfg: dict, tokens, dtype=torch.float32,
            start_pos: int = 0, kv: Optional[List] = None, last_only: bool = False) -> torch.Tensor:
    """Logits [T, vocab] for a token sequence ([1, vocab] with last_only).  With `kv` (list of per-layer [k, v])
    the call appends to the cache and attends over the whole context (used for step-by-step greedy decoding)."""
    dev = next(iter(w.values())).device
    tokens = torch.as_tensor(tokens, dtype=torch.long, device=dev)
    T = tokens.shape[0]
    H, D, nq, nkv = cfg["hidden"], cfg["head_dim"], cfg["n_q_heads"], cfg["n_kv_heads"]
    I = cfg["ffn"]
    G = nq // nkv
    eps, theta = cfg.get("rms_eps", 1e-5), cfg.get("rope_theta", 500000.0)
    pos = torch.arange(start_pos, start_pos + T, device=dev)
    h = w["embed"].to(dtype)[tokens]
    for l in range(cfg["n_layers"]):
        p = f"layers.{l}."
        x = _rmsnorm(h, w[p + "attn_norm"].to(dtype), eps)
        qkv = x @ w[p + "wqkv"].to(dtype).T
        if cfg.get("qkv_bias"):
            qkv = qkv + w[p + "bqkv"].to(dtype)
        q = qkv[:, : nq * D].view(T, nq, D)
        k = qkv[:, nq * D:(nq + nkv) * D].view(T, nkv, D)
        v = qkv[:, (nq + nkv) * D:].view(T, nkv, D)
        q, k = _rope(q, pos, theta), _rope(k, pos, theta)
        if kv is not None:
            if len(kv) <= l:
                kv.append([k, v])
            else:
                kv[l][0] = torch.cat([kv[l][0], k], 0)
                kv[l][1] = torch.cat([kv[l][1], v], 0)
            k, v = kv[l]
        kk = k.repeat_interleave(G, dim=1)
        vv = v.repeat_interleave(G, dim=1)
        s = torch.einsum("qhd,khd->hqk", q, kk) / math.sqrt(D)
        kpos = torch.arange(k.shape[0], device=dev)
        s = s.masked_fill(~(kpos[None, :] <= pos[:, None])[None], float("-inf"))
        a = torch.einsum("hqk,khd->qhd", torch.softmax(s.float(), -1).to(dtype), vv).reshape(T, nq * D)
        h = h + a @ w[p + "wo"].to(dtype).T
        x = _rmsnorm(h, w[p + "mlp_norm"].to(dtype), eps)
        gu = x @ w[p + "w_gate_up"].to(dtype).T
        h = h + (torch.nn.functional.silu(gu[:, :I]) * gu[:, I:]) @ w[p + "w_down"].to(dtype).T
    if last_only:
        h = h[-1:]
    x = _rmsnorm(h, w["final_norm"].to(dtype), eps)
    return x @ w["lm_head"].to(dtype).T


@torch.no_grad()
def greedy(w, cfg, prompt, n_new: int, dtype=torch.float32) -> List[int]:
    kv: List = []
    logits = forward(w, cfg, prompt, dtype, 0, kv)
    out = [int(logits[-1].argmax())]
    for i in range(n_new - 1):
        logits = forward(w, cfg, [out[-1]], dtype, len(prompt) + i, kv)
        out.append(int(logits[-1].argmax()))
    return out


def to_hf_state_dict(w: Dict[str, torch.Tensor], cfg: dict) -> Dict[str, torch.Tensor]:
    """Map our fused tensors onto HF Llama/Qwen2 parameter names (used only to pin this oracle)."""
    D, nq, nkv, I = cfg["head_dim"], cfg["n_q_heads"], cfg["n_kv_heads"], cfg["ffn"]
    sd = {"model.embed_tokens.weight": w["embed"], "model.norm.weight": w["final_norm"], "lm_head.weight": w["lm_head"]}
    for l in range(cfg["n_layers"]):
        p, hp = f"layers.{l}.", f"model.layers.{l}."
        qkv = w[p + "wqkv"]
        sd[hp + "self_attn.q_proj.weight"] = qkv[: nq * D]
        sd[hp + "self_attn.k_proj.weight"] = qkv[nq * D:(nq + nkv) * D]
        sd[hp + "self_attn.v_proj.weight"] = qkv[(nq + nkv) * D:]
        if cfg.get("qkv_bias"):
            b = w[p + "bqkv"]
            sd[hp + "self_attn.q_proj.bias"] = b[: nq * D]
            sd[hp + "self_attn.k_proj.bias"] = b[nq * D:(nq + nkv) * D]
            sd[hp + "self_attn.v_proj.bias"] = b[(nq + nkv) * D:]
        sd[hp + "self_attn.o_proj.weight"] = w[p + "wo"]
        sd[hp + "input_layernorm.weight"] = w[p + "attn_norm"]
        sd[hp + "post_attention_layernorm.weight"] = w[p + "mlp_norm"]
        sd[hp + "mlp.gate_proj.weight"] = w[p + "w_gate_up"][:I]
        sd[hp + "mlp.up_proj.weight"] = w[p + "w_gate_up"][I:]
        sd[hp + "mlp.down_proj.weight"] = w[p + "w_down"]
    return sd


def to_hf_phi3_state_dict(w: Dict[str, torch.Tensor], cfg: dict) -> Dict[str, torch.Tensor]:
    """HF Phi3 keeps qkv_proj = [q; k; v] and gate_up_proj = [gate; up] fused (modeling_phi3.py Phi3Attention /
    Phi3MLP), i.e. our wqkv / w_gate_up verbatim."""
    sd = {"model.embed_tokens.weight": w["embed"], "model.norm.weight": w["final_norm"], "lm_head.weight": w["lm_head"]}
    for l in range(cfg["n_layers"]):
        p, hp = f"layers.{l}.", f"model.layers.{l}."
        sd[hp + "self_attn.qkv_proj.weight"] = w[p + "wqkv"]
        sd[hp + "self_attn.o_proj.weight"] = w[p + "wo"]
        sd[hp + "input_layernorm.weight"] = w[p + "attn_norm"]
        sd[hp + "post_attention_layernorm.weight"] = w[p + "mlp_norm"]
        sd[hp + "mlp.gate_up_proj.weight"] = w[p + "w_gate_up"]
        sd[hp + "mlp.down_proj.weight"] = w[p + "w_down"]
    return sd


# named model geometries (BASELINE.json configs)
LLAMA3_8B = dict(vocab=128256, hidden=4096, ffn=14336, n_layers=32, n_q_heads=32, n_kv_heads=8, head_dim=128,
                 qkv_bias=0, rope_theta=500000.0, rms_eps=1e-5)
QWEN25_7B = dict(vocab=152064, hidden=3584, ffn=18944, n_layers=28, n_q_heads=28, n_kv_heads=4, head_dim=128,
                 qkv_bias=1, rope_theta=1000000.0, rms_eps=1e-6)
TINY_LLAMA = dict(vocab=512, hidden=512, ffn=1024, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=128,
                  qkv_bias=0, rope_theta=500000.0, rms_eps=1e-5)
TINY_QWEN = dict(vocab=512, hidden=512, ffn=1024, n_layers=2, n_q_heads=4, n_kv_heads=1, head_dim=128,
                 qkv_bias=1, rope_theta=1000000.0, rms_eps=1e-6)
PHI3_MINI = dict(vocab=32064, hidden=3072, ffn=8192, n_layers=32, n_q_heads=32, n_kv_heads=32, head_dim=96,
                 qkv_bias=0, rope_theta=10000.0, rms_eps=1e-5, family="phi3")
TINY_PHI3 = dict(vocab=512, hidden=1536, ffn=1024, n_layers=2, n_q_heads=16, n_kv_heads=16, head_dim=96,
                 qkv_bias=0, rope_theta=10000.0, rms_eps=1e-5, family="phi3")

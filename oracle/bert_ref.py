"""bert_ref.py - TEST INFRASTRUCTURE ONLY (imported by tests/ and __graft_entry__.smoke(); never by the product path).

Plain-torch restatement of the BERT encoder + CLS pooling that serves `/api/embed` (BASELINE.json configs[4]:
bge-small = BERT, hidden 384, 12 layers, 12 heads of 32, ffn 1536, vocab 30 522).  Like the decoder, this arithmetic
lives in the external Ollama / llama.cpp server the reference forwards to (call site
/root/reference/src/dispatcher.rs:287-290) and is not pinned to any version there; the published
algorithm restated is HuggingFace transformers 5.5.0 models/bert/modeling_bert.py: BertEmbeddings (word + token-type 0
+ absolute position, LayerNorm), BertSelfAttention (bidirectional, scale 1/sqrt(d)), BertSelfOutput / BertOutput
(dense + residual + post-LayerNorm), BertIntermediate (exact erf GELU); sentence embedding = L2-normalised [CLS] state
(the bge recipe).

Pinned against: HF BertModel outputs on seeded random-init weights (tests/golden/bert_tiny.json, generated in the build
container by tests/golden/make_bert_golden.py).

Weight naming (C ABI, include/ollamamq_b200.h section 2b): word_embed, pos_embed, type_embed, emb_ln_g, emb_ln_b,
layers.<i>.{wqkv, bqkv, wo, bo, attn_ln_g, attn_ln_b, w_up, b_up, w_down, b_down, mlp_ln_g, mlp_ln_b};
torch Linear [out, in] layout, q / k / v rows stacked in wqkv.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch


def tensor_shapes(cfg: dict) -> Dict[str, tuple]:
    H, I, V, L, P = cfg["hidden"], cfg["ffn"], cfg["vocab"], cfg["n_layers"], cfg["max_positions"]
    out = {"word_embed": (V, H), "pos_embed": (P, H), "type_embed": (cfg.get("type_vocab", 2), H),
           "emb_ln_g": (H,), "emb_ln_b": (H,)}
    for l in range(L):
        p = f"layers.{l}."
        out.update({p + "wqkv": (3 * H, H), p + "bqkv": (3 * H,), p + "wo": (H, H), p + "bo": (H,),
                    p + "attn_ln_g": (H,), p + "attn_ln_b": (H,), p + "w_up": (I, H), p + "b_up": (I,),
                    p + "w_down": (H, I), p + "b_down": (H,), p + "mlp_ln_g": (H,), p + "mlp_ln_b": (H,)})
    return out


def make_weights(cfg: dict, seed: int = 0, std: float = 0.05, device="cpu") -> Dict[str, torch.Tensor]:
    """Seeded random-init bf16 weights (N(0, std^2); LayerNorm gains 1 + 0.1 N, every bias 0.1 N)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = {}
    for name, shape in tensor_shapes(cfg).items():
        leaf = name.split(".")[-1]
        if leaf.endswith("ln_g"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf.endswith("ln_b") or leaf.startswith("b"):
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            t = std * torch.randn(shape, generator=g)
        w[name] = t.to(torch.bfloat16).to(device)
    return w


def _ln(x, g, b, eps):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)


@torch.no_grad()
def hidden_states(w: Dict[str, torch.Tensor], cfg: dict, tokens, dtype=torch.float32) -> torch.Tensor:
    """Last-layer hidden states [T, H] of ONE sequence (no padding, so no attention mask is needed)."""
    dev = next(iter(w.values())).device
    tokens = torch.as_tensor(tokens, dtype=torch.long, device=dev)
    T = tokens.shape[0]
    H, nh, D, eps = cfg["hidden"], cfg["n_heads"], cfg["head_dim"], cfg.get("ln_eps", 1e-12)
    f = lambda name: w[name].to(dtype)
    h = f("word_embed")[tokens] + f("type_embed")[0] + f("pos_embed")[:T]
    h = _ln(h, f("emb_ln_g"), f("emb_ln_b"), eps)
    for l in range(cfg["n_layers"]):
        p = f"layers.{l}."
        qkv = h @ f(p + "wqkv").T + f(p + "bqkv")
        q, k, v = (qkv[:, i * H:(i + 1) * H].view(T, nh, D) for i in range(3))
        s = torch.einsum("qhd,khd->hqk", q, k) / math.sqrt(D)
        a = torch.einsum("hqk,khd->qhd", torch.softmax(s.float(), -1).to(dtype), v).reshape(T, H)
        h = _ln(a @ f(p + "wo").T + f(p + "bo") + h, f(p + "attn_ln_g"), f(p + "attn_ln_b"), eps)
        u = torch.nn.functional.gelu(h @ f(p + "w_up").T + f(p + "b_up"))          # exact (erf) GELU
        h = _ln(u @ f(p + "w_down").T + f(p + "b_down") + h, f(p + "mlp_ln_g"), f(p + "mlp_ln_b"), eps)
    return h


@torch.no_grad()
def embed(w, cfg, sequences: Sequence[Sequence[int]], dtype=torch.float32) -> torch.Tensor:
    """[n_seq, H] L2-normalised [CLS] (first position) states - what /api/embed returns per input."""
    rows: List[torch.Tensor] = []
    for toks in sequences:
        c = hidden_states(w, cfg, toks, dtype)[0].float()
        rows.append(c / c.norm().clamp_min(1e-12))
    return torch.stack(rows)


def to_hf_state_dict(w: Dict[str, torch.Tensor], cfg: dict) -> Dict[str, torch.Tensor]:
    H = cfg["hidden"]
    sd = {"embeddings.word_embeddings.weight": w["word_embed"], "embeddings.position_embeddings.weight": w["pos_embed"],
          "embeddings.token_type_embeddings.weight": w["type_embed"], "embeddings.LayerNorm.weight": w["emb_ln_g"],
          "embeddings.LayerNorm.bias": w["emb_ln_b"]}
    for l in range(cfg["n_layers"]):
        p, hp = f"layers.{l}.", f"encoder.layer.{l}."
        for i, nm in enumerate(("query", "key", "value")):
            sd[hp + f"attention.self.{nm}.weight"] = w[p + "wqkv"][i * H:(i + 1) * H]
            sd[hp + f"attention.self.{nm}.bias"] = w[p + "bqkv"][i * H:(i + 1) * H]
        sd[hp + "attention.output.dense.weight"] = w[p + "wo"]
        sd[hp + "attention.output.dense.bias"] = w[p + "bo"]
        sd[hp + "attention.output.LayerNorm.weight"] = w[p + "attn_ln_g"]
        sd[hp + "attention.output.LayerNorm.bias"] = w[p + "attn_ln_b"]
        sd[hp + "intermediate.dense.weight"] = w[p + "w_up"]
        sd[hp + "intermediate.dense.bias"] = w[p + "b_up"]
        sd[hp + "output.dense.weight"] = w[p + "w_down"]
        sd[hp + "output.dense.bias"] = w[p + "b_down"]
        sd[hp + "output.LayerNorm.weight"] = w[p + "mlp_ln_g"]
        sd[hp + "output.LayerNorm.bias"] = w[p + "mlp_ln_b"]
    return sd


# BASELINE.json configs[4] and a small pin / test geometry
BGE_SMALL = dict(vocab=30522, hidden=384, ffn=1536, n_layers=12, n_heads=12, head_dim=32, max_positions=512,
                 type_vocab=2, ln_eps=1e-12)
TINY_BERT = dict(vocab=512, hidden=128, ffn=256, n_layers=2, n_heads=4, head_dim=32, max_positions=128, type_vocab=2,
                 ln_eps=1e-12)

"""Generates tests/golden/llama_tiny.json: HF transformers outputs that pin oracle/llama_ref.py.

Run in the build container (transformers 5.5.0, torch CPU) from the repo root:
    python tests/golden/make_llama_golden.py
The fixture stores the HF fp32 logits of seeded random-init tiny Llama / Qwen2 / Phi3 models on fixed token
sequences, plus a checksum of the seeded weights so RNG drift is detected instead of silently mis-pinning.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import llama_ref as R  # noqa: E402

import transformers  # noqa: E402
from transformers import (LlamaConfig, LlamaForCausalLM, Phi3Config, Phi3ForCausalLM, Qwen2Config,  # noqa: E402
                          Qwen2ForCausalLM)


def hf_model(cfg, w):
    common = dict(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], intermediate_size=cfg["ffn"],
                  num_hidden_layers=cfg["n_layers"], num_attention_heads=cfg["n_q_heads"],
                  num_key_value_heads=cfg["n_kv_heads"], rms_norm_eps=cfg["rms_eps"],
                  max_position_embeddings=4096, tie_word_embeddings=False)
    if cfg.get("family") == "phi3":
        assert cfg["hidden"] == cfg["n_q_heads"] * cfg["head_dim"]
        c = Phi3Config(**common, pad_token_id=0, bos_token_id=1, eos_token_id=2,
                       rope_parameters={"rope_type": "default", "rope_theta": cfg["rope_theta"]})
        m = Phi3ForCausalLM(c)
        sd = {k: v.float() for k, v in R.to_hf_phi3_state_dict(w, cfg).items()}
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not [k for k in missing if "rotary" not in k], missing
        assert not unexpected, unexpected
        return m.eval().float()
    if cfg.get("qkv_bias"):
        c = Qwen2Config(**common, rope_parameters={"rope_type": "default", "rope_theta": cfg["rope_theta"]})
        m = Qwen2ForCausalLM(c)
    else:
        c = LlamaConfig(**common, head_dim=cfg["head_dim"], attention_bias=False,
                        rope_parameters={"rope_type": "default", "rope_theta": cfg["rope_theta"]})
        m = LlamaForCausalLM(c)
    sd = {k: v.float() for k, v in R.to_hf_state_dict(w, cfg).items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "rotary" not in k], missing
    assert not unexpected, unexpected
    return m.eval().float()


def checksum(w):
    return float(sum(v.float().double().abs().sum() for v in w.values()))


out = {"generator": "tests/golden/make_llama_golden.py", "transformers": transformers.__version__,
       "torch": torch.__version__, "cases": []}
for name, cfg, seed in (("tiny_llama", R.TINY_LLAMA, 1234), ("tiny_qwen", R.TINY_QWEN, 4321),
                        ("tiny_phi3", R.TINY_PHI3, 2468)):
    w = R.make_weights(cfg, seed=seed)
    m = hf_model(cfg, w)
    g = torch.Generator().manual_seed(seed + 1)
    for T in (1, 19, 70):
        toks = torch.randint(0, cfg["vocab"], (T,), generator=g)
        with torch.no_grad():
            logits = m(toks[None]).logits[0].float()
        ours = R.forward(w, cfg, toks, torch.float32)
        err = (ours - logits).abs().max().item()
        print(name, T, "oracle-vs-HF max abs err", err, "max|logit|", logits.abs().max().item())
        assert err < 2e-4, err
        out["cases"].append({"model": name, "seed": seed, "weights_abs_sum": checksum(w), "tokens": toks.tolist(),
                             "last_logits": [round(float(x), 6) for x in logits[-1]],
                             "first_logits_head": [round(float(x), 6) for x in logits[0][:32]],
                             "argmax_all": logits.argmax(-1).tolist()})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "llama_tiny.json"), "w"))
print("wrote llama_tiny.json", os.path.getsize(os.path.join(ROOT, "tests", "golden", "llama_tiny.json")), "bytes")

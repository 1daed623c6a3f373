"""Generates tests/golden/bert_tiny.json: HF transformers BertModel outputs that pin oracle/bert_ref.py.

Run in the build container (transformers 5.5.0, torch CPU) from the repo root:
    python tests/golden/make_bert_golden.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bert_ref as B  # noqa: E402

import transformers  # noqa: E402
from transformers import BertConfig, BertModel  # noqa: E402


def hf_model(cfg, w):
    c = BertConfig(vocab_size=cfg["vocab"], hidden_size=cfg["hidden"], num_hidden_layers=cfg["n_layers"],
                   num_attention_heads=cfg["n_heads"], intermediate_size=cfg["ffn"], hidden_act="gelu",
                   hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                   max_position_embeddings=cfg["max_positions"], type_vocab_size=cfg["type_vocab"],
                   layer_norm_eps=cfg["ln_eps"])
    m = BertModel(c, add_pooling_layer=False)
    sd = {k: v.float() for k, v in B.to_hf_state_dict(w, cfg).items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "position_ids" not in k and "token_type_ids" not in k], missing
    assert not unexpected, unexpected
    return m.eval().float()


out = {"generator": "tests/golden/make_bert_golden.py", "transformers": transformers.__version__,
       "torch": torch.__version__, "cases": []}
cfg, seed = B.TINY_BERT, 97
w = B.make_weights(cfg, seed=seed)
m = hf_model(cfg, w)
g = torch.Generator().manual_seed(seed + 1)
for T in (1, 7, 33, 128):
    toks = torch.randint(0, cfg["vocab"], (T,), generator=g)
    with torch.no_grad():
        hs = m(toks[None]).last_hidden_state[0].float()
    ours = B.hidden_states(w, cfg, toks)
    err = (ours - hs).abs().max().item()
    print("tiny_bert", T, "oracle-vs-HF max abs err", err, "max|h|", hs.abs().max().item())
    assert err < 2e-4, err
    cls = hs[0] / hs[0].norm()
    out["cases"].append({"model": "tiny_bert", "seed": seed,
                         "weights_abs_sum": float(sum(v.float().double().abs().sum() for v in w.values())),
                         "tokens": toks.tolist(), "embedding": [round(float(x), 7) for x in cls],
                         "last_row_head": [round(float(x), 6) for x in hs[-1][:32]]})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bert_tiny.json"), "w"))
print("wrote bert_tiny.json", os.path.getsize(os.path.join(ROOT, "tests", "golden", "bert_tiny.json")), "bytes")

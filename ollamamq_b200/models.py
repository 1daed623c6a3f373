"""Model geometries of the BASELINE.json configs the worker supports (decoder-only, RMSNorm + RoPE + SwiGLU;
head_dim 128 / 96 / 64).
Plain dicts with the same keys as mq_model_cfg (include/ollamamq_b200.h)."""

LLAMA3_8B = dict(vocab=128256, hidden=4096, ffn=14336, n_layers=32, n_q_heads=32, n_kv_heads=8, head_dim=128,
                 qkv_bias=0, rope_theta=500000.0, rms_eps=1e-5)
QWEN25_7B = dict(vocab=152064, hidden=3584, ffn=18944, n_layers=28, n_q_heads=28, n_kv_heads=4, head_dim=128,
                 qkv_bias=1, rope_theta=1000000.0, rms_eps=1e-6)
PHI3_MINI = dict(vocab=32064, hidden=3072, ffn=8192, n_layers=32, n_q_heads=32, n_kv_heads=32, head_dim=96,
                 qkv_bias=0, rope_theta=10000.0, rms_eps=1e-5)
# BASELINE configs[4]: bge-small (BERT encoder) behind /api/embed - keys of mq_encoder_cfg
BGE_SMALL = dict(vocab=30522, hidden=384, ffn=1536, n_layers=12, n_heads=12, head_dim=32, max_positions=512,
                 type_vocab=2, ln_eps=1e-12)

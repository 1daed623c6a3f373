"""The synthetic trace of the headline benchmark (BASELINE.json configs[1]) - pure numpy, imports nothing of the product
and nothing of oracle/, so both bench arms and the parity test (tests/test_headline_gpu.py) replay the SAME inputs.

    64 users user00..user63, one request each, all arriving at t = 0; prompt = 512 token ids uniform in [0, vocab)
    from numpy.random.default_rng(user_idx); greedy decode of exactly 128 tokens."""
import zlib

USERS = 64
PROMPT_LEN = 512
GEN_LEN = 128
VOCAB = 128256  # Llama-3-8B


def prompts(users: int = USERS, prompt_len: int = PROMPT_LEN, vocab: int = VOCAB):
    import numpy as np
    return [np.random.default_rng(u).integers(0, vocab, prompt_len).astype("int32").tolist() for u in range(users)]


def token_checksum(tokens_by_user) -> str:
    """CRC-32 over the generated token ids of every user in user order (little-endian int32).  bench.py prints it for the
    first timed step; tests/test_headline_gpu.py checks the same tokens against the fp32 oracle and pins the value, so
    the path that is timed is the path that is tested."""
    import numpy as np
    crc = 0
    for toks in tokens_by_user:
        crc = zlib.crc32(np.asarray(toks, dtype="<i4").tobytes(), crc)
    return "%08x" % (crc & 0xFFFFFFFF)

"""Decode-chain kernels (csrc/gemm_dk.cuh + the RMSNorm fold of gemm.cuh / gemm_streamk.cuh) vs plain torch fp32
references of the same ops, through the C ABI's device-pointer test entry points.

What the chain replaces per layer (round-1 path): add_rmsnorm -> QKV GEMM (fp32 split-K planes) -> rope_kv -> ... ->
O GEMM (planes) -> add_rmsnorm -> gate/up -> down (planes).  Here: QKV planes with the rstd fold -> rope_kv -> ... ->
O (cluster split-K, DSMEM reduce, residual add, xg = bf16(h * gamma), sum h^2) -> gate/up (rstd fold) -> down (same as O).
Tolerances: one bf16 rounding of an fp32-accumulated value (rel 2^-8) plus accumulation-order noise; fp32 outputs 2e-3.
"""
import ctypes as C

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

PAGE = 16


def _lib():
    import ollamamq_b200 as m
    return m


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def dev():
    return torch.device("cuda:0")


def _relerr(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def _inv_freq(theta, D):
    return 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.float32) / D))


def _rope_ref(x, pos, inv_freq):
    ang = pos.float()[:, None] * inv_freq[None, :]
    cos = torch.cat([ang.cos(), ang.cos()], -1)[:, None, :]
    sin = torch.cat([ang.sin(), ang.sin()], -1)[:, None, :]
    D = x.shape[-1]
    rot = torch.cat([-x[..., D // 2:], x[..., : D // 2]], -1)
    return x * cos + rot * sin


def test_cluster_occupancy_is_reported():
    m = _lib()
    out = (C.c_int * 8)()
    assert m.lib.mq_debug_cluster_info(out) == 0, m.last_error()
    got = list(out)
    print("co-resident clusters of 1..8 CTAs:", got)
    assert got[0] >= 100 and got[1] >= 50, got  # a B200 has 148 SMs; pairs pack perfectly


def test_embed_chain():
    m = _lib()
    V, H, T = 1000, 4096, 37
    E = torch.randn(V, H, device=dev()).bfloat16()
    gamma = (1 + 0.1 * torch.randn(H, device=dev())).bfloat16()
    ids = torch.randint(0, V, (T,), device=dev(), dtype=torch.int32)
    h = torch.zeros(T, H, device=dev())
    xg = torch.zeros(T, H, device=dev(), dtype=torch.bfloat16)
    ssq = torch.zeros(T, device=dev())
    assert m.lib.mq_debug_embed_chain(P(ids), P(E), P(h), T, H, P(gamma), P(xg), P(ssq)) == 0, m.last_error()
    ref = E[ids.long()].float()
    assert torch.equal(h, ref)
    assert _relerr(xg, ref * gamma.float()) < 4e-3
    assert torch.allclose(ssq, ref.pow(2).sum(-1), rtol=1e-5)


# T, n_out, K, cs
RESID_CASES = [(64, 4096, 4096, 4), (64, 4096, 14336, 4), (64, 4096, 4096, 2), (16, 4096, 14336, 4), (48, 1024, 2048, 3),
               (16, 1024, 1024, 1), (5, 512, 512, 2), (33, 3584, 3584, 4), (64, 1000, 1024, 2), (32, 4096, 14336, 0)]


@pytest.mark.parametrize("T,n_out,K,cs", RESID_CASES)
def test_gemm_dk_resid(T, n_out, K, cs):
    """h += X W^T;  xg = bf16(h * gamma);  per-tile sum of h^2 - and bit-identical on a second run (deterministic)."""
    m = _lib()
    g = torch.Generator(device="cuda").manual_seed(T * 13 + n_out + K)
    xa = 64
    W = (torch.randn(n_out, K, device=dev(), generator=g) * 0.03).bfloat16()
    X = torch.randn(xa, K, device=dev(), generator=g).bfloat16()
    h0 = torch.randn(T, n_out, device=dev(), generator=g)
    gamma = (1 + 0.1 * torch.randn(n_out, device=dev(), generator=g)).bfloat16()
    tiles = (n_out + 127) // 128
    ref_h = h0 + X[:T].float() @ W.float().T
    outs = []
    for rep in range(2):
        h = h0.clone()
        xg = torch.full((T, n_out), float("nan"), device=dev(), dtype=torch.bfloat16)
        ssq = torch.full((tiles, 64), float("nan"), device=dev())
        rc = m.lib.mq_debug_gemm_dk_resid(P(W), n_out, K, P(X), xa, T, cs, P(h), P(gamma), P(xg), P(ssq), 64, 0, None)
        assert rc == 0, m.last_error()
        outs.append((h, xg, ssq))
    h, xg, ssq = outs[0]
    assert torch.isfinite(h).all() and torch.isfinite(xg.float()).all()
    assert _relerr(h, ref_h) < 2e-3
    assert _relerr(xg, h * gamma.float()) < 4e-3
    got_ssq = ssq[:, :T].sum(0)
    assert torch.isfinite(got_ssq).all()
    assert torch.allclose(got_ssq, h.pow(2).sum(-1), rtol=1e-4)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2][:, :T], outs[1][2][:, :T])


@pytest.mark.parametrize("T,I,K,rows,sk", [(64, 14336, 4096, -1, 0), (64, 14336, 4096, 0, 0), (16, 1024, 512, 64, 0),
                                            (48, 2816, 1024, 72, 0), (64, 1024, 4096, 0, 1)])
def test_gemm_silu_fold_and_balanced_tiles(T, I, K, rows, sk):
    """gate/up with the RMSNorm fold: act = silu(g * rstd) * (u * rstd), on 128-row and on balanced (< 128-row) tiles."""
    m = _lib()
    g = torch.Generator(device="cuda").manual_seed(T + I + K)
    W = (torch.randn(2 * I, K, device=dev(), generator=g) * 0.03).bfloat16()
    X = torch.randn(64, K, device=dev(), generator=g).bfloat16()
    parts = 3
    ssq = torch.rand(parts, 64, device=dev(), generator=g) * K
    eps = 1e-5
    rstd = torch.rsqrt(ssq.sum(0)[:T] / K + eps)[:, None]
    gate = (X[:T].float() @ W[:I].float().T) * rstd
    up = (X[:T].float() @ W[I:].float().T) * rstd
    ref = torch.nn.functional.silu(gate) * up
    out = torch.full((T, I), float("nan"), device=dev(), dtype=torch.bfloat16)
    rc = m.lib.mq_debug_gemm_fold(P(W), 2 * I, I, K, P(X), 64, T, 2, P(out), I, I, rows, sk, P(ssq), parts, 64, 1.0 / K, eps,
                                  0, None)
    assert rc == 0, m.last_error()
    assert torch.isfinite(out.float()).all(), "non-finite / unwritten outputs"
    assert _relerr(out, ref) < 8e-3


@pytest.mark.parametrize("T,V,K,sk", [(64, 128256, 4096, 1), (16, 32064, 3072, 1), (48, 5000, 1024, 0), (64, 8192, 2048, 0)])
def test_lm_head_fold(T, V, K, sk):
    """LM head on xg = bf16(h * gamma) with the final RMSNorm as the epilogue's per-token scale (stream-K and plain)."""
    m = _lib()
    g = torch.Generator(device="cuda").manual_seed(T + V)
    W = (torch.randn(V, K, device=dev(), generator=g) * 0.03).bfloat16()
    X = torch.randn(64, K, device=dev(), generator=g).bfloat16()
    parts = 4
    ssq = torch.rand(parts, 64, device=dev(), generator=g) * K
    eps = 1e-5
    rstd = torch.rsqrt(ssq.sum(0)[:T] / K + eps)[:, None]
    ref = (X[:T].float() @ W.float().T) * rstd
    out = torch.full((T, V), float("nan"), device=dev())
    rc = m.lib.mq_debug_gemm_fold(P(W), V, V, K, P(X), 64, T, 0, P(out), V, 0, 0, sk, P(ssq), parts, 64, 1.0 / K, eps, 0, None)
    assert rc == 0, m.last_error()
    assert torch.isfinite(out).all()
    assert _relerr(out, ref) < 2e-3


@pytest.mark.parametrize("T,n_out,K", [(4608, 4096, 4096), (300, 1024, 2048), (1000, 512, 14336), (257, 256, 512)])
def test_prefill_resid_epilogue_on_the_persistent_2cta_kernel(T, n_out, K):
    """Prefill O / down projection with the RMSNorm fold: h += X W^T, xg = bf16(h * gamma), per-128-feature-tile sum of h^2
    (ragged last token tile included), deterministic."""
    m = _lib()
    g = torch.Generator(device="cuda").manual_seed(T + n_out + K)
    W = (torch.randn(n_out, K, device=dev(), generator=g) * 0.03).bfloat16()
    X = torch.randn(T, K, device=dev(), generator=g).bfloat16()
    h0 = torch.randn(T, n_out, device=dev(), generator=g)
    gamma = (1 + 0.1 * torch.randn(n_out, device=dev(), generator=g)).bfloat16()
    tiles = n_out // 128
    stride = ((T + 15) // 16) * 16
    ref_h = h0 + X.float() @ W.float().T
    outs = []
    for rep in range(2):
        h = h0.clone()
        xg = torch.full((T, n_out), float("nan"), device=dev(), dtype=torch.bfloat16)
        ssq = torch.full((tiles, stride), float("nan"), device=dev())
        rc = m.lib.mq_debug_gemm_resid_prefill(P(W), n_out, K, P(X), T, T, P(h), P(gamma), P(xg), P(ssq), stride, None, 0, 0, 0.0,
                                               0.0, 0, None)
        assert rc == 0, m.last_error()
        outs.append((h, xg, ssq))
    h, xg, ssq = outs[0]
    assert torch.isfinite(h).all() and torch.isfinite(xg.float()).all()
    assert _relerr(h, ref_h) < 2e-3
    assert _relerr(xg, h * gamma.float()) < 4e-3
    got = ssq[:, :T].sum(0)
    assert torch.isfinite(got).all()
    assert torch.allclose(got, h.pow(2).sum(-1), rtol=1e-4)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][2][:, :T], outs[1][2][:, :T])

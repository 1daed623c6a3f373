// Weight-stationary-A tcgen05 GEMM for the forward pass (replaces the arithmetic behind the reference's
// backend call, /root/reference/src/dispatcher.rs:287-290).
//
//   out[t, f] = sum_k X[t, k] * W[f, k]          X: [T, K] bf16 row-major (activations)
//                                                W: [N_out, K] bf16 row-major (torch Linear layout)
//
// The UMMA "A" operand (M = 128 TMEM lanes) is a 128-row tile of W, the "B" operand (N = BN TMEM columns)
// is a BN-row tile of X.  One kernel therefore serves both regimes:
//   * decode  (T = running sequences <= 256): the weights stream through the tensor core at HBM rate while
//     the whole batch sits in one N tile; split-K (gridDim.z) fills the 148 SMs and the fp32 partial
//     planes are summed by the consumer kernel (rmsnorm / rope), so the residual add is fused for free;
//   * prefill (T = thousands of prompt tokens): BN = 256 tiles, tensor-pipe bound.
//
// Pipeline per CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> smem tile -> one TMA tensor store).  Weight tiles of the first ring pass are requested
// BEFORE griddepcontrol.wait, so under programmatic dependent launch the HBM stream of this GEMM starts
// while the previous kernel is still draining.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace mq {

enum GemmEpilogue : int {
  EPI_F32 = 0,       // out_f32[z][t][f] = acc                       (split-K partial planes)
  EPI_BF16 = 1,      // out_bf16[t][f]   = acc
  EPI_SILU_BF16 = 2, // out_bf16[t][f]   = silu(acc_gate) * acc_up   (two A tiles: rows f and f + a2_row_off)
  EPI_GELU_BF16 = 3, // out_bf16[t][f]   = gelu_erf(acc + bias[f])   (BERT intermediate; plain 1-CTA kernel only)
  EPI_BIAS_BF16 = 4, // out_bf16[t][f]   = acc + bias[f]             (encoder QKV; plain 1-CTA kernel only)
  EPI_RESID = 5      // h[t][f] += acc;  xg[t][f] = bf16(h * gamma_next[f]);  ssq[tile][t] = sum_f h^2   (prefill O / down
                     // projections with the RMSNorm fold; 2-CTA kernel only - gemm_2cta.cuh)
};

// RMSNorm fold (decode chain): the activation operand of a GEMM is xg = bf16(h * gamma) and the epilogue scales token
// column t of the accumulator by rstd[t] = rsqrt(sum_i ssq[i * stride + t] * inv_h + eps).  `ssq` holds per-weight-tile
// partial sums of h^2 written by the producer of h (gemm_dk.cuh DK_RESID, or the embedding gather: 1 part); they are
// added in index order, so the result does not depend on scheduling.  ssq == nullptr: no scaling.
struct RstdIn {
  const float* ssq;
  int parts;
  int stride;
  float inv_h;
  float eps;
};
__device__ __forceinline__ float rstd_of(const RstdIn& r, int t) {
  float s = 0.f;
  for (int i = 0; i < r.parts; ++i) s += __ldcg(r.ssq + (size_t)i * r.stride + t);
  return rsqrtf(s * r.inv_h + r.eps);
}

struct GemmParams {
  void* out;               // (kept for reference; the epilogue writes through the tmC tensor map)
  long long split_stride;  // elements between split-K planes of `out`
  int ldo;                 // leading dimension of out (elements)
  int T;                   // valid rows of X
  int n_out;               // valid output features
  int k_blocks;            // K / 64
  int kb_per_split;        // k-blocks handled by one blockIdx.z
  int a2_row_off;          // EPI_SILU_BF16: row offset of the "up" half inside W
  int m_tiles, n_tiles;    // tile grid; blockIdx.x = linear tile id, rasterised in groups of `group_m` weight tiles
  int group_m;             // so that one wave of 148 CTAs touches ~group_m weight tiles x ~148/group_m token tiles
  int tile_rows;           // weight rows per tile: 128, or fewer (multiple of 8) so that m_tiles ~ the SM count: an
                           // HBM-bound launch is as fast as its busiest SM (14336 gate/up rows: 112 tiles of 128 rows leave
                           // 36 SMs idle, 138 tiles of 104 rows do not).  The MMA still runs M = 128; lanes >= tile_rows
                           // hold garbage that is never read.
  unsigned long long w_policy;  // L2 policy for weight tiles: stream-once (decode) vs re-used inside a wave (prefill)
  RstdIn rs;                    // optional RMSNorm fold (see above)
  Trace tr;                     // optional timeline stamps (MQ_TRACE=1)
  const void* bias;             // EPI_GELU_BF16 / EPI_BIAS_BF16: bf16 [n_out] (nullable)
};

constexpr int kGemmThreads = 192;        // producer warp + MMA warp + 4 epilogue warps
constexpr int kGemmThreadsDecode = 320;  // ... + 8 epilogue warps (decode widths: the epilogue is on the critical path)
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB

__host__ __device__ constexpr int gemm_stage_bytes(int bn, int epi) {
  return kATileBytes * (epi == EPI_SILU_BF16 ? 2 : 1) + bn * kBlockK * 2;
}
// The ring fills the SM (up to 8 stages / 200 KiB), one CTA per SM.  (r01 A/B: a <= 120 KiB ring that lets the next
// kernel's CTA be co-resident was slower, 4.64 vs 4.36 ms per decode step - bytes in flight per SM win.)
__host__ __device__ constexpr int gemm_stages(int bn, int epi) {
  int s = (200 * 1024) / gemm_stage_bytes(bn, epi);
  return s > 8 ? 8 : (s < 2 ? 2 : s);
}
__host__ __device__ constexpr int gemm_out_tile_bytes(int bn, int epi) {
  return bn * kBlockM * (epi == EPI_F32 ? 4 : 2);  // epilogue staging tile [BN tokens][<= 128 features]
}
__host__ __device__ constexpr int gemm_smem_bytes(int bn, int epi) {
  return gemm_stages(bn, epi) * gemm_stage_bytes(bn, epi) + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*rstd[BN]*/;
}
__host__ __device__ constexpr uint32_t gemm_tmem_cols(int bn, int epi) {
  int need = bn * (epi == EPI_SILU_BF16 ? 2 : 1);
  return need <= 32 ? 32u : need <= 64 ? 64u : need <= 128 ? 128u : need <= 256 ? 256u : 512u;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreadsDecode, 1)
gemm_wx_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  constexpr bool kDual = (EPI == EPI_SILU_BF16);
  constexpr int STAGES = gemm_stages(BN, EPI);
  constexpr int STAGE_BYTES = gemm_stage_bytes(BN, EPI);
  constexpr int B_OFF = kATileBytes * (kDual ? 2 : 1);
  constexpr uint32_t TMEM_COLS = gemm_tmem_cols(BN, EPI);
  constexpr uint32_t IDESC = umma_idesc_bf16(kBlockM, BN);
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static_assert(!kDual || BN * 2 <= 512, "dual accumulator must fit TMEM");
  static_assert(gemm_out_tile_bytes(BN, EPI) <= STAGES * STAGE_BYTES, "epilogue tile is staged in the pipeline smem");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* rstd_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);  // [BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Grouped rasterisation.  With m-tiles fastest (r01 v1-v3) every wave of a prefill GEMM re-streamed the whole
  // weight matrix from HBM (ncu: 8.2 GB of DRAM reads for the 0.57 GB gate/up GEMM, L2 hit rate 46 %).
  int tile_m, tile_n;
  {
    const int pid = blockIdx.x;
    const int per_group = p.group_m * p.n_tiles;
    const int first_m = (pid / per_group) * p.group_m;
    const int gsz = min(p.m_tiles - first_m, p.group_m);
    const int r = pid % per_group;
    tile_m = first_m + r % gsz;
    tile_n = r / gsz;
  }
  const int R = p.tile_rows;
  const int m0 = tile_m * R;
  const int n0 = tile_n * BN;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int nkb = min(p.kb_per_split, p.k_blocks - kb0);
  const uint32_t stage_tx = (uint32_t)(R * kBlockK * 2 * (kDual ? 2 : 1) + BN * kBlockK * 2);

  if (warp == 0 && lane == 0) {
    trace_begin(p.tr);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Let the next kernel in the stream begin its own prologue (it still waits for our completion
  // through griddepcontrol.wait before touching anything we write).
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      const int npre = nkb < STAGES ? nkb : STAGES;
      for (int s = 0; s < npre; ++s) {
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], stage_tx);
        tma_load_2d(st, &tmA, &full_bar[s], (kb0 + s) * kBlockK, m0, p.w_policy);
        if (kDual) tma_load_2d(st + kATileBytes, &tmA, &full_bar[s], (kb0 + s) * kBlockK, m0 + p.a2_row_off, p.w_policy);
      }
      pdl_wait();  // activations are produced by the previous kernel
      trace_waited(p.tr);
      for (int s = 0; s < npre; ++s)
        tma_load_2d(smem + s * STAGE_BYTES + B_OFF, &tmB, &full_bar[s], (kb0 + s) * kBlockK, n0, kEvictLast);
      for (int kb = npre; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], stage_tx);
        tma_load_2d(st, &tmA, &full_bar[s], (kb0 + kb) * kBlockK, m0, p.w_policy);
        if (kDual) tma_load_2d(st + kATileBytes, &tmA, &full_bar[s], (kb0 + kb) * kBlockK, m0 + p.a2_row_off, p.w_policy);
        tma_load_2d(st + B_OFF, &tmB, &full_bar[s], (kb0 + kb) * kBlockK, n0, kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer (single thread) ----------------
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + B_OFF;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t db = umma_desc_sw128(b_addr + k * 32);
          const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
          umma_bf16(tmem_base, umma_desc_sw128(a_addr + k * 32), db, IDESC, acc);
          if (kDual) umma_bf16(tmem_base + BN, umma_desc_sw128(a_addr + kATileBytes + k * 32), db, IDESC, acc);
        }
        umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs retire
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ---------------- epilogue: TMEM lane = output feature, TMEM column = token ----------------
    // The accumulator is transposed on its way out: each thread owns one feature (TMEM lane) and walks the
    // token columns, writing a [token][tile_rows features] tile into the (now idle) pipeline smem; one TMA tensor
    // store then moves the whole tile to global memory, coalesced and clipped to the tensor bounds.
    // 4 or 8 epilogue warps (blockDim): two warps may share a TMEM lane quarter and split the token columns
    const int n_epi = (int)blockDim.x - 64;
    const int hid = (warp - 2) >> 2, n_half = n_epi >> 7;
    const bool fold = p.rs.ssq != nullptr;
    if (fold) {  // RMSNorm fold: per-token scale, computed while the mainloop streams (sums are from the previous kernel)
      pdl_wait();
      for (int t = threadIdx.x - 64; t < BN; t += n_epi) rstd_s[t] = (n0 + t < p.T) ? rstd_of(p.rs, n0 + t) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory");
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const bool live = row < R;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* stg = smem;  // every MMA has retired (tmem_full), so all stage buffers are free
    const uint32_t stg_a = smem_u32(stg), rstd_a = smem_u32(rstd_s);
    constexpr int ESZ = EPI == EPI_F32 ? 4 : 2;
#pragma unroll 1
    for (int c0 = hid * 16; c0 < BN; c0 += 16 * n_half) {
      if (n0 + c0 >= p.T) break;  // warp-uniform; columns past T are clipped by the store anyway
      uint32_t v[16];
      tmem_ld16(t_lane + c0, v);
      const uint32_t o = stg_a + (uint32_t)(c0 * R + row) * ESZ;  // explicit st.shared: see ptx.cuh
      if constexpr (kDual) {
        uint32_t u[16];
        tmem_ld16(t_lane + BN + c0, u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float sc = fold ? lds_f32(rstd_a + (uint32_t)(c0 + j) * 4u) : 1.f;
          const float g = __uint_as_float(v[j]) * sc;
          const float up = __uint_as_float(u[j]) * sc;
          if (live) sts_bf16(o + (uint32_t)(j * R) * 2u, __fdividef(g, 1.0f + __expf(-g)) * up);
        }
      } else {
        tmem_ld_wait();
        if constexpr (EPI == EPI_F32) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (live) sts_f32(o + (uint32_t)(j * R) * 4u, __uint_as_float(v[j]) * (fold ? lds_f32(rstd_a + (uint32_t)(c0 + j) * 4u) : 1.f));
        } else if constexpr (EPI == EPI_GELU_BF16 || EPI == EPI_BIAS_BF16) {
          const int f = m0 + row;  // this thread's output feature
          const float b = (p.bias && f < p.n_out) ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[f]) : 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = __uint_as_float(v[j]) + b;
            if (!live) continue;
            if constexpr (EPI == EPI_GELU_BF16) sts_bf16(o + (uint32_t)(j * R) * 2u, 0.5f * a * (1.0f + erff(a * 0.70710678118654752f)));
            else sts_bf16(o + (uint32_t)(j * R) * 2u, a);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (live) sts_bf16(o + (uint32_t)(j * R) * 2u, __uint_as_float(v[j]) * (fold ? lds_f32(rstd_a + (uint32_t)(c0 + j) * 4u) : 1.f));
        }
      }
    }
    fence_proxy_async();                                   // generic-proxy smem writes -> visible to the TMA engine
    asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory");  // the epilogue warps only
    if (warp == 2 && lane == 0) {
      if constexpr (EPI == EPI_F32) tma_store_3d(&tmC, stg, m0, n0, blockIdx.z);
      else tma_store_2d(&tmC, stg, m0, n0);
      tma_store_commit();
      tma_store_wait_read();                               // smem must stay intact until the engine has read it
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
  if (threadIdx.x == 0) trace_end(p.tr);
}

}  // namespace mq

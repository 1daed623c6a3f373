// Persistent stream-K variant of the weight-streaming tcgen05 GEMM, for the decode regime (T <= one token tile).
//
// Why: with one 128-row weight tile per CTA the decode GEMMs of Llama-3-8B run on 112 (gate/up), 128 (O, down)
// or 144 (QKV) of the 148 SMs, and an HBM-bound kernel is as fast as its busiest SM (r01 launch shares: gate/up
// at 78 % of its HBM floor).  Here the work is cut into units of (weight tile, 64-wide k-block); CTA c of P
// persistent CTAs owns the contiguous unit range [c*U/P, (c+1)*U/P), so every SM streams the same number of
// weight bytes whatever the shape.  A tile whose k-range is shared by several CTAs is finished in-kernel:
//   * every CTA walks its range from the END towards the start, so the segment that completes a tile (its
//     "finisher" segment) is processed last and the segments that only contribute to a tile are processed first;
//   * a contributor writes its fp32 partial accumulator to a per-CTA workspace slot and bumps the finisher's
//     arrival counter (each CTA has at most one contributor segment and finishes at most one shared tile);
//   * the finisher adds the partials to its own TMEM accumulator in the epilogue, so the consumer kernels see
//     ONE complete plane (no split-K planes to re-reduce) and SiLU(gate)*up can still be fused.
// All P <= #SM CTAs are co-resident (1 CTA/SM), and a finisher only ever waits for lower-numbered CTAs that never
// wait for it, so the spin cannot deadlock.  TMEM holds two accumulator buffers: the epilogue of one segment
// overlaps the MMAs of the next, and the TMA ring never drains at tile boundaries.
#pragma once
#include "gemm.cuh"

namespace mq {

struct StreamKParams {
  void* out;       // [T][ldo] fp32 or bf16
  int ldo;
  int T;           // valid activation rows (<= BN)
  int n_out;
  int k_blocks;    // K / 64
  int m_tiles;
  int a2_row_off;  // EPI_SILU_BF16: row offset of the "up" half inside W
  int n_ctas;      // P
  float* ws;       // [P][planes][BN*128] fp32 partial accumulators
  int* flags;      // [P] arrival counters, zero between launches
  unsigned long long w_policy;
  RstdIn rs;       // optional RMSNorm fold: token column t of the finished tile is scaled by rstd[t] (gemm.cuh)
  Trace tr;        // optional timeline stamps (MQ_TRACE=1)
};

// No smem staging tile here: with <= 64 token columns the finisher stores straight from registers (a warp writes
// 32 consecutive features of one token = 128 B fp32 / 64 B bf16), which keeps the TMA ring as deep as the
// split-K kernel's (r01: a 6-stage ring + staged store made every stream-K launch ~7 us slower than split-K).
__host__ __device__ constexpr int sk_stages(int bn, int epi) {
  int s = (200 * 1024) / gemm_stage_bytes(bn, epi);
  return s > 8 ? 8 : s;
}
__host__ __device__ constexpr int sk_smem_bytes(int bn, int epi) {
  return sk_stages(bn, epi) * gemm_stage_bytes(bn, epi) + 1024 + 256 + 256 /*rstd[64]*/;
}
__host__ __device__ constexpr uint32_t sk_tmem_cols(int bn, int epi) {
  int need = 2 * bn * (epi == EPI_SILU_BF16 ? 2 : 1);  // two accumulator buffers
  return need <= 32 ? 32u : need <= 64 ? 64u : need <= 128 ? 128u : need <= 256 ? 256u : 512u;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_streamk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const StreamKParams p) {
  constexpr bool kDual = (EPI == EPI_SILU_BF16);
  constexpr int STAGES = sk_stages(BN, EPI);
  constexpr int STAGE_BYTES = gemm_stage_bytes(BN, EPI);
  constexpr int B_OFF = kATileBytes * (kDual ? 2 : 1);
  constexpr int ACC_COLS = BN * (kDual ? 2 : 1);
  constexpr uint32_t TMEM_COLS = sk_tmem_cols(BN, EPI);
  constexpr uint32_t IDESC = umma_idesc_bf16(kBlockM, BN);
  constexpr int TILE = BN * kBlockM;  // elements of one partial plane
  static_assert(BN <= 64 && BN % 16 == 0, "stream-K serves the decode tile widths");
  static_assert(STAGES >= 3, "pipeline depth");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* rstd_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);  // [BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x, P = p.n_ctas, KB = p.k_blocks;
  const long long U = (long long)p.m_tiles * KB;
  const long long u0 = c * U / P, u1 = (c + 1) * U / P;
  const int total = (int)(u1 - u0);
  const int tile_first = total > 0 ? (int)(u0 / KB) : 0;
  const int tile_last = total > 0 ? (int)((u1 - 1) / KB) : -1;
  auto seg_a = [&](int t) { const long long s = (long long)t * KB; return (int)((u0 > s ? u0 : s) - s); };
  auto seg_b = [&](int t) { const long long s = (long long)t * KB, e = s + KB; return (int)((u1 < e ? u1 : e) - s); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (threadIdx.x == 0) trace_begin(p.tr);

  if (warp == 0) {
    if (lane == 0 && total > 0) {
      // ---------------- TMA producer: one continuous stream over all segments (tiles descending) ----------------
      int t = tile_last, kb = seg_a(t), kb_end = seg_b(t);
      auto advance = [&](int& tt, int& k, int& ke) {
        if (++k == ke) {
          --tt;
          if (tt >= tile_first) { k = seg_a(tt); ke = seg_b(tt); }
        }
      };
      const int npre = total < STAGES ? total : STAGES;
      for (int i = 0; i < npre; ++i) {  // weights first: they do not depend on the previous kernel
        uint8_t* st = smem + i * STAGE_BYTES;
        mbar_expect_tx(&full_bar[i], STAGE_BYTES);
        tma_load_2d(st, &tmA, &full_bar[i], kb * kBlockK, t * kBlockM, p.w_policy);
        if (kDual) tma_load_2d(st + kATileBytes, &tmA, &full_bar[i], kb * kBlockK, t * kBlockM + p.a2_row_off, p.w_policy);
        advance(t, kb, kb_end);
      }
      pdl_wait();
      trace_waited(p.tr);
      {
        int t2 = tile_last, k2 = seg_a(t2), ke2 = seg_b(t2);
        for (int i = 0; i < npre; ++i) {
          tma_load_2d(smem + i * STAGE_BYTES + B_OFF, &tmB, &full_bar[i], k2 * kBlockK, 0, kEvictLast);
          advance(t2, k2, ke2);
        }
      }
      for (int it = npre; it < total; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_2d(st, &tmA, &full_bar[s], kb * kBlockK, t * kBlockM, p.w_policy);
        if (kDual) tma_load_2d(st + kATileBytes, &tmA, &full_bar[s], kb * kBlockK, t * kBlockM + p.a2_row_off, p.w_policy);
        tma_load_2d(st + B_OFF, &tmB, &full_bar[s], kb * kBlockK, 0, kEvictLast);
        advance(t, kb, kb_end);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer: segment i accumulates into TMEM buffer i & 1 ----------------
      int it = 0, seg = 0;
      for (int t = tile_last; t >= tile_first; --t, ++seg) {
        const int a = seg_a(t), b = seg_b(t);
        const int buf = seg & 1, use = seg >> 1;
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);  // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t acc_base = tmem_base + buf * ACC_COLS;
        for (int kb = a; kb < b; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + B_OFF;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t db = umma_desc_sw128(b_addr + k * 32);
            const uint32_t acc = (kb > a || k > 0) ? 1u : 0u;
            umma_bf16(acc_base, umma_desc_sw128(a_addr + k * 32), db, IDESC, acc);
            if (kDual) umma_bf16(acc_base + BN, umma_desc_sw128(a_addr + kATileBytes + k * 32), db, IDESC, acc);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // ---------------- epilogue warps ----------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool leader = (warp == 2 && lane == 0);
    const bool fold = p.rs.ssq != nullptr;
    if (fold) {  // per-token RMSNorm scale, computed while the mainloop streams
      pdl_wait();
      if ((int)threadIdx.x - 64 < BN) rstd_s[threadIdx.x - 64] = ((int)threadIdx.x - 64 < p.T) ? rstd_of(p.rs, threadIdx.x - 64) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    float* my_ws = p.ws + (size_t)c * (kDual ? 2 : 1) * TILE;
    int seg = 0;
    for (int t = tile_last; t >= tile_first; --t, ++seg) {
      const int a = seg_a(t), b = seg_b(t);
      const int buf = seg & 1, use = seg >> 1;
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + buf * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      if (b < KB) {
        // ---- contributor: park the partial accumulator for the finisher of tile t
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          if (c0 >= p.T) break;
          uint32_t v[16];
          tmem_ld16(t_lane + c0, v);
          if constexpr (kDual) {
            uint32_t u[16];
            tmem_ld16(t_lane + BN + c0, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              my_ws[(c0 + j) * kBlockM + row] = __uint_as_float(v[j]);
              my_ws[TILE + (c0 + j) * kBlockM + row] = __uint_as_float(u[j]);
            }
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) my_ws[(c0 + j) * kBlockM + row] = __uint_as_float(v[j]);
          }
        }
        tc_fence_before();
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {
          mbar_arrive(&tempty_bar[buf]);
          int cf = c;
          while ((long long)(cf + 1) * U / P < (long long)(t + 1) * KB) ++cf;  // CTA holding the tile's last unit
          atomicAdd(p.flags + cf, 1);
        }
      } else {
        // ---- finisher (or sole owner) of tile t
        int c_lo = c;
        if (a > 0) {
          while (c_lo * U / P > (long long)t * KB) --c_lo;  // first CTA that touches tile t
          if (leader) {
            while (ld_acquire_gpu(p.flags + c) < c - c_lo) {
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          if (c0 >= p.T) break;
          float v[16], u[16];
          {
            uint32_t r[16];
            tmem_ld16(t_lane + c0, r);
            if constexpr (kDual) {
              uint32_t r2[16];
              tmem_ld16(t_lane + BN + c0, r2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) u[j] = __uint_as_float(r2[j]);
            } else {
              tmem_ld_wait();
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          }
          for (int cc = c_lo; cc < c; ++cc) {  // fixed order -> deterministic sums
            const float* w = p.ws + (size_t)cc * (kDual ? 2 : 1) * TILE + c0 * kBlockM + row;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              v[j] += __ldcg(w + j * kBlockM);
              if constexpr (kDual) u[j] += __ldcg(w + TILE + j * kBlockM);
            }
          }
          if (fold) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              v[j] *= rstd_s[c0 + j];
              if constexpr (kDual) u[j] *= rstd_s[c0 + j];
            }
          }
          const int f = t * kBlockM + row;
          if (f < p.n_out) {
            if constexpr (kDual) {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)c0 * p.ldo + f;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.T) o[(size_t)j * p.ldo] = __float2bfloat16(__fdividef(v[j], 1.0f + __expf(-v[j])) * u[j]);
            } else if constexpr (EPI == EPI_F32) {
              float* o = reinterpret_cast<float*>(p.out) + (size_t)c0 * p.ldo + f;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.T) o[(size_t)j * p.ldo] = v[j];
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)c0 * p.ldo + f;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.T) o[(size_t)j * p.ldo] = __float2bfloat16(v[j]);
            }
          }
        }
        tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {
          mbar_arrive(&tempty_bar[buf]);
          if (a > 0) p.flags[c] = 0;  // every contributor has arrived; re-arm for the next launch
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
  if (threadIdx.x == 0) trace_end(p.tr);
}

}  // namespace mq

// Projection + bias + residual + LayerNorm in one kernel, for the embedding worker's two "output" sublayers
// (BertSelfOutput / BertOutput: dense -> + residual -> LayerNorm; bge-small: 384 features).
// Part of the forward pass that stands where the reference forwards /api/embed to a remote backend
// (/root/reference/src/dispatcher.rs:287-312).
//
// The other tcgen05 GEMMs of this library put WEIGHT rows on the TMEM lanes (a thread owns one output feature across the
// tile's tokens).  A LayerNorm needs the opposite: every feature of one token.  Here the TOKENS are the M operand: a CTA
// pair computes 256 tokens x N_OUT features with tcgen05.mma.cta_group::2 (M = 256; N = 256 + (N_OUT - 256) as two MMAs
// per k-step; N_OUT = 256 or 384 fp32 columns of tensor memory - 512 would fit TMEM but leave one pipeline stage beside
// the row tile in shared memory), so thread = token and its TMEM lane holds the complete output row:
//   pass 1  x = acc + bias + residual, x written BACK into tensor memory, partial sum / sum of squares
//   pass 2  y = (x - mean) * rstd * gamma + beta  ->  bf16, in place over the residual
// The residual / result tile (128 tokens x N_OUT bf16 per CTA) travels through shared memory by TMA, as N_OUT / 64
// sub-tiles [128 rows][64 columns] with the 128-byte swizzle (a lane reads / writes 16 bytes of its own row: eight rows
// hit eight different 16-byte slots, conflict-free): the load is issued while the tile's MMAs run, the store is one bulk
// tensor store per sub-tile.  First version: every lane read and wrote its own 768-byte row in global memory - 32
// half-used sectors per instruction, 12 000 sector requests per tile, ~27 000 cycles of epilogue per tile against
// 5 600 cycles of MMAs (r02 ncu: the top stall was the load / store queue).
// 16 epilogue warps: four per TMEM lane quarter, a quarter of the columns each; the four partial (sum, sum of squares) of a
// row meet in shared memory.  What this replaces per sublayer: a GEMM that wrote `sub`, and a LayerNorm kernel that read
// `sub` and the residual back (r02 launch list, 32 768 tokens: O 22.1 + LN 13.7, down 52.1 + LN 13.4 us per layer); the
// 384-feature GEMMs also wasted a quarter of every 256-feature MMA tile, and now pull 80 KB instead of 112 KB of operands
// per k-block and 256 tokens through L2.
// One accumulator buffer (N_OUT columns fill most of TMEM): the MMAs of tile i + 1 wait for the epilogue of tile i.
#pragma once
#include "gemm_2cta.cuh"

namespace mq {

struct RowLnParams {
  __nv_bfloat16* x;             // [T][N_OUT]: residual in, LayerNorm(residual + X W^T + bias) out
  const __nv_bfloat16* bias;    // [N_OUT]
  const __nv_bfloat16* gamma;   // [N_OUT]
  const __nv_bfloat16* beta;    // [N_OUT]
  float* h32;                   // nullable: fp32 copy of the result [T][N_OUT]
  int T, k_blocks, n_tiles, n_pairs;
  float eps;
};

template <int N_OUT> __host__ __device__ constexpr int rl_nb() { return N_OUT - 256; }         // columns of the second MMA
template <int N_OUT> __host__ __device__ constexpr int rl_b_rows() { return 128 + rl_nb<N_OUT>() / 2; }  // weight rows per CTA
template <int N_OUT> __host__ __device__ constexpr int rl_stage_bytes() { return kATileBytes + rl_b_rows<N_OUT>() * kBlockK * 2; }
template <int N_OUT> __host__ __device__ constexpr int rl_row_tile_bytes() { return 128 * N_OUT * 2; }  // residual in / result out
template <int N_OUT> __host__ __device__ constexpr int rl_tail_bytes() {
  return 256 /*barriers*/ + 3 * N_OUT * 4 /*bias, gamma, beta as fp32*/ + 4 * 128 * 8 /*row partials*/;
}
template <int N_OUT> __host__ __device__ constexpr int rl_stages() {
  int s = (227 * 1024 - 1024 - rl_row_tile_bytes<N_OUT>() - rl_tail_bytes<N_OUT>()) / rl_stage_bytes<N_OUT>();
  return s > 8 ? 8 : s;
}
template <int N_OUT> __host__ __device__ constexpr int rl_smem_bytes() {
  return rl_stages<N_OUT>() * rl_stage_bytes<N_OUT>() + rl_row_tile_bytes<N_OUT>() + 1024 + rl_tail_bytes<N_OUT>();
}
constexpr int kRlThreads = 64 + 512;  // producer warp, MMA warp, 16 epilogue warps

__device__ __forceinline__ void tmem_st16_rl(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

template <int N_OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kRlThreads, 1)
gemm_rowln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmX, const RowLnParams p) {
  static_assert(N_OUT == 256 || N_OUT == 384, "256 + {0, 128} feature columns (512 would leave one pipeline stage beside the row tile)");
  constexpr int NB = rl_nb<N_OUT>();
  constexpr int STAGES = rl_stages<N_OUT>();
  constexpr int STAGE_BYTES = rl_stage_bytes<N_OUT>();
  constexpr int B0_OFF = kATileBytes;                   // this CTA's 128 weight rows of features [0, 256)
  constexpr int B1_OFF = kATileBytes + 128 * kBlockK * 2;  // its NB / 2 rows of features [256, N_OUT)
  constexpr uint32_t IDESC0 = umma_idesc_bf16(256, 256);
  constexpr uint32_t IDESC1 = umma_idesc_bf16(256, NB > 0 ? NB : 16);
  constexpr int CPW = N_OUT / 4;                        // feature columns per epilogue warp
  static_assert(CPW % 16 == 0, "16-column TMEM chunks");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  static_assert(STAGES >= 3, "pipeline depth");
  constexpr int ROW_TILE = rl_row_tile_bytes<N_OUT>();
  uint8_t* rowt = smem + STAGES * STAGE_BYTES;   // N_OUT / 64 sub-tiles [128 rows][64 cols] bf16, 128-byte swizzle
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + ROW_TILE);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // accumulator complete (multicast commit: both CTAs)
  uint64_t* tempty_bar = tfull_bar + 1;       // leader only: both CTAs have drained the accumulator
  uint64_t* res_bar = tempty_bar + 1;         // the residual rows of the current tile have landed in `rowt`
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + ROW_TILE + 256);
  float* gamma_s = bias_s + N_OUT;
  float* beta_s = gamma_s + N_OUT;
  float2* part_s = reinterpret_cast<float2*>(beta_s + N_OUT);  // [4 column parts][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int nkb = p.k_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmX);
    mbar_init(res_bar, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx + the peer producer's remote arrive
      mbar_init(&empty_bar[s], 1);  // multicast commit from the leader's MMA thread
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 2);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
  for (int i = threadIdx.x; i < N_OUT; i += kRlThreads) {  // weights of the model: no dependency on the previous kernel
    bias_s[i] = p.bias ? __bfloat162float(p.bias[i]) : 0.f;
    gamma_s[i] = __bfloat162float(p.gamma[i]);
    beta_s[i] = __bfloat162float(p.beta[i]);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs): own 128 token rows + own share of the weight rows ----------------
      const int my_tiles = pair < p.n_tiles ? (p.n_tiles - pair + p.n_pairs - 1) / p.n_pairs : 0;
      const int total = my_tiles * nkb;
      auto load_w = [&](int it) {
        const int kb = it % nkb, s = it % STAGES;
        if (leader_cta) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
        else mbar_arrive_remote(&full_bar[s], 0);
        uint8_t* st = smem + s * STAGE_BYTES;
        tma_load_2d_2cta(st + B0_OFF, &tmB, &full_bar[s], kb * kBlockK, (int)rank * 128, kEvictLast);
        tma_load_2d_2cta(st + B0_OFF + 64 * kBlockK * 2, &tmB, &full_bar[s], kb * kBlockK, (int)rank * 128 + 64, kEvictLast);
        if (NB > 0) {
#pragma unroll
          for (int r0 = 0; r0 < NB / 2; r0 += 64)
            tma_load_2d_2cta(st + B1_OFF + r0 * kBlockK * 2, &tmB, &full_bar[s], kb * kBlockK, 256 + (int)rank * (NB / 2) + r0,
                             kEvictLast);
        }
      };
      auto load_x = [&](int it) {
        const int tile = pair + (it / nkb) * p.n_pairs, kb = it % nkb, s = it % STAGES;
        tma_load_2d_2cta(smem + s * STAGE_BYTES, &tmA, &full_bar[s], kb * kBlockK, tile * 256 + (int)rank * 128, kEvictFirst);
      };
      const int npre = total < STAGES ? total : STAGES;
      for (int it = 0; it < npre; ++it) load_w(it);
      pdl_wait();  // the activations come from the previous kernel
      for (int it = 0; it < npre; ++it) load_x(it);
      for (int it = npre; it < total; ++it) {
        mbar_wait(&empty_bar[it % STAGES], ((it / STAGES) & 1) ^ 1);
        load_w(it);
        load_x(it);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader_cta) {
      // ---------------- MMA issuer: leader CTA only ----------------
      int it = 0, i = 0;
      for (int t = pair; t < p.n_tiles; t += p.n_pairs, ++i) {
        mbar_wait(tempty_bar, (i & 1) ^ 1);  // both CTAs have drained the accumulator (first tile: passes at once)
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * 32);
            const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
            umma_bf16_2cta(tmem_base, da, umma_desc_sw128(a_addr + B0_OFF + k * 32), IDESC0, acc);
            if (NB > 0) umma_bf16_2cta(tmem_base + 256, da, umma_desc_sw128(a_addr + B1_OFF + k * 32), IDESC1, acc);
          }
          umma_commit_2cta(&empty_bar[s]);
        }
        umma_commit_2cta(tfull_bar);
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): TMEM lane = token; registers <-> TMEM <-> the row tile in shared memory ----
    constexpr int NCH = CPW / 16;           // 16-column chunks per warp: 4 / 6
    constexpr int NSUB = N_OUT / 64;        // sub-tiles of the row tile
    const int q = warp & 3;                 // TMEM lane quarter
    const int part = (warp - 2) >> 2;       // which quarter of the feature columns
    const int row = q * 32 + lane;          // token inside this CTA's 128
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (uint32_t)(part * CPW);
    // explicit shared-window addresses: pointers rebuilt from the aligned dynamic-smem base compile to generic LD.E
    const uint32_t bias_a = smem_u32(bias_s) + (uint32_t)(part * CPW) * 4u, gamma_a = smem_u32(gamma_s) + (uint32_t)(part * CPW) * 4u,
                   beta_a = smem_u32(beta_s) + (uint32_t)(part * CPW) * 4u, part_a = smem_u32(part_s);
    const uint32_t rowt_a = smem_u32(rowt) + (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    // 16 bytes = 8 columns of this thread's row: column c (multiple of 8) of the [tokens][N_OUT] tile
    auto slot = [&](int c) { return rowt_a + (uint32_t)(c >> 6) * 16384u + ((((uint32_t)(c & 63) >> 3) ^ sw) << 4); };
    const bool io = warp == 2 && lane == 0;  // issues the row tile's TMA loads and stores
    auto load_rows = [&](int t) {
      mbar_expect_tx(res_bar, ROW_TILE);
#pragma unroll
      for (int sb = 0; sb < NSUB; ++sb)  // rows past T are zero-filled
        tma_load_2d(rowt + sb * 16384, &tmX, res_bar, sb * 64, t * 256 + (int)rank * 128, kEvictFirst);
    };
    pdl_wait();  // the residual stream is written by earlier kernels
    if (io && pair < p.n_tiles) load_rows(pair);
    int i = 0;
    for (int t = pair; t < p.n_tiles; t += p.n_pairs, ++i) {
      const int tok = t * 256 + (int)rank * 128 + row;
      const bool ok = tok < p.T;
      mbar_wait(tfull_bar, i & 1);
      mbar_wait(res_bar, i & 1);
      tc_fence_after();
      // pass 1: x = acc + bias + residual, back into tensor memory; partial sums
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 2
      for (int ch = 0; ch < NCH; ++ch) {
        uint32_t v[16];
        tmem_ld16(t_lane + ch * 16, v);
        const int c = part * CPW + ch * 16;
        const uint4 r0 = lds_u32x4(slot(c)), r1 = lds_u32x4(slot(c + 8));
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        float bb[16];
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 b4 = lds_f32x4(bias_a + (uint32_t)(ch * 16 + k4 * 4) * 4u);
          bb[k4 * 4] = b4.x; bb[k4 * 4 + 1] = b4.y; bb[k4 * 4 + 2] = b4.z; bb[k4 * 4 + 3] = b4.w;
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float res = __uint_as_float((j & 1) ? (rr[j >> 1] & 0xffff0000u) : (rr[j >> 1] << 16));
          const float xv = __uint_as_float(v[j]) + bb[j] + res;
          s1 += xv;
          s2 = fmaf(xv, xv, s2);
          v[j] = __float_as_uint(xv);
        }
        tmem_st16_rl(t_lane + ch * 16, v);
      }
      sts_f32x2(part_a + (uint32_t)(part * 128 + row) * 8u, s1, s2);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("bar.sync 1, 512;" ::: "memory");
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // fixed order: the same statistics in all four warps of a row
        const float2 pp = lds_f32x2(part_a + (uint32_t)(k * 128 + row) * 8u);
        m1 += pp.x;
        m2 += pp.y;
      }
      const float mean = m1 * (1.0f / N_OUT);
      const float rstd = rsqrtf(fmaxf(m2 * (1.0f / N_OUT) - mean * mean, 0.f) + p.eps);
      // pass 2: normalise, bf16, over the residual in the row tile
#pragma unroll 2
      for (int ch = 0; ch < NCH; ++ch) {
        uint32_t v[16];
        tmem_ld16(t_lane + ch * 16, v);
        const int c = part * CPW + ch * 16;
        float gg[16], be[16];
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 g4 = lds_f32x4(gamma_a + (uint32_t)(ch * 16 + k4 * 4) * 4u), b4 = lds_f32x4(beta_a + (uint32_t)(ch * 16 + k4 * 4) * 4u);
          gg[k4 * 4] = g4.x; gg[k4 * 4 + 1] = g4.y; gg[k4 * 4 + 2] = g4.z; gg[k4 * 4 + 3] = g4.w;
          be[k4 * 4] = b4.x; be[k4 * 4 + 1] = b4.y; be[k4 * 4 + 2] = b4.z; be[k4 * 4 + 3] = b4.w;
        }
        tmem_ld_wait();
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) y[j] = fmaf((__uint_as_float(v[j]) - mean) * rstd, gg[j], be[j]);
        uint4 w0, w1;
        w0.x = pack_bf16(y[0], y[1]); w0.y = pack_bf16(y[2], y[3]); w0.z = pack_bf16(y[4], y[5]); w0.w = pack_bf16(y[6], y[7]);
        w1.x = pack_bf16(y[8], y[9]); w1.y = pack_bf16(y[10], y[11]); w1.z = pack_bf16(y[12], y[13]); w1.w = pack_bf16(y[14], y[15]);
        sts_u32x4(slot(c), w0);
        sts_u32x4(slot(c + 8), w1);
        if (p.h32 && ok) {   // (the model's last LayerNorm only)
          float4* hp = reinterpret_cast<float4*>(p.h32 + (size_t)tok * N_OUT + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) hp[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        }
      }
      // the accumulator has been read twice and is free; the row tile holds the result
      fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA engine
      tc_fence_before();
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (io) {
        if (leader_cta) mbar_arrive(tempty_bar);
        else mbar_arrive_remote(tempty_bar, 0);
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb)  // rows past T are clipped by the tensor map
          tma_store_2d(&tmX, rowt + sb * 16384, sb * 64, t * 256 + (int)rank * 128);
        tma_store_commit();
        tma_store_wait_read();                              // the row tile may be overwritten ...
        if (t + p.n_pairs < p.n_tiles) load_rows(t + p.n_pairs);  // ... by the next tile's residual rows, under its MMAs
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may free TMEM / exit while the pair's MMAs or epilogues still use its memory
  if (warp == 1) tmem_dealloc_2cta<512>(tmem_base);
}

}  // namespace mq

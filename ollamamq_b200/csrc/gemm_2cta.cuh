// 2-CTA (cta_group::2) variant of the prefill tcgen05 GEMM: one 256-feature x 256-token tile per CTA PAIR.
//
// Why (r01 ncu, prefill tiles): the 1-CTA kernel keeps the tensor pipe only 57-79 % busy and a persistent,
// epilogue-overlapped variant changed nothing, i.e. the MMAs wait on operands, not on epilogues.  Per 128x256x16
// MMA an SM reads 12 KiB of operands from shared memory while TMA writes another 12 KiB per 128 cycles into the
// same memory.  With cta_group::2 the pair computes a 256x256 tile: each CTA stages its OWN 128 weight rows and
// only HALF of the token tile (the tensor cores read the other half from the peer's shared memory), so the bytes
// written into each SM's shared memory (and pulled through L2) per MMA drop by a third (single accumulator) or a
// quarter (gate/up dual accumulator), and the stage count rises from 4/3 to 6/4.
//
// Protocol (CTA rank 0 = leader):
//   * both producers issue cp.async.bulk.tensor ... .cta_group::2 loads that complete_tx on the LEADER's full barrier
//     (count 2: leader arrive.expect_tx(2 x stage bytes) + one remote arrive from the peer's producer);
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256) and releases a stage with a
//     multicast tcgen05.commit that arrives on the empty barrier of BOTH CTAs; the final commit arrives on both
//     tmem_full barriers;
//   * each CTA drains its own 128 TMEM lanes (its half of the features).
//
// Round 2: PERSISTENT, two TMEM accumulator buffers.  r01's ncu showed the tensor pipe 61-67 % active: every 256x256
// tile paid its pipeline fill and its epilogue (TMEM -> registers -> smem -> TMA store) un-overlapped, 3.9 waves per
// GEMM with a ragged last one.  Now one CTA pair per SM pair walks tiles pair, pair + P, ...: the producers keep the
// TMA ring full across tile boundaries, the MMA thread flips between two accumulator buffers (2 x 256 TMEM columns;
// the dual SiLU epilogue uses 128-token tiles so that gate + up of both buffers fit the 512 columns), and the eight
// epilogue warps drain buffer b (straight from registers to global memory - no staging tile, so the ring keeps its
// depth) while the tensor cores fill buffer b ^ 1.  tmem_empty[b] (count 2: one arrive per CTA of the pair, the
// peer's through mapa) tells the leader's MMA thread that both halves of buffer b have been read.
#pragma once
#include "gemm.cuh"

namespace mq {

struct TwoCtaParams {
  void* out;       // [T][ldo] bf16 (EPI_BF16 / EPI_SILU_BF16 / EPI_BIAS_BF16 / EPI_GELU_BF16) or fp32 (EPI_F32)
  int ldo;
  const __nv_bfloat16* bias;  // EPI_BIAS_BF16 / EPI_GELU_BF16: [n_out] (nullable)
  int T, n_out, k_blocks, a2_row_off;
  int m_tiles, n_tiles, group_m;  // tiles of 256 features x c2_bn(epi) tokens
  int n_pairs;                    // persistent CTA pairs launched
  unsigned long long w_policy;
  RstdIn rs;                      // RMSNorm fold of the activation operand: token column t is scaled by rstd[t] (gemm.cuh)
  // EPI_RESID: out = the fp32 residual stream h (ldo = its leading dimension), updated in place
  const __nv_bfloat16* gamma_next;  // [n_out] weight of the NEXT RMSNorm
  __nv_bfloat16* xg;                // [T][ldx] bf16(h * gamma_next)
  int ldx;
  float* ssq_out;                   // [n_out / 128][ssq_stride] per-128-feature-tile partial sums of h^2
  int ssq_stride;
};

// token columns per tile.  The dual SiLU epilogue keeps 256 too: gate + up then fill all 512 TMEM columns, i.e. ONE
// accumulator buffer (no MMA / epilogue overlap) - a 128-token tile would double-buffer but reads 12 KB of operands per
// MFLOP from shared memory instead of 8 and is bound by that (measured: prefill 432 -> 459 ms).
__host__ __device__ constexpr int c2_bn(int epi) { return 256; }
__host__ __device__ constexpr int c2_nbuf(int epi) { return epi == EPI_SILU_BF16 ? 1 : 2; }
__host__ __device__ constexpr int c2_tail_bytes() { return 256 /*barriers*/ + 256 * 4 /*rstd*/ + 4 * 256 * 4 /*h^2 partials*/; }
__host__ __device__ constexpr int c2_stage_bytes(int epi) {
  return kATileBytes * (epi == EPI_SILU_BF16 ? 2 : 1) + (c2_bn(epi) / 2) * kBlockK * 2;  // own weight rows + own half of the tokens
}
// The bias / GELU epilogues (encoder GEMMs: K = 384 or 1536, i.e. 6 - 24 k-blocks per tile) stage their 128-feature x
// 256-token half of the tile in shared memory and write it with ONE bulk tensor store: with so few k-blocks per tile
// the epilogue, not the MMAs, sets the tile time, and 2-byte stores straight from registers (64 B per warp instruction)
// took 14 000 cycles per tile against 3 000 cycles of MMAs.  The staging tile costs two of the six pipeline stages,
// which a 6-k-block tile does not miss.
__host__ __device__ constexpr bool c2_staged(int epi) { return epi == EPI_BIAS_BF16 || epi == EPI_GELU_BF16; }
__host__ __device__ constexpr int c2_stage_tile_bytes(int epi) { return c2_staged(epi) ? c2_bn(epi) * kBlockM * 2 : 0; }
__host__ __device__ constexpr int c2_stages(int epi) {
  int s = (200 * 1024 - c2_stage_tile_bytes(epi)) / c2_stage_bytes(epi);
  return s > 8 ? 8 : s;
}
__host__ __device__ constexpr int c2_smem_bytes(int epi) {
  return c2_stages(epi) * c2_stage_bytes(epi) + c2_stage_tile_bytes(epi) + 1024 + c2_tail_bytes();
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader's copy
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const void* tmap, uint64_t* bar_leader_local, int c0,
                                                 int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar_leader_local) & kPeerBitMask),
        "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {  // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// producer warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter); the staged epilogues run 16 (four per quarter,
// 64 token columns each): their tiles have 6 - 24 k-blocks, the epilogue sets the tile time and is a chain of
// TMEM-load / ALU latencies that only more warps hide
__host__ __device__ constexpr int c2_epi_warps(int epi) { return c2_staged(epi) ? 16 : 8; }
__host__ __device__ constexpr int c2_threads(int epi) { return 64 + 32 * c2_epi_warps(epi); }

// erf-GELU with the Abramowitz-Stegun 7.1.26 rational form (|error of erf| < 1.5e-7, far below the bf16 output's 2^-9):
// 14 instructions and two MUFU results per element where erff() compiles to ~35 with divergent range branches - the
// encoder's up-projection epilogue is instruction-bound (6 k-blocks of MMAs per 256 x 256 tile).
__device__ __forceinline__ float gelu_erf_fast(float a) {
  const float x = a * 0.70710678118654752f;
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float erf_x = copysignf(erf_abs, x);
  const float ha = 0.5f * a;
  return fmaf(ha, erf_x, ha);
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(c2_threads(EPI), 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const TwoCtaParams p) {
  constexpr bool kDual = (EPI == EPI_SILU_BF16);
  constexpr int BN = c2_bn(EPI);                // token columns of the pair's accumulator
  constexpr int STAGES = c2_stages(EPI);
  constexpr int STAGE_BYTES = c2_stage_bytes(EPI);
  constexpr int B_OFF = kATileBytes * (kDual ? 2 : 1);
  constexpr int ACC_COLS = BN * (kDual ? 2 : 1);  // TMEM columns of one accumulator buffer
  constexpr int NBUF = c2_nbuf(EPI);              // 2: the epilogue of tile i overlaps the MMAs of tile i + 1
  constexpr bool kStaged = c2_staged(EPI);        // output through a shared-memory tile + one bulk tensor store
  constexpr int STG_BYTES = c2_stage_tile_bytes(EPI);
  constexpr uint32_t TMEM_COLS = 512u;
  constexpr uint32_t IDESC = umma_idesc_bf16(256, BN);
  static_assert(NBUF * ACC_COLS <= 512, "the accumulator buffers must fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + STAGES * STAGE_BYTES;     // [256 tokens][128 features] bf16 (kStaged)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STG_BYTES);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator buffer complete (multicast commit: both CTAs)
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] used in the leader only: both CTAs have drained the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* rstd_s = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + STG_BYTES + 256);  // [256] rstd of the tile's tokens
  float* ssq_s = rstd_s + 256;                                                  // [4 lane quarters][256 tokens]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_tiles_total = p.m_tiles * p.n_tiles;
  const int nkb = p.k_blocks;
  auto tile_of = [&](int t, int* tile_m, int* tile_n) {  // grouped rasterisation (see gemm.cuh)
    const int per_group = p.group_m * p.n_tiles;
    const int first_m = (t / per_group) * p.group_m;
    const int gsz = min(p.m_tiles - first_m, p.group_m);
    const int r = t % per_group;
    *tile_m = first_m + r % gsz;
    *tile_n = r / gsz;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx + the peer producer's remote arrive
      mbar_init(&empty_bar[s], 1);  // multicast commit from the leader's MMA thread
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 2);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation visible to both CTAs before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs): own weight rows + own half of the token tile, all tiles ----------------
      const int my_tiles = pair < n_tiles_total ? (n_tiles_total - pair + p.n_pairs - 1) / p.n_pairs : 0;
      const int total = my_tiles * nkb;  // k-blocks this producer issues (ring position = it)
      auto coords = [&](int it, int* kb, int* m0, int* nb0) {
        int tile_m, tile_n;
        tile_of(pair + (it / nkb) * p.n_pairs, &tile_m, &tile_n);
        *kb = it % nkb;
        *m0 = tile_m * 256 + (int)rank * kBlockM;   // my 128 weight rows
        *nb0 = tile_n * BN + (int)rank * (BN / 2);  // my half of the token tile
      };
      auto load_a = [&](int it) {
        int kb, m0, nb0;
        coords(it, &kb, &m0, &nb0);
        const int s = it % STAGES;
        if (leader_cta) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
        else mbar_arrive_remote(&full_bar[s], 0);
        uint8_t* st = smem + s * STAGE_BYTES;
        tma_load_2d_2cta(st, &tmA, &full_bar[s], kb * kBlockK, m0, p.w_policy);
        if (kDual) tma_load_2d_2cta(st + kATileBytes, &tmA, &full_bar[s], kb * kBlockK, m0 + p.a2_row_off, p.w_policy);
      };
      auto load_b = [&](int it) {
        int kb, m0, nb0;
        coords(it, &kb, &m0, &nb0);
        tma_load_2d_2cta(smem + (it % STAGES) * STAGE_BYTES + B_OFF, &tmB, &full_bar[it % STAGES], kb * kBlockK, nb0, kEvictLast);
      };
      const int npre = total < STAGES ? total : STAGES;
      for (int it = 0; it < npre; ++it) load_a(it);  // weights of the first ring pass: no dependency on the previous kernel
      pdl_wait();                                    // activations are produced by the previous kernel
      for (int it = 0; it < npre; ++it) load_b(it);
      for (int it = npre; it < total; ++it) {
        mbar_wait(&empty_bar[it % STAGES], ((it / STAGES) & 1) ^ 1);
        load_a(it);
        load_b(it);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader_cta) {
      // ---------------- MMA issuer: leader CTA only; tile i accumulates into buffer i & 1 ----------------
      int it = 0, i = 0;
      for (int t = pair; t < n_tiles_total; t += p.n_pairs, ++i) {
        const int buf = i % NBUF, use = i / NBUF;
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);  // both CTAs have drained this buffer (first use: passes at once)
        tc_fence_after();
        const uint32_t acc_base = tmem_base + (uint32_t)(buf * ACC_COLS);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + B_OFF;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t db = umma_desc_sw128(b_addr + k * 32);
            const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
            umma_bf16_2cta(acc_base, umma_desc_sw128(a_addr + k * 32), db, IDESC, acc);
            if (kDual) umma_bf16_2cta(acc_base + BN, umma_desc_sw128(a_addr + kATileBytes + k * 32), db, IDESC, acc);
          }
          umma_commit_2cta(&empty_bar[s]);  // stage s reusable in BOTH CTAs once these MMAs retire
        }
        umma_commit_2cta(&tfull_bar[buf]);  // accumulator buffer complete, in both CTAs
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): my 128 TMEM lanes = my 128 features; registers -> global ----------------
    const int q = warp & 3;                 // TMEM lane quarter (warps w and w + 4 share one)
    constexpr int NEPI = 32 * c2_epi_warps(EPI);   // epilogue threads (named barrier 1)
    constexpr int CPW = BN / (c2_epi_warps(EPI) / 4);  // token columns per warp: 128 (8 warps) or 64 (16)
    const int chalf = (warp - 2) >> 2;      // which slice of the token columns this warp drains
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;        // 0..NEPI-1
    const bool fold = p.rs.ssq != nullptr;
    if (fold || EPI == EPI_RESID) pdl_wait();  // rstd partials / the residual stream come from earlier kernels
    int i = 0;
    for (int t = pair; t < n_tiles_total; t += p.n_pairs, ++i) {
      int tile_m, tile_n;
      tile_of(t, &tile_m, &tile_n);
      const int f = tile_m * 256 + (int)rank * kBlockM + row;  // this thread's output feature
      const bool f_ok = f < p.n_out;  // n_out need not fill the last 256-row tile (encoder: 384 / 1152 features)
      float bias_f = 0.f;
      if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_GELU_BF16)
        bias_f = (p.bias != nullptr && f_ok) ? __bfloat162float(p.bias[f]) : 0.f;
      const int n0 = tile_n * BN;
      const int buf = i % NBUF, use = i / NBUF;
      if (fold) {  // per-token RMSNorm scale of this tile's 256 tokens (under the MMAs of the tile)
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");  // previous tile's readers are done with rstd_s
        if (et < 256) rstd_s[et] = (n0 + et < p.T) ? rstd_of(p.rs, n0 + et) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");
      }
      float gm = 0.f;
      float hnext[16];  // EPI_RESID: residual values of the chunk about to be processed
      if constexpr (EPI == EPI_RESID) {
        gm = __bfloat162float(p.gamma_next[f]);
        const int c0 = chalf * CPW;
        const float* hp = reinterpret_cast<const float*>(p.out) + (size_t)(n0 + c0) * p.ldo + f;
#pragma unroll
        for (int j = 0; j < 16; ++j) hnext[j] = (n0 + c0 + j < p.T) ? hp[(size_t)j * p.ldo] : 0.f;
      }
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + (uint32_t)(buf * ACC_COLS) + (static_cast<uint32_t>(q * 32) << 16);
      if constexpr (kStaged) {
        // Phase 1, TMEM -> registers: every value of this thread's 64 token columns is converted and packed (bf16 pairs
        // of neighbouring tokens) before anything touches the staging tile - the previous tile's bulk store may still be
        // reading it, and that read (~1 400 cycles per tile when the warps waited for it up front) now hides under the
        // conversion.  (The rstd fold is not offered by these epilogues.)
        constexpr int NCH = CPW / 16;
        const int cbeg = chalf * CPW;
        const int cend = min((chalf + 1) * CPW, (p.T - n0 + 15) & ~15);  // columns past T are clipped by the store
        uint32_t pk[NCH * 8];
        static_assert(NCH == 4, "64 token columns per epilogue warp: 32 packed registers");
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int c0 = cbeg + ch * 16;
          if (c0 < cend) {  // warp-uniform
            uint32_t v[16];  // (one load in flight per warp: a second buffer spills at the 96 registers 18 warps leave)
            tmem_ld16(t_lane + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float a0 = __uint_as_float(v[j]) + bias_f, a1 = __uint_as_float(v[j + 1]) + bias_f;
              if constexpr (EPI == EPI_GELU_BF16) { a0 = gelu_erf_fast(a0); a1 = gelu_erf_fast(a1); }
              pk[ch * 8 + j / 2] = pack_bf16(a0, a1);
            }
          }
        }
        // the accumulator is in registers: hand the TMEM buffer back; and the staging tile must be free
        if (i > 0 && warp == 2 && lane == 0) tma_store_wait_read();
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");
        if (warp == 2 && lane == 0) {
          if (leader_cta) mbar_arrive(&tempty_bar[buf]);
          else mbar_arrive_remote(&tempty_bar[buf], 0);
        }
        // Phase 2, registers -> staging tile [token][128 features] -> one bulk tensor store
        const uint32_t o = smem_u32(stg) + (uint32_t)(cbeg * kBlockM + row) * 2u;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          if (cbeg + ch * 16 < cend) {
#pragma unroll
            for (int q2 = 0; q2 < 8; ++q2) {
              const uint32_t w = pk[ch * 8 + q2];
              sts_u16(o + (uint32_t)((ch * 16 + 2 * q2) * kBlockM) * 2u, w);
              sts_u16(o + (uint32_t)((ch * 16 + 2 * q2 + 1) * kBlockM) * 2u, w >> 16);
            }
          }
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA engine
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");
        if (warp == 2 && lane == 0) {  // rows past n_out / columns past T are clipped by the tensor map
          tma_store_2d(&tmC, stg, tile_m * 256 + (int)rank * kBlockM, n0);
          tma_store_commit();
        }
      } else {
#pragma unroll 1
      for (int c0 = chalf * CPW; c0 < (chalf + 1) * CPW; c0 += 16) {
        if (n0 + c0 >= p.T) {
          if constexpr (EPI == EPI_RESID) {  // columns past T: zero partials so the tile's sums stay defined
            if (lane == 0)
              for (int j = 0; j < 16; ++j) ssq_s[q * 256 + c0 + j] = 0.f;
            continue;
          } else {
            break;
          }
        }
        uint32_t v[16];
        tmem_ld16(t_lane + c0, v);
        if constexpr (kDual) {
          uint32_t u[16];
          tmem_ld16(t_lane + BN + c0, u);
          tmem_ld_wait();
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)(n0 + c0) * p.ldo + f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float sc = fold ? rstd_s[c0 + j] : 1.f;
            const float g = __uint_as_float(v[j]) * sc;
            if (n0 + c0 + j < p.T) o[(size_t)j * p.ldo] = __float2bfloat16(__fdividef(g, 1.0f + __expf(-g)) * (__uint_as_float(u[j]) * sc));
          }
        } else if constexpr (EPI == EPI_RESID) {
          float* hp = reinterpret_cast<float*>(p.out) + (size_t)(n0 + c0) * p.ldo + f;
          float hv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) hv[j] = hnext[j];
          // the residual rows of the NEXT chunk are requested before this one is touched: one L2 round trip per
          // chunk on the critical path made this epilogue slower than the kernel it replaces
          {
            const int c1 = c0 + 16;
            const float* hn = hp + (size_t)16 * p.ldo;
            const bool more = c1 < (chalf + 1) * CPW;
#pragma unroll
            for (int j = 0; j < 16; ++j) hnext[j] = (more && n0 + c1 + j < p.T) ? hn[(size_t)j * p.ldo] : 0.f;
          }
          tmem_ld_wait();
          __nv_bfloat16* xp = p.xg + (size_t)(n0 + c0) * p.ldx + f;
          float sq[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const bool ok = n0 + c0 + j < p.T;
            hv[j] = ok ? hv[j] + __uint_as_float(v[j]) : 0.f;
            if (ok) {
              hp[(size_t)j * p.ldo] = hv[j];
              xp[(size_t)j * p.ldx] = __float2bfloat16(hv[j] * gm);
            }
            sq[j] = hv[j] * hv[j];
          }
          // 16 column sums over the warp's 32 features with 16 + 15 shuffles instead of 80: at every step a lane keeps
          // half of its values and receives the partner's other half (fixed pattern: deterministic, slot-independent)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool up = lane & 16;
            const float send = up ? sq[j] : sq[j + 8], keep = up ? sq[j + 8] : sq[j];
            sq[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool up = lane & 8;
            const float send = up ? sq[j] : sq[j + 4], keep = up ? sq[j + 4] : sq[j];
            sq[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const bool up = lane & 4;
            const float send = up ? sq[j] : sq[j + 2], keep = up ? sq[j + 2] : sq[j];
            sq[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
          {
            const bool up = lane & 2;
            const float send = up ? sq[0] : sq[1], keep = up ? sq[1] : sq[0];
            sq[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
          sq[0] += __shfl_xor_sync(0xffffffffu, sq[0], 1);
          // lane l now holds the sum of column ((l >> 4) & 1) * 8 + ((l >> 3) & 1) * 4 + ((l >> 2) & 1) * 2 + ((l >> 1) & 1)
          if ((lane & 1) == 0) {
            const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            ssq_s[q * 256 + c0 + col] = sq[0];
          }
        } else {
          tmem_ld_wait();
          if constexpr (EPI == EPI_F32) {
            float* o = reinterpret_cast<float*>(p.out) + (size_t)(n0 + c0) * p.ldo + f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (f_ok && n0 + c0 + j < p.T) o[(size_t)j * p.ldo] = __uint_as_float(v[j]) * (fold ? rstd_s[c0 + j] : 1.f);
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)(n0 + c0) * p.ldo + f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (f_ok && n0 + c0 + j < p.T) o[(size_t)j * p.ldo] = __float2bfloat16(__uint_as_float(v[j]) * (fold ? rstd_s[c0 + j] : 1.f));
          }
        }
      }
      }  // !kStaged
      if constexpr (!kStaged) {
        // this CTA's half of buffer `buf` is in registers / on its way to memory: hand the buffer back to the MMA thread
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");
        if (warp == 2 && lane == 0) {
          if (leader_cta) mbar_arrive(&tempty_bar[buf]);
          else mbar_arrive_remote(&tempty_bar[buf], 0);
        }
      }
      if constexpr (EPI == EPI_RESID) {  // this CTA's 128 features of the tile: one partial per token, quarters in order
        if (n0 + et < p.T)
          p.ssq_out[(size_t)(tile_m * 2 + (int)rank) * p.ssq_stride + n0 + et] =
              (ssq_s[et] + ssq_s[256 + et]) + (ssq_s[512 + et] + ssq_s[768 + et]);
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");  // ssq_s is rewritten by the next tile
      }
    }
  }

  if (kStaged && warp == 2 && lane == 0) tma_store_wait_read();  // the staging tile must outlive the last bulk store's read
  tc_fence_before();
  cluster_sync_all();  // neither CTA may free TMEM / exit while the pair's MMAs or epilogues still use its memory
  if (warp == 1) tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
}

}  // namespace mq

// Persistent variant of the prefill tcgen05 GEMM (single accumulator, BN = 128 / 256 token tiles).
//
// Why: in gemm_wx_kernel every CTA pays barrier init + TMEM alloc + first-TMA latency per tile and its epilogue
// runs after the last MMA with the tensor pipe idle (ncu r01: sm__pipe_tensor_cycles_active 66-79 % inside the
// O / down / QKV prefill kernels).  Here 148 CTAs stay resident and walk the (super-tile rasterised) tile list;
// TMEM holds TWO accumulator buffers (2 x BN columns), so warps 2..5 drain tile i (tcgen05.ld -> bf16 -> global)
// while the MMA thread is already issuing tile i+1 and the TMA ring never drains at tile boundaries.
#pragma once
#include "gemm.cuh"

namespace mq {

struct PersistParams {
  void* out;  // [T][ldo] bf16 (EPI_BF16) or fp32 (EPI_F32)
  int ldo, T, n_out, k_blocks, m_tiles, n_tiles, group_m, n_ctas;
  unsigned long long w_policy;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const PersistParams p) {
  static_assert(EPI != EPI_SILU_BF16 && 2 * BN <= 512, "single accumulator, two TMEM buffers");
  constexpr int STAGES = gemm_stages(BN, EPI);
  constexpr int STAGE_BYTES = gemm_stage_bytes(BN, EPI);
  constexpr int B_OFF = kATileBytes;
  constexpr uint32_t TMEM_COLS = 2 * BN <= 256 ? 256u : 512u;
  constexpr uint32_t IDESC = umma_idesc_bf16(kBlockM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x, P = p.n_ctas, KB = p.k_blocks;
  const int total_tiles = p.m_tiles * p.n_tiles;
  auto tile_of = [&](int pid, int& tm, int& tn) {  // same super-tile rasterisation as gemm_wx_kernel
    const int per_group = p.group_m * p.n_tiles;
    const int first_m = (pid / per_group) * p.group_m;
    const int gsz = min(p.m_tiles - first_m, p.group_m);
    const int r = pid % per_group;
    tm = first_m + r % gsz;
    tn = r / gsz;
  };

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

  if (warp == 0) {
    if (lane == 0 && c < total_tiles) {
      // ---------------- TMA producer: one continuous stream over this CTA's tiles ----------------
      const int my_tiles = (total_tiles - c + P - 1) / P;
      const int total = my_tiles * KB;
      int tile = c, kb = 0, tm, tn;
      tile_of(tile, tm, tn);
      auto advance = [&]() {
        if (++kb == KB) {
          kb = 0;
          tile += P;
          if (tile < total_tiles) tile_of(tile, tm, tn);
        }
      };
      const int npre = total < STAGES ? total : STAGES;
      for (int i = 0; i < npre; ++i) {  // weights first: they do not depend on the previous kernel
        mbar_expect_tx(&full_bar[i], STAGE_BYTES);
        tma_load_2d(smem + i * STAGE_BYTES, &tmA, &full_bar[i], kb * kBlockK, tm * kBlockM, p.w_policy);
        advance();
      }
      pdl_wait();
      tile = c; kb = 0;
      tile_of(tile, tm, tn);
      for (int i = 0; i < npre; ++i) {
        tma_load_2d(smem + i * STAGE_BYTES + B_OFF, &tmB, &full_bar[i], kb * kBlockK, tn * BN, kEvictLast);
        advance();
      }
      for (int it = npre; it < total; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_2d(st, &tmA, &full_bar[s], kb * kBlockK, tm * kBlockM, p.w_policy);
        tma_load_2d(st + B_OFF, &tmB, &full_bar[s], kb * kBlockK, tn * BN, kEvictLast);
        advance();
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer: tile i accumulates into TMEM buffer i & 1 ----------------
      int it = 0, i = 0;
      for (int tile = c; tile < total_tiles; tile += P, ++i) {
        const int buf = i & 1, use = i >> 1;
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc_base = tmem_base + buf * BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t b_addr = a_addr + B_OFF;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(acc_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), IDESC,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // ---------------- epilogue warps: drain buffer i & 1 while the next tile is being computed ----------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool leader = (warp == 2 && lane == 0);
    int i = 0;
    for (int tile = c; tile < total_tiles; tile += P, ++i) {
      const int buf = i & 1, use = i >> 1;
      int tm, tn;
      tile_of(tile, tm, tn);
      const int f = tm * kBlockM + row, n0 = tn * BN;
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        if (n0 + c0 >= p.T) break;  // warp-uniform
        uint32_t v[16];
        tmem_ld16(t_lane + c0, v);
        tmem_ld_wait();
        if (f < p.n_out) {
          if constexpr (EPI == EPI_F32) {
            float* o = reinterpret_cast<float*>(p.out) + (size_t)(n0 + c0) * p.ldo + f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c0 + j < p.T) o[(size_t)j * p.ldo] = __uint_as_float(v[j]);
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)(n0 + c0) * p.ldo + f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c0 + j < p.T) o[(size_t)j * p.ldo] = __float2bfloat16(__uint_as_float(v[j]));
          }
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (leader) mbar_arrive(&tempty_bar[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace mq

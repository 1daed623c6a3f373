// Prefill attention on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim 128.
// Part of the forward pass that stands where the reference calls a remote backend (/root/reference/src/dispatcher.rs:287-290);
// BASELINE.json north_star: "prefill as tcgen05 tensor-core GEMMs fed by TMA into shared memory".  Round 1 ran prefill
// attention on the legacy mma.sync pipe (33 % tensor-active, 7.4 % of a prefill pass).
//
// One CTA per (query tile, kv head), TWO CTAs resident per SM (96 KB of shared memory, 256 TMEM columns, <= 168
// registers each): inside a CTA the tensor pipe and the softmax warps alternate, and the SM overlaps the MMAs of one CTA
// with the exponentials of the other.  A query tile is 128 rows = 128/G tokens x the G query heads of the GQA group
// (row = token * G + head), so K / V tiles are shared by the whole group.  Per 128-token KV tile j:
//
//   warp 0 (one thread)   TMA: the 8 pages of K and of V of the tile (separate single-stage buffers, each refilled as soon
//                         as its MMA retires), each page = one 16 x 128 block of the paged cache fetched as two
//                         {64 d, 16 token} boxes with the 128-byte swizzle -> [128 tokens][64 d] x 2 sub-tiles = the
//                         canonical K-major (K) / MN-major (V) UMMA operand layouts; Q once, as {64 d, G heads,
//                         128/G tokens} boxes of the [T][n_q][128] activation
//   warp 1 (one thread)   S  = Q . K_j^T     tcgen05.mma, A and B from shared memory, fp32 in TMEM
//                         O += P . V_j       tcgen05.mma, A = P from TENSOR MEMORY (bf16), B = V MN-major; O stays in TMEM
//   warps 2-5 (128 thr)   one query row per thread (TMEM lane = row: no shuffles anywhere): two passes over its 128
//                         scores (row max, then p = exp2(s - m)), P written back into the S buffer's first 64 columns
//                         as packed bf16 (tcgen05.st).  The running maximum is only raised - and the O row in TMEM
//                         rescaled by exp2(m_old - m_new) - when a warp sees a score more than 2^8 above it (the
//                         probabilities then stay below 256: exact in fp32 sums, 8 bits of headroom in bf16).
//
// TMEM: columns [0,128) S / P, [128,256) O.
#include "kernels.cuh"
#include "gemm.cuh"
#include "attn_tc.cuh"
#include <cudaTypedefs.h>
#include <mutex>

namespace mq {

constexpr int kTcRows = 128;       // query rows per CTA
constexpr int kTcKv = 128;         // kv tokens per tile (8 pages)
constexpr int kTcD = 128;          // head_dim
constexpr int kTcSub = 128 * 128;  // bytes of one [128 rows][64 elements] bf16 sub-tile
constexpr int kTcSmem = 2 * kTcSub /*Q*/ + 2 * kTcSub /*K*/ + 2 * kTcSub /*V*/ + 1024 /*align*/ + 256 /*barriers*/;  // 97.25 KB
constexpr float kTcRescale = 8.f;  // log2 units: raise the running maximum only when a score tops it by more than 2^8

struct AttnTcParams {
  CUtensorMap tmQ;   // [T][n_q][128] bf16: box {64, G, 128 / G}
  CUtensorMap tmK;   // [pages][n_kv][16][128] bf16: box {64, 16, 1, 1}
  CUtensorMap tmV;
  const int* block_table;
  int max_pages;
  const int4* tiles;  // {tok0, ntok, slot, pos0} per query tile
  __nv_bfloat16* out; // [T][n_q * 128]
  int n_q, n_kv;
  float scale_log2;
};

__global__ void __launch_bounds__(192, 2) prefill_attn_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                      // 2 sub-tiles each
  uint8_t* Ks = smem + 2 * kTcSub;
  uint8_t* Vs = smem + 4 * kTcSub;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTcSub);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;             // count 4 (one per softmax warp)
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvh = blockIdx.y;
  const int G = p.n_q / p.n_kv;
  const int4 tile = p.tiles[blockIdx.x];
  const int tok0 = tile.x, ntok = tile.y, slot = tile.z, pos0 = tile.w;
  const int kv_end = pos0 + ntok;                       // causal: the tile's last token sees positions < kv_end
  const int n_tiles = (kv_end + kTcKv - 1) / kTcKv;
  const int* btab = p.block_table + (size_t)slot * p.max_pages;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], i == 6 ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;
  pdl_launch_dependents();
  pdl_wait();  // q and this pass's K / V rows come from the rope kernel

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      mbar_expect_tx(q_full, 2 * kTcSub);
      // rows >= ntok * G of the tile belong to other sequences (or lie past T): they are computed and never stored
      tma_load_3d(Qs, &p.tmQ, q_full, 0, kvh * G, tok0);
      tma_load_3d(Qs + kTcSub, &p.tmQ, q_full, 64, kvh * G, tok0);
      auto load = [&](const CUtensorMap* tm, uint8_t* dst, uint64_t* bar, int j) {
        mbar_expect_tx(bar, 2 * kTcSub);
        for (int pg = 0; pg < kTcKv / kPageSize; ++pg) {
          const int t = j * kTcKv + pg * kPageSize;
          const int page = t < kv_end ? btab[t / kPageSize] : 0;  // past the end: the scratch page (masked below)
          tma_load_4d(dst + pg * 2048, tm, bar, 0, 0, kvh, page);
          tma_load_4d(dst + kTcSub + pg * 2048, tm, bar, 64, 0, kvh, page);
        }
      };
      for (int j = 0; j < n_tiles; ++j) {
        if (j >= 1) mbar_wait(k_empty, (j - 1) & 1);   // S_{j-1} = Q K_{j-1}^T has retired
        load(&p.tmK, Ks, k_full, j);
        if (j >= 1) mbar_wait(v_empty, (j - 1) & 1);   // O += P_{j-1} V_{j-1} has retired
        load(&p.tmV, Vs, v_full, j);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t IDESC_S = umma_idesc_bf16(kTcRows, kTcKv);     // S = Q K^T: both operands K-major
      constexpr uint32_t IDESC_O = umma_idesc_bf16_bmn(kTcRows, kTcD);  // O = P V: V is MN-major
      const uint32_t q_addr = smem_u32(Qs), k_addr = smem_u32(Ks), v_addr = smem_u32(Vs);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(k_full, j & 1);
        tc_fence_after();
        // (the S / P columns are free: P_{j-1} V_{j-1} was issued before this and the pipe runs in order)
#pragma unroll
        for (int k = 0; k < kTcD / 16; ++k) {  // 8 k-steps over d: 4 per 64-wide sub-tile
          const uint32_t off = (uint32_t)(k >> 2) * kTcSub + (uint32_t)(k & 3) * 32;
          umma_bf16(tmem_base, umma_desc_sw128(q_addr + off), umma_desc_sw128(k_addr + off), IDESC_S, k != 0);
        }
        umma_commit(s_full);   // also tells the softmax warps that O holds every earlier P V (commit covers all prior MMAs)
        umma_commit(k_empty);
        mbar_wait(p_full, j & 1);                          // P_j is in TMEM (and O has been rescaled if it had to be)
        mbar_wait(v_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kTcKv / 16; ++k)               // 8 k-steps over the kv tokens: 16 rows = 2 KB each
          umma_bf16_ts(tmem_o, tmem_base + (uint32_t)(k * 8), umma_desc_sw128_mn(v_addr + k * 2048, kTcSub), IDESC_O,
                       (j | k) != 0);
        umma_commit(v_empty);
        if (j + 1 == n_tiles) umma_commit(o_full);
      }
    }
  } else {
    // ---------------- softmax + output: one query row per thread ----------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_o = t_lane + 128;
    const int qpos = pos0 + row / G;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int t0 = j * kTcKv;
      const bool edge = t0 + kTcKv > qpos + 1;             // some column of this tile is masked for this row
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // the diagonal tile masks columns per row; every other tile takes the unmasked loops (warp-uniform choice)
      const bool edge_w = __any_sync(0xffffffffu, edge);
      const int lim = qpos - t0;                          // columns <= lim are visible to this row
      // pass 1: row maximum (64 columns per TMEM round trip)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < kTcKv; c0 += 64) {
        uint32_t v[64];
        tmem_ld32(t_lane + (uint32_t)c0, v);
        tmem_ld32(t_lane + (uint32_t)(c0 + 32), v + 32);
        tmem_ld_wait();
        if (!edge_w) {
#pragma unroll
          for (int i = 0; i < 64; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) mx = fmaxf(mx, c0 + i <= lim ? __uint_as_float(v[i]) : -INFINITY);
        }
      }
      mx *= p.scale_log2;   // finite: column t0 <= qpos for every tile this row walks
      // raise the running maximum (and rescale the O row in TMEM) only when some row of this warp needs it
      const bool raise = __any_sync(0xffffffffu, mx > m_run + kTcRescale);
      if (raise) {
        const float m_new = fmaxf(m_run, mx);
        const float alpha = ex2_ftz(m_run - m_new);        // 0 on the first tile (m_run = -inf): O is not read then
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {
#pragma unroll 1
          for (int c0 = 0; c0 < kTcD; c0 += 64) {
            uint32_t v[64];
            tmem_ld32(t_o + (uint32_t)c0, v);
            tmem_ld32(t_o + (uint32_t)(c0 + 32), v + 32);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(t_o + (uint32_t)c0, v);
            tmem_st32(t_o + (uint32_t)(c0 + 32), v + 32);
          }
        }
      }
      // pass 2: probabilities -> packed bf16 into the first 64 columns of the same buffer (columns already consumed)
      float lsum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < kTcKv; c0 += 64) {
        uint32_t v[64], pk[32];
        tmem_ld32(t_lane + (uint32_t)c0, v);
        tmem_ld32(t_lane + (uint32_t)(c0 + 32), v + 32);
        tmem_ld_wait();
        if (!edge_w) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float p0 = ex2_ftz(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m_run));
            const float p1 = ex2_ftz(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m_run));
            lsum += p0 + p1;
            pk[i] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float p0 = ex2_ftz(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m_run));
            float p1 = ex2_ftz(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m_run));
            if (c0 + 2 * i > lim) p0 = 0.f;
            if (c0 + 2 * i + 1 > lim) p1 = 0.f;
            lsum += p0 + p1;
            pk[i] = pack_bf16(p0, p1);
          }
        }
        tmem_st32(t_lane + (uint32_t)(c0 / 2), pk);  // 64 probabilities = 32 packed columns
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      l_run += lsum;
    }
    // ---- normalise and store this row (256 contiguous bytes)
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    __nv_bfloat16* po = p.out + ((size_t)(tok0 + row / G) * p.n_q + kvh * G + row % G) * kTcD;
#pragma unroll 1
    for (int c0 = 0; c0 < kTcD; c0 += 64) {
      uint32_t v[64];
      tmem_ld32(t_o + (uint32_t)c0, v);
      tmem_ld32(t_o + (uint32_t)(c0 + 32), v + 32);
      tmem_ld_wait();
      if (row < ntok * G) {
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
          *reinterpret_cast<uint4*>(po + c0 + i) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

bool attn_tc_supported(int head_dim, int n_q, int n_kv) {
  const int G = n_kv > 0 ? n_q / n_kv : 0;
  return head_dim == kTcD && G >= 1 && G <= 8 && kTcRows % G == 0 && n_q % n_kv == 0;
}

bool attn_tc_encode_q(CUtensorMap* out, const void* q, int rows, int n_q, int G) {
  auto fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)kTcD, (cuuint64_t)n_q, (cuuint64_t)rows};
  const cuuint64_t strides[2] = {(cuuint64_t)kTcD * 2, (cuuint64_t)n_q * kTcD * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)G, (cuuint32_t)(kTcRows / G)};
  const cuuint32_t es[3] = {1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(q), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool attn_tc_encode_kv(CUtensorMap* out, const void* cache, int n_pages, int n_kv) {
  auto fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)kTcD, (cuuint64_t)kPageSize, (cuuint64_t)n_kv, (cuuint64_t)n_pages};
  const cuuint64_t strides[3] = {(cuuint64_t)kTcD * 2, (cuuint64_t)kPageSize * kTcD * 2, (cuuint64_t)n_kv * kPageSize * kTcD * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)kPageSize, 1, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(cache), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void attn_tc_set_attrs() {
  cudaFuncSetAttribute(prefill_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
}

cudaError_t launch_attn_prefill_tc(const LaunchCfg& lc, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                   const AttnParams& a, int n_tiles) {
  AttnTcParams p;
  p.tmQ = tmQ; p.tmK = tmK; p.tmV = tmV;
  p.block_table = a.block_table; p.max_pages = a.max_pages; p.tiles = a.tiles; p.out = a.out; p.n_q = a.n_q; p.n_kv = a.n_kv;
  p.scale_log2 = a.scale_log2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_tiles, a.n_kv, 1);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = kTcSmem;
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, prefill_attn_tc_kernel, p);
}

}  // namespace mq

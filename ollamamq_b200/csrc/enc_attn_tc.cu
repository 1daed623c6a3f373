// Bidirectional self-attention of the embedding worker on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim 32,
// sequences of up to 512 tokens packed back to back in one [T][3H] activation (q | k | v column blocks).
// Part of the forward pass that stands where the reference forwards /api/embed to a remote backend
// (/root/reference/src/dispatcher.rs:287-312); BASELINE configs[4] (bge-small: 12 heads of 32).
//
// Round 1 ran this on the mma.sync flash kernel: 70 us per layer for 8 192 tokens, the largest item of a pass (d = 32 means
// one exponential per 64 flops - the softmax, not the MMAs, is the bound: 16 MUFU results per clock and SM).  The shape
// of this kernel follows from that bound: as many softmax warps per SM as fit, everything else out of their way.
//
// One CTA per (sequence, up to 256 query rows, head); two CTAs per SM (80 KB of shared memory, 256 TMEM columns,
// 320 threads of <= 102 registers each).  K and V of the whole sequence sit in shared memory (32 KB each), loaded in
// 64-key boxes {32 d, 64 rows} with the 64-byte swizzle = the canonical K-major (Q, K) / MN-major (V) UMMA layouts.
//   warp 0 (one thread)   TMA: Q (two 128-row tiles), then K_j, V_j for j = 0 .. len/64, one mbarrier per box
//   warp 1 (one thread)   per query tile t and 64-key tile j:   S_t  = Q_t . K_j^T   (M 128, N 64, K 32: two MMAs)
//                                                               O_t += P_t . V_j     (M 128, N 32, K 64: four MMAs,
//                         A = P from tensor memory); the two query tiles alternate, so the pipe works on one while the
//                         softmax warps of the other are busy
//   warps 2-5 / 6-9       softmax of query tile 0 / 1: one row per thread (TMEM lane = row), the 64 scores of a tile in
//                         registers (one tcgen05.ld round trip), p = exp2(s * scale - m) written back as packed bf16 over
//                         the first 32 columns of S; the running maximum is raised (and the 32-column O row rescaled in
//                         TMEM) only when a warp sees a score more than 2^8 above it
// TMEM: query tile t at columns 128 t: [0,64) S / P, [64,96) O.
#include "kernels.cuh"
#include "gemm.cuh"
#include "attn_tc.cuh"
#include <cudaTypedefs.h>
#include <mutex>

namespace mq {

constexpr int kEaD = 32;             // head_dim
constexpr int kEaKv = 64;            // keys per tile
constexpr int kEaMaxSeq = 512;
constexpr int kEaBox = kEaKv * kEaD * 2;                       // 4 KB: one {32 d, 64 rows} box
constexpr int kEaQBytes = 2 * 2 * kEaBox;                      // two 128-row query tiles
constexpr int kEaKBytes = kEaMaxSeq / kEaKv * kEaBox;          // 32 KB
constexpr int kEaSmem = kEaQBytes + 2 * kEaKBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kEaThreads = 320;
constexpr float kEaRescale = 8.f;

struct EncAttnParams {
  CUtensorMap tm;        // the [rows][3H] activation: box {32, 64}, 64-byte swizzle
  const int4* items;     // {row0, n_rows (<= 256), seq_first_row, seq_len} per CTA column
  __nv_bfloat16* out;    // [T][H]
  int H;
  float scale_log2;
};

__global__ void __launch_bounds__(kEaThreads, 2) enc_attn_tc_kernel(const __grid_constant__ EncAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;
  uint8_t* Ks = smem + kEaQBytes;
  uint8_t* Vs = Ks + kEaKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + kEaKBytes);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [8] single use
  uint64_t* v_full = bars + 9;      // [8] single use
  uint64_t* s_full = bars + 17;     // [2] per query tile
  uint64_t* p_full = bars + 19;     // [2] count 4
  uint64_t* o_full = bars + 21;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int4 item = p.items[blockIdx.x];
  const int row0 = item.x, n_rows = item.y, seq0 = item.z, len = item.w;
  const int n_kv = (len + kEaKv - 1) / kEaKv;     // 1 .. 8
  const int n_qt = n_rows > 128 ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm);
    for (int i = 0; i < 23; ++i) mbar_init(&bars[i], (i == 19 || i == 20) ? 4 : 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // the packed q | k | v rows come from the QKV GEMM

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      // rows past the sequence (the next sequence's, or stale rows of the buffer) are loaded and masked / never stored
      mbar_expect_tx(q_full, (uint32_t)(n_qt * 2 * kEaBox));
      for (int b = 0; b < n_qt * 2; ++b) tma_load_2d(Qs + b * kEaBox, &p.tm, q_full, head * kEaD, row0 + b * kEaKv, kEvictFirst);
      for (int j = 0; j < n_kv; ++j) {
        mbar_expect_tx(&k_full[j], kEaBox);
        tma_load_2d(Ks + j * kEaBox, &p.tm, &k_full[j], p.H + head * kEaD, seq0 + j * kEaKv, kEvictLast);
        mbar_expect_tx(&v_full[j], kEaBox);
        tma_load_2d(Vs + j * kEaBox, &p.tm, &v_full[j], 2 * p.H + head * kEaD, seq0 + j * kEaKv, kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t IDESC_S = umma_idesc_bf16(128, kEaKv);        // S = Q K^T: both operands K-major
      constexpr uint32_t IDESC_O = umma_idesc_bf16_bmn(128, kEaD);     // O = P V: V is MN-major
      const uint32_t q_addr = smem_u32(Qs), k_addr = smem_u32(Ks), v_addr = smem_u32(Vs);
      auto mma_s = [&](int t, int j) {
        mbar_wait(&k_full[j], 0);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kEaD / 16; ++k)
          umma_bf16(tmem_base + (uint32_t)(t * 128), umma_desc_sw64(q_addr + t * 2 * kEaBox + k * 32),
                    umma_desc_sw64(k_addr + j * kEaBox + k * 32), IDESC_S, k != 0);
        umma_commit(&s_full[t]);   // (also: every earlier P V of this CTA has retired)
      };
      mbar_wait(q_full, 0);
      for (int t = 0; t < n_qt; ++t) mma_s(t, 0);
      for (int j = 0; j < n_kv; ++j) {
        for (int t = 0; t < n_qt; ++t) {
          mbar_wait(&p_full[t], j & 1);                    // P_t,j is in TMEM (and O_t rescaled if it had to be)
          mbar_wait(&v_full[j], 0);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kEaKv / 16; ++k)             // 16 keys = 16 rows of 64 B
            umma_bf16_ts(tmem_base + (uint32_t)(t * 128 + 64), tmem_base + (uint32_t)(t * 128 + k * 8),
                         umma_desc_sw64(v_addr + j * kEaBox + k * 1024), IDESC_O, (j | k) != 0);
          if (j + 1 < n_kv) mma_s(t, j + 1);               // overwrites S_t: issued after P_t,j V_j, the pipe runs in order
          else umma_commit(&o_full[t]);
        }
      }
    }
  } else {
    // ---------------- softmax + output: one query row per thread ----------------
    const int t = (warp - 2) >> 2;          // query tile of this warp group
    if (t < n_qt) {
      const int q = warp & 3;               // TMEM lane quarter
      const int row = q * 32 + lane;        // row inside the query tile
      const uint32_t t_s = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (uint32_t)(t * 128);
      const uint32_t t_o = t_s + 64;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_kv; ++j) {
        const int valid = len - j * kEaKv;  // keys of this tile that exist (>= 1; < 64 only in the last tile)
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t v[64];
        tmem_ld32(t_s, v);
        tmem_ld32(t_s + 32, v + 32);
        tmem_ld_wait();
        float mx = -INFINITY;
        if (valid >= kEaKv) {
#pragma unroll
          for (int i = 0; i < 64; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            if (i >= valid) v[i] = 0xff800000u;  // -inf: exp2 gives exactly 0
            mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        mx *= p.scale_log2;
        if (__any_sync(0xffffffffu, mx > m_run + kEaRescale)) {
          const float m_new = fmaxf(m_run, mx);
          const float alpha = ex2_ftz(m_run - m_new);      // 0 on the first tile (m_run = -inf): O is not read then
          m_run = m_new;
          l_run *= alpha;
          if (j > 0) {
#pragma unroll 1
            for (int c = 0; c < kEaD; c += 16) {
              uint32_t o[16];
              tmem_ld16(t_o + (uint32_t)c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st16(t_o + (uint32_t)c, o);
            }
          }
        }
        float lsum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p0 = ex2_ftz(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m_run));
          const float p1 = ex2_ftz(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m_run));
          lsum += p0 + p1;
          v[i] = pack_bf16(p0, p1);
        }
        tmem_st32(t_s, v);                                  // 64 probabilities = 32 packed columns over S[0,32)
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        l_run += lsum;
      }
      // ---- normalise and store this row: 32 bf16 = 64 contiguous bytes
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(t_o, o);
      tmem_ld_wait();
      const int r = t * 128 + row;
      if (r < n_rows) {
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        __nv_bfloat16* po = p.out + (size_t)(row0 + r) * p.H + head * kEaD;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(po + i) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host
bool enc_attn_tc_supported(int head_dim, int max_seq, int hidden) {
  return head_dim == kEaD && max_seq <= kEaMaxSeq && hidden % 8 == 0;
}

bool enc_attn_tc_encode(CUtensorMap* out, const void* qkv, int rows, int H) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)3 * H, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)3 * H * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kEaD, (cuuint32_t)kEaKv};
  const cuuint32_t es[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void enc_attn_tc_set_attrs() {
  cudaFuncSetAttribute(enc_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEaSmem);
}

cudaError_t launch_enc_attn_tc(const LaunchCfg& lc, const CUtensorMap& tm, const int4* items, int n_items, int n_heads, int H,
                               __nv_bfloat16* out, float scale_log2) {
  EncAttnParams p;
  p.tm = tm; p.items = items; p.out = out; p.H = H; p.scale_log2 = scale_log2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_items, n_heads, 1);
  cfg.blockDim = dim3(kEaThreads);
  cfg.dynamicSmemBytes = kEaSmem;
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, enc_attn_tc_kernel, p);
}

}  // namespace mq

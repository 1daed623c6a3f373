// Decode-chain GEMM: split-K across a thread-block CLUSTER, reduced through distributed shared memory, with the
// consumer-side work of the old reduce kernels fused into the epilogue.  (Stands where the reference has the remote
// backend's arithmetic, /root/reference/src/dispatcher.rs:287-290; round-1 VERDICT task 3.)
//
//   out[t, f] = sum_k X[t, k] * W[f, k]        T <= 64 running sequences (one token each), W streamed once from HBM
//
// Round 1 ran this as split-K over blockIdx.z writing fp32 planes to global memory, and three more kernels per layer
// (add_rmsnorm x 2, rope_kv) re-read the planes: 13.6 us of a 122 us layer spent in non-streaming reduce kernels.
// Here the CS CTAs of a cluster own the CS k-ranges of ONE weight tile:
//   1. mainloop as in gemm.cuh (TMA producer warp, single-thread tcgen05.mma issuer, accumulator in TMEM);
//   2. reduce-scatter: every epilogue thread (TMEM lane = output feature) sends the columns (tokens) of its partial
//      row to the CTA that owns that token range - st.shared::cluster into a dedicated receive buffer (never
//      aliased with the TMA ring: a faster peer may write while this CTA is still in its mainloop);
//   3. one barrier.cluster (release / acquire);
//   4. every CTA finishes ITS token range for all 128 features of the tile, partial sums added in rank order
//      (deterministic), then the fused epilogue:
//        h[t,f] += sum;  xg[t,f] = bf16(h * gamma_next[f]);  ssq_out[tile][t] = sum_f h^2  (per-tile partial of the
//        next RMSNorm) - the O and down projections.  (A DK_QKV variant with RoPE and the paged-KV write in the epilogue
//        was built, parity-tested and measured: 48 head tiles fit only 2-CTA clusters = 96 SMs, 4.3 us per layer slower
//        than the 144-CTA plane GEMM + the small rope kernel, so QKV stays on that path - profiles/r02_ablations.md.)
// RMSNorm itself never runs as a kernel: x = h * rstd * gamma feeds only GEMMs, and rstd[t] is a per-token scalar,
// so the producer emits xg = bf16(h * gamma) and the consumer scales its accumulator column t by
// rstd[t] = rsqrt(sum_tiles ssq[tile][t] / H + eps)  (RstdIn; summed in tile order: deterministic).
#pragma once
#include "gemm.cuh"
#include <type_traits>

namespace mq {

struct DkParams {
  int T;             // valid activation rows (<= BN)
  int n_out;         // valid output features
  int tile_rows;     // weight rows per tile (<= 128, multiple of 8)
  int k_blocks;      // K / 64
  int kb_per_split;  // k-blocks per cluster rank
  unsigned long long w_policy;
  float* h;          // [T][ldh] fp32 residual stream, updated in place
  int ldh;
  const __nv_bfloat16* gamma_next;  // [n_out] weight of the NEXT RMSNorm
  __nv_bfloat16* xg;                // [T][ldx] bf16(h * gamma_next): the next GEMM's activation operand
  int ldx;
  float* ssq_out;    // [m_tiles][ssq_stride]
  int ssq_stride;
  Trace tr;
  unsigned long long* dbg;  // optional: phase stamps of CTA 0 (%globaltimer ns), tools/dk_bench.py (nullptr in production)
};
__device__ __forceinline__ void dk_stamp(const DkParams& p, int k) {
  if (p.dbg && blockIdx.x == 0) p.dbg[k] = globaltimer_ns();
}

constexpr int kDkMaxCluster = 8;  // portable cluster size limit
constexpr int kDkThreads = 64 + 256;  // TMA producer warp, MMA issuer warp, 8 epilogue warps (two per TMEM lane quarter:
                                      // the epilogue runs on single-warp schedulers, so halving its per-thread work
                                      // halves its time - measured 2.4 -> ... us per launch)
constexpr int kDkMaxTok = 32;  // tokens one CTA finishes (its preload buffer): T > 32 needs CS >= 2
// shared memory: [TMA ring][receive buffer: CS slots x tok_per tokens x 128 fp32][preload: MAXTOK x 128 fp32][misc]
__host__ __device__ constexpr int dk_maxtok(int bn) { return bn < kDkMaxTok ? bn : kDkMaxTok; }
__host__ __device__ constexpr int dk_recv_bytes(int bn) { return (bn + 8) * kBlockM * 4; }
// preload / h^2 buffer: token rows are 132 floats apart and every 32-feature quarter is shifted by one more float, so that
// the per-token sum of squares (4 threads per token, 32 features each) walks its values in a FIXED order - the same for
// every token slot - without bank conflicts.  (A lane-rotated start was conflict-free too, but made the rounding of a
// token's RMSNorm depend on its slot: the trace's tokens changed with the order the users were submitted in.)
constexpr int kDkPreStride = 132;
__host__ __device__ constexpr int dk_pre_bytes(int bn) { return (dk_maxtok(bn) * kDkPreStride + 8) * 4; }
__device__ __forceinline__ uint32_t dk_pre_idx(int tok, int row) { return (uint32_t)(tok * kDkPreStride + row + (row >> 5)); }
__host__ __device__ constexpr int dk_stages(int bn) {
  int s = (222 * 1024 - dk_recv_bytes(bn) - dk_pre_bytes(bn)) / gemm_stage_bytes(bn, EPI_F32);
  return s > 8 ? 8 : s;
}
__host__ __device__ constexpr int dk_smem_bytes(int bn) {
  return dk_stages(bn) * gemm_stage_bytes(bn, EPI_F32) + dk_recv_bytes(bn) + dk_pre_bytes(bn) + 1024 /*align*/ +
         1024 /*barriers, per-token scalars*/;
}

// Code-size note (measured, B200): a first version kept the per-token preloads in registers and unrolled the token
// loops 32x - 200 KB of SASS per instance, and the finalize phase ran at ~1 us per token on instruction fetch alone.
// Everything per-token now lives in shared memory and the loops are rolled.
template <int BN>
__global__ void __launch_bounds__(kDkThreads, 1)
gemm_dk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DkParams p) {
  constexpr int STAGES = dk_stages(BN);
  constexpr int STAGE_BYTES = gemm_stage_bytes(BN, EPI_F32);
  constexpr int B_OFF = kATileBytes;
  constexpr uint32_t TMEM_COLS = gemm_tmem_cols(BN, EPI_F32);
  constexpr uint32_t IDESC = umma_idesc_bf16(kBlockM, BN);
  constexpr int MAXTOK = dk_maxtok(BN);
  static_assert(BN == 16 || BN == 32 || BN == 64, "decode tile widths");
  static_assert(STAGES >= 4, "pipeline depth");
  static_assert(BN * kBlockM * 4 <= STAGES * STAGE_BYTES, "the partial tile is staged in the (idle) ring");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stage_t = reinterpret_cast<float*>(smem);                       // [T][128] partial tile (after the mainloop)
  float* recv = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);   // [CS][tok_per][128], slot = sender rank
  float* pre = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + dk_recv_bytes(BN));  // [MAXTOK][128]
  uint8_t* tail = smem + STAGES * STAGE_BYTES + dk_recv_bytes(BN) + dk_pre_bytes(BN);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* recv_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(recv_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int CS = (int)cluster_nctarank();
  const int rank = (int)cluster_ctarank();
  const int tile_m = blockIdx.x / CS;
  const int R = p.tile_rows;
  const int m0 = tile_m * R;
  const int kb0 = rank * p.kb_per_split;
  const int nkb = max(0, min(p.kb_per_split, p.k_blocks - kb0));
  const int tok_per = (p.T + CS - 1) / CS;         // tokens a rank finishes (<= MAXTOK, checked on the host)
  const int t0 = rank * tok_per;
  const int ntok = max(0, min(tok_per, p.T - t0));

  if (warp == 0 && lane == 0) {
    trace_begin(p.tr);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(recv_bar, 1);
    fence_mbar_init();
    // partial tiles of all CS ranks (this one included: a local bulk copy) for this rank's tokens; posted before any
    // peer can send (cluster barrier #1 below)
    if (ntok > 0) mbar_expect_tx(recv_bar, (uint32_t)(CS * ntok * kBlockM * 4));
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  // cluster barrier #1 (split): "my receive barrier is armed" - peers wait for it right before they send
  cluster_arrive_release();

  const int q = warp & 3;              // TMEM lane quarter an epilogue warp may read
  const int row = q * 32 + lane;       // epilogue: feature row of the tile
  const int f = m0 + row;
  const bool frow = warp >= 2 && row < R && f < p.n_out;
  const int et = threadIdx.x - 64;     // epilogue thread index 0..255
  const int hid = et >> 7;             // which of the two warps of this lane quarter: takes every other token / column chunk

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: weights of the first ring pass before the dependency wait ----------------
      const int npre = nkb < STAGES ? nkb : STAGES;
      for (int s = 0; s < npre; ++s) {
        mbar_expect_tx(&full_bar[s], R * kBlockK * 2 + BN * kBlockK * 2);
        tma_load_2d(smem + s * STAGE_BYTES, &tmA, &full_bar[s], (kb0 + s) * kBlockK, m0, p.w_policy);
      }
      pdl_wait();
      trace_waited(p.tr);
      for (int s = 0; s < npre; ++s)
        tma_load_2d(smem + s * STAGE_BYTES + B_OFF, &tmB, &full_bar[s], (kb0 + s) * kBlockK, 0, kEvictLast);
      for (int kb = npre; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], R * kBlockK * 2 + BN * kBlockK * 2);
        tma_load_2d(st, &tmA, &full_bar[s], (kb0 + kb) * kBlockK, m0, p.w_policy);
        tma_load_2d(st + B_OFF, &tmB, &full_bar[s], (kb0 + kb) * kBlockK, 0, kEvictLast);
      }
      dk_stamp(p, 8);
    }
    __syncwarp();
    cluster_wait_acquire();
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
        for (int k = 0; k < kBlockK / 16; ++k)
          umma_bf16(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), IDESC,
                    (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
      dk_stamp(p, 9);
    }
    __syncwarp();
    cluster_wait_acquire();
  } else {
    // ---------------- epilogue warps: TMEM lane = output feature, TMEM column = token ----------------
    // ---- phase 0 (under the mainloop): everything the fused epilogue needs from EARLIER kernels, into shared memory
    if (et == 0) dk_stamp(p, 0);
    pdl_wait();
    if (et == 0) dk_stamp(p, 1);
    float gm = 0.f;  // gamma_next[f]
    if (frow) {
      gm = __bfloat162float(p.gamma_next[f]);
      const uint32_t dsts = smem_u32(pre);
      const float* src = p.h + (size_t)t0 * p.ldh + f;
      for (int i = hid; i < ntok; i += 2) cp_async4(dsts + dk_pre_idx(i, row) * 4u, src + (size_t)i * p.ldh);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // ---- phase 1: partial accumulator -> [token][feature] tile in the idle ring -> bulk copies to the owners
    if (et == 0) dk_stamp(p, 2);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (et == 0) dk_stamp(p, 3);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int c0 = hid * 16; c0 < BN; c0 += 32) {
      if (c0 >= p.T) break;
      uint32_t v[16];
      if (nkb > 0) {
        tmem_ld16(t_lane + c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;  // a rank without k-blocks contributes zeros
      }
      const uint32_t o = smem_u32(stage_t) + (uint32_t)(c0 * kBlockM + row) * 4u;
#pragma unroll
      for (int j = 0; j < 16; ++j) sts_f32(o + (uint32_t)(j * kBlockM) * 4u, __uint_as_float(v[j]));
    }
    tc_fence_before();
    fence_proxy_async();                            // generic-proxy smem writes -> visible to the bulk-copy engine
    cluster_wait_acquire();                         // barrier #1: every peer's receive barrier is armed
    asm volatile("bar.sync 1, 256;" ::: "memory");  // whole tile staged
    if (et == 0) {
      for (int r = 0; r < CS; ++r) {
        const int n_r = max(0, min(tok_per, p.T - r * tok_per));
        if (n_r == 0) continue;
        dsmem_bulk_copy(mapa_shared(smem_u32(recv + (size_t)rank * tok_per * kBlockM), (uint32_t)r),
                        smem_u32(stage_t + (size_t)r * tok_per * kBlockM), (uint32_t)(n_r * kBlockM * 4),
                        mapa_shared(smem_u32(recv_bar), (uint32_t)r));
      }
      dk_stamp(p, 4);
    }
    // ---- phase 2: the peers' partials for this rank's tokens, and this rank's own preloads
    cp_async_wait_all();
    if (et == 0) dk_stamp(p, 12);
    if (ntok > 0) mbar_wait(recv_bar, 0);
    if (et == 0) dk_stamp(p, 13);
    asm volatile("bar.sync 1, 256;" ::: "memory");  // (cp.async data of the other epilogue threads)
    if (et == 0) dk_stamp(p, 5);
    // ---- phase 3: finish this rank's tokens: partials added in rank order (deterministic)
    // (token loops are rolled and the rank loop is a plain strided walk: the first cut spent ~230 instructions per
    //  token here, 3.7 us for 16 tokens on four single-warp schedulers)
    const uint32_t slot_bytes = (uint32_t)(tok_per * kBlockM) * 4u;  // bytes between the slots of two sender ranks
    // The token loops are instantiated per cluster size (dispatch OUTSIDE the loop): with a run-time rank loop the body
    // was ~70 instructions per token on a single-warp scheduler (290 cycles per token measured); straight-line it is ~25.
    auto finalize = [&](auto cs_tag) {
      constexpr int kCS = decltype(cs_tag)::value;  // 0 = generic (run-time CS)
      auto rank_sum = [&](uint32_t a) {
        if constexpr (kCS == 0) {
          float v[kDkMaxCluster];
#pragma unroll
          for (int r = 0; r < kDkMaxCluster; ++r) v[r] = r < CS ? lds_f32(a + (uint32_t)r * slot_bytes) : 0.f;
          float s = v[0];
#pragma unroll
          for (int r = 1; r < kDkMaxCluster; ++r) s += v[r];
          return s;
        } else {
          float v[kCS];
#pragma unroll
          for (int r = 0; r < kCS; ++r) v[r] = lds_f32(a + (uint32_t)r * slot_bytes);  // all loads before the first add
          float s = v[0];
#pragma unroll
          for (int r = 1; r < kCS; ++r) s += v[r];                                     // rank order: deterministic
          return s;
        }
      };
      {
        float* hp = p.h + (size_t)(t0 + hid) * p.ldh + f;
        __nv_bfloat16* xp = p.xg + (size_t)(t0 + hid) * p.ldx + f;
        const size_t hstep = 2 * (size_t)p.ldh, xstep = 2 * (size_t)p.ldx;
        uint32_t ra = smem_u32(recv) + (uint32_t)(hid * kBlockM + row) * 4u, pa = smem_u32(pre) + dk_pre_idx(hid, row) * 4u;
        // batches of 4 tokens: every shared-memory load of a batch is issued before its first use (the loads are volatile
        // asm, i.e. kept in program order - token by token they formed one dependent chain per token)
        for (int i0 = hid; i0 < ntok; i0 += 8) {
          float sv[4], pv[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const bool ok = i0 + 2 * b < ntok;
            sv[b] = ok ? rank_sum(ra + (uint32_t)b * (2 * kBlockM * 4)) : 0.f;
            pv[b] = ok ? lds_f32(pa + (uint32_t)b * (2 * kDkPreStride * 4)) : 0.f;
          }
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (i0 + 2 * b < ntok) {
              const float hv = frow ? pv[b] + sv[b] : 0.f;
              if (frow) {
                hp[(size_t)b * hstep] = hv;
                xp[(size_t)b * xstep] = __float2bfloat16(hv * gm);
              }
              sts_f32(pa + (uint32_t)b * (2 * kDkPreStride * 4), hv * hv);
            }
          }
          ra += 4 * 2 * kBlockM * 4; pa += 4 * 2 * kDkPreStride * 4; hp += 4 * hstep; xp += 4 * xstep;
        }
      }
    };
    switch (CS) {
      case 1: finalize(std::integral_constant<int, 1>{}); break;
      case 2: finalize(std::integral_constant<int, 2>{}); break;
      case 3: finalize(std::integral_constant<int, 3>{}); break;
      case 4: finalize(std::integral_constant<int, 4>{}); break;
      default: finalize(std::integral_constant<int, 0>{}); break;
    }
    {
      if (et == 0) dk_stamp(p, 10);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // sum of squares per token: 4 threads per token, 32 features each, in feature order (see kDkPreStride)
      // (all lanes run the shuffles: ntok need not be a multiple of the 8 tokens a warp covers)
      const int ti = et >> 2, part = et & 3;  // (ti >= 32 >= ntok for the second half of the threads)
      float ss = 0.f;
      if (ti < ntok) {
        const uint32_t base = smem_u32(pre) + dk_pre_idx(ti, part * 32) * 4u;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) ss += lds_f32(base + (uint32_t)k * 4u);
      }
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      if (ti < ntok && part == 0) p.ssq_out[(size_t)tile_m * p.ssq_stride + t0 + ti] = ss;
    }
    if (et == 0) dk_stamp(p, 6);
  }

  // cluster barrier #2: no CTA leaves while a peer may still be reading its staged tile (outgoing bulk copies) -
  // every peer has passed its receive wait by the time it arrives here
  __syncwarp();
  cluster_arrive_release();
  cluster_wait_acquire();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
  if (threadIdx.x == 0) trace_end(p.tr);
}

}  // namespace mq

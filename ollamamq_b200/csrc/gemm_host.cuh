// Host-side plan for one tcgen05 GEMM launch (TMA maps + params are built once and re-used every step).
#pragma once
#include "gemm.cuh"
#include "gemm_streamk.cuh"
#include "gemm_persist.cuh"
#include "gemm_2cta.cuh"
#include "kernels.cuh"

namespace mq {

struct GemmPlan {
  CUtensorMap tmA;  // weights  [w_rows, K], box {64, 128}
  CUtensorMap tmB;  // activations [x_rows, K], box {64, bn}
  CUtensorMap tmC;  // output [T, n_out] (bf16, 2-D) or [splits, T, n_out] (fp32, 3-D), box {128, bn(,1)}, no swizzle
  GemmParams p;
  int bn;
  int epi;
  int splits;
  bool deep;  // pipeline depth variant (see gemm_stages)
  bool twocta;       // prefill regime: cta_group::2 kernel (gemm_2cta.cuh); tmB then has a 128-row box
  TwoCtaParams c2;
  bool persist;      // prefill regime, single accumulator: persistent double-buffered kernel (gemm_persist.cuh)
  PersistParams pk;
  bool streamk;      // decode regime: persistent stream-K kernel (gemm_streamk.cuh), output is one complete plane
  StreamKParams sk;
};

// Scratch for the stream-K fix-up: per-CTA partial accumulators + arrival counters (zeroed once, self re-arming).
struct StreamKWorkspace {
  float* ws = nullptr;
  int* flags = nullptr;
  int n_ctas = 0;  // persistent CTAs = SM count of the device
  bool force = false;  // use stream-K for every decode-width GEMM planned with this workspace (tests, MQ_STREAMK=1)
};
int streamk_workspace_alloc(StreamKWorkspace* w);   // cudaMalloc on the current device
void streamk_workspace_free(StreamKWorkspace* w);
bool streamk_enabled();                             // MQ_STREAMK=0 turns it off (A/B switch)

// Encode a row-major bf16 [rows, cols] tensor with a {64, box_rows} box and the 128-byte swizzle.
bool tmap_encode_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

// Token-tile width for T activation rows.
int gemm_pick_bn(int T);

// out[t, f] (+ split planes) = X[t, :] . W[f, :]
//   w_rows       rows of the weight tensor (n_out, or 2*n_out for the gate|up tensor)
//   x_rows_alloc rows the activation buffer really has (TMA bounds; rows >= T read as-is, >= alloc as zero)
bool gemm_plan(GemmPlan* g, const void* W, int w_rows, int n_out, int K, const void* X, int x_rows_alloc, int T,
               int epi, void* out, int ldo, int splits, long long split_stride, int a2_row_off,
               const StreamKWorkspace* sk = nullptr);  // with sk (and T <= 64): stream-K, `splits` is ignored (1 plane)

// Attach the fused residual-add + RMSNorm prologue (decode only; split-K kernel only): see GemmParams::norm_*.
void gemm_plan_fuse_norm(GemmPlan* g, float* h, const float* partial, int n_planes, long long plane_stride,
                         const void* gamma, void* x, int H, float eps, int* counter);
cudaError_t gemm_launch(const GemmPlan& g, const LaunchCfg& lc);
void gemm_set_attrs();

}  // namespace mq

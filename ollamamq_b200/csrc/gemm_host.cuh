// Host-side plan for one tcgen05 GEMM launch (TMA maps + params are built once and re-used every step).
#pragma once
#include "gemm.cuh"
#include "gemm_streamk.cuh"
#include "gemm_persist.cuh"
#include "gemm_2cta.cuh"
#include "gemm_dk.cuh"
#include "gemm_rowln.cuh"
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
               const StreamKWorkspace* sk = nullptr,  // with sk (and T <= 64): stream-K, `splits` is ignored (1 plane)
               int tile_rows = 0);                    // weight rows per tile (0 = 128); see GemmParams::tile_rows
// Tile height that spreads n_out weight rows over (almost) all SMs of the device: multiple of 8, 64..128.
int gemm_balanced_rows(int n_out);
// EPI_BIAS_BF16 / EPI_GELU_BF16: bf16 [n_out] bias added before the activation (whichever kernel the plan picked).
void gemm_plan_set_bias(GemmPlan* g, const void* bias);
// RMSNorm fold: scale token column t of the result by rstd[t] (GemmParams::rs / StreamKParams::rs).
void gemm_plan_set_rstd(GemmPlan* g, const RstdIn& rs);
// EPI_RESID (prefill O / down with the fold; only servable by the 2-CTA kernel: n_out % 256 == 0, T > 128): `out` of
// gemm_plan is the fp32 residual stream.  false: the plan is not a 2-CTA EPI_RESID plan.
bool gemm_plan_set_resid(GemmPlan* g, const void* gamma_next, void* xg, int ldx, float* ssq_out, int ssq_stride);

// ---- decode chain (gemm_dk.cuh): cluster split-K GEMM with the residual add / next-norm partials fused
struct DkPlan {
  CUtensorMap tmA;  // weights [w_rows, K], box {64, tile_rows}
  CUtensorMap tmB;  // activations [x_rows, K], box {64, bn}
  DkParams p;       // the caller fills the epilogue fields after dk_plan()
  int bn, cs, m_tiles;
};
int dk_max_clusters(int cs);                          // co-resident clusters of `cs` CTAs on this device (occupancy query)
int dk_pick_cluster(int m_tiles, int k_blocks, int T);  // 0 = shape not servable
// cs <= 0: pick.  false: shape not servable by the chain kernel (T > 64, too many tiles, ...)
bool dk_plan(DkPlan* g, const void* W, int w_rows, int n_out, int K, const void* X, int x_rows_alloc, int T,
             int tile_rows, int cs);
cudaError_t dk_launch(const DkPlan& g, const LaunchCfg& lc);
cudaError_t gemm_launch(const GemmPlan& g, const LaunchCfg& lc);

// ---- projection + bias + residual + LayerNorm (gemm_rowln.cuh): x = LayerNorm(x + X W^T + bias) * gamma + beta, in place
// over the bf16 residual stream x [T][n_out]; n_out in {256, 384}, K % 64 == 0, X rows allocated to a multiple of 128.
struct RowLnPlan {
  CUtensorMap tmA;  // activations [x_rows, K], box {64, 128}
  CUtensorMap tmB;  // weights [n_out, K], box {64, 64}
  CUtensorMap tmX;  // residual / result [T, n_out], box {64, 128}
  RowLnParams p;
  int n_out;
};
bool rowln_supported(int n_out, int K);
bool rowln_plan(RowLnPlan* g, const void* W, int n_out, int K, const void* X, int x_rows_alloc, int T, void* x_resid,
                const void* bias, const void* gamma, const void* beta, float eps, float* h32);
cudaError_t rowln_launch(const RowLnPlan& g, const LaunchCfg& lc);
void gemm_set_attrs();

}  // namespace mq

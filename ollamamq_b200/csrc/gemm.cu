// Host side of the tcgen05 GEMM: TMA descriptor construction + template dispatch.
#include "gemm_host.cuh"
#include <cudaTypedefs.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace mq {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // Resolved through the runtime so libollamamq_b200.so carries no link-time dependency on libcuda.so
    // (the library must dlopen on a CPU-only box for the symbol-export test).
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

bool tmap_encode_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  auto fn = get_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};  // bytes, dim 1
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// output tile map: row-major [.., T, ldo] with the feature dimension innermost; rows >= T and features >= n_out are
// outside the tensor and therefore dropped by the TMA store
static bool tmap_encode_out(CUtensorMap* out, void* base, int epi, int n_out, int T, int ldo, int splits,
                            long long split_stride, int bn, int tile_rows) {
  auto fn = get_encode_fn();
  if (!fn) return false;
  const cuuint32_t estr[3] = {1, 1, 1};
  if (epi == EPI_F32 || epi == EPI_RESID) {
    const cuuint64_t dims[3] = {(cuuint64_t)n_out, (cuuint64_t)T, (cuuint64_t)splits};
    const cuuint64_t strides[2] = {(cuuint64_t)ldo * 4, (cuuint64_t)(splits > 1 ? split_stride : (long long)T * ldo) * 4};
    const cuuint32_t box[3] = {(cuuint32_t)tile_rows, (cuuint32_t)bn, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)n_out, (cuuint64_t)T};
  const cuuint64_t strides[1] = {(cuuint64_t)ldo * 2};
  const cuuint32_t box[2] = {(cuuint32_t)tile_rows, (cuuint32_t)bn};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int EPI>
static cudaError_t launch_one(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmParams& p,
                              dim3 grid, const LaunchCfg& lc) {
  constexpr int smem = gemm_smem_bytes(BN, EPI);
  // dynamic-smem opt-in happens once per device in gemm_set_attrs() (never inside a graph capture)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(BN <= 64 ? kGemmThreadsDecode : kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_wx_kernel<BN, EPI>, a, b, c, p);
}

template <int EPI>
static cudaError_t launch_bn(int bn, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                             const GemmParams& p, dim3 grid, const LaunchCfg& lc) {
  switch (bn) {
    case 16: return launch_one<16, EPI>(a, b, c, p, grid, lc);
    case 32: return launch_one<32, EPI>(a, b, c, p, grid, lc);
    case 64: return launch_one<64, EPI>(a, b, c, p, grid, lc);
    case 128: return launch_one<128, EPI>(a, b, c, p, grid, lc);
    case 256: return launch_one<256, EPI>(a, b, c, p, grid, lc);
    default: return cudaErrorInvalidValue;
  }
}

template <int BN, int EPI>
static cudaError_t launch_sk(const GemmPlan& g, const LaunchCfg& lc) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.sk.n_ctas);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = sk_smem_bytes(BN, EPI);
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_streamk_kernel<BN, EPI>, g.tmA, g.tmB, g.sk);
}
template <int EPI>
static cudaError_t launch_sk_bn(const GemmPlan& g, const LaunchCfg& lc) {
  switch (g.bn) {
    case 16: return launch_sk<16, EPI>(g, lc);
    case 32: return launch_sk<32, EPI>(g, lc);
    case 64: return launch_sk<64, EPI>(g, lc);
    default: return cudaErrorInvalidValue;
  }
}

template <int BN, int EPI>
static cudaError_t launch_pk(const GemmPlan& g, const LaunchCfg& lc) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.pk.n_ctas);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = gemm_smem_bytes(BN, EPI);
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_persist_kernel<BN, EPI>, g.tmA, g.tmB, g.pk);
}

template <int EPI>
static cudaError_t launch_c2(const GemmPlan& g, const LaunchCfg& lc) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * g.c2.n_pairs);  // persistent pairs; cluster dims (2,1,1) are compiled into the kernel
  cfg.blockDim = dim3(c2_threads(EPI));
  cfg.dynamicSmemBytes = c2_smem_bytes(EPI);
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_2cta_kernel<EPI>, g.tmA, g.tmB, g.tmC, g.c2);
}

cudaError_t gemm_launch(const GemmPlan& g, const LaunchCfg& lc) {
  if (g.twocta) switch (g.epi) {
      case EPI_F32: return launch_c2<EPI_F32>(g, lc);
      case EPI_BF16: return launch_c2<EPI_BF16>(g, lc);
      case EPI_SILU_BF16: return launch_c2<EPI_SILU_BF16>(g, lc);
      case EPI_RESID: return launch_c2<EPI_RESID>(g, lc);
      case EPI_BIAS_BF16: return launch_c2<EPI_BIAS_BF16>(g, lc);
      case EPI_GELU_BF16: return launch_c2<EPI_GELU_BF16>(g, lc);
      default: return cudaErrorInvalidValue;
    }
  if (g.persist) {
    if (g.bn == 256) return g.epi == EPI_F32 ? launch_pk<256, EPI_F32>(g, lc) : launch_pk<256, EPI_BF16>(g, lc);
    return g.epi == EPI_F32 ? launch_pk<128, EPI_F32>(g, lc) : launch_pk<128, EPI_BF16>(g, lc);
  }
  if (g.streamk) switch (g.epi) {
      case EPI_F32: return launch_sk_bn<EPI_F32>(g, lc);
      case EPI_BF16: return launch_sk_bn<EPI_BF16>(g, lc);
      case EPI_SILU_BF16: return launch_sk_bn<EPI_SILU_BF16>(g, lc);
      default: return cudaErrorInvalidValue;
    }
  dim3 grid(g.p.m_tiles * g.p.n_tiles, 1, g.splits);
  switch (g.epi) {
    case EPI_F32: return launch_bn<EPI_F32>(g.bn, g.tmA, g.tmB, g.tmC, g.p, grid, lc);
    case EPI_BF16: return launch_bn<EPI_BF16>(g.bn, g.tmA, g.tmB, g.tmC, g.p, grid, lc);
    case EPI_SILU_BF16: return launch_bn<EPI_SILU_BF16>(g.bn, g.tmA, g.tmB, g.tmC, g.p, grid, lc);
    case EPI_GELU_BF16: return launch_bn<EPI_GELU_BF16>(g.bn, g.tmA, g.tmB, g.tmC, g.p, grid, lc);
    case EPI_BIAS_BF16: return launch_bn<EPI_BIAS_BF16>(g.bn, g.tmA, g.tmB, g.tmC, g.p, grid, lc);
    default: return cudaErrorInvalidValue;
  }
}

template <int BN, int EPI>
static void set_attrs_sk() {
  cudaFuncSetAttribute(gemm_streamk_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, sk_smem_bytes(BN, EPI));
}
template <int BN, int EPI>
static void set_attrs_one() {
  cudaFuncSetAttribute(gemm_wx_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes(BN, EPI));
}
template <int EPI>
static void set_attrs_epi() {
  set_attrs_one<16, EPI>();
  set_attrs_one<32, EPI>();
  set_attrs_one<64, EPI>();
  set_attrs_one<128, EPI>();
  set_attrs_one<256, EPI>();
}
// Opt every instantiation into its dynamic shared memory size up front (per device), so nothing but
// launches happens while a decode step is being captured into a CUDA graph.
void dk_set_attrs();
void gemm_set_attrs() {
  dk_set_attrs();
  set_attrs_epi<EPI_F32>();
  set_attrs_epi<EPI_BF16>();
  set_attrs_epi<EPI_SILU_BF16>();
  set_attrs_epi<EPI_GELU_BF16>();
  set_attrs_epi<EPI_BIAS_BF16>();
  set_attrs_sk<16, EPI_F32>(); set_attrs_sk<32, EPI_F32>(); set_attrs_sk<64, EPI_F32>();
  set_attrs_sk<16, EPI_BF16>(); set_attrs_sk<32, EPI_BF16>(); set_attrs_sk<64, EPI_BF16>();
  set_attrs_sk<16, EPI_SILU_BF16>(); set_attrs_sk<32, EPI_SILU_BF16>(); set_attrs_sk<64, EPI_SILU_BF16>();
  cudaFuncSetAttribute(gemm_2cta_kernel<EPI_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, c2_smem_bytes(EPI_F32));
  cudaFuncSetAttribute(gemm_2cta_kernel<EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, c2_smem_bytes(EPI_BF16));
  cudaFuncSetAttribute(gemm_2cta_kernel<EPI_SILU_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, c2_smem_bytes(EPI_SILU_BF16));
  cudaFuncSetAttribute(gemm_2cta_kernel<EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, c2_smem_bytes(EPI_RESID));
  cudaFuncSetAttribute(gemm_2cta_kernel<EPI_BIAS_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, c2_smem_bytes(EPI_BIAS_BF16));
  cudaFuncSetAttribute(gemm_2cta_kernel<EPI_GELU_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, c2_smem_bytes(EPI_GELU_BF16));
  cudaFuncSetAttribute(gemm_rowln_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, rl_smem_bytes<256>());
  cudaFuncSetAttribute(gemm_rowln_kernel<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, rl_smem_bytes<384>());
  cudaFuncSetAttribute(gemm_persist_kernel<256, EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes(256, EPI_BF16));
  cudaFuncSetAttribute(gemm_persist_kernel<256, EPI_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes(256, EPI_F32));
  cudaFuncSetAttribute(gemm_persist_kernel<128, EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes(128, EPI_BF16));
  cudaFuncSetAttribute(gemm_persist_kernel<128, EPI_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes(128, EPI_F32));
}

// MQ_2CTA=0 turns the cta_group::2 prefill kernel off (A/B switch)
static bool twocta_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MQ_2CTA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// MQ_PERSIST=0 turns the persistent prefill kernel off (A/B switch)
static bool persist_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MQ_PERSIST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
static int device_sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// MQ_STREAMK: "1" = every decode GEMM, "0" = none, unset = auto.  Measured on B200 (r01, Llama-3-8B, B=64):
// the in-kernel fix-up tail (contributor epilogue -> fence -> flag -> finisher reads partials) costs ~6 us per
// launch, more than the balance gain on the per-layer GEMMs (QKV 22.3 vs 16.5 us, O 21.9 vs 13.8, down 33 vs 27,
// gate/up 53 vs 49) but less than it on the LM head, where every CTA owns ~6.8 whole tiles (162 vs 185 us).
// Auto therefore uses stream-K only when a CTA owns at least two whole weight tiles.
static int streamk_mode() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("MQ_STREAMK");
    v = !e ? -1 : (e[0] == '0' ? 0 : 1);
  }
  return v;
}
bool streamk_enabled() { return streamk_mode() != 0; }
static bool streamk_pick(int m_tiles, const StreamKWorkspace* sk) {
  const int m = streamk_mode();
  return sk->force || m == 1 || (m == -1 && m_tiles >= 2 * sk->n_ctas);
}
int streamk_workspace_alloc(StreamKWorkspace* w) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  w->n_ctas = sms;
  const size_t ws_bytes = (size_t)sms * 2 * 64 * kBlockM * sizeof(float);  // dual planes, BN <= 64
  if (cudaMalloc((void**)&w->ws, ws_bytes) != cudaSuccess) return -1;
  if (cudaMalloc((void**)&w->flags, sms * sizeof(int)) != cudaSuccess) return -1;
  if (cudaMemset(w->flags, 0, sms * sizeof(int)) != cudaSuccess) return -1;
  return 0;
}
void streamk_workspace_free(StreamKWorkspace* w) {
  if (w->ws) cudaFree(w->ws);
  if (w->flags) cudaFree(w->flags);
  w->ws = nullptr;
  w->flags = nullptr;
}

int gemm_pick_bn(int T) {
  if (T <= 16) return 16;
  if (T <= 32) return 32;
  if (T <= 64) return 64;
  if (T <= 128) return 128;
  return 256;
}

int gemm_balanced_rows(int n_out) {
  const int sms = device_sm_count();
  if ((n_out + kBlockM - 1) / kBlockM >= sms) return kBlockM;  // more 128-row tiles than SMs anyway
  int r = ((n_out + sms - 1) / sms + 7) & ~7;
  return r < 64 ? 64 : (r > kBlockM ? kBlockM : r);
}

bool gemm_plan(GemmPlan* g, const void* W, int w_rows, int n_out, int K, const void* X, int x_rows_alloc, int T,
               int epi, void* out, int ldo, int splits, long long split_stride, int a2_row_off,
               const StreamKWorkspace* sk, int tile_rows) {
  if (K % kBlockK != 0) return false;
  const int kb = K / kBlockK;
  g->streamk = sk != nullptr && sk->ws != nullptr && T <= 64 && epi != EPI_GELU_BF16 && epi != EPI_BIAS_BF16 && streamk_pick((n_out + kBlockM - 1) / kBlockM, sk);
  if (g->streamk) splits = 1;  // the kernel finishes shared tiles itself: one complete plane
  if (splits < 1) return false;
  const int kbps = (kb + splits - 1) / splits;  // uneven split-K: the last plane may get fewer k-blocks
  if ((splits - 1) * kbps >= kb) return false;   // but never zero
  g->bn = gemm_pick_bn(T);
  // one token tile, no split-K and fewer weight tiles than half the SMs (Phi-3 gate/up at 256 slots: 64 tiles): halve
  // the token tile so two CTAs share each weight tile through L2 - same HBM bytes, twice the SMs streaming them
  if (T > 128 && T <= 256 && splits == 1 && (n_out + kBlockM - 1) / kBlockM * 2 <= device_sm_count()) g->bn = 128;
  g->epi = epi;
  g->splits = splits;
  // tile_rows < 128 only for the plain decode-width kernel (one token tile): see GemmParams::tile_rows
  if (tile_rows <= 0 || tile_rows > kBlockM || tile_rows % 8 != 0 || g->streamk || T > g->bn) tile_rows = kBlockM;
  if (!tmap_encode_2d(&g->tmA, W, (uint64_t)w_rows, (uint64_t)K, (uint32_t)tile_rows)) return false;
  if (!tmap_encode_2d(&g->tmB, X, (uint64_t)x_rows_alloc, (uint64_t)K, (uint32_t)g->bn)) return false;
  if (!tmap_encode_out(&g->tmC, out, epi, n_out, T, ldo, splits, split_stride, g->bn, tile_rows)) return false;
  g->p = GemmParams{};
  g->p.out = out;
  g->p.split_stride = split_stride;
  g->p.ldo = ldo;
  g->p.T = T;
  g->p.n_out = n_out;
  g->p.k_blocks = kb;
  g->p.kb_per_split = kbps;
  g->p.a2_row_off = a2_row_off;
  g->p.tile_rows = tile_rows;
  g->p.m_tiles = (n_out + tile_rows - 1) / tile_rows;
  g->p.n_tiles = (T + g->bn - 1) / g->bn;
  // super-tile height: minimise the bytes one wave of 148 CTAs must pull through L2,
  //   group_m * (weight tile bytes) + (148 / group_m) * (activation tile bytes)
  const double wt = (double)kBlockM * (epi == EPI_SILU_BF16 ? 2 : 1), xt = (double)g->bn;
  int gm = (int)(sqrt(148.0 * xt / wt) + 0.5);
  if (gm < 1) gm = 1;
  if (gm > g->p.m_tiles) gm = g->p.m_tiles;
  g->p.group_m = g->p.n_tiles == 1 ? g->p.m_tiles : gm;
  g->p.w_policy = g->p.n_tiles == 1 ? kEvictFirst : kEvictNormal;  // decode streams weights exactly once
  // a ragged last 256-feature tile is fine for the plain epilogues (stores are guarded); the dual / residual ones keep
  // whole tiles (their per-128-feature partials are indexed by tile)
  const bool whole = n_out % 256 == 0;
  g->twocta = !g->streamk && tile_rows == kBlockM && g->bn == 256 && splits == 1 && twocta_enabled() &&
              (whole || (epi != EPI_SILU_BF16 && epi != EPI_RESID && n_out % 8 == 0));
  if (g->twocta) {
    // the pair computes 256 features x c2_bn tokens: every CTA stages only its own half of the activation tile
    const int bn2 = c2_bn(epi);
    if (!tmap_encode_2d(&g->tmB, X, (uint64_t)x_rows_alloc, (uint64_t)K, (uint32_t)(bn2 / 2))) return false;
    g->c2 = TwoCtaParams{};
    g->c2.out = out; g->c2.ldo = ldo;
    g->c2.T = T; g->c2.n_out = n_out; g->c2.k_blocks = kb; g->c2.a2_row_off = a2_row_off;
    g->c2.m_tiles = (n_out + 255) / 256;
    g->c2.n_tiles = (T + bn2 - 1) / bn2;
    const int pairs = device_sm_count() / 2;
    int gm2 = (int)(sqrt((double)pairs * bn2 / (epi == EPI_SILU_BF16 ? 512.0 : 256.0)) + 0.5);
    if (gm2 > g->c2.m_tiles) gm2 = g->c2.m_tiles;
    g->c2.group_m = gm2 < 1 ? 1 : gm2;
    g->c2.n_pairs = std::min(pairs, g->c2.m_tiles * g->c2.n_tiles);
    g->c2.w_policy = g->p.w_policy;
  }
  g->persist = false;
  if (!g->twocta && !g->streamk && tile_rows == kBlockM && epi != EPI_SILU_BF16 && epi != EPI_GELU_BF16 && epi != EPI_BIAS_BF16 && g->bn >= 128 && splits == 1 && persist_enabled()) {
    const int sms = device_sm_count();
    const int tiles = g->p.m_tiles * g->p.n_tiles;
    if (tiles >= 2 * sms) {  // at least two tiles per CTA, otherwise there is nothing to overlap
      g->persist = true;
      g->pk = PersistParams{};
      g->pk.out = out; g->pk.ldo = ldo; g->pk.T = T; g->pk.n_out = n_out; g->pk.k_blocks = kb;
      g->pk.m_tiles = g->p.m_tiles; g->pk.n_tiles = g->p.n_tiles; g->pk.group_m = g->p.group_m;
      g->pk.n_ctas = sms; g->pk.w_policy = g->p.w_policy;
    }
  }
  if (g->streamk) {
    const long long U = (long long)g->p.m_tiles * kb;
    g->sk = StreamKParams{};  // (also clears the optional trace slot)
    g->sk.out = out;
    g->sk.ldo = ldo;
    g->sk.T = T;
    g->sk.n_out = n_out;
    g->sk.k_blocks = kb;
    g->sk.m_tiles = g->p.m_tiles;
    g->sk.a2_row_off = a2_row_off;
    g->sk.n_ctas = (int)(U < sk->n_ctas ? U : sk->n_ctas);  // every CTA owns >= 1 unit (the fix-up relies on it)
    g->sk.ws = sk->ws;
    g->sk.flags = sk->flags;
    g->sk.w_policy = kEvictFirst;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------ gemm_rowln.cuh
bool rowln_supported(int n_out, int K) {
  return (n_out == 256 || n_out == 384) && K > 0 && K % kBlockK == 0 && twocta_enabled();
}
bool rowln_plan(RowLnPlan* g, const void* W, int n_out, int K, const void* X, int x_rows_alloc, int T, void* x_resid,
                const void* bias, const void* gamma, const void* beta, float eps, float* h32) {
  if (!rowln_supported(n_out, K) || T < 1) return false;
  if (!tmap_encode_2d(&g->tmA, X, (uint64_t)x_rows_alloc, (uint64_t)K, 128u)) return false;
  if (!tmap_encode_2d(&g->tmB, W, (uint64_t)n_out, (uint64_t)K, 64u)) return false;
  if (!tmap_encode_2d(&g->tmX, x_resid, (uint64_t)T, (uint64_t)n_out, 128u)) return false;  // rows past T: zero-filled / clipped
  g->n_out = n_out;
  g->p = RowLnParams{};
  g->p.x = (__nv_bfloat16*)x_resid;
  g->p.bias = (const __nv_bfloat16*)bias;
  g->p.gamma = (const __nv_bfloat16*)gamma;
  g->p.beta = (const __nv_bfloat16*)beta;
  g->p.h32 = h32;
  g->p.T = T;
  g->p.k_blocks = K / kBlockK;
  g->p.n_tiles = (T + 255) / 256;
  g->p.n_pairs = std::min(device_sm_count() / 2, g->p.n_tiles);
  g->p.eps = eps;
  return true;
}
template <int N_OUT>
static cudaError_t launch_rowln(const RowLnPlan& g, const LaunchCfg& lc) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * g.p.n_pairs);  // persistent pairs; cluster dims (2,1,1) are compiled into the kernel
  cfg.blockDim = dim3(kRlThreads);
  cfg.dynamicSmemBytes = rl_smem_bytes<N_OUT>();
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_rowln_kernel<N_OUT>, g.tmA, g.tmB, g.tmX, g.p);
}
cudaError_t rowln_launch(const RowLnPlan& g, const LaunchCfg& lc) {
  switch (g.n_out) {
    case 256: return launch_rowln<256>(g, lc);
    case 384: return launch_rowln<384>(g, lc);
    default: return cudaErrorInvalidValue;
  }
}

void gemm_plan_set_bias(GemmPlan* g, const void* bias) {
  g->p.bias = bias;
  g->c2.bias = (const __nv_bfloat16*)bias;
}

// ------------------------------------------------------------------------------------------------ decode chain
void gemm_plan_set_rstd(GemmPlan* g, const RstdIn& rs) {
  g->p.rs = rs;
  g->sk.rs = rs;
  g->c2.rs = rs;
}
// EPI_RESID plans (2-CTA kernel): `out` given to gemm_plan is the fp32 residual stream h
bool gemm_plan_set_resid(GemmPlan* g, const void* gamma_next, void* xg, int ldx, float* ssq_out, int ssq_stride) {
  if (!g->twocta || g->epi != EPI_RESID) return false;
  g->c2.gamma_next = (const __nv_bfloat16*)gamma_next;
  g->c2.xg = (__nv_bfloat16*)xg;
  g->c2.ldx = ldx;
  g->c2.ssq_out = ssq_out;
  g->c2.ssq_stride = ssq_stride;
  return true;
}

template <int BN>
static int dk_query_clusters(int cs) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * 64);
  cfg.blockDim = dim3(kDkThreads);
  cfg.dynamicSmemBytes = dk_smem_bytes(BN);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_dk_kernel<BN>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
// clusters of `cs` CTAs (one per SM, ~200 KiB of shared memory each) the device runs at once; cached per process
int dk_max_clusters(int cs) {
  static int cache[kDkMaxCluster + 1] = {0};
  static std::once_flag once;
  std::call_once(once, [] {
    for (int c = 1; c <= kDkMaxCluster; ++c) cache[c] = c == 1 ? device_sm_count() : dk_query_clusters<64>(c);
  });
  return (cs >= 1 && cs <= kDkMaxCluster) ? cache[cs] : 0;
}

// Cluster size (= K splits) for a decode-chain GEMM: the most CTAs that are all resident at once, with >= 4 k-blocks
// per rank, and few enough tokens per rank for the register-resident epilogue (kDkMaxTok).
int dk_pick_cluster(int m_tiles, int k_blocks, int T) {
  int best = 0, best_ctas = 0;
  for (int cs = 1; cs <= kDkMaxCluster; ++cs) {
    if (m_tiles > dk_max_clusters(cs)) continue;
    if (cs > 1 && k_blocks / cs < 4) continue;
    if ((T + cs - 1) / cs > kDkMaxTok) continue;
    if (m_tiles * cs > best_ctas) { best = cs; best_ctas = m_tiles * cs; }
  }
  return best;  // 0: shape not servable by the chain kernel (caller falls back to the plane-based path)
}

bool dk_plan(DkPlan* g, const void* W, int w_rows, int n_out, int K, const void* X, int x_rows_alloc, int T,
             int tile_rows, int cs) {
  if (K % kBlockK != 0 || T < 1 || T > 64 || tile_rows < 8 || tile_rows > kBlockM || tile_rows % 8 != 0) return false;
  const int kb = K / kBlockK;
  const int m_tiles = (n_out + tile_rows - 1) / tile_rows;
  if (cs <= 0) cs = dk_pick_cluster(m_tiles, kb, T);
  if (cs < 1 || cs > kDkMaxCluster || (T + cs - 1) / cs > kDkMaxTok) return false;
  g->bn = gemm_pick_bn(T);
  g->cs = cs;
  g->m_tiles = m_tiles;
  if (!tmap_encode_2d(&g->tmA, W, (uint64_t)w_rows, (uint64_t)K, (uint32_t)tile_rows)) return false;
  if (!tmap_encode_2d(&g->tmB, X, (uint64_t)x_rows_alloc, (uint64_t)K, (uint32_t)g->bn)) return false;
  g->p = DkParams{};
  g->p.T = T;
  g->p.n_out = n_out;
  g->p.tile_rows = tile_rows;
  g->p.k_blocks = kb;
  g->p.kb_per_split = (kb + cs - 1) / cs;
  g->p.w_policy = kEvictFirst;  // decode streams every weight exactly once
  return true;
}

template <int BN>
static cudaError_t launch_dk(const DkPlan& g, const LaunchCfg& lc) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.m_tiles * g.cs);
  cfg.blockDim = dim3(kDkThreads);
  cfg.dynamicSmemBytes = dk_smem_bytes(BN);
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_dk_kernel<BN>, g.tmA, g.tmB, g.p);
}
cudaError_t dk_launch(const DkPlan& g, const LaunchCfg& lc) {
  switch (g.bn) {
    case 16: return launch_dk<16>(g, lc);
    case 32: return launch_dk<32>(g, lc);
    case 64: return launch_dk<64>(g, lc);
    default: return cudaErrorInvalidValue;
  }
}
template <int BN>
static void set_attrs_dk() {
  cudaFuncSetAttribute(gemm_dk_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, dk_smem_bytes(BN));
}
void dk_set_attrs() {
  set_attrs_dk<16>(); set_attrs_dk<32>(); set_attrs_dk<64>();
}

}  // namespace mq

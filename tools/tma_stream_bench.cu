// Microbenchmark behind the "pre-tiled weights" decision: how fast can 148 CTAs stream a weight
// matrix through TMA into shared memory, (A) from the row-major [N, K] layout with a {64 x 128} box (each of the 128
// rows of a tile is a separate 128-byte run, 8 KiB apart), (B) from a tile-contiguous layout (every box = one
// contiguous 16 KiB run)?  No MMA, no epilogue: just the producer ring of the decode GEMM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/tma_stream_bench tools/tma_stream_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../ollamamq_b200/csrc/ptx.cuh"
using namespace mq;

constexpr int STAGES = 12, TILE = 16384;

__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, int tiles_per_cta, int kb_per_tile, int tiled, int n_ctas) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * TILE);
  uint64_t* empty = full + STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int total = tiles_per_cta * kb_per_tile;
  if (threadIdx.x == 0) {
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
      const int tile = blockIdx.x + (it / kb_per_tile) * n_ctas, kb = it % kb_per_tile;
      mbar_expect_tx(&full[s], TILE);
      if (tiled) tma_load_2d(smem + s * TILE, &tm, &full[s], 0, (tile * kb_per_tile + kb) * 128, kEvictFirst);
      else tma_load_2d(smem + s * TILE, &tm, &full[s], kb * 64, tile * 128, kEvictFirst);
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full[s], (it / STAGES) & 1);
      mbar_arrive(&empty[s]);
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 4096, N = argc > 2 ? atoi(argv[2]) : 28672 * 4;  // ~0.94 GB
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qr);
  EncodeFn enc = (EncodeFn)fnp;
  uint16_t* w;
  cudaMalloc(&w, (size_t)N * K * 2);
  cudaMemset(w, 1, (size_t)N * K * 2);
  const int kb = K / 64, m_tiles = N / 128, n_ctas = 148, tiles_per_cta = m_tiles / n_ctas;
  const cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  CUtensorMap tmA, tmB;
  { const cuuint64_t d[2] = {(cuuint64_t)K, (cuuint64_t)N}, st[1] = {(cuuint64_t)K * 2};
    enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, d, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  { const cuuint64_t d[2] = {64, (cuuint64_t)N * kb}, st[1] = {128};
    enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, d, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  const int smem = STAGES * TILE + 1024 + 256;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)tiles_per_cta * n_ctas * kb * TILE;
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      stream_kernel<<<n_ctas, 64, smem>>>(mode ? tmB : tmA, tiles_per_cta, kb, mode, n_ctas);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%s K=%d N=%d: %.1f MB in %.3f ms = %.1f GB/s (%s)\n", mode ? "tile-contiguous" : "row-major      ", K, N,
             bytes / 1e6, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}

// Hardware probe (not product code): the conv kernel's producer pattern in isolation -- per stage one im2col load
// (A: 128 px x 64 ch) plus one tiled load (B: nB rows x 64) into a ring, consumer frees stages immediately
// (mode 0) or after a busy-wait of `delay` cycles (mode 1, mimics the MMA time).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../planer_b200/csrc/ptx.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(64, 1)
mix_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int nB, int stages,
           int iters, int num_tiles, int OW, int OH, int kstages, int delay, int use_a, int use_b) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = 16384, b_bytes = nB * 128, stage_bytes = a_bytes + b_bytes;
  const uint32_t bars = base + stages * stage_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(bars + 8 * i, 1); ptx::mbar_init(bars + 8 * (stages + i), 1); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0, ph = 0;
    int tile = blockIdx.x, j = 0;
    for (int it = 0; it < iters; ++it) {
      const int m0 = tile * 128, q0 = m0 % OW, t1 = m0 / OW, p0 = t1 % OH, img = t1 / OH;
      const int tap = j % 9;
      while (!ptx::mbar_try_wait(bars + 8 * (stages + s), ph ^ 1)) {}
      ptx::mbar_arrive_expect_tx(bars + 8 * s, (use_a ? a_bytes : 0) + (use_b ? b_bytes : 0));
      if (use_a) ptx::tma_load_im2col_4d(base + s * stage_bytes, &mapA, bars + 8 * s, 0, q0 - 1, p0 - 1, img, (uint16_t)(tap % 3), (uint16_t)(tap / 3));
      if (use_b) ptx::tma_load_2d(base + s * stage_bytes + a_bytes, &mapB, bars + 8 * s, j * 64, 0);
      if (++j == kstages) { j = 0; tile += gridDim.x; if (tile >= num_tiles) tile = blockIdx.x; }
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      while (!ptx::mbar_try_wait(bars + 8 * s, ph)) {}
      if (delay) { long long t = clock64(); while (clock64() - t < delay) {} }
      ptx::mbar_arrive(bars + 8 * (stages + s));
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void *fnT = nullptr, *fnI = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnT, cudaEnableDefault, &q));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fnI, cudaEnableDefault, &q));
  void *bufA, *bufB; CK(cudaMalloc(&bufA, 64 << 20)); CK(cudaMemset(bufA, 1, 64 << 20));
  CK(cudaMalloc(&bufB, 8 << 20)); CK(cudaMemset(bufB, 1, 8 << 20));
  CK(cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225000));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int N = 128;
  CUtensorMap mapA;
  { cuuint64_t dims[4] = {64, 56, 56, (cuuint64_t)N}; cuuint64_t strides[3] = {128, 128 * 56, 128 * 56 * 56};
    int lower[2] = {-1, -1}, upper[2] = {-1, -1}; cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult cr = ((EncodeIm2colFn)fnI)(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, bufA, dims, strides, lower, upper, 64, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr) { printf("encA %d\n", cr); return 1; } }
  printf("nB stages ctas delay useA useB | GB/s_total B/clk/SM cycles_per_stage\n");
  for (int nB : {64, 256}) {
    const int K = 2304;   // weight matrix [nB rows][K] fp16, box 64 x nB
    CUtensorMap mapB;
    { cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)nB}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
      cuuint32_t bx[2] = {64, (cuuint32_t)nB}, es[2] = {1, 1};
      CUresult cr = ((EncodeTiledFn)fnT)(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, bufB, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr) { printf("encB %d\n", cr); return 1; } }
    const int stage_bytes = 16384 + nB * 128;
    const int stages = nB == 64 ? 8 : 4;
    for (int nc : {148, 74})
      for (int delay : {0, 512})
        for (int mode = 0; mode < 3; ++mode) {
          const int use_a = mode != 2, use_b = mode != 1;
          const int iters = 36 * 100;
          size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
          mix_kernel<<<nc, 64, smem>>>(mapA, mapB, nB, stages, 72, N * 56 * 56 / 128, 56, 56, 36, delay, use_a, use_b);
          CK(cudaDeviceSynchronize());
          cudaEventRecord(e0);
          mix_kernel<<<nc, 64, smem>>>(mapA, mapB, nB, stages, iters, N * 56 * 56 / 128, 56, 56, 36, delay, use_a, use_b);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          double bytes = (double)nc * iters * ((use_a ? 16384 : 0) + (use_b ? nB * 128 : 0));
          double gbs = bytes / (ms * 1e-3) / 1e9;
          printf("%3d %3d %4d %4d %d %d | %9.1f %7.1f %8.0f\n", nB, stages, nc, delay, use_a, use_b, gbs, gbs / nc / 1.9, ms * 1e-3 * 1.9e9 / iters);
        }
  }
  return 0;
}

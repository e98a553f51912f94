// Hardware probe (not product code): how many bytes per clock can ONE SM pull through its TMA unit?
// Each CTA streams [64 x rows] fp16 boxes (128B-swizzled) from a global buffer into a ring of shared-memory
// stages; a consumer thread frees every stage at once.  Sweeps box height, ring depth, CTA count, working-set
// size (L2-resident vs DRAM) and the number of CTAs that read the SAME addresses (sharing in L2).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tma_stream_probe tools/tma_stream_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../planer_b200/csrc/ptx.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap map, int box_rows, int stages, int iters, int total_rows, int share,
              unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t stage_bytes = box_rows * 128;
  const uint32_t bars = base + stages * stage_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(bars + 8 * i, 1); ptx::mbar_init(bars + 8 * (stages + i), 1); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const int group = blockIdx.x / share;           // CTAs of one group read identical addresses
  const int ngroups = (gridDim.x + share - 1) / share;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    uint32_t s = 0, ph = 0;
    int row = (int)(((long long)group * (total_rows / ngroups)) / box_rows) * box_rows;
    for (int it = 0; it < iters; ++it) {
      while (!ptx::mbar_try_wait(bars + 8 * (stages + s), ph ^ 1)) {}
      ptx::mbar_arrive_expect_tx(bars + 8 * s, stage_bytes);
      ptx::tma_load_2d(base + s * stage_bytes, &map, bars + 8 * s, 0, row);
      row += box_rows;
      if (row + box_rows > total_rows) row = 0;
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      while (!ptx::mbar_try_wait(bars + 8 * s, ph)) {}
      ptx::mbar_arrive(bars + 8 * (stages + s));
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

// im2col variant: [128 pixels x 64 ch] boxes of a 3x3/p1 convolution over an NHWC tensor (N,56,56,64)
__global__ void __launch_bounds__(64, 1)
im2col_kernel(const __grid_constant__ CUtensorMap map, int stages, int iters, int num_tiles, int OW, int OH) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t stage_bytes = 16384;
  const uint32_t bars = base + stages * stage_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(bars + 8 * i, 1); ptx::mbar_init(bars + 8 * (stages + i), 1); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0, ph = 0;
    int tile = blockIdx.x, tap = 0;
    for (int it = 0; it < iters; ++it) {
      const int m0 = tile * 128, q0 = m0 % OW, t1 = m0 / OW, p0 = t1 % OH, img = t1 / OH;
      while (!ptx::mbar_try_wait(bars + 8 * (stages + s), ph ^ 1)) {}
      ptx::mbar_arrive_expect_tx(bars + 8 * s, stage_bytes);
      ptx::tma_load_im2col_4d(base + s * stage_bytes, &map, bars + 8 * s, 0, q0 - 1, p0 - 1, img, (uint16_t)(tap % 3),
                              (uint16_t)(tap / 3));
      if (++tap == 9) { tap = 0; tile += gridDim.x; if (tile >= num_tiles) tile = blockIdx.x; }
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < iters; ++it) {
      while (!ptx::mbar_try_wait(bars + 8 * s, ph)) {}
      ptx::mbar_arrive(bars + 8 * (stages + s));
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  const size_t max_bytes = 2ull << 30;
  void* buf; CK(cudaMalloc(&buf, max_bytes)); CK(cudaMemset(buf, 1, max_bytes));
  unsigned long long* dcyc; CK(cudaMalloc(&dcyc, 148 * 8));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220000));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  {
    void* fn2 = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn2, cudaEnableDefault, &q));
    CK(cudaFuncSetAttribute(im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220000));
    printf("im2col 3x3/p1 over (N,56,56,64) fp16:  N stages ctas  GB/s_total  B_per_clk_per_SM ms\n");
    for (int N : {16, 128}) {
      CUtensorMap map;
      cuuint64_t dims[4] = {64, 56, 56, (cuuint64_t)N};
      cuuint64_t strides[3] = {128, 128 * 56, 128 * 56 * 56};
      int lower[2] = {-1, -1}, upper[2] = {-1, -1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult cr = ((EncodeIm2colFn)fn2)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, buf, dims, strides, lower, upper, 64, 128,
                                         es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) { printf("im2col encode failed %d\n", (int)cr); return 1; }
      const int num_tiles = N * 56 * 56 / 128;
      for (int st : {2, 4, 8, 12})
        for (int nc : {148, 74}) {
          const int iters = 9 * 200;
          size_t smem = (size_t)st * 16384 + 1024 + 256;
          im2col_kernel<<<nc, 64, smem>>>(map, st, 90, num_tiles, 56, 56);
          CK(cudaDeviceSynchronize());
          cudaEventRecord(e0);
          im2col_kernel<<<nc, 64, smem>>>(map, st, iters, num_tiles, 56, 56);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          double gbs = (double)nc * iters * 16384 / (ms * 1e-3) / 1e9;
          printf("im2col %4d %3d %4d  %10.1f  %8.1f  %.3f\n", N, st, nc, gbs, gbs / nc / 1.9, ms);
        }
    }
  }
  const size_t sizes[] = {48ull << 20};
  const int boxes[] = {128};
  const int stage_opts[] = {2, 4, 8};
  const int ctas[] = {148, 74};
  const int shares[] = {1, 4, 148};
  printf("ws_MB box_rows stages ctas share  GB/s_total  B_per_clk_per_SM(@1.9GHz) ms\n");
  for (size_t ws : sizes) {
    const int total_rows = (int)(ws / 128);
    for (int box : boxes) {
      CUtensorMap map;
      cuuint64_t dims[2] = {64, (cuuint64_t)total_rows}; cuuint64_t strides[1] = {128};
      cuuint32_t bx[2] = {64, (cuuint32_t)box}, es[2] = {1, 1};
      CUresult cr = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, bx, es,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }
      for (int st : stage_opts) {
        if ((size_t)st * box * 128 > 200000) continue;
        for (int nc : ctas)
          for (int sh : shares) {
            if (sh > 1 && (box != 128 || st != 4)) continue;
            const int iters = (int)((256ull << 20) / ((size_t)box * 128));   // 256 MB per CTA
            size_t smem = (size_t)st * box * 128 + 1024 + 256;
            stream_kernel<<<nc, 64, smem>>>(map, box, st, 64, total_rows, sh, dcyc);   // warm-up
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            stream_kernel<<<nc, 64, smem>>>(map, box, st, iters, total_rows, sh, dcyc);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double bytes = (double)nc * iters * box * 128;
            double gbs = bytes / (ms * 1e-3) / 1e9;
            printf("%5zu %7d %6d %5d %5d  %10.1f  %8.1f  %.3f\n", ws >> 20, box, st, nc, sh, gbs, gbs / nc / 1.9, ms);
          }
      }
    }
  }
  return 0;
}

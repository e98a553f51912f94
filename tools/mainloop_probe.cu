// Hardware probe (not product code): the conv main loop in isolation -- TMA (im2col A + tiled B) producer, tcgen05.mma
// consumer, no epilogue -- for cta_group 1 and 2, to separate TMA, MMA and shared-memory contention effects.
//   mode bits: 1 = issue TMA loads, 2 = issue MMAs.   (mode 2 alone re-uses stale smem: pure MMA rate with commits)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../planer_b200/csrc/ptx.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int CG>
__global__ void __launch_bounds__(128, 1)
loop_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int N, int stages,
            int iters, int num_tiles, int mode, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const uint32_t b_bytes = (N / CG) * 128, stage_bytes = 16384 + b_bytes;
  const uint32_t sA = base, sB = base + stages * 16384;
  const uint32_t bars = base + stages * stage_bytes;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bp + stages * stage_bytes + 512);
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(bars + 8 * i, 1); ptx::mbar_init(bars + 8 * (stages + i), 1); }
    ptx::mbar_init(bars + 8 * 2 * stages, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) { if (CG == 2) { ptx::tmem_alloc_pair(ptx::smem_u32(slot), 512); ptx::tmem_relinquish_pair(); } else { ptx::tmem_alloc(ptx::smem_u32(slot), 512); ptx::tmem_relinquish(); } }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  const int unit = CG == 2 ? blockIdx.x >> 1 : blockIdx.x, nunits = CG == 2 ? gridDim.x >> 1 : gridDim.x;
  long long t0 = clock64();
  if (warp == 0) {
    uint32_t s = 0, ph = 0; int tile = unit, j = 0;
    for (int it = 0; it < iters; ++it) {
      const int m0 = (tile * CG + rank) * 128, q0 = m0 % 56, t1 = m0 / 56, p0 = t1 % 56, img = t1 / 56, tap = j % 9;
      while (!ptx::mbar_try_wait(bars + 8 * (stages + s), ph ^ 1)) {}
      if (ptx::elect_one()) {
        const uint32_t full = bars + 8 * s;
        if (mode & 1) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(full, CG * stage_bytes);
          if (CG == 2) { ptx::tma_load_im2col_4d_pair(sA + s * 16384, &mapA, full, 0, q0 - 1, p0 - 1, img, tap % 3, tap / 3);
                         ptx::tma_load_2d_pair(sB + s * b_bytes, &mapB, full, j * 64, rank * (N / 2)); }
          else { ptx::tma_load_im2col_4d(sA + s * 16384, &mapA, full, 0, q0 - 1, p0 - 1, img, tap % 3, tap / 3);
                 ptx::tma_load_2d(sB + s * b_bytes, &mapB, full, j * 64, 0); }
        } else if (rank == 0) ptx::mbar_arrive(full);
      }
      __syncwarp();
      if (++j == 36) { j = 0; tile += nunits; if ((tile * CG + CG) * 128 > num_tiles * 128) tile = unit; }
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && rank == 0) {
    uint32_t s = 0, ph = 0;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    const uint64_t a0 = ptx::make_smem_desc(sA, 1024, 2), b0 = ptx::make_smem_desc(sB, 1024, 2);
    for (int it = 0; it < iters; ++it) {
      while (!ptx::mbar_try_wait(bars + 8 * s, ph)) {}
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        if (mode & 2) {
          uint64_t da = a0 + s * (16384 >> 4), db = b0 + s * (b_bytes >> 4);
          for (int k = 0; k < 4; ++k) { if (CG == 2) ptx::umma_f16_pair(tmem, da, db, idesc, 1); else ptx::umma_f16(tmem, da, db, idesc, 1); da += 2; db += 2; }
          if (CG == 2) ptx::umma_commit_pair(bars + 8 * (stages + s), 3); else ptx::umma_commit(bars + 8 * (stages + s));
        } else {
          ptx::mbar_arrive(bars + 8 * (stages + s));
          if (CG == 2) ptx::mbar_arrive_remote(bars + 8 * (stages + s), 1);
        }
      }
      __syncwarp();
      if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
    }
    if (ptx::elect_one()) { if (CG == 2) ptx::umma_commit_pair(bars + 8 * 2 * stages, 1); else ptx::umma_commit(bars + 8 * 2 * stages); }
    __syncwarp();
    while (!ptx::mbar_try_wait(bars + 8 * 2 * stages, 0)) {}
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 2) { if (CG == 2) ptx::tmem_dealloc_pair(tmem, 512); else ptx::tmem_dealloc(tmem, 512); }
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CG> void run(void* fnI, void* fnT, void* bufA, void* bufB, unsigned long long* d, int N, int mode) {
  const int NB = 128;
  CUtensorMap mapA, mapB;
  { cuuint64_t dims[4] = {64, 56, 56, (cuuint64_t)NB}; cuuint64_t strides[3] = {128, 128 * 56, 128 * 56 * 56};
    int lower[2] = {-1, -1}, upper[2] = {-1, -1}; cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult cr = ((EncodeIm2colFn)fnI)(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, bufA, dims, strides, lower, upper, 64, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr) { printf("encA %d\n", cr); exit(1); } }
  { const int K = 2304; cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t bx[2] = {64, (cuuint32_t)(N / CG)}, es[2] = {1, 1};
    CUresult cr = ((EncodeTiledFn)fnT)(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, bufB, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr) { printf("encB %d\n", cr); exit(1); } }
  const int stage_bytes = 16384 + (N / CG) * 128;
  int stages = 200000 / stage_bytes; if (stages > 8) stages = 8;
  const int iters = 36 * 40;
  size_t smem = (size_t)stages * stage_bytes + 2048;
  CK(cudaFuncSetAttribute(loop_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225000));
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaLaunchKernelEx(&cfg, loop_kernel<CG>, mapA, mapB, N, stages, iters, NB * 56 * 56 / 128, mode, d));
    CK(cudaDeviceSynchronize());
  }
  unsigned long long h[148]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  double mx = 0; for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
  printf("CG=%d N=%3d stages=%d mode=%d (tma=%d mma=%d): %7.0f cycles/stage   (MMA ideal %d)\n", CG, N, stages, mode, mode & 1, (mode >> 1) & 1, mx / iters, 4 * N / 2);
}

int main() {
  void *fnT = nullptr, *fnI = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnT, cudaEnableDefault, &q));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fnI, cudaEnableDefault, &q));
  void *bufA, *bufB; CK(cudaMalloc(&bufA, 64 << 20)); CK(cudaMemset(bufA, 0, 64 << 20));
  CK(cudaMalloc(&bufB, 8 << 20)); CK(cudaMemset(bufB, 0, 8 << 20));
  unsigned long long* d; CK(cudaMalloc(&d, 148 * 8));
  for (int N : {64, 128, 256})
    for (int mode : {1, 2, 3}) {
      run<1>(fnI, fnT, bufA, bufB, d, N, mode);
      run<2>(fnI, fnT, bufA, bufB, d, N, mode);
    }
  return 0;
}

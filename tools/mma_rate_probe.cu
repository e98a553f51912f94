// Hardware probe (not product code): tcgen05.mma issue/throughput for M=128, K=16, N in {64,128,256}, operands resident
// in shared memory (128B-swizzled K-major), with and without a tcgen05.commit after every 4 MMAs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../planer_b200/csrc/ptx.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 1)
mma_kernel(int N, int iters, int commit_every, int nstages, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const uint32_t stage_bytes = 16384 + N * 128;
  const uint32_t bars = base + nstages * stage_bytes;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bp + nstages * stage_bytes + 256);
  for (uint32_t i = threadIdx.x; i < nstages * stage_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(bp)[i] = 0x3c003c00u;
  ptx::fence_proxy_async_smem();
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(bars + 8 * i, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(slot), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = clock64();
    uint32_t ph[8] = {0,0,0,0,0,0,0,0};
    for (int it = 0; it < iters; ++it) {
      const int s = it % nstages;
      const uint32_t a = base + s * stage_bytes, b = a + 16384;
      for (int k = 0; k < 4; ++k)
        ptx::umma_f16(tmem + (it & 1) * N, ptx::make_smem_desc(a + k * 32, 1024, 2), ptx::make_smem_desc(b + k * 32, 1024, 2), idesc, (it | k) ? 1u : 0u);
      if (commit_every) ptx::umma_commit(bars + 8 * (it & 3));
    }
    long long t1 = clock64();
    ptx::umma_commit(bars + 8 * 7);
    while (!ptx::mbar_try_wait(bars + 8 * 7, 0)) {}
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0;
    (void)ph;
  }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d; CK(cudaMalloc(&d, 148 * 16));
  CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225000));
  printf("N ctas commit stages | issue_cycles_per_mma total_cycles_per_mma (expected N/2)\n");
  for (int N : {64, 128, 256})
    for (int nc : {1, 148})
      for (int ce : {0, 1})
        for (int ns : {1, 4}) {
          const int iters = 2000;
          size_t smem = (size_t)ns * (16384 + N * 128) + 2048;
          mma_kernel<<<nc, 128, smem>>>(N, iters, ce, ns, d);
          CK(cudaDeviceSynchronize());
          unsigned long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
          printf("%3d %3d %d %d | %8.1f %8.1f (%d)\n", N, nc, ce, ns, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), N / 2);
        }
  return 0;
}

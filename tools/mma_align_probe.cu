// Hardware probe (not product code): what sets the ~100-cycle floor of a tcgen05.mma with M=128, N<=128?
//   * nacc  : MMAs are issued round-robin over `nacc` independent TMEM accumulators (1 = one dependent chain)
//   * layout: shared-memory operand layout, K-major: 0 = 128B swizzle (64-wide k panels), 1 = 64B swizzle, 2 = 32B swizzle
//             (16-wide k panels: the K=16 slice of an operand is one dense 4 KB block)
//   * ts    : A operand read from TMEM instead of shared memory
//   * same_a: every MMA of a group of 4 uses the SAME A descriptor (does the hardware keep A?)
//   * aoff  : the A descriptor starts `aoff` rows (128 B each) into the buffer: the shift-GEMM conv reads its taps from
//             row-shifted views that are NOT aligned to the 1024-byte swizzle atom -- does that cost shared-memory bandwidth?
// Timing only; operand values are a constant.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../planer_b200/csrc/ptx.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
mma_kernel(int N, int iters, int nacc, int layout, int ts, int same_a, int aoff, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const uint32_t a_bytes = 16384 + 16384, b_bytes = 256 * 128;
  const uint32_t bars = base + a_bytes + b_bytes;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bp + a_bytes + b_bytes + 256);
  for (uint32_t i = threadIdx.x; i < (a_bytes + b_bytes) / 4; i += 128) reinterpret_cast<uint32_t*>(bp)[i] = 0x3c003c00u;
  ptx::fence_proxy_async_smem();
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(bars + 8 * i, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(slot), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    // K-major panels: 128B swizzle: row pitch 128 B, 8-row group 1024 B, k-step +32 B inside the row
    //                 64B: row pitch 64 B, group 512 B, k-step: +32 B (2 steps per panel), then next panel
    //                 32B: row pitch 32 B, group 256 B, k-step = next panel
    uint32_t sbo, ltype;
    const uint32_t a0 = base + (uint32_t)aoff * 128u, b0 = base + a_bytes;
    if (layout == 0) { sbo = 1024; ltype = 2; }
    else if (layout == 1) { sbo = 512; ltype = 4; }
    else { sbo = 256; ltype = 6; }
    uint64_t ad[4], bd[4];
    for (int k = 0; k < 4; ++k) {
      uint32_t ao, bo;
      if (layout == 0) { ao = bo = k * 32; }
      else if (layout == 1) { ao = (k & 1) * 32 + (k >> 1) * (128 * 64); bo = (k & 1) * 32 + (k >> 1) * (256 * 64); }
      else { ao = k * (128 * 32); bo = k * (256 * 32); }
      ad[k] = ptx::make_smem_desc(a0 + (same_a ? 0 : ao), sbo, ltype);
      bd[k] = ptx::make_smem_desc(b0 + bo, sbo, ltype);
    }
    const uint32_t a_tmem = tmem + 448;      // 8 columns per K=16 slice; garbage contents are fine for timing
    uint32_t dd[4], at[4];
    for (int k = 0; k < 4; ++k) { dd[k] = tmem + (uint32_t)((k % nacc) * N); at[k] = a_tmem + (same_a ? 0 : k * 8); }
    long long t0 = clock64();
    if (ts) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(dd[k], at[k], bd[k], idesc, 1u);
      }
    } else {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(dd[k], ad[k], bd[k], idesc, 1u);
      }
    }
    long long t1 = clock64();
    ptx::umma_commit(bars + 8 * 7);
    while (!ptx::mbar_try_wait(bars + 8 * 7, 0)) {}
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d; CK(cudaMalloc(&d, 148 * 16));
  CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225000));
  const int iters = 2000;
  printf("N aoff | cycles_per_mma (tensor floor N/2; aligned model 32 + N/4)\n");
  for (int N : {64, 128, 192, 256})
    for (int aoff : {0, 1, 2, 4, 8, 57, 58, 64}) {
      size_t smem = 32768 + 256 * 128 + 2048;
      mma_kernel<<<148, 128, smem>>>(N, iters, 1, 0, 0, 0, aoff, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%3d %d | CUDA error %s\n", N, aoff, cudaGetErrorString(e)); return 1; }
      unsigned long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
      printf("%3d %2d | %8.1f   (%d, %d)\n", N, aoff, (double)h[1] / (iters * 4), N / 2, 32 + N / 4);
      fflush(stdout);
    }
  return 0;
}

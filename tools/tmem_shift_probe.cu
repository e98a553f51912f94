// Hardware probe (not product code): what does tcgen05.shift.down do to an fp32 accumulator block in tensor memory?
// Writes value(lane, col) = lane * 1000 + col into 64 columns x 128 lanes with tcgen05.st, issues `nshift` tcgen05.shift.down
// at column offset `coff` (one per 8-column = 32-byte element when `per8` is set), commits to an mbarrier, reads everything back.
// Prints, per column group, which lane each row's data came from -- direction, width and whether rows cross the 32-lane
// quarters -- and the cycles from first shift to barrier completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_shift_probe tools/tmem_shift_probe.cu && tools/tmem_shift_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../planer_b200/csrc/ptx.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int coff, int nshift, int stride, uint32_t* out, long long* cyc) {
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 64); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
  const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    for (int e = 0; e < 32; ++e) v[e] = (uint32_t)((warp * 32 + lane) * 1000 + c0 + e);
    tmem_st_32x32b_x32(t_row + c0, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < nshift; ++i)
      asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tmem + (uint32_t)(coff + i * stride)) : "memory");
    ptx::umma_commit(ptx::smem_u32(&bar));
    while (!ptx::mbar_try_wait(ptx::smem_u32(&bar), 0)) {}
    cyc[0] = clock64() - t0;
  }
  __syncthreads();
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(t_row + c0, v);
    ptx::tmem_ld_wait();
    for (int e = 0; e < 32; ++e) out[(warp * 32 + lane) * 64 + c0 + e] = v[e];
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 64); }
}

int main() {
  uint32_t* d; long long* dc;
  CK(cudaMalloc(&d, 128 * 64 * 4)); CK(cudaMalloc(&dc, 8));
  static uint32_t h[128 * 64];
  const int cfgs[][3] = {{0, 1, 0}, {8, 1, 0}, {32, 1, 0}, {0, 2, 0}, {0, 8, 8}, {0, 4, 8}, {16, 3, 0}};
  for (auto& c : cfgs) {
    CK(cudaMemset(d, 0xff, 128 * 64 * 4));
    probe<<<1, 128>>>(c[0], c[1], c[2], d, dc);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    printf("== column offset %d, %d shift(s), column stride %d between them: %lld clk issue -> barrier\n", c[0], c[1], c[2], cyc);
    // per 8-column group: is it changed, and by what lane displacement (row r now holds data of lane r + d)
    for (int g = 0; g < 8; ++g) {
      int changed = 0, disp = 9999, consistent = 1, cols_ok = 1;
      for (int r = 0; r < 128; ++r) {
        for (int e = 0; e < 8; ++e) {
          const uint32_t v = h[r * 64 + g * 8 + e];
          const int src_lane = (int)(v / 1000), src_col = (int)(v % 1000);
          if (src_col != g * 8 + e) cols_ok = 0;
          if (src_lane != r) {
            ++changed;
            if (disp == 9999) disp = src_lane - r; else if (disp != src_lane - r) consistent = 0;
          }
        }
      }
      if (!changed) { printf("   cols %2d-%2d: unchanged\n", g * 8, g * 8 + 7); continue; }
      printf("   cols %2d-%2d: %d of 1024 values moved, row r holds lane r%+d (%s, columns %s); rows 0,1,31,32,33,126,127 hold lanes", g * 8,
             g * 8 + 7, changed, disp, consistent ? "uniform" : "NOT uniform", cols_ok ? "kept" : "MIXED");
      const int rows[] = {0, 1, 31, 32, 33, 126, 127};
      for (int r : rows) printf(" %d", (int)(h[r * 64 + g * 8] / 1000));
      printf("\n");
    }
  }
  return 0;
}

// Pointwise (1x1 / stride 1 / no padding) fp16 convolution as a plain GEMM  y[M, Cout] = x[M, Cin] . W^T  over the
// M = N*H*W pixel rows: the filter (or a slice of 128 / 64 output channels of it, pinned to the CTA) stays RESIDENT in shared
// memory, the activations stream through a deep TMA ring.
//
// Replaces, for the 1x1 bottleneck layers of Darknet-style networks (YOLOv3 at batch 32: 64 -> 32 @208x208, 128 -> 64 @104x104,
// 256 -> 128 @52x52, and on few rows 512 -> 256 @26x26, 1024 -> 512 @13x13), the same reference code as conv_shift.cu: Conv2d
// (planer/layer.py:22-26 + planer/util.py:17-44) -> BatchNorm (planer/layer.py:125-127) -> LeakyReLU / ReLU
// (planer/layer.py:44-51).
//
// Why a separate kernel.  These layers have FOUR to SIXTEEN tcgen05.mma per 128-pixel tile and are bound by HBM (41 us of
// traffic for 64 -> 32 @208 x32), but ran at 40 % of that roofline through conv_shift.cu: role counters showed every role
// spending ~2300 clk per tile whatever the tile's work -- 700 clk of 64-bit divisions in the producer (fixed there too),
// and ~570 clk per call of the four-MMA block in a kernel of 19 k SASS instructions whose per-tile paths run cold in the
// instruction cache (profiles/r02_kernel_experiments.md 12).  This kernel is ~1 k instructions: a tile is row block t of
// the [M, Cin] matrix (no geometry, no divisions), the weights are loaded once per CTA, the A ring is eight 16 KB stages
// deep (128 KB in flight per SM hides the DRAM latency at full bandwidth), and the epilogue is the packed-fp16 one of the
// other conv kernels (fp32 accumulator rounded to fp16 once, transposed through shared memory, scale / shift as one HFMA2,
// activation, 16-byte stores).
//
// CTA = 12 warps, persistent over tiles (tile t = rows [128 t, 128 t + 128)):
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM owner   warp 3: stages scale / shift
//   warps 4-11: epilogue, two column groups x four TMEM lane quarters (the groups take tiles in turn when Cout <= 32)
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 384;
constexpr int kMaxStages = 8;
constexpr int kMaxChunks = 16;                       // Cin <= 1024
constexpr uint32_t kAStage = kTileM * 128;           // 128 rows x 64 channels x fp16
constexpr uint32_t kStageBytes = 8 * 2048;           // epilogue transposition stage: 32 rows x 64 B per warp
constexpr long long kWatchdogCycles = 4000000000ll;

struct PwParams {
  int M, Cout, cchunks, n_tile, num_tiles, stages;
  int num_n;         // output-channel blocks of n_tile channels; block b % num_n is PINNED to CTA b (its filter slice stays resident)
  uint32_t b_chunk_bytes, idesc, tmem_cols;
  __half* y; int yld, ycoff;
  const float* scale; const float* shift;
  int act; float alpha;
  int* err;
};

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int role) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFF) == 0) {
      if (*reinterpret_cast<volatile int*>(err) != 0) return;
      if (clock64() - t0 > kWatchdogCycles) {
        if (atomicCAS(err, 0, 5) == 0) { err[1] = blockIdx.x; err[2] = role; err[3] = (int)parity; }
        __threadfence();
        return;
      }
    }
  }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// kAct: 1 = ReLU, 2 = LeakyReLU with 0 <= alpha <= 1, 0 = any (runtime switch)
template <int kAct>
__global__ void __launch_bounds__(kThreads, 1)
conv_pw_f16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const PwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t stages = (uint32_t)p.stages;
  const uint32_t sA = base;
  const uint32_t sB = sA + stages * kAStage;
  const uint32_t off_stage = stages * kAStage + (uint32_t)p.cchunks * p.b_chunk_bytes;
  const uint32_t off_ss = off_stage + kStageBytes;                  // scale[n_tile] | shift[n_tile] in fp16 (<= 1 KB)
  const uint32_t sBar = base + off_ss + 1024;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * kMaxStages;
  const uint32_t bar_bfull = sBar + 16 * kMaxStages;
  const uint32_t bar_tfull = bar_bfull + 8, bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + off_ss + 1024 + 16 * kMaxStages + 8 + 32);
  uint8_t* stage = base_ptr + off_stage;
  __half* ssh = reinterpret_cast<__half*>(base_ptr + off_ss);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA b serves output-channel block b % num_n (its filter slice is loaded once) and the row tiles b / num_n + k * (grid / num_n)
  const int n_idx = (int)blockIdx.x % p.num_n, n0 = n_idx * p.n_tile;
  const int t_first = (int)blockIdx.x / p.num_n, t_step = (int)gridDim.x / p.num_n;
  const int nchunks = p.n_tile >> 5;
  const bool alternate = nchunks == 1;               // one 32-column chunk per tile: the two column groups take tiles in turn

  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&mapA); ptx::prefetch_tmap(&mapB); }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < stages; ++i) { ptx::mbar_init(bar_full + 8 * i, 1); ptx::mbar_init(bar_empty + 8 * i, 1); }
    ptx::mbar_init(bar_bfull, 1);
    for (uint32_t a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);
      ptx::mbar_init(bar_tempty + 8 * a, alternate ? 128 : 256);    // the epilogue threads that drain this accumulator
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), p.tmem_cols); ptx::tmem_relinquish(); }
  if (warp == 3) {
    for (int i = lane; i < p.n_tile; i += 32) {
      float sc = 0.f, sf = 0.f;
      if (n0 + i < p.Cout) { sc = p.scale ? __ldg(p.scale + n0 + i) : 1.f; sf = p.shift ? __ldg(p.shift + n0 + i) : 0.f; }
      ssh[i] = __float2half_rn(sc);
      ssh[p.n_tile + i] = __float2half_rn(sf);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  ptx::grid_launch_dependents();                     // programmatic dependent launch (common.cuh)

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (t_first < p.num_tiles) {
      // the filter does not depend on the previous kernel: requested before the dependency wait
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(bar_bfull, (uint32_t)p.cchunks * p.b_chunk_bytes);
        for (int cc = 0; cc < p.cchunks; ++cc) ptx::tma_load_2d(sB + cc * p.b_chunk_bytes, &mapB, bar_bfull, cc * 64, n0);
      }
      __syncwarp();
    }
    ptx::grid_dependency_wait();
    uint32_t s = 0, ph = 0;
    for (int tile = t_first; tile < p.num_tiles; tile += t_step) {
      const int row0 = tile * kTileM;
      for (int cc = 0; cc < p.cchunks; ++cc) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1, p.err, 0);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bar_full + 8 * s, kAStage);          // rows beyond M are zero-filled and counted
          ptx::tma_load_2d(sA + s * kAStage, &mapA, bar_full + 8 * s, cc * 64, row0);
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =========================================
    uint32_t s = 0, ph = 0, it = 0;
    const uint32_t desc_hi = (uint32_t)(ptx::make_smem_desc(0, 1024, 2) >> 32);
    const uint32_t a_lo0 = (uint32_t)ptx::make_smem_desc(sA, 1024, 2);
    const uint32_t b_lo0 = (uint32_t)ptx::make_smem_desc(sB, 1024, 2);
    const uint32_t a_step = kAStage >> 4, b_step = p.b_chunk_bytes >> 4;
    const bool elected = ptx::elect_one();
    const int cchunks = p.cchunks;
    for (int tile = t_first; tile < p.num_tiles; tile += t_step, ++it) {
      const uint32_t a = it & 1, tph = (it >> 1) & 1;
      mbar_wait(bar_tempty + 8 * a, tph ^ 1, p.err, 1);
      if (it == 0) mbar_wait(bar_bfull, 0, p.err, 5);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * (uint32_t)p.n_tile;
      for (int cc = 0; cc < cchunks; ++cc) {
        mbar_wait(bar_full + 8 * s, ph, p.err, 2);
        ptx::tc_fence_after();
        if (elected) {
          ptx::umma_f16_x4<1>(d_tmem, a_lo0 + s * a_step, b_lo0 + (uint32_t)cc * b_step, desc_hi, p.idesc, cc ? 1u : 0u);
          ptx::umma_commit(bar_empty + 8 * s);
          if (cc == cchunks - 1) ptx::umma_commit(bar_tfull + 8 * a);
        }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue ===========================================
    ptx::grid_dependency_wait();                     // output rows may still be read by the previous kernel
    const int ew = warp & 3;                         // TMEM lane quarter
    const int eg = (warp - 4) >> 2;                  // column group
    const int half_ = (nchunks + 1) >> 1;
    const int c_begin = alternate ? 0 : eg * half_ * 32, c_end = alternate ? 32 : min(nchunks, (eg + 1) * half_) * 32;
    uint8_t* st_o = stage + (warp - 4) * 2048;
    const uint32_t my_sw = (uint32_t)((lane >> 1) & 3);
    const int piece = lane & 3, rrow = lane >> 2;
    uint8_t* st_w = st_o + lane * 64;
    const uint8_t* st_r = st_o + rrow * 64 + ((piece ^ ((rrow >> 1) & 3)) << 4);
    const __half2 zero2 = __float2half2_rn(0.f), alpha2 = __float2half2_rn(p.alpha);
    auto act2 = [&](__half2 x) {
      if (kAct == 1) return __hmax2(x, zero2);
      if (kAct == 2) return __hmax2(x, __hmul2(x, alpha2));
      const float2 f = __half22float2(x);
      return __floats2half2_rn(plnr_apply_act(f.x, p.act, p.alpha), plnr_apply_act(f.y, p.act, p.alpha));
    };
    uint32_t it = 0;
    for (int tile = t_first; tile < p.num_tiles; tile += t_step, ++it) {
      const uint32_t a = it & 1, tph = (it >> 1) & 1;
      if (alternate && (it & 1u) != (uint32_t)eg) continue;           // the other group's tile (it drains accumulator a alone)
      const int row_base = tile * kTileM + ew * 32;
      mbar_wait(bar_tfull + 8 * a, tph, p.err, 3);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + a * (uint32_t)p.n_tile;
      if (c_begin >= c_end) {                        // a group with no columns (n_tile == 32 handled by `alternate`)
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_tempty + 8 * a);
      }
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(t_row + c0, v);
        ptx::tmem_ld_wait();
        if (c0 + 32 >= c_end) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(bar_tempty + 8 * a);      // accumulator fully read by this thread
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 pk;
          pk.x = pack_half2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
          pk.y = pack_half2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
          pk.z = pack_half2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
          pk.w = pack_half2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
          *reinterpret_cast<uint4*>(st_w + (((uint32_t)q ^ my_sw) << 4)) = pk;
        }
        const int ch = c0 + piece * 8;               // this lane's 8 channels in each of its four rows
        const uint4 sc4 = *reinterpret_cast<const uint4*>(ssh + ch);
        const uint4 sf4 = *reinterpret_cast<const uint4*>(ssh + p.n_tile + ch);
        const __half2* sch = reinterpret_cast<const __half2*>(&sc4);
        const __half2* sfh = reinterpret_cast<const __half2*>(&sf4);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 val = *reinterpret_cast<const uint4*>(st_r + i * 512);
          const int row = row_base + 8 * i + rrow;
          if (row < p.M && n0 + ch + 8 <= p.Cout) {
            __half2* vh = reinterpret_cast<__half2*>(&val);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __half2 x;
              if (kAct == 1) x = __hfma2_relu(vh[e], sch[e], sfh[e]);
              else x = act2(__hfma2(vh[e], sch[e], sfh[e]));
              vh[e] = x;
            }
            *reinterpret_cast<uint4*>(p.y + (size_t)row * p.yld + p.ycoff + n0 + ch) = val;
          }
        }
        __syncwarp();                                // the stage is rewritten by the next chunk
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static int g_driver_version = 0;

static int resolve_driver() {
  if (g_encode_tiled) return PLNR_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    plnr_set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return PLNR_ERR_DRIVER;
  }
  g_encode_tiled = (EncodeTiledFn)fn;
  cudaDriverGetVersion(&g_driver_version);
  return PLNR_OK;
}

static void small_tensor_fixup(CUtensorMap* m, uint64_t tensor_bytes) {
  if (g_driver_version <= 13010 && tensor_bytes < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct PwPlan { bool ok; int n_tile, num_n, cchunks, stages; uint32_t b_chunk_bytes; size_t smem_bytes; };

static PwPlan make_plan(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y, const plnr_epilogue* ep) {
  PwPlan pl;
  memset(&pl, 0, sizeof(pl));
  const char* on = getenv("PLNR_PW");                 // PLNR_PW=0: these layers run conv_shift.cu (A/B timing, parity tests)
  if (on && !atoi(on)) return pl;
  if (d->dtype != PLNR_F16 || d->groups != 1 || d->algo == PLNR_ALGO_DIRECT) return pl;
  if (d->kh != 1 || d->kw != 1 || d->stride_h != 1 || d->stride_w != 1) return pl;
  if (d->pad_t || d->pad_l || d->pad_b || d->pad_r) return pl;
  if (ep && (ep->residual || ep->out_nchw || ep->out_f32)) return pl;
  if (x->c % 64 != 0 || x->c > 64 * kMaxChunks || x->ld % 8 != 0 || x->coff % 8 != 0 || (reinterpret_cast<uintptr_t>(x->ptr) & 15)) return pl;
  if (y->c % 8 != 0 || y->ld % 8 != 0 || y->coff % 8 != 0 || (reinterpret_cast<uintptr_t>(y->ptr) & 15)) return pl;
  const long long M = (long long)x->n * x->h * x->w;
  if (M < 1 || M >= (1ll << 31) - 256) return pl;
  pl.cchunks = x->c / 64;
  // output-channel block: all channels when their filter fits beside the A ring, else slices of 128 / 64 channels pinned to
  // CTAs (the slice stays resident; the activations, the small operand of such layers, are re-read per slice from L2)
  const size_t kMaxB = 128 * 1024;
  pl.n_tile = round_up(y->c, 32) <= 256 ? round_up(y->c, 32) : 256;
  while (pl.n_tile > 64 && (size_t)pl.cchunks * pl.n_tile * 128 > kMaxB) pl.n_tile = pl.n_tile > 128 ? 128 : 64;
  pl.num_n = (y->c + pl.n_tile - 1) / pl.n_tile;
  pl.b_chunk_bytes = (uint32_t)pl.n_tile * 128u;
  const size_t b_bytes = (size_t)pl.cchunks * pl.b_chunk_bytes;
  if (b_bytes > kMaxB || pl.num_n > 16) return pl;
  // Large filters are worth it only where the rows are few: with many row tiles per CTA conv_shift.cu's streamed N = 256 CTA
  // pairs run at the tensor floor, while few tiles leave it re-streaming the whole filter per tile (YOLOv3 512 -> 256 @26x26:
  // 23 us there for 6 us of traffic)
  if ((size_t)x->c * y->c * 2 > 64 * 1024 && M > 65536) return pl;
  const size_t fixed = kStageBytes + 1024 + 16 * kMaxStages + 8 + 32 + 64 + 1024;
  size_t room = 232448 - fixed - b_bytes;
  int stages = (int)(room / kAStage);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 3) return pl;
  pl.stages = stages;
  pl.smem_bytes = fixed + b_bytes + (size_t)stages * kAStage;
  pl.ok = true;
  return pl;
}

}  // namespace

bool plnr_conv2d_pw_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y, const plnr_epilogue* ep) {
  return make_plan(d, x, y, ep).ok;
}

int plnr_conv2d_pw(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w, const plnr_tensor* y,
                   const plnr_epilogue* ep) {
  int rc = resolve_driver();
  if (rc != PLNR_OK) return rc;
  const PwPlan pl = make_plan(d, x, y, ep);
  PLNR_REQUIRE(pl.ok, "conv2d(pointwise): problem not eligible");
  PLNR_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv2d(pointwise): packed weights must be 16-byte aligned");

  PwParams p;
  memset(&p, 0, sizeof(p));
  p.M = x->n * x->h * x->w; p.Cout = y->c; p.cchunks = pl.cchunks; p.n_tile = pl.n_tile;
  p.num_tiles = (p.M + kTileM - 1) / kTileM;
  p.stages = pl.stages;
  p.num_n = pl.num_n;
  p.b_chunk_bytes = pl.b_chunk_bytes;
  p.idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
  uint32_t cols = 32;
  while (cols < 2u * p.n_tile) cols <<= 1;
  p.tmem_cols = cols;
  p.y = (__half*)y->ptr; p.yld = y->ld; p.ycoff = y->coff;
  if (ep) { p.scale = ep->scale; p.shift = ep->shift; p.act = ep->act; p.alpha = ep->alpha; }
  p.err = ctx->dev_error;

  CUtensorMap mapA, mapB;
  {
    // the activations as a matrix [M rows][Cin] with row pitch ld: box = 64 channels x 128 rows, 128B swizzle
    const cuuint64_t dims[2] = {(cuuint64_t)x->c, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)x->ld * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)kTileM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)((__half*)x->ptr + x->coff), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { plnr_set_error("cuTensorMapEncodeTiled(A, pointwise) failed with CUresult %d", (int)r); return PLNR_ERR_DRIVER; }
    small_tensor_fixup(&mapA, (uint64_t)p.M * x->ld * 2);
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)x->c, (cuuint64_t)y->c};
    const cuuint64_t strides[1] = {(cuuint64_t)x->c * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)p.n_tile};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { plnr_set_error("cuTensorMapEncodeTiled(B, pointwise) failed with CUresult %d", (int)r); return PLNR_ERR_DRIVER; }
    small_tensor_fixup(&mapB, (uint64_t)x->c * y->c * 2);
  }

  const int kact = p.act == PLNR_ACT_RELU ? 1 : (p.act == PLNR_ACT_LEAKY && p.alpha >= 0.f && p.alpha <= 1.f ? 2 : 0);
  typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const PwParams);
  static const KernFn kerns[3] = {conv_pw_f16_kernel<0>, conv_pw_f16_kernel<1>, conv_pw_f16_kernel<2>};
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[kact]) {
    PLNR_CHECK_CUDA(cudaFuncSetAttribute(kerns[kact], cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set[kact] = true;
  }
  int per_n = ctx->sm_count / pl.num_n;                       // every output-channel block gets the same number of CTAs
  if (per_n > p.num_tiles) per_n = p.num_tiles;
  if (per_n < 1) per_n = 1;
  const int grid = per_n * pl.num_n;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = plnr_pdl_enabled() ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kerns[kact], mapA, mapB, p);
  if (le != cudaSuccess) {
    plnr_set_error("launch of conv_pw_f16_kernel failed: %s", cudaGetErrorString(le));
    return PLNR_ERR_CUDA;
  }
  return plnr_after_launch(ctx, "conv2d_pw");
}

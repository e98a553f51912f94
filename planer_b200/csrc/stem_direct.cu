// First layer of an image network with a SMALL filter bank: 3x3 / stride-1 / pad-1 convolution of a 1..4-channel NCHW image
// (fp16 or uint8) to <= 32 channels, + scale / shift (bias, folded BatchNorm) + ReLU / LeakyReLU, written pixel-major (NHWC).
//
// Replaces, for YOLOv3's 3 -> 32 stem, the reference chain Conv2d (planer/layer.py:22-26 + planer/util.py:17-44) ->
// BatchNorm (planer/layer.py:125-127) -> LeakyReLU (planer/layer.py:48-51) and this library's earlier two launches
// (plnr_stem_pack + a tensor-core conv with K = 48, N = 32: 29 TFLOP/s, 12 % of the YOLOv3 step).  With K = 27 and
// N = 32 the layer is 9.6 GFLOP against 16.6 MB read + 354 MB written per 32 images: too small for a 128-row tcgen05 tile
// pipeline, too many FLOPs for HFMA2 (the first version: 0.33 ms, instruction-bound).  It runs on the WARP-level tensor-core
// instruction, straight from the NCHW image, next to the stores:
//   * the filter travels in the KERNEL PARAMETERS (27 taps x 32 channels of fp16 = 1.7 KB), is staged once per block in shared
//     memory and lives in 16 registers per lane as mma.sync B fragments;
//   * the 27 input values of a pixel are 2-byte global loads gathered directly into A fragments (rows = 16 adjacent pixels),
//     served by L1 (neighbouring pixels share 6 of 9 positions);
//   * all 27 products accumulate in fp32, the sum is rounded to fp16 once (the reference's conv output is an fp16 array), then
//     scale / shift as one HFMA2 and the activation in fp16, like every other conv epilogue of this library;
//   * a pixel's 32 channels are 64 contiguous bytes, written by the four lanes of a quad as 16-byte pieces (stem3_tap).
#include "common.cuh"

namespace {

constexpr int kMaxCout = 32, kTaps = 27;

struct Stem3Params {
  const void* x;
  int N, C, H, W;
  __half* y; int yld, ycoff, Cout;
  int act; float alpha;
  alignas(16) __half2 w[kTaps][kMaxCout / 2];       // [(c*3 + r)*3 + s][channel pair], zero beyond C / Cout (16-byte constant loads)
  __half2 scale[kMaxCout / 2], shift[kMaxCout / 2];
};

template <typename Tin> __device__ __forceinline__ __half ld_h(const Tin* p);
template <> __device__ __forceinline__ __half ld_h<__half>(const __half* p) { return __ldg(p); }
template <> __device__ __forceinline__ __half ld_h<uint8_t>(const uint8_t* p) { return __ushort2half_rn((unsigned short)__ldg(p)); }

// One WARP = 16 horizontally adjacent output pixels x 32 channels per trip, on the warp-level tensor-core instruction
// (mma.sync.m16n8k16, fp16 x fp16 -> fp32): the 27 taps are the K dimension (padded to 32 = two K=16 steps), the 32 output
// channels four N=8 blocks.  The first version of this kernel (four pixels x 16 channels per thread, HFMA2 with the filter in the
// constant bank) executed 2200 instructions per thread of which 906 were HFMA2 and ran at 29 TFLOP/s, 0.33 ms for 32 x 416 x 416
// (ncu: issue slots 58 % busy, FMA pipe 59 %): instruction-bound.  Here a warp spends ~215 instructions per 16 pixels: 16
// two-byte loads per lane gather the A fragments (rows = pixels, neighbours hit L1), 8 mma.sync, and the epilogue of the other conv
// kernels (fp32 sum rounded to fp16 once, scale / shift as one HFMA2, activation) on the C fragments.  The filter's B fragments
// (16 registers) are read once per warp from shared memory, where the block stages the kernel parameters.  80 registers, three
// blocks per SM: 0.24 ms for 32 x 416 x 416 (0.33 before); that version was bound by the L1 / LSU tag rate of its gathers (ncu:
// 49 % excessive sectors -- every load request touched four lines, every 4-byte store request eight half-used sectors).  The
// fragment orders below (stem3_tap) fix both without shared memory or shuffles: 0.18 ms.
constexpr int kWarpsPerBlock = 8;

// Order of the contraction index and of the output channels inside the fragments.  ANY order works as long as the A gather
// and the filter staging agree, so both are chosen for the memory system (v4; profiles/r02_kernel_experiments.md 17):
//   * k = 16 ks + 8 hh + 2 tig + e.  The load instruction j = 4 ks + 2 hh + e of a warp covers lanes tig = 0..3 of eight
//     adjacent pixels; it reads input row j = (channel j / 3, filter row j % 3) at horizontal taps s = tig for tig < 3 -- ten
//     adjacent 2-byte values, one or two sectors of ONE line, where the natural order k = (c*3 + r)*3 + s touched three or
//     four lines per instruction.  The ninth row (c = 2, r = 2) rides in the tig = 3 lanes of j = 0..2, the other tig = 3
//     slots are zero padding.
//   * MMA column 8 nb + 2 tig + e holds output channel 8 tig + 2 nb + e: a lane's four N blocks are EIGHT CONSECUTIVE
//     channels = one 16-byte store, a quad writes a whole 64-byte pixel and a warp instruction 512 contiguous bytes (before:
//     eight 4-byte stores per lane, each warp instruction half-filling eight sectors).
__device__ __forceinline__ bool stem3_tap(int j, int tig, int& c, int& r, int& sx) {
  int row, s;
  if (tig < 3) { row = j; s = tig; }
  else { row = 8; s = j; if (j >= 3) { c = r = sx = 0; return false; } }
  c = row / 3; r = row - 3 * c; sx = s;
  return true;
}

template <typename Tin, int kAct>
__global__ void __launch_bounds__(32 * kWarpsPerBlock, 3) stem3x3_kernel(const __grid_constant__ Stem3Params p, uint32_t num_tiles,
                                                                      uint32_t tiles_per_row, FastDiv div_tpr, FastDiv div_h) {
  __shared__ __half wsm[32][kMaxCout + 8];                   // [k (27 taps, zero-padded to 32)][channel], pitch 40: conflict-light
  for (int i = threadIdx.x; i < 32 * kMaxCout; i += blockDim.x) {
    const int k = i / kMaxCout, n = i - k * kMaxCout;
    int c, r, sx;
    const bool ok = stem3_tap(4 * (k >> 4) + 2 * ((k >> 3) & 1) + (k & 1), (k >> 1) & 3, c, r, sx);
    wsm[k][n] = ok ? reinterpret_cast<const __half*>(&p.w[(c * 3 + r) * 3 + sx][0])[n] : __float2half_rn(0.f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;                 // fragment coordinates of mma.m16n8k16
  // B fragments: b[ks][nb][0] = {W[16 ks + 2 tig][ch], W[16 ks + 2 tig + 1][ch]}, b[ks][nb][1] = the same at k + 8, where
  // ch = 8 (gid / 2) + 2 nb + gid % 2 is the output channel of MMA column 8 nb + gid
  uint32_t bfr[2][4][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k = 16 * ks + 2 * tig + 8 * hh;
        const int ch = 8 * (gid >> 1) + 2 * nb + (gid & 1);
        const __half2 v = __halves2half2(wsm[k][ch], wsm[k + 1][ch]);
        bfr[ks][nb][hh] = *reinterpret_cast<const uint32_t*>(&v);
      }
  // this lane's eight k indices (two K=16 steps x {2 tig, 2 tig + 1, 2 tig + 8, 2 tig + 9}) decoded once: input offset
  // relative to the output pixel, row / column displacement for the zero padding
  // (kept small: koff + one bit mask; the border path re-derives row / column displacements from k)
  int koff[8];
  uint32_t kmask = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int c, r, sx;
    const bool ok = stem3_tap(j, tig, c, r, sx) && c < p.C;
    kmask |= (ok ? 1u : 0u) << j;
    koff[j] = ok ? (c * p.H + (r - 1)) * p.W + (sx - 1) : 0;      // padded k: a valid address whose value is discarded
    if (!ok && tig == 3 && j >= 3 && j / 3 < p.C) koff[j] = ((j / 3) * p.H + (j % 3 - 1)) * p.W;   // ... in the sector its quad reads anyway
  }
  // scale / shift of this lane's channel pairs (8 tig + 2 nb, + 1)
  __shared__ __half2 ssm[2][kMaxCout / 2];
  if (threadIdx.x < kMaxCout / 2) { ssm[0][threadIdx.x] = p.scale[threadIdx.x]; ssm[1][threadIdx.x] = p.shift[threadIdx.x]; }
  __syncthreads();
  const __half2 alpha2 = __float2half2_rn(p.alpha);
  const __half zero = __float2half_rn(0.f);
  const size_t img_stride = (size_t)p.C * p.H * p.W;

  for (uint32_t tile = blockIdx.x * kWarpsPerBlock + warp; tile < num_tiles; tile += gridDim.x * kWarpsPerBlock) {
    // tile -> (image, row, 16-pixel group) by multiply-high divisions (the 64-bit % and / of the first version of this loop
    // were a quarter of its 444 instructions)
    const uint32_t t2 = fast_div(tile, div_tpr);
    const int tw = (int)(tile - t2 * tiles_per_row);
    const int n = (int)fast_div(t2, div_h);
    const int h = (int)(t2 - (uint32_t)n * (uint32_t)p.H);
    const int w0 = tw * 16;
    // interior tiles (91 % at 416 x 416) need no padding tests: all 27 taps of all 16 pixels are inside the image
    const bool interior = h >= 1 && h + 1 < p.H && w0 >= 1 && w0 + 16 < p.W;
    const Tin* xc = reinterpret_cast<const Tin*>(p.x) + (size_t)n * img_stride + (size_t)h * p.W + w0;
    // A fragments: registers {a0, a1, a2, a3} of step ks = (pixel gid, k 2tig..), (pixel gid + 8, same k), (gid, k + 8), (gid + 8, k + 8)
    uint32_t afr[2][4];
    if (interior) {
      // every load is unconditional (no branches): a padded k reads the pixel itself and is replaced by zero
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const int pix = gid + 8 * pp;
            __half v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = 4 * ks + 2 * hh + e;
              const __half t = ld_h<Tin>(xc + pix + koff[j]);
              v[e] = ((kmask >> j) & 1u) ? t : zero;
            }
            const __half2 pk = __halves2half2(v[0], v[1]);
            afr[ks][2 * hh + pp] = *reinterpret_cast<const uint32_t*>(&pk);
          }
    } else {
      // image border: coordinates clamped into the image (the load stays unconditional), taps outside contribute zero
      const Tin* xi = reinterpret_cast<const Tin*>(p.x) + (size_t)n * img_stride;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const int pix = gid + 8 * pp;
            __half v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = 4 * ks + 2 * hh + e;
              int c, r, sx;
              stem3_tap(j, tig, c, r, sx);
              const bool kvalid = (kmask >> j) & 1u;
              const int ih = h + r - 1, iw = w0 + pix + sx - 1;
              const bool ok = kvalid && (unsigned)ih < (unsigned)p.H && (unsigned)iw < (unsigned)p.W;
              const int ihc = min(max(ih, 0), p.H - 1), iwc = min(max(iw, 0), p.W - 1);
              const int cofs = kvalid ? c * p.H * p.W : 0;                   // channel plane offset of this k (0 when padded)
              const __half t = ld_h<Tin>(xi + cofs + (size_t)ihc * p.W + iwc);
              v[e] = ok ? t : zero;
            }
            const __half2 pk = __halves2half2(v[0], v[1]);
            afr[ks][2 * hh + pp] = *reinterpret_cast<const uint32_t*>(&pk);
          }
    }
    float acc[4][4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nb][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[nb][0]), "+f"(acc[nb][1]), "+f"(acc[nb][2]), "+f"(acc[nb][3])
                     : "r"(afr[ks][0]), "r"(afr[ks][1]), "r"(afr[ks][2]), "r"(afr[ks][3]), "r"(bfr[ks][nb][0]), "r"(bfr[ks][nb][1]));
    }
    // C fragment: (pixel gid, channels 8 tig + 2 nb, + 1) in acc[nb][0..1], (pixel gid + 8, same channels) in acc[nb][2..3]:
    // the four N blocks of a lane are eight consecutive channels, one 16-byte store per pixel
    const long long pix0 = ((long long)n * p.H + h) * p.W + w0;
    if (8 * tig < p.Cout) {
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const int pix = gid + 8 * pp;
        if (w0 + pix < p.W) {
          uint4 out;
          uint32_t* o = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
          for (int nb = 0; nb < 4; ++nb) {
            __half2 v = __floats2half2_rn(acc[nb][2 * pp], acc[nb][2 * pp + 1]);      // conv output rounded to fp16 once
            const __half2 sc = ssm[0][4 * tig + nb], sf = ssm[1][4 * tig + nb];
            if (kAct == 1) v = __hfma2_relu(v, sc, sf);
            else {
              v = __hfma2(v, sc, sf);
              if (kAct == 2) v = __hmax2(v, __hmul2(v, alpha2));
              else if (kAct == 3) {
                const float2 f = __half22float2(v);
                v = __floats2half2_rn(plnr_apply_act(f.x, p.act, p.alpha), plnr_apply_act(f.y, p.act, p.alpha));
              }
            }
            o[nb] = *reinterpret_cast<const uint32_t*>(&v);
          }
          *reinterpret_cast<uint4*>(p.y + (size_t)(pix0 + pix) * p.yld + p.ycoff + 8 * tig) = out;
        }
      }
    }
  }
}

}  // namespace

extern "C" int plnr_stem3x3_supported(int dtype, int c, int cout, int kh, int kw, int stride, int pad_t, int pad_l, int pad_b,
                                      int pad_r, int dil) {
  if (dtype != PLNR_F16) return 0;
  if (c < 1 || c > 3 || cout < 8 || cout > kMaxCout || cout % 8 != 0) return 0;
  if (kh != 3 || kw != 3 || stride != 1 || dil != 1) return 0;
  return (pad_t == 1 && pad_l == 1 && pad_b == 1 && pad_r == 1) ? 1 : 0;
}

// w_oihw: HOST pointer to the OIHW fp16 filter (cout, c, 3, 3); scale / shift: HOST fp32 [cout] or NULL.  They travel in the
// kernel parameters, so the call is safe inside CUDA-graph capture (an earlier version cached a device read-back per
// POINTER: a freed and re-used address then served a stale filter).
extern "C" int plnr_stem3x3_fwd(plnr_ctx* ctx, const void* x, int x_dtype, int n, int c, int h, int w, const void* w_oihw,
                                const float* scale, const float* shift, int act, float alpha, const plnr_tensor* y) {
  PLNR_REQUIRE(ctx && x && w_oihw && y && y->ptr, "stem3x3: NULL argument");
  PLNR_REQUIRE(x_dtype == PLNR_F16 || x_dtype == PLNR_U8, "stem3x3: the image must be fp16 or uint8");
  PLNR_REQUIRE(plnr_stem3x3_supported(PLNR_F16, c, y->c, 3, 3, 1, 1, 1, 1, 1, 1), "stem3x3: unsupported problem (c=%d cout=%d)", c, y->c);
  PLNR_REQUIRE(y->n == n && y->h == h && y->w == w, "stem3x3: output is (%d,%d,%d), expected (%d,%d,%d)", y->n, y->h, y->w, n, h, w);
  PLNR_REQUIRE((reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0 && y->ld % 8 == 0 && y->coff % 8 == 0,
               "stem3x3: output rows must be 16-byte aligned");
  {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, w_oihw) == cudaSuccess)
      PLNR_REQUIRE(pa.type != cudaMemoryTypeDevice, "stem3x3: w_oihw must be a HOST pointer (ABI 2)");
    cudaGetLastError();
  }
  const int cout = y->c;
  const __half* hw = reinterpret_cast<const __half*>(w_oihw);
  const float* sc = scale;
  const float* sf = shift;

  Stem3Params p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.N = n; p.C = c; p.H = h; p.W = w;
  p.y = (__half*)y->ptr; p.yld = y->ld; p.ycoff = y->coff; p.Cout = cout;
  p.act = act; p.alpha = alpha;
  for (int ci = 0; ci < c; ++ci)
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s)
        for (int co = 0; co < cout; ++co)
          reinterpret_cast<__half*>(&p.w[(ci * 3 + r) * 3 + s][0])[co] = hw[((size_t)co * c + ci) * 9 + r * 3 + s];
  for (int co = 0; co < kMaxCout; ++co) {
    reinterpret_cast<__half*>(&p.scale[0])[co] = __float2half_rn(co < cout ? (sc ? sc[co] : 1.f) : 0.f);
    reinterpret_cast<__half*>(&p.shift[0])[co] = __float2half_rn(co < cout ? (sf ? sf[co] : 0.f) : 0.f);
  }
  const int tiles_per_row = (w + 15) / 16;
  const long long num_tiles = (long long)n * h * tiles_per_row;                    // one warp trip = 16 pixels of one row
  PLNR_REQUIRE(num_tiles > 0 && num_tiles < (1ll << 31), "stem3x3: bad extents");
  long long blocks = (num_tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long)ctx->sm_count * 32;                             // grid-stride: the B fragments are set up once per warp
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  const int kact = act == PLNR_ACT_RELU ? 1 : (act == PLNR_ACT_LEAKY && alpha >= 0.f && alpha <= 1.f ? 2 : (act == PLNR_ACT_NONE ? 0 : 3));
#define LAUNCH(T, A) stem3x3_kernel<T, A><<<grid, 32 * kWarpsPerBlock, 0, ctx->stream>>>(p, (uint32_t)num_tiles, (uint32_t)tiles_per_row, make_fastdiv((uint32_t)tiles_per_row), make_fastdiv((uint32_t)h))
#define LAUNCH_T(T) do { if (kact == 1) LAUNCH(T, 1); else if (kact == 2) LAUNCH(T, 2); else if (kact == 0) LAUNCH(T, 0); else LAUNCH(T, 3); } while (0)
  if (x_dtype == PLNR_U8) LAUNCH_T(uint8_t); else LAUNCH_T(__half);
#undef LAUNCH_T
#undef LAUNCH
  return plnr_after_launch(ctx, "stem3x3");
}

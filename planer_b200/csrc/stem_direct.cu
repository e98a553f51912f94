// First layer of an image network with a SMALL filter bank: 3x3 / stride-1 / pad-1 convolution of a 1..4-channel NCHW image
// (fp16 or uint8) to <= 32 channels, + scale / shift (bias, folded BatchNorm) + ReLU / LeakyReLU, written pixel-major (NHWC).
//
// Replaces, for YOLOv3's 3 -> 32 stem, the reference chain Conv2d (planer/layer.py:22-26 + planer/util.py:17-44) ->
// BatchNorm (planer/layer.py:125-127) -> LeakyReLU (planer/layer.py:48-51) and this library's earlier two launches
// (plnr_stem_pack + a tensor-core conv with K = 48, N = 32: 29 TFLOP/s, 12 % of the YOLOv3 step).  With K = 27 and
// N = 32 the layer is 9.6 GFLOP against 16.6 MB read + 354 MB written per 32 images: it belongs on the CUDA cores, next to
// the stores.  One thread = four adjacent output pixels x 16 channels:
//   * the filter lives in the KERNEL PARAMETERS (27 taps x 16 channel pairs of fp16 = 1.7 KB), so every HFMA2 takes its
//     weight operand straight from the constant bank -- no shared memory, no weight loads;
//   * the 27 input values of a pixel are 2-byte global loads, coalesced across the warp (consecutive pixels) and served by L1
//     (neighbouring pixels share 6 of 9 positions);
//   * products of one filter row (9 terms) accumulate in packed fp16, the three rows are added in fp32, the sum is rounded to
//     fp16 once (the reference's conv output is an fp16 array), then scale / shift as one HFMA2 and the activation in fp16,
//     like every other conv epilogue of this library;
//   * a pixel's 32 channels are 64 contiguous bytes, written by the two threads of a pixel group as 2 x 2 16-byte stores.
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

// One thread = FOUR horizontally adjacent output pixels x 16 channels (one half of the filter bank): a weight pair fetched
// from the constant bank feeds four HFMA2 instead of one (the one-pixel version spent 285 LDC on 467 HFMA2 per pixel and ran
// at 26 TFLOP/s), and the six input values of a filter row serve all four pixels.
constexpr int kPix = 4;

template <typename Tin, int kAct>
__global__ void __launch_bounds__(256, 2) stem3x3_kernel(const __grid_constant__ Stem3Params p) {
  const int half_ = threadIdx.x & 1;                         // channels [16 half_, 16 half_ + 16)
  const long long grp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  const int gw = (p.W + kPix - 1) / kPix;                    // pixel groups per row
  const long long total = (long long)p.N * p.H * gw;
  if (grp >= total) return;
  const int g_ = (int)(grp % gw);
  const long long t = grp / gw;
  const int h = (int)(t % p.H), n = (int)(t / p.H);
  const int w0 = g_ * kPix;
  const Tin* x = reinterpret_cast<const Tin*>(p.x) + (size_t)n * p.C * p.H * p.W;
  const __half zero = __float2half_rn(0.f);
  float tot[kPix][16];
#pragma unroll
  for (int i = 0; i < kPix; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) tot[i][j] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int ih = h + r - 1;
    const bool rok = ih >= 0 && ih < p.H;
    __half2 acc[kPix][8];
#pragma unroll
    for (int i = 0; i < kPix; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = __half2half2(zero);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < p.C) {
        const Tin* row = x + ((size_t)c * p.H + (rok ? ih : 0)) * p.W;
        __half2 xv[kPix + 2];                                // columns w0 - 1 .. w0 + 4, broadcast to both halves of a pair
#pragma unroll
        for (int q = 0; q < kPix + 2; ++q) {
          const int iw = w0 + q - 1;
          xv[q] = __half2half2((rok && iw >= 0 && iw < p.W) ? ld_h<Tin>(row + iw) : zero);
        }
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          // this thread's 8 weight pairs of the tap: 16-byte constant-bank loads
          const uint4* wt = reinterpret_cast<const uint4*>(&p.w[(c * 3 + r) * 3 + s][half_ * 8]);
          const uint4 wa = wt[0], wb = wt[1];
          const __half2 wv[8] = {*reinterpret_cast<const __half2*>(&wa.x), *reinterpret_cast<const __half2*>(&wa.y),
                                 *reinterpret_cast<const __half2*>(&wa.z), *reinterpret_cast<const __half2*>(&wa.w),
                                 *reinterpret_cast<const __half2*>(&wb.x), *reinterpret_cast<const __half2*>(&wb.y),
                                 *reinterpret_cast<const __half2*>(&wb.z), *reinterpret_cast<const __half2*>(&wb.w)};
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < kPix; ++i) acc[i][j] = __hfma2(xv[i + s], wv[j], acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kPix; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 f = __half22float2(acc[i][j]);
        tot[i][2 * j] += f.x; tot[i][2 * j + 1] += f.y;
      }
  }
  const __half2 alpha2 = __float2half2_rn(p.alpha);
  const long long pix0 = ((long long)n * p.H + h) * p.W + w0;
#pragma unroll
  for (int i = 0; i < kPix; ++i) {
    if (w0 + i < p.W) {
      __half* yrow = p.y + (size_t)(pix0 + i) * p.yld + p.ycoff + half_ * 16;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (half_ * 16 + q * 8 < p.Cout) {
          uint4 out;
          __half2* oh = reinterpret_cast<__half2*>(&out);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = q * 4 + e;
            __half2 v = __floats2half2_rn(tot[i][2 * j], tot[i][2 * j + 1]);          // conv output rounded to fp16 once
            const __half2 sc = p.scale[half_ * 8 + j], sf = p.shift[half_ * 8 + j];
            if (kAct == 1) v = __hfma2_relu(v, sc, sf);
            else {
              v = __hfma2(v, sc, sf);
              if (kAct == 2) v = __hmax2(v, __hmul2(v, alpha2));
              else if (kAct == 3) {
                const float2 f = __half22float2(v);
                v = __floats2half2_rn(plnr_apply_act(f.x, p.act, p.alpha), plnr_apply_act(f.y, p.act, p.alpha));
              }
            }
            oh[e] = v;
          }
          *reinterpret_cast<uint4*>(yrow + q * 8) = out;
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
  const long long total = (long long)n * h * ((w + kPix - 1) / kPix) * 2;          // threads: (group of four pixels) x (channel half)
  PLNR_REQUIRE(total > 0 && (total + 255) / 256 < (1ll << 31), "stem3x3: bad extents");
  const unsigned grid = (unsigned)((total + 255) / 256);
  const int kact = act == PLNR_ACT_RELU ? 1 : (act == PLNR_ACT_LEAKY && alpha >= 0.f && alpha <= 1.f ? 2 : (act == PLNR_ACT_NONE ? 0 : 3));
#define LAUNCH(T, A) stem3x3_kernel<T, A><<<grid, 256, 0, ctx->stream>>>(p)
#define LAUNCH_T(T) do { if (kact == 1) LAUNCH(T, 1); else if (kact == 2) LAUNCH(T, 2); else if (kact == 0) LAUNCH(T, 0); else LAUNCH(T, 3); } while (0)
  if (x_dtype == PLNR_U8) LAUNCH_T(uint8_t); else LAUNCH_T(__half);
#undef LAUNCH_T
#undef LAUNCH
  return plnr_after_launch(ctx, "stem3x3");
}

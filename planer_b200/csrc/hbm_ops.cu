// HBM-bound companions of the conv/dense path: layout transforms, casts, weight packing, maxpool,
// nearest upsample, channel-slice copy (concat), the elementwise family and global average pool.
// All are coalesced, 16-byte-vectorised streaming kernels over the pixel-major (NHWC) layout; their
// roofline is HBM bandwidth (algorithmic bytes = (elements_in + elements_out) * sizeof(dtype)).
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T, int V> struct alignas(sizeof(T) * V) Vec { T v[V]; };

template <typename T> struct VecWidth { static constexpr int value = 16 / sizeof(T); };

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static inline int grid_for(int64_t work, int sm_count) {
  int64_t blocks = (work + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)sm_count * 16;  // grid-stride loops: a few resident waves are enough
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// NCHW <-> pixel-major transposes through a 32x33 shared tile (coalesced on both sides)
// ---------------------------------------------------------------------------------------------
template <typename Tin, typename Tout>
__global__ void nchw_to_nhwc_kernel(const Tin* __restrict__ x, Tout* __restrict__ y, int C, int HW, int Cy, int ld,
                                    int coff) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < HW) v = ld_f(x + ((size_t)n * C + c) * HW + p);
    tile[j][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    int p = p0 + j, c = c0 + tx;
    if (p < HW && c < Cy) st_f(y + ((size_t)n * HW + p) * ld + coff + c, tile[tx][j]);
  }
}

template <typename Tin, typename Tout>
__global__ void nhwc_to_nchw_kernel(const Tin* __restrict__ x, Tout* __restrict__ y, int C, int HW, int ld, int coff) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    int p = p0 + j, c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < C) v = ld_f(x + ((size_t)n * HW + p) * ld + coff + c);
    tile[j][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j, p = p0 + tx;
    if (c < C && p < HW) st_f(y + ((size_t)n * C + c) * HW + p, tile[tx][j]);
  }
}

// fp16 pixel-major -> NCHW without shared memory: a thread loads an 8-pixel x 8-channel block as eight 16-byte vectors (one
// per pixel), transposes it in registers (32 byte-permutes) and stores eight 16-byte vectors (one per channel: 8 consecutive
// pixels).  A warp covers 8 pixel blocks x 4 channel groups, so a load instruction touches 64-byte runs and a store instruction
// 128-byte runs: every sector moved is used in full.  (The 32x33 float tile above moves 2-byte elements: 0.28 of the copy
// bandwidth; this one is bound by HBM.)  Stores fall back to narrower vectors when the NCHW row is not 16-byte aligned
// (H*W not a multiple of 8).
__global__ void __launch_bounds__(256) nhwc_to_nchw_h16_kernel(const __half* __restrict__ x, __half* __restrict__ y, int C, int HW,
                                                                int ld, int coff, int PB8, int CG4, long long total) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int lane = (int)(idx & 31);
  long long t = idx >> 5;
  const int cg4 = (int)(t % CG4); t /= CG4;
  const int pb8 = (int)(t % PB8);
  const int n = (int)(t / PB8);
  const int p0 = (pb8 * 8 + (lane & 7)) * 8;             // first pixel of this thread's block
  const int c0 = (cg4 * 4 + (lane >> 3)) * 8;            // first channel
  if (p0 >= HW || c0 >= C) return;
  const __half* xb = x + ((size_t)n * HW + p0) * ld + coff + c0;
  uint4 in[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    in[j] = make_uint4(0u, 0u, 0u, 0u);
    if (p0 + j < HW) in[j] = *reinterpret_cast<const uint4*>(xb + (size_t)j * ld);
  }
  const int npix = min(8, HW - p0);
#pragma unroll
  for (int k = 0; k < 8; ++k) {                          // channel c0 + k = half (k & 1) of word k / 2 of every pixel
    uint32_t o[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const uint32_t w0 = reinterpret_cast<const uint32_t*>(&in[2 * a])[k >> 1];
      const uint32_t w1 = reinterpret_cast<const uint32_t*>(&in[2 * a + 1])[k >> 1];
      o[a] = __byte_perm(w0, w1, (k & 1) ? 0x7632 : 0x5410);
    }
    if (c0 + k >= C) break;                              // channel count not a multiple of 8: the row's pad channels are dropped
    __half* dst = y + ((size_t)n * C + c0 + k) * HW + p0;
    const uintptr_t ad = reinterpret_cast<uintptr_t>(dst);
    if (npix == 8 && (ad & 15) == 0) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    } else if (npix == 8 && (ad & 7) == 0) {
      *reinterpret_cast<uint2*>(dst) = make_uint2(o[0], o[1]);
      *reinterpret_cast<uint2*>(dst + 4) = make_uint2(o[2], o[3]);
    } else if (npix == 8 && (ad & 3) == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a) *reinterpret_cast<uint32_t*>(dst + 2 * a) = o[a];
    } else {
      const __half* oh = reinterpret_cast<const __half*>(o);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < npix) dst[j] = oh[j];
    }
  }
}

// Stem packing: NCHW input (few channels) -> pixel-major tensor whose "channels" are the horizontal filter taps x
// vertical stride phases x input channels of one output column, so that a k x k / stride-s convolution becomes a
// (T x 1) stride-1 convolution with 16/32/64 channels (one TMA row per pixel instead of k*k tiny ones).
//   y[n, h2, ow, (ph*kw + sx)*C + c] = x[n, c, s*h2 + ph, s*ow + sx - pad_l]      (0 outside the image)
template <typename Tin>
__global__ void __launch_bounds__(256) stem_pack_kernel(const Tin* __restrict__ x, __half* __restrict__ y, int N, int C,
                                                         int H, int W, int H2, int OW, int CP, int kw, int s, int pad_l,
                                                         int Wp) {
  // one CTA per packed row (n, h2): stage the s*C source rows in shared memory (coalesced reads), then emit the
  // OW * CP packed values as 16-byte stores (coalesced writes).  tab[ch] = offset of packed channel ch in `rows`.
  extern __shared__ __half stem_smem[];
  __half* rows = stem_smem;                       // [s*C][Wp], column j holds input column j - pad_l
  int* tab = reinterpret_cast<int*>(stem_smem + ((s * C * Wp + 7) / 8) * 8);
  const int n = blockIdx.x / H2, h2 = blockIdx.x % H2;
  const int real = s * kw * C;
  for (int i = threadIdx.x; i < s * C * Wp; i += blockDim.x) {
    const int j = i % Wp, rc = i / Wp;
    const int c = rc % C, ph = rc / C;
    const int ih = s * h2 + ph, iw = j - pad_l;
    float v = 0.f;
    if (ih < H && iw >= 0 && iw < W) v = ld_f(x + (((size_t)n * C + c) * H + ih) * W + iw);
    rows[i] = __float2half_rn(v);
  }
  for (int ch = threadIdx.x; ch < CP; ch += blockDim.x) {
    int off = -1;
    if (ch < real) {
      const int c = ch % C, tap = ch / C;
      const int sx = tap % kw, ph = tap / kw;
      off = (ph * C + c) * Wp + sx;
    }
    tab[ch] = off;
  }
  __syncthreads();
  const int groups = CP / 8;
  __half* yrow = y + ((size_t)n * H2 + h2) * OW * CP;
  for (int i = threadIdx.x; i < OW * groups; i += blockDim.x) {
    const int g = i % groups, ow = i / groups;
    Vec<__half, 8> o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int off = tab[g * 8 + j];
      o.v[j] = off >= 0 ? rows[off + s * ow] : __float2half_rn(0.f);
    }
    *reinterpret_cast<Vec<__half, 8>*>(yrow + (size_t)ow * CP + g * 8) = o;
  }
}

template <typename Tin, typename Tout>
__global__ void cast_kernel(const Tin* __restrict__ x, Tout* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st_f(y + i, ld_f(x + i));
}

template <typename Tin, typename Tout>
__global__ void pack_weight_kernel(const Tin* __restrict__ w, Tout* __restrict__ out, int cout, int cin_g, int kh,
                                   int kw, int cin_pad) {
  int64_t total = (int64_t)cout * kh * kw * cin_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ci = (int)(i % cin_pad);
    int64_t t = i / cin_pad;
    int s = (int)(t % kw);
    t /= kw;
    int r = (int)(t % kh);
    int co = (int)(t / kh);
    float v = 0.f;
    if (ci < cin_g) v = ld_f(w + (((size_t)co * cin_g + ci) * kh + r) * kw + s);
    st_f(out + i, v);
  }
}

template <typename T>
__global__ void fold_affine_kernel(const T* __restrict__ bias, const T* __restrict__ bn_k, const T* __restrict__ bn_b,
                                   float* __restrict__ scale, float* __restrict__ shift, int c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float k = bn_k ? ld_f(bn_k + i) : 1.f;
  float b = bn_b ? ld_f(bn_b + i) : 0.f;
  float bi = bias ? ld_f(bias + i) : 0.f;
  scale[i] = k;
  shift[i] = bi * k + b;
}

// ---------------------------------------------------------------------------------------------
// maxpool: zero padding + -1e4 floor (planer/util.py:79-95)
// ---------------------------------------------------------------------------------------------
template <typename T, int V, int KH, int KW>
__global__ void maxpool_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int xld,
                               int xcoff, int OH, int OW, int yld, int ycoff, int kh_rt, int kw_rt, int pt, int pl,
                               int sh, int sw) {
  // KH/KW > 0: compile-time window, every tap is an unconditional (clamped) 16-byte load so that all KH*KW loads
  // are in flight together; out-of-image taps contribute 0 (the reference zero-pads), the accumulator starts at -1e4.
  const int kh = KH > 0 ? KH : kh_rt, kw = KW > 0 ? KW : kw_rt;
  const int CV = C / V;
  const int64_t total = (int64_t)N * OH * OW * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    int64_t t = i / CV;
    int ow = (int)(t % OW);
    t /= OW;
    int oh = (int)(t % OH);
    int n = (int)(t / OH);
    float m[V];
#pragma unroll
    for (int k = 0; k < V; ++k) m[k] = -1e4f;
    const T* xb = x + (size_t)n * H * W * xld + xcoff + cv * V;
    if (KH > 0) {
      Vec<T, V> v[(KH > 0 ? KH : 1) * (KW > 0 ? KW : 1)];
      bool ok[(KH > 0 ? KH : 1) * (KW > 0 ? KW : 1)];
#pragma unroll
      for (int r = 0; r < KH; ++r)
#pragma unroll
        for (int q = 0; q < KW; ++q) {
          const int ih = oh * sh + r - pt, iw = ow * sw + q - pl;
          ok[r * KW + q] = ih >= 0 && ih < H && iw >= 0 && iw < W;
          const int ihc = min(max(ih, 0), H - 1), iwc = min(max(iw, 0), W - 1);
          v[r * KW + q] = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)ihc * W + iwc) * xld);
        }
#pragma unroll
      for (int j = 0; j < KH * KW; ++j)
#pragma unroll
        for (int k = 0; k < V; ++k) m[k] = fmaxf(m[k], ok[j] ? ld_f(&v[j].v[k]) : 0.f);
    } else {
      for (int r = 0; r < kh; ++r) {
        int ih = oh * sh + r - pt;
        for (int q = 0; q < kw; ++q) {
          int iw = ow * sw + q - pl;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)ih * W + iw) * xld);
#pragma unroll
            for (int k = 0; k < V; ++k) m[k] = fmaxf(m[k], ld_f(&v.v[k]));
          } else {
#pragma unroll
            for (int k = 0; k < V; ++k) m[k] = fmaxf(m[k], 0.f);  // the reference pads with ZERO, not -inf
          }
        }
      }
    }
    Vec<T, V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) st_f(&o.v[k], m[k]);
    *reinterpret_cast<Vec<T, V>*>(y + (((size_t)n * OH + oh) * OW + ow) * yld + ycoff + cv * V) = o;
  }
}


// 3x3 / stride-2 max pooling, register blocked: a thread owns TWO adjacent output columns and a strip of RB output
// rows of one 16-byte channel group.  Every input row of the strip is loaded once (5 columns serve both outputs) and its
// two horizontal 3-maxima are folded into the running vertical maxima, the row shared by two consecutive outputs being
// used for both: 5.6 sixteen-byte loads per output instead of 9 (the generic kernel is load-issue bound, not DRAM
// bound).  Semantics are the reference's (planer/util.py:79-95): taps outside the image contribute 0, floor -1e4.
template <typename T, int V> struct VMax;
template <int V> struct VMax<float, V> {
  static __device__ __forceinline__ Vec<float, V> max2(const Vec<float, V>& a, const Vec<float, V>& b) {
    Vec<float, V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = fmaxf(a.v[k], b.v[k]);
    return o;
  }
  static __device__ __forceinline__ Vec<float, V> splat(float f) {
    Vec<float, V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = f;
    return o;
  }
};
template <int V> struct VMax<__half, V> {
  static __device__ __forceinline__ Vec<__half, V> max2(const Vec<__half, V>& a, const Vec<__half, V>& b) {
    Vec<__half, V> o;
    const __half2* pa = reinterpret_cast<const __half2*>(&a); const __half2* pb = reinterpret_cast<const __half2*>(&b);
    __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < V / 2; ++k) po[k] = __hmax2(pa[k], pb[k]);
    return o;
  }
  static __device__ __forceinline__ Vec<__half, V> splat(float f) {
    Vec<__half, V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = __float2half_rn(f);
    return o;
  }
};

template <typename T, int V, int RB>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int xld,
                                    int xcoff, int OH, int OW, int yld, int ycoff, int pt, int pl) {
  using VM = VMax<T, V>;
  const int CV = C / V, OWP = (OW + 1) / 2, OHS = (OH + RB - 1) / RB;
  const int64_t total = (int64_t)N * OHS * OWP * CV;
  const Vec<T, V> zero = VM::splat(0.f), floor_v = VM::splat(-1e4f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    int64_t t = i / CV;
    const int owp = (int)(t % OWP); t /= OWP;
    const int ohs = (int)(t % OHS);
    const int n = (int)(t / OHS);
    const int ow0 = 2 * owp, oh0 = ohs * RB;
    const int nout = min(RB, OH - oh0);
    const T* xb = x + (size_t)n * H * W * xld + xcoff + cv * V;
    T* yb = y + (size_t)n * OH * OW * yld + ycoff + cv * V;
    const int iw0 = 2 * ow0 - pl;                       // five input columns iw0 .. iw0 + 4
    bool cok[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) cok[q] = iw0 + q >= 0 && iw0 + q < W;
    Vec<T, V> m0 = floor_v, m1 = floor_v;
    for (int r = 0; r <= 2 * nout; ++r) {               // input rows of the strip: 2*oh0 - pt + r
      const int ih = 2 * oh0 - pt + r;
      Vec<T, V> h0 = zero, h1 = zero;                   // a row outside the image contributes 0 through every tap
      if (ih >= 0 && ih < H) {
        Vec<T, V> c[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int iwc = min(max(iw0 + q, 0), W - 1);
          c[q] = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)ih * W + iwc) * xld);
          if (!cok[q]) c[q] = zero;
        }
        h0 = VM::max2(VM::max2(c[0], c[1]), c[2]);
        h1 = VM::max2(VM::max2(c[2], c[3]), c[4]);
      }
      m0 = VM::max2(m0, h0); m1 = VM::max2(m1, h1);
      if (r >= 2 && (r & 1) == 0) {                     // third row of output oh0 + r/2 - 1: emit, restart with this row
        const int oh = oh0 + (r >> 1) - 1;
        *reinterpret_cast<Vec<T, V>*>(yb + ((size_t)oh * OW + ow0) * yld) = m0;
        if (ow0 + 1 < OW) *reinterpret_cast<Vec<T, V>*>(yb + ((size_t)oh * OW + ow0 + 1) * yld) = m1;
        m0 = VM::max2(floor_v, h0); m1 = VM::max2(floor_v, h1);
      }
    }
  }
}

// 3x3 / stride-2 AVERAGE pooling with the same register blocking (a thread owns two adjacent output columns and a strip of
// RB output rows; five loads per input row serve both): zero padding, divisor always 9 (planer/util.py:97-100), fp32 sums.
template <typename T, int V, int RB>
__global__ void avgpool3x3s2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int xld,
                                    int xcoff, int OH, int OW, int yld, int ycoff, int pt, int pl) {
  const int CV = C / V, OWP = (OW + 1) / 2, OHS = (OH + RB - 1) / RB;
  const int64_t total = (int64_t)N * OHS * OWP * CV;
  const float inv = 1.f / 9.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    int64_t t = i / CV;
    const int owp = (int)(t % OWP); t /= OWP;
    const int ohs = (int)(t % OHS);
    const int n = (int)(t / OHS);
    const int ow0 = 2 * owp, oh0 = ohs * RB;
    const int nout = min(RB, OH - oh0);
    const T* xb = x + (size_t)n * H * W * xld + xcoff + cv * V;
    T* yb = y + (size_t)n * OH * OW * yld + ycoff + cv * V;
    const int iw0 = 2 * ow0 - pl;
    bool cok[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) cok[q] = iw0 + q >= 0 && iw0 + q < W;
    float m0[V], m1[V];
#pragma unroll
    for (int k = 0; k < V; ++k) m0[k] = m1[k] = 0.f;
    for (int r = 0; r <= 2 * nout; ++r) {
      const int ih = 2 * oh0 - pt + r;
      float h0[V], h1[V];
#pragma unroll
      for (int k = 0; k < V; ++k) h0[k] = h1[k] = 0.f;
      if (ih >= 0 && ih < H) {
        Vec<T, V> c[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int iwc = min(max(iw0 + q, 0), W - 1);
          c[q] = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)ih * W + iwc) * xld);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float f0 = cok[0] ? ld_f(&c[0].v[k]) : 0.f, f1 = cok[1] ? ld_f(&c[1].v[k]) : 0.f, f2 = cok[2] ? ld_f(&c[2].v[k]) : 0.f;
          const float f3 = cok[3] ? ld_f(&c[3].v[k]) : 0.f, f4 = cok[4] ? ld_f(&c[4].v[k]) : 0.f;
          h0[k] = f0 + f1 + f2; h1[k] = f2 + f3 + f4;
        }
      }
#pragma unroll
      for (int k = 0; k < V; ++k) { m0[k] += h0[k]; m1[k] += h1[k]; }
      if (r >= 2 && (r & 1) == 0) {
        const int oh = oh0 + (r >> 1) - 1;
        Vec<T, V> o0, o1;
#pragma unroll
        for (int k = 0; k < V; ++k) { st_f(&o0.v[k], m0[k] * inv); st_f(&o1.v[k], m1[k] * inv); }
        *reinterpret_cast<Vec<T, V>*>(yb + ((size_t)oh * OW + ow0) * yld) = o0;
        if (ow0 + 1 < OW) *reinterpret_cast<Vec<T, V>*>(yb + ((size_t)oh * OW + ow0 + 1) * yld) = o1;
#pragma unroll
        for (int k = 0; k < V; ++k) { m0[k] = h0[k]; m1[k] = h1[k]; }
      }
    }
  }
}

template <typename T, int V>
__global__ void upsample_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int xld,
                                int xcoff, int yld, int ycoff, int fh, int fw) {
  // one thread per INPUT vector: a single 16-byte load feeds the fh x fw stores of its replicas (the output side is
  // fh*fw times larger: the kernel is store-bound, so nothing is read twice)
  const int CV = C / V, OW = W * fw;
  const int64_t total = (int64_t)N * H * W * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    int64_t t = i / CV;
    int iw = (int)(t % W);
    t /= W;
    int ih = (int)(t % H);
    int n = (int)(t / H);
    const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(x + (((size_t)n * H + ih) * W + iw) * xld + xcoff + cv * V);
    T* yb = y + (((size_t)n * H * fh + (size_t)ih * fh) * OW + (size_t)iw * fw) * yld + ycoff + cv * V;
    for (int r = 0; r < fh; ++r)
      for (int q = 0; q < fw; ++q) *reinterpret_cast<Vec<T, V>*>(yb + ((size_t)r * OW + q) * yld) = v;
  }
}

// Bilinear upsample by integer factors (planer/util.py:121-153: upsample_blinear): the image is edge-replicated by one
// pixel, every 2x2 neighbourhood of the padded image yields an fh x fw block of outputs as a 4-tap blend with weights
// wmat[tap][a*fw + b] (the reference's make_upmat, computed in fp16 on the host exactly as it does and passed as fp32),
// and the result is cropped by (fh/2, fw/2).  One thread per output vector; the 4 taps hit L1/L2.
template <typename T, int V>
__global__ void upsample_linear_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ wmat, int N, int H,
                                       int W, int C, int xld, int xcoff, int yld, int ycoff, int fh, int fw) {
  // one thread per CELL of the edge-replicated image (2x2 neighbourhood) and 16-byte channel vector: its four corner loads
  // feed the fh x fw outputs of the cell (one thread per OUTPUT vector re-loaded the corners fh*fw times and was bound by
  // load issue at 0.27 of the copy bandwidth, tools/hbm_bench.py)
  const int CV = C / V, OH = H * fh, OW = W * fw, kk = fh * fw, HC = H + 1, WC = W + 1;
  // blockIdx.y walks cell rows (n, cy), x the (cell column, channel vector) pairs of a row
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= WC * CV) return;
  const int cx = t / CV, cv = t - cx * CV;
  for (int64_t row = blockIdx.y; row < (int64_t)N * HC; row += gridDim.y) {
    const int n = (int)(row / HC), cy = (int)(row - (int64_t)n * HC);
    const int r0 = max(cy - 1, 0), r1 = min(cy, H - 1), c0 = max(cx - 1, 0), c1 = min(cx, W - 1);
    const T* xb = x + (size_t)n * H * W * xld + xcoff + cv * V;
    const Vec<T, V> p00 = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)r0 * W + c0) * xld);
    const Vec<T, V> p01 = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)r0 * W + c1) * xld);
    const Vec<T, V> p10 = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)r1 * W + c0) * xld);
    const Vec<T, V> p11 = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)r1 * W + c1) * xld);
    float f00[V], f01[V], f10[V], f11[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { f00[k] = ld_f(&p00.v[k]); f01[k] = ld_f(&p01.v[k]); f10[k] = ld_f(&p10.v[k]); f11[k] = ld_f(&p11.v[k]); }
    for (int a = 0; a < fh; ++a) {
      const int oy = cy * fh + a - fh / 2;                 // the crop of planer/util.py:153
      if (oy < 0 || oy >= OH) continue;
      for (int b = 0; b < fw; ++b) {
        const int ox = cx * fw + b - fw / 2;
        if (ox < 0 || ox >= OW) continue;
        const int wi = a * fw + b;
        const float w0 = wmat[wi], w1 = wmat[kk + wi], w2 = wmat[2 * kk + wi], w3 = wmat[3 * kk + wi];
        Vec<T, V> o;
#pragma unroll
        for (int k = 0; k < V; ++k) st_f(&o.v[k], fmaf(f11[k], w3, fmaf(f10[k], w2, fmaf(f01[k], w1, f00[k] * w0))));
        *reinterpret_cast<Vec<T, V>*>(y + (((size_t)n * OH + oy) * OW + ox) * yld + ycoff + cv * V) = o;
      }
    }
  }
}

// Bilinear resize to an arbitrary size (planer/util.py:194-210, upsample_size): separable gather + lerp, columns first.  The
// source indices and weights of every output row / column come from the host, computed in the image dtype exactly as the
// reference does (a float16 image gets float16 coordinates there); tables: lo index, weight of lo + 1, one minus that weight.
template <typename T, int V>
__global__ void resize_linear_kernel(const T* __restrict__ x, T* __restrict__ y, const int* __restrict__ rlo, const float* __restrict__ rw,
                                     const float* __restrict__ rw1, const int* __restrict__ clo, const float* __restrict__ cw,
                                     const float* __restrict__ cw1, int N, int H, int W, int C, int xld, int xcoff, int OH, int OW,
                                     int yld, int ycoff) {
  const int CV = C / V;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= OW * CV) return;
  const int ox = t / CV, cv = t - ox * CV;
  const int c0 = clo[ox], c1 = min(c0 + 1, W - 1);
  const float fc = cw[ox], gc = cw1[ox];
  for (int64_t row = blockIdx.y; row < (int64_t)N * OH; row += gridDim.y) {
    const int n = (int)(row / OH), oy = (int)(row - (int64_t)n * OH);
    const int r0 = rlo[oy], r1 = min(r0 + 1, H - 1);
    const float fr = rw[oy], gr = rw1[oy];
    const T* xb = x + (size_t)n * H * W * xld + xcoff + cv * V;
    const Vec<T, V> p00 = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)r0 * W + c0) * xld);
    const Vec<T, V> p01 = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)r0 * W + c1) * xld);
    const Vec<T, V> p10 = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)r1 * W + c0) * xld);
    const Vec<T, V> p11 = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)r1 * W + c1) * xld);
    Vec<T, V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      T top, bot;                                            // the column pass is rounded to the image dtype (buf in util.py:208)
      st_f(&top, ld_f(&p00.v[k]) * gc + ld_f(&p01.v[k]) * fc);
      st_f(&bot, ld_f(&p10.v[k]) * gc + ld_f(&p11.v[k]) * fc);
      st_f(&o.v[k], ld_f(&top) * gr + ld_f(&bot) * fr);
    }
    *reinterpret_cast<Vec<T, V>*>(y + (((size_t)n * OH + oy) * OW + ox) * yld + ycoff + cv * V) = o;
  }
}

template <typename T, int V>
__global__ void copy_channels_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t npix, int C, int xld,
                                     int xcoff, int yld, int ycoff) {
  const int CV = C / V;
  const int64_t total = npix * CV;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    int64_t p = i / CV;
    *reinterpret_cast<Vec<T, V>*>(y + p * yld + ycoff + cv * V) =
        *reinterpret_cast<const Vec<T, V>*>(x + p * xld + xcoff + cv * V);
  }
}

// sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5 on fp16 pairs: ONE MUFU op per two elements (tanh.approx.f16x2, abs. error
// 2^-11) instead of ex2 + rcp per element -- with expf the fp16 sigmoid kernel is MUFU-bound at 66 % of the HBM rate
__device__ __forceinline__ uint32_t sigmoid_h2(uint32_t v) {
  const __half2 half = __float2half2_rn(0.5f);
  __half2 h = __hmul2(*reinterpret_cast<__half2*>(&v), half);
  uint32_t t, hu = *reinterpret_cast<uint32_t*>(&h);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(hu));
  __half2 r = __hfma2(*reinterpret_cast<__half2*>(&t), half, half);
  return *reinterpret_cast<uint32_t*>(&r);
}

template <typename T, int V>
__global__ void eltwise_kernel(int op, const T* __restrict__ x, const T* __restrict__ p0, const T* __restrict__ p1,
                               T* __restrict__ y, int64_t total_vec, int C, float alpha) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const Vec<T, V> a = *reinterpret_cast<const Vec<T, V>*>(x + i * V);
    Vec<T, V> o;
    if (op == PLNR_EW_ADD) {
      const Vec<T, V> b = *reinterpret_cast<const Vec<T, V>*>(p0 + i * V);
#pragma unroll
      for (int k = 0; k < V; ++k) st_f(&o.v[k], ld_f(&a.v[k]) + ld_f(&b.v[k]));
    } else if (op == PLNR_EW_SCALE_SHIFT) {
      int c = (int)((i * V) % C);
#pragma unroll
      for (int k = 0; k < V; ++k) st_f(&o.v[k], ld_f(&a.v[k]) * ld_f(p0 + c + k) + ld_f(p1 + c + k));
    } else if (op == PLNR_EW_SIGMOID && sizeof(T) == 2 && V == 8) {
      const uint4 in = *reinterpret_cast<const uint4*>(&a);
      uint4 out;
      out.x = sigmoid_h2(in.x); out.y = sigmoid_h2(in.y); out.z = sigmoid_h2(in.z); out.w = sigmoid_h2(in.w);
      *reinterpret_cast<uint4*>(&o) = out;
    } else {
      int act = op == PLNR_EW_RELU ? PLNR_ACT_RELU : (op == PLNR_EW_LEAKY ? PLNR_ACT_LEAKY : PLNR_ACT_SIGMOID);
#pragma unroll
      for (int k = 0; k < V; ++k) st_f(&o.v[k], plnr_apply_act(ld_f(&a.v[k]), act, alpha));
    }
    *reinterpret_cast<Vec<T, V>*>(y + i * V) = o;
  }
}

// Unary operators with two scalar parameters: CLIP y = max(min(x, b), a) (planer/layer.py:247-251),
// HARDSIGMOID y = max(min(x*a + b, 1), 0) (planer/layer.py:66-69).  fp16 data is computed in fp32 and rounded once.
template <typename T, int V>
__global__ void unary2_kernel(int op, const T* __restrict__ x, T* __restrict__ y, int64_t total_vec, float a, float b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(x + i * V);
    Vec<T, V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float f = ld_f(&v.v[k]);
      st_f(&o.v[k], op == PLNR_EW_CLIP ? fmaxf(fminf(f, b), a) : fmaxf(fminf(fmaf(f, a, b), 1.f), 0.f));
    }
    *reinterpret_cast<Vec<T, V>*>(y + i * V) = o;
  }
}

// Softmax over the last, contiguous axis of a dense (rows, c) array (planer/layer.py:141-146: y = x - max; e = exp(y);
// y - log(sum e); exp) -- channel softmax of pixel-major activations, or the class axis of 2-D logits.  One warp per row.
template <typename T>
__global__ void softmax_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int c) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const T* xr = x + r * c;
    float m = -3.0e38f;
    for (int i = lane; i < c; i += 32) m = fmaxf(m, ld_f(xr + i));
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    float sum = 0.f;
    for (int i = lane; i < c; i += 32) sum += __expf(ld_f(xr + i) - m);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const float lg = __logf(sum);
    for (int i = lane; i < c; i += 32) st_f(y + r * c + i, __expf(ld_f(xr + i) - m - lg));
  }
}

// Softmax of short rows (c = G * V elements, G a power of two <= 32): G adjacent lanes hold one row in registers -- one
// 16-byte load and one 16-byte store per lane, max and sum by xor-shuffles inside the lane group; 32 / G rows per warp.
// (The warp-per-row kernel above reads a 64-channel row three times through 2-byte loads: 0.19 of the copy bandwidth.)
template <typename T, int V>
__global__ void softmax_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int G) {
  const int64_t total = rows * G;
  const int64_t n_iter = (total + (int64_t)gridDim.x * blockDim.x - 1) / ((int64_t)gridDim.x * blockDim.x);
  for (int64_t it = 0; it < n_iter; ++it) {                 // whole warps stay in the loop: the shuffles are warp-wide
    const int64_t i = it * (int64_t)gridDim.x * blockDim.x + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool ok = i < total;
    float f[V];
    if (ok) {
      const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(x + i * V);
#pragma unroll
      for (int k = 0; k < V; ++k) f[k] = ld_f(&v.v[k]);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) f[k] = 0.f;
    }
    float m = f[0];
#pragma unroll
    for (int k = 1; k < V; ++k) m = fmaxf(m, f[k]);
    for (int d = 1; d < G; d <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) sum += __expf(f[k] - m);
    for (int d = 1; d < G; d <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const float lg = __logf(sum);
    if (ok) {
      Vec<T, V> o;
#pragma unroll
      for (int k = 0; k < V; ++k) st_f(&o.v[k], __expf(f[k] - m - lg));
      *reinterpret_cast<Vec<T, V>*>(y + i * V) = o;
    }
  }
}

template <typename T, int V>
__global__ void gap_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int HW, int C, int xld, int xcoff) {
  const int CV = C / V;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * CV) return;
  int cv = i % CV, n = i / CV;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  for (int p = 0; p < HW; ++p) {
    const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(x + ((size_t)n * HW + p) * xld + xcoff + cv * V);
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += ld_f(&v.v[k]);
  }
  Vec<T, V> o;
  const float inv = 1.f / (float)HW;
#pragma unroll
  for (int k = 0; k < V; ++k) st_f(&o.v[k], acc[k] * inv);
  *reinterpret_cast<Vec<T, V>*>(y + (size_t)n * C + cv * V) = o;
}


// ---------------------------------------------------------------------------------------------
// averagepool: zero padding, the divisor is ALWAYS kh*kw (planer/util.py:97-100: pool(np.add) then rst /= c[0]*c[1])
// ---------------------------------------------------------------------------------------------
template <typename T, int V, int KH, int KW>
__global__ void avgpool_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int xld,
                               int xcoff, int OH, int OW, int yld, int ycoff, int kh_rt, int kw_rt, int pt, int pl, int sh, int sw) {
  // blockIdx.y walks output rows (n, oh), blockIdx.x * blockDim.x + threadIdx.x the (ow, channel vector) pairs of a row: no
  // 64-bit divisions per thread; KH/KW > 0: compile-time window, every tap an unconditional clamped load, all in flight
  const int kh = KH > 0 ? KH : kh_rt, kw = KW > 0 ? KW : kw_rt;
  const int CV = C / V;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= OW * CV) return;
  const int ow = t / CV, cv = t - ow * CV;
  const float inv = 1.f / (float)(kh * kw);
  for (int64_t row = blockIdx.y; row < (int64_t)N * OH; row += gridDim.y) {
    const int n = (int)(row / OH), oh = (int)(row - (int64_t)n * OH);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    const T* xb = x + (size_t)n * H * W * xld + xcoff + cv * V;
    if (KH > 0) {
      Vec<T, V> v[(KH > 0 ? KH : 1) * (KW > 0 ? KW : 1)];
      bool ok[(KH > 0 ? KH : 1) * (KW > 0 ? KW : 1)];
#pragma unroll
      for (int r = 0; r < KH; ++r)
#pragma unroll
        for (int q = 0; q < KW; ++q) {
          const int ih = oh * sh + r - pt, iw = ow * sw + q - pl;
          ok[r * KW + q] = ih >= 0 && ih < H && iw >= 0 && iw < W;
          const int ihc = min(max(ih, 0), H - 1), iwc = min(max(iw, 0), W - 1);
          v[r * KW + q] = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)ihc * W + iwc) * xld);
        }
#pragma unroll
      for (int j = 0; j < KH * KW; ++j)
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += ok[j] ? ld_f(&v[j].v[k]) : 0.f;
    } else {
      for (int r = 0; r < kh; ++r) {
        const int ih = oh * sh + r - pt;
        if (ih < 0 || ih >= H) continue;
        for (int q = 0; q < kw; ++q) {
          const int iw = ow * sw + q - pl;
          if (iw < 0 || iw >= W) continue;
          const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(xb + ((size_t)ih * W + iw) * xld);
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k] += ld_f(&v.v[k]);
        }
      }
    }
    Vec<T, V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) st_f(&o.v[k], acc[k] * inv);
    *reinterpret_cast<Vec<T, V>*>(y + (((size_t)n * OH + oh) * OW + ow) * yld + ycoff + cv * V) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose2d (planer/layer.py:28-34) = zero-stuffing + an ordinary stride-1 convolution with the filter transposed
// (in <-> out channels) and flipped in both spatial axes.  zero_stuff: y[n, lo_h + i*sh, lo_w + j*sw, :] = x[n, i, j, :],
// zero elsewhere (one pass: every output pixel is written once).  flip_weight: K (ci, co, kh, kw) -> (co, ci, kh, kw)
// with K'[o, i, r, s] = K[i, o, kh-1-r, kw-1-s], done once when the executor is built.
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void zero_stuff_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int xld, int xcoff,
                                  int OH, int OW, int yld, int ycoff, int lo_h, int lo_w, int sh, int sw) {
  // blockIdx.y walks output rows (n, oh): whether a row carries input pixels at all is decided once per row, and a thread
  // needs one small division (by the channel-vector count) instead of three 64-bit ones
  const int CV = C / V;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= OW * CV) return;
  const int ow = t / CV, cv = t - ow * CV;
  const int b = ow - lo_w;
  const bool col_hit = b >= 0 && b % sw == 0 && b / sw < W;
  const int iw = col_hit ? b / sw : 0;
  for (int64_t row = blockIdx.y; row < (int64_t)N * OH; row += gridDim.y) {
    const int n = (int)(row / OH), oh = (int)(row - (int64_t)n * OH);
    const int a = oh - lo_h;
    Vec<T, V> v;
#pragma unroll
    for (int k = 0; k < V; ++k) st_f(&v.v[k], 0.f);
    if (col_hit && a >= 0 && a % sh == 0 && a / sh < H)
      v = *reinterpret_cast<const Vec<T, V>*>(x + (((size_t)n * H + a / sh) * W + iw) * xld + xcoff + cv * V);
    *reinterpret_cast<Vec<T, V>*>(y + (((size_t)n * OH + oh) * OW + ow) * yld + ycoff + cv * V) = v;
  }
}

template <typename T>
__global__ void flip_weight_kernel(const T* __restrict__ w, T* __restrict__ out, int ci, int co, int kh, int kw) {
  const int64_t total = (int64_t)ci * co * kh * kw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int s = (int)(i % kw);
    int64_t t = i / kw;
    int r = (int)(t % kh);
    t /= kh;
    int c = (int)(t % ci);
    int o = (int)(t / ci);
    out[i] = w[(((size_t)c * co + o) * kh + (kh - 1 - r)) * kw + (kw - 1 - s)];
  }
}

// GlobalAveragePool -> Flatten -> Dense in one launch (planer/layer.py:77-78, :59, :15-18).  A CTA takes IMGS images and a
// slice of the output features: (1) the pooled vectors (fp32) of its images go to shared memory: SPLIT adjacent lanes
// share one 16-byte channel group of one image and sum interleaved pixel rows, with all PB row loads of a lane in flight
// at once; (2) each warp walks output features of the slice, OPW weight rows x two 16-byte pieces per lane in flight,
// IMGS dot products at a time, then a shuffle reduction and the fused scale/shift/activation.
// The tail of a network is ~65 MFLOP on 6 MB of input: neither FLOPs nor bandwidth but the NUMBER OF DEPENDENT LOAD
// ROUNDS bounds it (ncu: the first version spent its time in long-scoreboard stalls behind 4 pixel rounds + 6 weight
// rounds per CTA, 23 us) -- here 1-2 + 2 rounds.
template <typename T, int V, int IMGS, int SPLIT, int PB, int OPW>
__global__ void __launch_bounds__(512) gap_dense_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         T* __restrict__ y, int N, int HW, int C, int xld, int xcoff,
                                                         int OUT, int och, int act, float alpha) {
  extern __shared__ float pooled[];                 // [IMGS][C]
  const int n0 = blockIdx.x * IMGS, o0 = blockIdx.y * och;
  const int CV = C / V;
  const float inv = 1.f / (float)HW;
  const int items = IMGS * CV * SPLIT;
  for (int idx = threadIdx.x; idx < ((items + 31) & ~31); idx += blockDim.x) {      // whole warps: shuffles below
    const int item = idx / SPLIT, part = idx - item * SPLIT;
    const int img = item / CV, cv = item - img * CV;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    if (idx < items && n0 + img < N) {
      const T* xp = x + (size_t)(n0 + img) * HW * xld + xcoff + cv * V;
      for (int p0 = part; p0 < HW; p0 += PB * SPLIT) {
        Vec<T, V> v[PB];
#pragma unroll
        for (int j = 0; j < PB; ++j)
          if (p0 + j * SPLIT < HW) v[j] = *reinterpret_cast<const Vec<T, V>*>(xp + (size_t)(p0 + j * SPLIT) * xld);
#pragma unroll
        for (int j = 0; j < PB; ++j)
          if (p0 + j * SPLIT < HW) {
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] += ld_f(&v[j].v[k]);
          }
      }
    }
    // the SPLIT partial sums sit in adjacent lanes (SPLIT divides 32 and the loop bound is a multiple of SPLIT)
#pragma unroll
    for (int d = 1; d < SPLIT; d <<= 1) {
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
    }
    if (part == 0 && idx < items) {
#pragma unroll
      for (int k = 0; k < V; ++k) pooled[img * C + cv * V + k] = acc[k] * inv;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int o_end = min(o0 + och, OUT);
  for (int ob = o0 + warp * OPW; ob < o_end; ob += nwarps * OPW) {
    float acc[OPW][IMGS];
#pragma unroll
    for (int j = 0; j < OPW; ++j)
#pragma unroll
      for (int i = 0; i < IMGS; ++i) acc[j][i] = 0.f;
    for (int cv0 = lane; cv0 < CV; cv0 += 64) {
      Vec<T, V> wv[2][OPW];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < OPW; ++j)
          if (cv0 + 32 * u < CV)
            wv[u][j] = *reinterpret_cast<const Vec<T, V>*>(w + (size_t)min(ob + j, OUT - 1) * C + (cv0 + 32 * u) * V);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int cv = cv0 + 32 * u;
        if (cv < CV) {
#pragma unroll
          for (int i = 0; i < IMGS; ++i) {
            // 16-byte shared-memory reads: scalar reads at this 4*V-byte lane stride would be V-way bank conflicts
            float pf[V];
#pragma unroll
            for (int k = 0; k < V; k += 4)
              *reinterpret_cast<float4*>(&pf[k]) = *reinterpret_cast<const float4*>(pooled + i * C + cv * V + k);
#pragma unroll
            for (int j = 0; j < OPW; ++j)
#pragma unroll
              for (int k = 0; k < V; ++k) acc[j][i] = fmaf(ld_f(&wv[u][j].v[k]), pf[k], acc[j][i]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < OPW; ++j) {
#pragma unroll
      for (int i = 0; i < IMGS; ++i) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc[j][i] += __shfl_xor_sync(0xffffffffu, acc[j][i], d);
      }
      const int o = ob + j;
      if (o < o_end && lane < IMGS && n0 + lane < N) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < IMGS; ++i) v = lane == i ? acc[j][i] : v;
        v = fmaf(v, scale ? scale[o] : 1.f, shift ? shift[o] : 0.f);
        st_f(y + (size_t)(n0 + lane) * OUT + o, plnr_apply_act(v, act, alpha));
      }
    }
  }
}

// The same tail with the CTA's slice of the weight matrix BULK-COPIED to shared memory (one cp.async.bulk of och*C elements,
// issued before the pooling phase and waited for after it): the dense phase then has no global-memory round trips at all.
// ncu on the kernel above: its time is the number of dependent load rounds (pixels, then weights), not bytes or FLOPs.
template <typename T, int V, int IMGS, int SPLIT, int PB, int OPW>
__global__ void __launch_bounds__(512) gap_dense_smem_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                              const float* __restrict__ scale, const float* __restrict__ shift,
                                                              T* __restrict__ y, int N, int HW, int C, int xld, int xcoff,
                                                              int OUT, int och, int act, float alpha) {
  extern __shared__ __align__(128) uint8_t gd_smem[];
  float* pooled = reinterpret_cast<float*>(gd_smem);                                   // [IMGS][C]
  T* wsm = reinterpret_cast<T*>(gd_smem + (size_t)IMGS * C * sizeof(float));           // [och][C]
  const uint32_t bar = ptx::smem_u32(gd_smem + (size_t)IMGS * C * sizeof(float) + (size_t)och * C * sizeof(T));
  const int n0 = blockIdx.x * IMGS, o0 = blockIdx.y * och;
  const int o_end = min(o0 + och, OUT);
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)((size_t)(o_end - o0) * C * sizeof(T));
    ptx::mbar_arrive_expect_tx(bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ptx::smem_u32(wsm)), "l"(reinterpret_cast<uint64_t>(w + (size_t)o0 * C)), "r"(bytes), "r"(bar) : "memory");
  }
  const int CV = C / V;
  const float inv = 1.f / (float)HW;
  const int items = IMGS * CV * SPLIT;
  for (int idx = threadIdx.x; idx < ((items + 31) & ~31); idx += blockDim.x) {      // whole warps: shuffles below
    const int item = idx / SPLIT, part = idx - item * SPLIT;
    const int img = item / CV, cv = item - img * CV;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    if (idx < items && n0 + img < N) {
      const T* xp = x + (size_t)(n0 + img) * HW * xld + xcoff + cv * V;
      for (int p0 = part; p0 < HW; p0 += PB * SPLIT) {
        Vec<T, V> v[PB];
#pragma unroll
        for (int j = 0; j < PB; ++j)
          if (p0 + j * SPLIT < HW) v[j] = *reinterpret_cast<const Vec<T, V>*>(xp + (size_t)(p0 + j * SPLIT) * xld);
#pragma unroll
        for (int j = 0; j < PB; ++j)
          if (p0 + j * SPLIT < HW) {
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] += ld_f(&v[j].v[k]);
          }
      }
    }
#pragma unroll
    for (int d = 1; d < SPLIT; d <<= 1) {
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
    }
    if (part == 0 && idx < items) {
#pragma unroll
      for (int k = 0; k < V; ++k) pooled[img * C + cv * V + k] = acc[k] * inv;
    }
  }
  __syncthreads();
  while (!ptx::mbar_try_wait(bar, 0)) {}
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int ob = o0 + warp * OPW; ob < o_end; ob += nwarps * OPW) {
    float acc[OPW][IMGS];
#pragma unroll
    for (int j = 0; j < OPW; ++j)
#pragma unroll
      for (int i = 0; i < IMGS; ++i) acc[j][i] = 0.f;
    for (int cv = lane; cv < CV; cv += 32) {
      Vec<T, V> wv[OPW];
#pragma unroll
      for (int j = 0; j < OPW; ++j)
        wv[j] = *reinterpret_cast<const Vec<T, V>*>(wsm + (size_t)(min(ob + j, o_end - 1) - o0) * C + cv * V);
#pragma unroll
      for (int i = 0; i < IMGS; ++i) {
        float pf[V];
#pragma unroll
        for (int k = 0; k < V; k += 4)
          *reinterpret_cast<float4*>(&pf[k]) = *reinterpret_cast<const float4*>(pooled + i * C + cv * V + k);
#pragma unroll
        for (int j = 0; j < OPW; ++j)
#pragma unroll
          for (int k = 0; k < V; ++k) acc[j][i] = fmaf(ld_f(&wv[j].v[k]), pf[k], acc[j][i]);
      }
    }
#pragma unroll
    for (int j = 0; j < OPW; ++j) {
#pragma unroll
      for (int i = 0; i < IMGS; ++i) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc[j][i] += __shfl_xor_sync(0xffffffffu, acc[j][i], d);
      }
      const int o = ob + j;
      if (o < o_end && lane < IMGS && n0 + lane < N) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < IMGS; ++i) v = lane == i ? acc[j][i] : v;
        v = fmaf(v, scale ? scale[o] : 1.f, shift ? shift[o] : 0.f);
        st_f(y + (size_t)(n0 + lane) * OUT + o, plnr_apply_act(v, act, alpha));
      }
    }
  }
}

// fp16 tail on the (legacy) warp-level tensor-core instruction: the CUDA-core versions above execute 8.6 M warp instructions
// (ncu: issue slots 52 % busy -- 2 M FMA, 0.5 M half->float conversions of the weights, 1.3 M shuffle-reduction steps), i.e.
// they are INSTRUCTION-bound at ~20 us for 65 MFLOP.  Here a CTA pools 8 images to fp16 (the reference's GlobalAveragePool
// returns an fp16 array too) and each warp computes a 16-feature x 8-image block with mma.sync.m16n8k16 (fp32 accumulate):
// D[f, n] = sum_k W[f, k] * pooled[n, k].  A fragments come straight from the (L2-resident, row-major) weight matrix, B
// fragments from the pooled vectors in shared memory (row pitch C + 8 halves: conflict-free).  tcgen05 would need a
// 128-row tile for an 8-column problem; this op is 0.014 % of the step's FLOPs.
template <int IMGS>
__global__ void __launch_bounds__(512) gap_dense_mma_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             __half* __restrict__ y, int N, int HW, int C, int xld, int xcoff,
                                                             int OUT, int och, int act, float alpha) {
  static_assert(IMGS == 8, "the n8 dimension of mma.m16n8k16 is the image group");
  extern __shared__ __align__(16) uint8_t gm_smem[];
  __half* pooled = reinterpret_cast<__half*>(gm_smem);          // [8][C + 8]
  const int pitch = C + 8;
  const int n0 = blockIdx.x * IMGS, o0 = blockIdx.y * och;
  const int o_end = min(o0 + och, OUT);
  constexpr int V = 8, SPLIT = 2, PB = 13;
  const int CV = C / V;
  const float inv = 1.f / (float)HW;
  const int items = IMGS * CV * SPLIT;
  for (int idx = threadIdx.x; idx < ((items + 31) & ~31); idx += blockDim.x) {
    const int item = idx / SPLIT, part = idx - item * SPLIT;
    const int img = item / CV, cv = item - img * CV;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    if (idx < items && n0 + img < N) {
      const __half* xp = x + (size_t)(n0 + img) * HW * xld + xcoff + cv * V;
      for (int p0 = part; p0 < HW; p0 += PB * SPLIT) {
        Vec<__half, V> v[PB];
#pragma unroll
        for (int j = 0; j < PB; ++j)
          if (p0 + j * SPLIT < HW) v[j] = *reinterpret_cast<const Vec<__half, V>*>(xp + (size_t)(p0 + j * SPLIT) * xld);
#pragma unroll
        for (int j = 0; j < PB; ++j)
          if (p0 + j * SPLIT < HW) {
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] += ld_f(&v[j].v[k]);
          }
      }
    }
#pragma unroll
    for (int d = 1; d < SPLIT; d <<= 1) {
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
    }
    if (part == 0 && idx < items) {
      Vec<__half, V> o;
#pragma unroll
      for (int k = 0; k < V; ++k) o.v[k] = __float2half_rn(acc[k] * inv);
      *reinterpret_cast<Vec<__half, V>*>(pooled + img * pitch + cv * V) = o;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;               // fragment coordinates of mma.m16n8k16
  for (int m0 = o0 + warp * 16; m0 < o_end; m0 += nwarps * 16) {
    const int r0 = min(m0 + gid, OUT - 1), r1 = min(m0 + gid + 8, OUT - 1);
    const uint32_t* w0 = reinterpret_cast<const uint32_t*>(w + (size_t)r0 * C) + tig;
    const uint32_t* w1 = reinterpret_cast<const uint32_t*>(w + (size_t)r1 * C) + tig;
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(pooled + gid * pitch) + tig;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    constexpr int U = 8;                                   // k-steps of 16 whose loads are in flight together
    for (int k0 = 0; k0 < C; k0 += 16 * U) {
      uint32_t a[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kw = (k0 + 16 * u) >> 1;                 // 32-bit word index of column k
        if (k0 + 16 * u < C) {
          a[u][0] = __ldg(w0 + kw); a[u][1] = __ldg(w1 + kw); a[u][2] = __ldg(w0 + kw + 4); a[u][3] = __ldg(w1 + kw + 4);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k0 + 16 * u < C) {
          const int kw = (k0 + 16 * u) >> 1;
          const uint32_t b0 = pb[kw], b1 = pb[kw + 4];
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
                       : "r"(a[u][0]), "r"(a[u][1]), "r"(a[u][2]), "r"(a[u][3]), "r"(b0), "r"(b1));
        }
      }
    }
    // D fragment: (feature m0 + gid, images 2*tig, 2*tig + 1) and (feature m0 + gid + 8, same images)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int f = m0 + gid + 8 * h;
      if (f < o_end) {
        const float sc = scale ? scale[f] : 1.f, sf = shift ? shift[f] : 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = n0 + 2 * tig + j;
          if (n < N) y[(size_t)n * OUT + f] = __float2half_rn(plnr_apply_act(fmaf(h ? (j ? d3 : d2) : (j ? d1 : d0), sc, sf), act, alpha));
        }
      }
    }
  }
}

// The same tail when GlobalAveragePool was folded into the producing convolution's epilogue (plnr_epilogue.pool_sum,
// conv_shift.cu): `pre` holds fp32 partial sums [image][part][channel].  Phase 1 adds the parts, divides by the number of pooled
// positions and rounds to fp16 (the reference's GlobalAveragePool returns an fp16 array: planer/layer.py:77-78) -- 16 KB of
// reads per CTA instead of re-pooling 400 KB of activations.  Phase 2 is gap_dense_mma_kernel's product with the K range of a
// 16-feature block split over two warps when there are warps to spare (all loads of a warp in flight at once: the phase is
// a chain of L2 latencies, not of bytes), partial accumulators added through shared memory in a fixed order.
template <int IMGS>
__global__ void __launch_bounds__(512) pooled_dense_mma_kernel(const float* __restrict__ pre, const __half* __restrict__ w,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                __half* __restrict__ y, int N, int P, float inv, int C, int OUT,
                                                                int och, int act, float alpha) {
  static_assert(IMGS == 8, "the n8 dimension of mma.m16n8k16 is the image group");
  extern __shared__ __align__(16) uint8_t pd_smem[];
  __half* pooled = reinterpret_cast<__half*>(pd_smem);          // [8][C + 8]
  const int pitch = C + 8;
  float* red = reinterpret_cast<float*>(pd_smem + (((size_t)IMGS * pitch * sizeof(__half)) + 15) / 16 * 16);   // [warp][4][32]
  const int n0 = blockIdx.x * IMGS, o0 = blockIdx.y * och;
  const int o_end = min(o0 + och, OUT);
  const int C4 = C >> 2;
  for (int idx = threadIdx.x; idx < IMGS * C4; idx += blockDim.x) {
    const int img = idx / C4, c4 = idx - img * C4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + img < N) {
      const float4* src = reinterpret_cast<const float4*>(pre + (size_t)(n0 + img) * P * C) + c4;
      for (int pp = 0; pp < P; ++pp) {
        const float4 v = __ldcg(src + (size_t)pp * C4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    __half2* dst = reinterpret_cast<__half2*>(pooled + img * pitch + c4 * 4);
    dst[0] = __floats2half2_rn(acc.x * inv, acc.y * inv);
    dst[1] = __floats2half2_rn(acc.z * inv, acc.w * inv);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;               // fragment coordinates of mma.m16n8k16
  const int blocks = (o_end - o0 + 15) >> 4;
  const int ks = (nwarps >= 2 * blocks && C % 32 == 0) ? 2 : 1;      // K halves per 16-feature block
  const int Kper = C / ks;
  for (int base = 0; base < blocks * ks; base += nwarps) {
    const int wi = base + warp;
    const bool active = wi < blocks * ks;
    const int blk = wi / ks, kh = wi - blk * ks;
    const int m0 = o0 + blk * 16;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    if (active) {
      const int r0 = min(m0 + gid, OUT - 1), r1 = min(m0 + gid + 8, OUT - 1);
      const uint32_t* w0 = reinterpret_cast<const uint32_t*>(w + (size_t)r0 * C + kh * Kper) + tig;
      const uint32_t* w1 = reinterpret_cast<const uint32_t*>(w + (size_t)r1 * C + kh * Kper) + tig;
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(pooled + gid * pitch + kh * Kper) + tig;
      constexpr int U = 16;                                 // k-steps of 16 whose loads are in flight together
      for (int k0 = 0; k0 < Kper; k0 += 16 * U) {
        uint32_t a[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int kw = (k0 + 16 * u) >> 1;                // 32-bit word index of column k
          if (k0 + 16 * u < Kper) {
            a[u][0] = __ldg(w0 + kw); a[u][1] = __ldg(w1 + kw); a[u][2] = __ldg(w0 + kw + 4); a[u][3] = __ldg(w1 + kw + 4);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (k0 + 16 * u < Kper) {
            const int kw = (k0 + 16 * u) >> 1;
            const uint32_t b0 = pb[kw], b1 = pb[kw + 4];
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
                         : "r"(a[u][0]), "r"(a[u][1]), "r"(a[u][2]), "r"(a[u][3]), "r"(b0), "r"(b1));
          }
        }
      }
    }
    if (ks == 2) {                                          // upper K half -> shared memory -> added by the lower half's warp
      if (active && kh == 1) {
        float* r = red + warp * 128;
        r[lane] = d0; r[32 + lane] = d1; r[64 + lane] = d2; r[96 + lane] = d3;
      }
      __syncthreads();
      if (active && kh == 0) {
        const float* r = red + (warp + 1) * 128;
        d0 += r[lane]; d1 += r[32 + lane]; d2 += r[64 + lane]; d3 += r[96 + lane];
      }
      __syncthreads();
    }
    if (active && kh == 0) {
      // D fragment: (feature m0 + gid, images 2*tig, 2*tig + 1) and (feature m0 + gid + 8, same images)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int f = m0 + gid + 8 * h;
        if (f < o_end) {
          const float sc = scale ? scale[f] : 1.f, sf = shift ? shift[f] : 0.f;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int n = n0 + 2 * tig + j;
            if (n < N) y[(size_t)n * OUT + f] = __float2half_rn(plnr_apply_act(fmaf(h ? (j ? d3 : d2) : (j ? d1 : d0), sc, sf), act, alpha));
          }
        }
      }
    }
  }
}

static inline bool view_vec_ok(const plnr_tensor* t, int V, size_t esz) {
  return t->c % V == 0 && t->ld % V == 0 && t->coff % V == 0 && aligned16(t->ptr) && (V * esz == 16);
}

#define DISPATCH_T(dt, ...)                                  \
  if ((dt) == PLNR_F16) { using T = __half; __VA_ARGS__ }    \
  else { using T = float; __VA_ARGS__ }

}  // namespace

extern "C" {

int plnr_nchw_to_nhwc(plnr_ctx* ctx, const void* x, int x_dtype, int c_src, const plnr_tensor* y, int y_dtype) {
  PLNR_REQUIRE(ctx && x && y && y->ptr, "nchw_to_nhwc: NULL argument");
  PLNR_REQUIRE(c_src >= 1 && c_src <= y->c && y->ld >= y->coff + y->c, "nchw_to_nhwc: bad channel counts");
  const int HW = y->h * y->w;
  dim3 grid((HW + 31) / 32, (y->c + 31) / 32, y->n), block(32, 8);
  PLNR_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "nchw_to_nhwc: batch/channel count too large for one launch");
#define LAUNCH(TI, TO)                                                                                       \
  nchw_to_nhwc_kernel<TI, TO><<<grid, block, 0, ctx->stream>>>((const TI*)x, (TO*)y->ptr, c_src, HW, y->c, y->ld, \
                                                               y->coff)
  PLNR_REQUIRE(y_dtype == PLNR_F32 || y_dtype == PLNR_F16, "nchw_to_nhwc: the output dtype must be F32 or F16");
  if (x_dtype == PLNR_U8 && y_dtype == PLNR_F16) LAUNCH(uint8_t, __half);        // uint8 images: value-preserving cast
  else if (x_dtype == PLNR_U8) LAUNCH(uint8_t, float);
  else if (x_dtype == PLNR_F32 && y_dtype == PLNR_F32) LAUNCH(float, float);
  else if (x_dtype == PLNR_F32 && y_dtype == PLNR_F16) LAUNCH(float, __half);
  else if (x_dtype == PLNR_F16 && y_dtype == PLNR_F16) LAUNCH(__half, __half);
  else LAUNCH(__half, float);
#undef LAUNCH
  return plnr_after_launch(ctx, "nchw_to_nhwc");
}

int plnr_stem_pack(plnr_ctx* ctx, const void* x, int x_dtype, int n, int c, int h, int w, const plnr_tensor* y, int kw,
                   int stride, int pad_l) {
  PLNR_REQUIRE(ctx && x && y && y->ptr, "stem_pack: NULL argument");
  PLNR_REQUIRE(stride >= 1 && kw >= 1 && c >= 1 && stride * kw * c <= y->c && y->c % 8 == 0 && y->ld == y->c &&
                   y->coff == 0 && y->n == n, "stem_pack: packed channel count %d cannot hold %d x %d x %d taps",
               y->c, stride, kw, c);
  PLNR_REQUIRE(aligned16(y->ptr), "stem_pack: output must be 16-byte aligned");
  const int Wp = stride * (y->w - 1) + kw;                 // widest source column index + 1 (in padded coordinates)
  const size_t smem = (size_t)((stride * c * Wp + 7) / 8) * 8 * 2 + (size_t)y->c * 4;
  PLNR_REQUIRE(smem <= 48 * 1024, "stem_pack: source rows do not fit shared memory (%zu bytes)", smem);
  const int grid = n * y->h;
  if (x_dtype == PLNR_U8)
    stem_pack_kernel<uint8_t><<<grid, 256, smem, ctx->stream>>>((const uint8_t*)x, (__half*)y->ptr, n, c, h, w, y->h, y->w,
                                                                y->c, kw, stride, pad_l, Wp);
  else if (x_dtype == PLNR_F16)
    stem_pack_kernel<__half><<<grid, 256, smem, ctx->stream>>>((const __half*)x, (__half*)y->ptr, n, c, h, w, y->h, y->w,
                                                               y->c, kw, stride, pad_l, Wp);
  else
    stem_pack_kernel<float><<<grid, 256, smem, ctx->stream>>>((const float*)x, (__half*)y->ptr, n, c, h, w, y->h, y->w,
                                                              y->c, kw, stride, pad_l, Wp);
  return plnr_after_launch(ctx, "stem_pack");
}

int plnr_nhwc_to_nchw(plnr_ctx* ctx, const plnr_tensor* x, int x_dtype, void* y, int y_dtype) {
  PLNR_REQUIRE(ctx && x && x->ptr && y, "nhwc_to_nchw: NULL argument");
  const int HW = x->h * x->w;
  // (a channel count that is not a multiple of 8 -- YOLO's 255-channel heads in rows padded to 256 -- reads the pad channels
  // of the last 16-byte vector, which must lie inside the row: coff + roundup(c, 8) <= ld)
  if (x_dtype == PLNR_F16 && y_dtype == PLNR_F16 && x->ld % 8 == 0 && x->coff % 8 == 0 && aligned16(x->ptr) &&
      x->coff + (x->c + 7) / 8 * 8 <= x->ld && (reinterpret_cast<uintptr_t>(y) & 1) == 0 && !getenv("PLNR_TRANSPOSE_TILE")) {
    const int PB8 = ((HW + 7) / 8 + 7) / 8, CG4 = ((x->c + 7) / 8 + 3) / 4;
    const long long total = (long long)x->n * PB8 * CG4 * 32;
    PLNR_REQUIRE((total + 255) / 256 < (1ll << 31), "nhwc_to_nchw: tensor too large for one launch");
    nhwc_to_nchw_h16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const __half*)x->ptr, (__half*)y, x->c, HW,
                                                                                    x->ld, x->coff, PB8, CG4, total);
    return plnr_after_launch(ctx, "nhwc_to_nchw");
  }
  dim3 grid((HW + 31) / 32, (x->c + 31) / 32, x->n), block(32, 8);
  PLNR_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "nhwc_to_nchw: batch/channel count too large for one launch");
#define LAUNCH(TI, TO) \
  nhwc_to_nchw_kernel<TI, TO><<<grid, block, 0, ctx->stream>>>((const TI*)x->ptr, (TO*)y, x->c, HW, x->ld, x->coff)
  if (x_dtype == PLNR_F32 && y_dtype == PLNR_F32) LAUNCH(float, float);
  else if (x_dtype == PLNR_F32 && y_dtype == PLNR_F16) LAUNCH(float, __half);
  else if (x_dtype == PLNR_F16 && y_dtype == PLNR_F16) LAUNCH(__half, __half);
  else LAUNCH(__half, float);
#undef LAUNCH
  return plnr_after_launch(ctx, "nhwc_to_nchw");
}

int plnr_cast(plnr_ctx* ctx, const void* x, int x_dtype, void* y, int y_dtype, int64_t n) {
  PLNR_REQUIRE(ctx && x && y && n >= 0, "cast: bad argument");
  if (n == 0) return PLNR_OK;
  int grid = grid_for(n, ctx->sm_count);
  PLNR_REQUIRE(y_dtype == PLNR_F32 || y_dtype == PLNR_F16, "cast: the output dtype must be F32 or F16");
  if (x_dtype == PLNR_U8 && y_dtype == PLNR_F16)
    cast_kernel<uint8_t, __half><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)x, (__half*)y, n);
  else if (x_dtype == PLNR_U8)
    cast_kernel<uint8_t, float><<<grid, kThreads, 0, ctx->stream>>>((const uint8_t*)x, (float*)y, n);
  else if (x_dtype == PLNR_F32 && y_dtype == PLNR_F16)
    cast_kernel<float, __half><<<grid, kThreads, 0, ctx->stream>>>((const float*)x, (__half*)y, n);
  else if (x_dtype == PLNR_F16 && y_dtype == PLNR_F32)
    cast_kernel<__half, float><<<grid, kThreads, 0, ctx->stream>>>((const __half*)x, (float*)y, n);
  else if (x_dtype == PLNR_F32)
    cast_kernel<float, float><<<grid, kThreads, 0, ctx->stream>>>((const float*)x, (float*)y, n);
  else
    cast_kernel<__half, __half><<<grid, kThreads, 0, ctx->stream>>>((const __half*)x, (__half*)y, n);
  return plnr_after_launch(ctx, "cast");
}

int plnr_pack_conv_weight(plnr_ctx* ctx, const void* w, int w_dtype, void* out, int out_dtype, int cout, int cin_g,
                          int kh, int kw, int cin_pad) {
  PLNR_REQUIRE(ctx && w && out, "pack_conv_weight: NULL argument");
  PLNR_REQUIRE(cin_pad >= cin_g && cout >= 1 && kh >= 1 && kw >= 1, "pack_conv_weight: bad extents");
  int64_t total = (int64_t)cout * kh * kw * cin_pad;
  int grid = grid_for(total, ctx->sm_count);
#define LAUNCH(TI, TO) \
  pack_weight_kernel<TI, TO><<<grid, kThreads, 0, ctx->stream>>>((const TI*)w, (TO*)out, cout, cin_g, kh, kw, cin_pad)
  if (w_dtype == PLNR_F32 && out_dtype == PLNR_F32) LAUNCH(float, float);
  else if (w_dtype == PLNR_F32 && out_dtype == PLNR_F16) LAUNCH(float, __half);
  else if (w_dtype == PLNR_F16 && out_dtype == PLNR_F16) LAUNCH(__half, __half);
  else LAUNCH(__half, float);
#undef LAUNCH
  return plnr_after_launch(ctx, "pack_conv_weight");
}

int plnr_fold_affine(plnr_ctx* ctx, const void* bias, const void* bn_k, const void* bn_b, int dtype, float* scale,
                     float* shift, int c) {
  PLNR_REQUIRE(ctx && scale && shift && c >= 1, "fold_affine: bad argument");
  int grid = (c + kThreads - 1) / kThreads;
  DISPATCH_T(dtype, fold_affine_kernel<T><<<grid, kThreads, 0, ctx->stream>>>((const T*)bias, (const T*)bn_k,
                                                                              (const T*)bn_b, scale, shift, c);)
  return plnr_after_launch(ctx, "fold_affine");
}

int plnr_maxpool2d(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, int kh, int kw, int pad_t,
                   int pad_l, int stride_h, int stride_w) {
  PLNR_REQUIRE(ctx && x && y && x->ptr && y->ptr, "maxpool2d: NULL argument");
  PLNR_REQUIRE(x->n == y->n && x->c == y->c, "maxpool2d: batch/channel mismatch");
  PLNR_REQUIRE(kh >= 1 && kw >= 1 && stride_h >= 1 && stride_w >= 1 && pad_t >= 0 && pad_l >= 0, "maxpool2d: bad window");
  int64_t work;
#define MP_LAUNCH(VV, KH_, KW_)                                                                                         \
  maxpool_kernel<T, VV, KH_, KW_><<<grid_for(work, ctx->sm_count * 2), kThreads, 0, ctx->stream>>>(                    \
      (const T*)x->ptr, (T*)y->ptr, x->n, x->h, x->w, x->c, x->ld, x->coff, y->h, y->w, y->ld, y->coff, kh, kw, pad_t, \
      pad_l, stride_h, stride_w)
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    if (view_vec_ok(x, V, sizeof(T)) && view_vec_ok(y, V, sizeof(T))) {
      work = (int64_t)y->n * y->h * y->w * (y->c / V);
      if (kh == 3 && kw == 3 && stride_h == 2 && stride_w == 2) {
        constexpr int RB = 4;
        const int64_t w2 = (int64_t)y->n * ((y->h + RB - 1) / RB) * ((y->w + 1) / 2) * (y->c / V);
        maxpool3x3s2_kernel<T, V, RB><<<grid_for(w2, ctx->sm_count * 2), kThreads, 0, ctx->stream>>>(
            (const T*)x->ptr, (T*)y->ptr, x->n, x->h, x->w, x->c, x->ld, x->coff, y->h, y->w, y->ld, y->coff, pad_t, pad_l);
      } else if (kh == 3 && kw == 3) MP_LAUNCH(V, 3, 3);
      else if (kh == 2 && kw == 2) MP_LAUNCH(V, 2, 2);
      else MP_LAUNCH(V, 0, 0);
    } else {
      work = (int64_t)y->n * y->h * y->w * y->c;
      MP_LAUNCH(1, 0, 0);
    }
  })
#undef MP_LAUNCH
  return plnr_after_launch(ctx, "maxpool2d");
}

static inline dim3 row_grid(int row_items, int64_t rows, int sm_count) {
  // x covers one output row; y = enough row walkers for ~16 resident CTAs per SM, each walking rows with stride gridDim.y
  // (one block per row made 100 k one-vector blocks for a 55-pixel row: block scheduling, not HBM, bounded the kernel)
  const int gx = (row_items + kThreads - 1) / kThreads;
  int64_t gy = ((int64_t)sm_count * 16 + gx - 1) / gx;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  return dim3((unsigned)gx, (unsigned)gy);
}

int plnr_avgpool2d(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, int kh, int kw, int pad_t,
                   int pad_l, int stride_h, int stride_w) {
  PLNR_REQUIRE(ctx && x && y && x->ptr && y->ptr, "avgpool2d: NULL argument");
  PLNR_REQUIRE(x->n == y->n && x->c == y->c, "avgpool2d: batch/channel mismatch");
  PLNR_REQUIRE(kh >= 1 && kw >= 1 && stride_h >= 1 && stride_w >= 1 && pad_t >= 0 && pad_l >= 0, "avgpool2d: bad window");
#define AP_LAUNCH(VV, KH_, KW_)                                                                                            \
  avgpool_kernel<T, VV, KH_, KW_><<<row_grid(y->w * (y->c / VV), (int64_t)y->n * y->h, ctx->sm_count), kThreads, 0, ctx->stream>>>(        \
      (const T*)x->ptr, (T*)y->ptr, x->n, x->h, x->w, x->c, x->ld, x->coff, y->h, y->w, y->ld, y->coff, kh, kw, pad_t, pad_l, \
      stride_h, stride_w)
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    if (view_vec_ok(x, V, sizeof(T)) && view_vec_ok(y, V, sizeof(T))) {
      if (kh == 3 && kw == 3 && stride_h == 2 && stride_w == 2) {
        constexpr int RB = 4;
        const int64_t w2 = (int64_t)y->n * ((y->h + RB - 1) / RB) * ((y->w + 1) / 2) * (y->c / V);
        avgpool3x3s2_kernel<T, V, RB><<<grid_for(w2, ctx->sm_count * 2), kThreads, 0, ctx->stream>>>(
            (const T*)x->ptr, (T*)y->ptr, x->n, x->h, x->w, x->c, x->ld, x->coff, y->h, y->w, y->ld, y->coff, pad_t, pad_l);
      } else if (kh == 2 && kw == 2) AP_LAUNCH(V, 2, 2);
      else if (kh == 3 && kw == 3) AP_LAUNCH(V, 3, 3);
      else AP_LAUNCH(V, 0, 0);
    } else {
      AP_LAUNCH(1, 0, 0);
    }
  })
#undef AP_LAUNCH
  return plnr_after_launch(ctx, "avgpool2d");
}

int plnr_zero_stuff(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, int lo_h, int lo_w, int stride_h,
                    int stride_w) {
  PLNR_REQUIRE(ctx && x && y && x->ptr && y->ptr, "zero_stuff: NULL argument");
  PLNR_REQUIRE(x->n == y->n && x->c == y->c && stride_h >= 1 && stride_w >= 1, "zero_stuff: shape mismatch");
  PLNR_REQUIRE(lo_h >= 0 && lo_w >= 0 && lo_h + (x->h - 1) * stride_h < y->h && lo_w + (x->w - 1) * stride_w < y->w,
               "zero_stuff: the stuffed input does not fit the output (negative crop is not supported)");
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    if (view_vec_ok(x, V, sizeof(T)) && view_vec_ok(y, V, sizeof(T))) {
      zero_stuff_kernel<T, V><<<row_grid(y->w * (y->c / V), (int64_t)y->n * y->h, ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, x->n, x->h, x->w, x->c, x->ld, x->coff, y->h, y->w, y->ld, y->coff, lo_h, lo_w,
          stride_h, stride_w);
    } else {
      zero_stuff_kernel<T, 1><<<row_grid(y->w * y->c, (int64_t)y->n * y->h, ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, x->n, x->h, x->w, x->c, x->ld, x->coff, y->h, y->w, y->ld, y->coff, lo_h, lo_w,
          stride_h, stride_w);
    }
  })
  return plnr_after_launch(ctx, "zero_stuff");
}

int plnr_flip_weight(plnr_ctx* ctx, int dtype, const void* w, void* out, int ci, int co, int kh, int kw) {
  PLNR_REQUIRE(ctx && w && out && ci >= 1 && co >= 1 && kh >= 1 && kw >= 1, "flip_weight: bad argument");
  const int64_t work = (int64_t)ci * co * kh * kw;
  DISPATCH_T(dtype, flip_weight_kernel<T><<<grid_for(work, ctx->sm_count * 2), kThreads, 0, ctx->stream>>>(
                        (const T*)w, (T*)out, ci, co, kh, kw);)
  return plnr_after_launch(ctx, "flip_weight");
}

int plnr_upsample_nearest(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, int fh, int fw) {
  PLNR_REQUIRE(ctx && x && y && x->ptr && y->ptr, "upsample_nearest: NULL argument");
  PLNR_REQUIRE(fh >= 1 && fw >= 1 && y->h == x->h * fh && y->w == x->w * fw && y->n == x->n && y->c == x->c,
               "upsample_nearest: output extents do not match factors (%d,%d)", fh, fw);
  int64_t work;
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    if (view_vec_ok(x, V, sizeof(T)) && view_vec_ok(y, V, sizeof(T))) {
      work = (int64_t)x->n * x->h * x->w * (x->c / V);
      upsample_kernel<T, V><<<grid_for(work, ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, x->n, x->h, x->w, x->c, x->ld, x->coff, y->ld, y->coff, fh, fw);
    } else {
      work = (int64_t)x->n * x->h * x->w * x->c;
      upsample_kernel<T, 1><<<grid_for(work, ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, x->n, x->h, x->w, x->c, x->ld, x->coff, y->ld, y->coff, fh, fw);
    }
  })
  return plnr_after_launch(ctx, "upsample_nearest");
}

int plnr_upsample_linear(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, int fh, int fw, const float* wmat) {
  PLNR_REQUIRE(ctx && x && y && x->ptr && y->ptr && wmat, "upsample_linear: NULL argument");
  PLNR_REQUIRE(fh >= 2 && fw >= 2 && y->h == x->h * fh && y->w == x->w * fw && y->n == x->n && y->c == x->c,
               "upsample_linear: factors must be >= 2 in both axes and match the output extents (%d,%d)", fh, fw);
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    if (view_vec_ok(x, V, sizeof(T)) && view_vec_ok(y, V, sizeof(T))) {
      upsample_linear_kernel<T, V><<<row_grid((x->w + 1) * (x->c / V), (int64_t)x->n * (x->h + 1), ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, wmat, x->n, x->h, x->w, x->c, x->ld, x->coff, y->ld, y->coff, fh, fw);
    } else {
      upsample_linear_kernel<T, 1><<<row_grid((x->w + 1) * x->c, (int64_t)x->n * (x->h + 1), ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, wmat, x->n, x->h, x->w, x->c, x->ld, x->coff, y->ld, y->coff, fh, fw);
    }
  })
  return plnr_after_launch(ctx, "upsample_linear");
}

int plnr_resize_linear(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, const int* row_lo, const float* row_w,
                       const float* row_w1, const int* col_lo, const float* col_w, const float* col_w1) {
  PLNR_REQUIRE(ctx && x && y && x->ptr && y->ptr && row_lo && row_w && row_w1 && col_lo && col_w && col_w1,
               "resize_linear: NULL argument");
  PLNR_REQUIRE(y->n == x->n && y->c == x->c && x->h >= 1 && x->w >= 1, "resize_linear: batch/channel mismatch");
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    if (view_vec_ok(x, V, sizeof(T)) && view_vec_ok(y, V, sizeof(T))) {
      resize_linear_kernel<T, V><<<row_grid(y->w * (y->c / V), (int64_t)y->n * y->h, ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, row_lo, row_w, row_w1, col_lo, col_w, col_w1, x->n, x->h, x->w, x->c, x->ld, x->coff,
          y->h, y->w, y->ld, y->coff);
    } else {
      resize_linear_kernel<T, 1><<<row_grid(y->w * y->c, (int64_t)y->n * y->h, ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, row_lo, row_w, row_w1, col_lo, col_w, col_w1, x->n, x->h, x->w, x->c, x->ld, x->coff,
          y->h, y->w, y->ld, y->coff);
    }
  })
  return plnr_after_launch(ctx, "resize_linear");
}

int plnr_copy_channels(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y) {
  PLNR_REQUIRE(ctx && x && y && x->ptr && y->ptr, "copy_channels: NULL argument");
  PLNR_REQUIRE(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c, "copy_channels: shape mismatch");
  const int64_t npix = (int64_t)x->n * x->h * x->w;
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    if (view_vec_ok(x, V, sizeof(T)) && view_vec_ok(y, V, sizeof(T))) {
      copy_channels_kernel<T, V><<<grid_for(npix * (x->c / V), ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, npix, x->c, x->ld, x->coff, y->ld, y->coff);
    } else {
      copy_channels_kernel<T, 1><<<grid_for(npix * x->c, ctx->sm_count), kThreads, 0, ctx->stream>>>(
          (const T*)x->ptr, (T*)y->ptr, npix, x->c, x->ld, x->coff, y->ld, y->coff);
    }
  })
  return plnr_after_launch(ctx, "copy_channels");
}

int plnr_eltwise(plnr_ctx* ctx, int op, int dtype, const void* x, const void* p0, const void* p1, void* y,
                 int64_t npix, int c, float alpha) {
  PLNR_REQUIRE(ctx && x && y, "eltwise: NULL argument");
  PLNR_REQUIRE(op >= PLNR_EW_RELU && op <= PLNR_EW_SCALE_SHIFT, "eltwise: unknown op %d", op);
  PLNR_REQUIRE(op != PLNR_EW_ADD || p0, "eltwise: ADD needs a second operand");
  PLNR_REQUIRE(op != PLNR_EW_SCALE_SHIFT || (p0 && p1), "eltwise: SCALE_SHIFT needs K and B");
  const int64_t total = npix * c;
  if (total == 0) return PLNR_OK;
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    bool vec = (c % V == 0) && aligned16(x) && aligned16(y) && (op != PLNR_EW_ADD || aligned16(p0));
    if (vec)
      eltwise_kernel<T, V><<<grid_for(total / V, ctx->sm_count), kThreads, 0, ctx->stream>>>(
          op, (const T*)x, (const T*)p0, (const T*)p1, (T*)y, total / V, c, alpha);
    else
      eltwise_kernel<T, 1><<<grid_for(total, ctx->sm_count), kThreads, 0, ctx->stream>>>(
          op, (const T*)x, (const T*)p0, (const T*)p1, (T*)y, total, c, alpha);
  })
  return plnr_after_launch(ctx, "eltwise");
}

int plnr_unary2(plnr_ctx* ctx, int op, int dtype, const void* x, void* y, int64_t n, float a, float b) {
  PLNR_REQUIRE(ctx && x && y && n >= 0, "unary2: bad argument");
  PLNR_REQUIRE(op == PLNR_EW_CLIP || op == PLNR_EW_HARDSIGMOID, "unary2: unknown op %d", op);
  if (n == 0) return PLNR_OK;
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    if (n % V == 0 && aligned16(x) && aligned16(y))
      unary2_kernel<T, V><<<grid_for(n / V, ctx->sm_count * 2), kThreads, 0, ctx->stream>>>(op, (const T*)x, (T*)y, n / V, a, b);
    else
      unary2_kernel<T, 1><<<grid_for(n, ctx->sm_count * 2), kThreads, 0, ctx->stream>>>(op, (const T*)x, (T*)y, n, a, b);
  })
  return plnr_after_launch(ctx, "unary2");
}

int plnr_softmax(plnr_ctx* ctx, int dtype, const void* x, void* y, int64_t rows, int c) {
  PLNR_REQUIRE(ctx && x && y && rows >= 0 && c >= 1, "softmax: bad argument");
  if (rows == 0) return PLNR_OK;
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    const int G = c / V;
    if (c % V == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0 && aligned16(x) && aligned16(y))
      softmax_rows_kernel<T, V><<<grid_for(rows * G, ctx->sm_count * 4), kThreads, 0, ctx->stream>>>((const T*)x, (T*)y, rows, G);
    else
      softmax_kernel<T><<<grid_for(rows * 32, ctx->sm_count * 4), kThreads, 0, ctx->stream>>>((const T*)x, (T*)y, rows, c);
  })
  return plnr_after_launch(ctx, "softmax");
}

int plnr_global_avgpool(plnr_ctx* ctx, int dtype, const plnr_tensor* x, void* y) {
  PLNR_REQUIRE(ctx && x && x->ptr && y, "global_avgpool: NULL argument");
  const int HW = x->h * x->w;
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    if (view_vec_ok(x, V, sizeof(T)) && aligned16(y)) {
      int work = x->n * (x->c / V);
      gap_kernel<T, V><<<(work + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>((const T*)x->ptr, (T*)y, x->n,
                                                                                     HW, x->c, x->ld, x->coff);
    } else {
      int work = x->n * x->c;
      gap_kernel<T, 1><<<(work + kThreads - 1) / kThreads, kThreads, 0, ctx->stream>>>((const T*)x->ptr, (T*)y, x->n,
                                                                                     HW, x->c, x->ld, x->coff);
    }
  })
  return plnr_after_launch(ctx, "global_avgpool");
}

int plnr_gap_dense_fwd(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const void* w, const float* scale,
                       const float* shift, void* y, int out_features, int act, float alpha) {
  PLNR_REQUIRE(ctx && x && x->ptr && w && y, "gap_dense: NULL argument");
  const int HW = x->h * x->w;
  constexpr int IMGS = 4, SPLIT = 2, PB = 13, OPW = 4;    // OPW = 8 spills in the fp16 instantiation (ptxas -v)
  DISPATCH_T(dtype, {
    constexpr int V = VecWidth<T>::value;
    PLNR_REQUIRE(view_vec_ok(x, V, sizeof(T)) && aligned16(w) && x->c % V == 0,
                 "gap_dense: channels (%d) must be a multiple of %d and pointers 16-byte aligned", x->c, V);
    const size_t pooled_bytes = (size_t)IMGS * x->c * sizeof(float);
    PLNR_REQUIRE(pooled_bytes <= 96 * 1024, "gap_dense: %d channels do not fit the pooled-vector stage", x->c);
    const size_t row_bytes = (size_t)x->c * sizeof(T);
    // (1) weights in shared memory: images in groups of 8, output features in slices of <= 160 KB of weight rows, about one
    //     CTA per SM; (2) otherwise the streaming kernel.
    constexpr int IMGS8 = 8;
    const int gx8 = (x->n + IMGS8 - 1) / IMGS8;
    int gy8 = ctx->sm_count / gx8;
    if (gy8 < 1) gy8 = 1;
    if (gy8 > 16) gy8 = 16;
    int och8 = (out_features + gy8 - 1) / gy8;
    och8 = (och8 + OPW - 1) / OPW * OPW;
    const size_t smem8 = (size_t)IMGS8 * x->c * sizeof(float) + (size_t)och8 * row_bytes + 16;
    const char* nomma = getenv("PLNR_GAP_DENSE_NO_MMA");
    if (sizeof(T) == 2 && x->c % 16 == 0 && (size_t)IMGS8 * (x->c + 8) * 2 <= 48 * 1024 && !(nomma && atoi(nomma))) {
      // fp16: 16 image groups x feature slices of a multiple of 16, about one CTA per SM
      int gym = ctx->sm_count / gx8;
      if (gym < 1) gym = 1;
      int ochm = (out_features + gym - 1) / gym;
      ochm = (ochm + 15) / 16 * 16;
      gym = (out_features + ochm - 1) / ochm;
      const size_t smemm = (size_t)IMGS8 * (x->c + 8) * sizeof(__half);
      gap_dense_mma_kernel<IMGS8><<<dim3(gx8, gym), 512, smemm, ctx->stream>>>(
          (const __half*)x->ptr, (const __half*)w, scale, shift, (__half*)y, x->n, HW, x->c, x->ld, x->coff, out_features, ochm,
          act, alpha);
      return plnr_after_launch(ctx, "gap_dense");
    }
    const char* nos = getenv("PLNR_GAP_DENSE_STREAM");
    if (smem8 <= 200 * 1024 && !(nos && atoi(nos))) {
      gy8 = (out_features + och8 - 1) / och8;
      auto kern8 = gap_dense_smem_kernel<T, V, IMGS8, SPLIT, PB, OPW>;
      PLNR_CHECK_CUDA(cudaFuncSetAttribute(kern8, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
      kern8<<<dim3(gx8, gy8), 512, smem8, ctx->stream>>>(
          (const T*)x->ptr, (const T*)w, scale, shift, (T*)y, x->n, HW, x->c, x->ld, x->coff, out_features, och8, act, alpha);
      return plnr_after_launch(ctx, "gap_dense");
    }
    const size_t smem = pooled_bytes;
    // ~one CTA per SM: images in groups of IMGS, output features in slices.  Every slice re-reads its images' pixels and
    // every image group re-reads the weights (both from L2): 32 x 4 for ResNet-18 at batch 128 keeps the two about equal.
    const int gx = (x->n + IMGS - 1) / IMGS;
    int gy = ctx->sm_count / gx;                     // never more CTAs than SMs: a CTA fills an SM's register file
    if (gy < 1) gy = 1;
    if (gy > 8) gy = 8;
    int och = (out_features + gy - 1) / gy;
    och = (och + OPW - 1) / OPW * OPW;
    if (och < 64) och = 64;
    gy = (out_features + och - 1) / och;
    auto kern = gap_dense_kernel<T, V, IMGS, SPLIT, PB, OPW>;
    if (smem > 48 * 1024) {
      PLNR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    }
    kern<<<dim3(gx, gy), 512, smem, ctx->stream>>>(
        (const T*)x->ptr, (const T*)w, scale, shift, (T*)y, x->n, HW, x->c, x->ld, x->coff, out_features, och, act, alpha);
  })
  return plnr_after_launch(ctx, "gap_dense");
}

int plnr_pooled_dense_fwd(plnr_ctx* ctx, const float* pool, int n, int parts, int c, int hw, const void* w, const float* scale,
                          const float* shift, void* y, int out_features, int act, float alpha) {
  PLNR_REQUIRE(ctx && pool && w && y, "pooled_dense: NULL argument");
  PLNR_REQUIRE(n >= 1 && parts >= 1 && hw >= 1 && out_features >= 1, "pooled_dense: empty problem");
  PLNR_REQUIRE(c % 16 == 0 && aligned16(w) && aligned16(pool), "pooled_dense: channels (%d) must be a multiple of 16 and pointers "
               "16-byte aligned", c);
  constexpr int IMGS8 = 8;
  const size_t smem = ((size_t)IMGS8 * (c + 8) * sizeof(__half) + 15) / 16 * 16 + 16 * 128 * sizeof(float);
  PLNR_REQUIRE(smem <= 48 * 1024, "pooled_dense: %d channels do not fit the pooled-vector stage", c);
  const int gx = (n + IMGS8 - 1) / IMGS8;
  int gy = ctx->sm_count / gx;
  if (gy < 1) gy = 1;
  int och = (out_features + gy - 1) / gy;
  och = (och + 15) / 16 * 16;
  if (och < 128 && out_features > och) och = 128 < out_features ? 128 : (out_features + 15) / 16 * 16;   // eight blocks x two K halves = 16 warps
  gy = (out_features + och - 1) / och;
  pooled_dense_mma_kernel<IMGS8><<<dim3(gx, gy), 512, smem, ctx->stream>>>(
      pool, (const __half*)w, scale, shift, (__half*)y, n, parts, 1.f / (float)hw, c, out_features, och, act, alpha);
  return plnr_after_launch(ctx, "pooled_dense");
}

}  // extern "C"

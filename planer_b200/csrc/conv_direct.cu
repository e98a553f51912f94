// Direct (CUDA-core) implicit-GEMM convolution over the pixel-major layout.
//
// This is the fp32 product path (BASELINE config 2: fp32 within 1e-3 of the numpy reference needs
// true fp32 multiplies -- tcgen05 has no fp32 kind, and TF32 truncation fails the bar)
// and the path for problems the tensor-core kernel does not take (groups > 1, Cin % 16 != 0).
// Same GEMM as the reference (planer/util.py:41-43: M=Co, K=C*kh*kw, N=N*oh*ow) but the im2col matrix
// is never materialised: each 64-pixel x 16-k slab is gathered straight into shared memory.
// fp32 accumulation for both dtypes; the fused epilogue is shared with the tensor-core kernel.
#include "common.cuh"

namespace {

constexpr int BM = 64;   // output pixels per CTA
constexpr int BN = 64;   // output channels per CTA
constexpr int BK = 16;   // contraction slab
constexpr int NT = 256;  // threads: 16 x 16, each owns a 4 x 4 micro-tile

struct DirectParams {
  const void* x; const void* w; void* y;
  int N, H, W, xld, xcoff;
  int OH, OW, yld, ycoff;
  int Cg, Cog;           // channels per group (input / output)
  int kh, kw, pt, pl, sh, sw, dh, dw;
  int Ktot;              // kh*kw*Cg
  int M;                 // N*OH*OW
  const float* scale; const float* shift;
  const void* res; int rld, rcoff;
  int act; float alpha; int res_after;
};

template <typename T>
__global__ void __launch_bounds__(NT) conv_direct_kernel(const DirectParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const T* __restrict__ x = (const T*)p.x;
  const T* __restrict__ w = (const T*)p.w;
  const int tid = threadIdx.x;
  const int g = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // loader role: thread -> (row, 4 consecutive k)
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int m = m0 + lrow;
  int pn = 0, poh = 0, pow_ = 0;
  const bool mvalid = m < p.M;
  if (mvalid) {
    pow_ = m % p.OW;
    int t = m / p.OW;
    poh = t % p.OH;
    pn = t / p.OH;
  }
  const int co_l = n0 + lrow;  // output channel (within group) this thread loads weights for
  const bool covalid = co_l < p.Cog;
  const T* wrow = w + ((size_t)(g * p.Cog + (covalid ? co_l : 0))) * p.Ktot;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.Ktot; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + lk + e;
      float av = 0.f, bv = 0.f;
      if (k < p.Ktot) {
        if (mvalid) {
          const int tap = k / p.Cg, c = k - tap * p.Cg;
          const int r = tap / p.kw, s = tap - r * p.kw;
          const int ih = poh * p.sh + r * p.dh - p.pt, iw = pow_ * p.sw + s * p.dw - p.pl;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
            av = ld_f(x + (((size_t)pn * p.H + ih) * p.W + iw) * p.xld + p.xcoff + g * p.Cg + c);
        }
        if (covalid) bv = ld_f(wrow + k);
      }
      As[lk + e][lrow] = av;
      Bs[lk + e][lrow] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  T* __restrict__ y = (T*)p.y;
  const T* __restrict__ res = (const T*)p.res;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + ty * 4 + i;
    if (mm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cg = n0 + tx * 4 + j;
      if (cg >= p.Cog) continue;
      const int co = g * p.Cog + cg;
      float v = acc[i][j];
      if (p.scale) v *= p.scale[co];
      if (p.shift) v += p.shift[co];
      const float rv = res ? ld_f(res + (size_t)mm * p.rld + p.rcoff + co) : 0.f;
      if (!p.res_after) v += rv;
      v = plnr_apply_act(v, p.act, p.alpha);
      if (p.res_after) v += rv;
      st_f(y + (size_t)mm * p.yld + p.ycoff + co, v);
    }
  }
}

// fp32 fast path (BASELINE config 2): 128 pixels x 128 channels per CTA, 8 x 8 outputs per thread, contraction slabs of 16
// gathered with ONE tap decode and one 16-byte load per four channels (needs Cg % 4 == 0 and 16-byte aligned rows), shared
// memory double buffered.  FFMA-bound: 64 FFMA per four 16-byte shared-memory loads.  Same arithmetic as the generic kernel
// above (fp32 multiplies, fp32 accumulation in k order within a slab), same fused epilogue.
constexpr int FM = 128, FN = 128, FK = 16;

__global__ void __launch_bounds__(256) conv_direct_f32x8_kernel(const DirectParams p) {
  __shared__ __align__(16) float As[2][FK][FM + 4];
  __shared__ __align__(16) float Bs[2][FK][FN + 4];
  const float* __restrict__ x = (const float*)p.x;
  const float* __restrict__ w = (const float*)p.w;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * FM, n0 = blockIdx.y * FN;
  // loader role: thread -> rows (tid / 4) and (tid / 4 + 64), four consecutive k starting at (tid % 4) * 4
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  int pn[2], ph[2], pw[2];
  bool mvalid[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + lrow + 64 * i;
    mvalid[i] = m < p.M;
    const int mm = mvalid[i] ? m : 0;
    pw[i] = mm % p.OW;
    const int t = mm / p.OW;
    ph[i] = (t % p.OH) * p.sh - p.pt;
    pn[i] = t / p.OH;
    pw[i] = pw[i] * p.sw - p.pl;
  }
  const float* wrow[2];
  bool covalid[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int co = n0 + lrow + 64 * i;
    covalid[i] = co < p.Cog;
    wrow[i] = w + (size_t)(covalid[i] ? co : 0) * p.Ktot;
  }
  const int ty = tid >> 4, tx = tid & 15;      // thread -> rows ty*4 + {0..3} and 64 + ty*4 + {0..3}; channels likewise with tx
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
    const int k = k0 + lk;
    const bool kvalid = k < p.Ktot;
    const int tap = kvalid ? k / p.Cg : 0, c = k - tap * p.Cg;
    const int r = tap / p.kw, s_ = tap - r * p.kw;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ih = ph[i] + r * p.dh, iw = pw[i] + s_ * p.dw;
      if (kvalid && mvalid[i] && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
        ra[i] = __ldg(reinterpret_cast<const float4*>(x + (((size_t)pn[i] * p.H + ih) * p.W + iw) * p.xld + p.xcoff + c));
      if (kvalid && covalid[i]) rb[i] = __ldg(reinterpret_cast<const float4*>(wrow[i] + k));
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = lrow + 64 * i;
      As[buf][lk + 0][row] = ra[i].x; As[buf][lk + 1][row] = ra[i].y; As[buf][lk + 2][row] = ra[i].z; As[buf][lk + 3][row] = ra[i].w;
      Bs[buf][lk + 0][row] = rb[i].x; Bs[buf][lk + 1][row] = rb[i].y; Bs[buf][lk + 2][row] = rb[i].z; Bs[buf][lk + 3][row] = rb[i].w;
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < p.Ktot; k0 += FK) {
    const bool more = k0 + FK < p.Ktot;
    if (more) gload(k0 + FK);                 // global loads of the next slab fly during this slab's FFMAs
#pragma unroll
    for (int kk = 0; kk < FK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  float* __restrict__ y = (float*)p.y;
  const float* __restrict__ res = (const float*)p.res;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int mm = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (mm >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int co = n0 + jh * 64 + tx * 4;
      if (co >= p.Cog) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][jh * 4 + j];
      const bool full = co + 4 <= p.Cog;
      float rv[4] = {0.f, 0.f, 0.f, 0.f};
      if (res) {
        if (full) { const float4 r4 = *reinterpret_cast<const float4*>(res + (size_t)mm * p.rld + p.rcoff + co); rv[0] = r4.x; rv[1] = r4.y; rv[2] = r4.z; rv[3] = r4.w; }
        else for (int j = 0; j < 4 && co + j < p.Cog; ++j) rv[j] = res[(size_t)mm * p.rld + p.rcoff + co + j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (co + j < p.Cog) {
          if (p.scale) v[j] *= p.scale[co + j];
          if (p.shift) v[j] += p.shift[co + j];
          if (!p.res_after) v[j] += rv[j];
          v[j] = plnr_apply_act(v[j], p.act, p.alpha);
          if (p.res_after) v[j] += rv[j];
        }
      }
      if (full) *reinterpret_cast<float4*>(y + (size_t)mm * p.yld + p.ycoff + co) = make_float4(v[0], v[1], v[2], v[3]);
      else for (int j = 0; j < 4 && co + j < p.Cog; ++j) y[(size_t)mm * p.yld + p.ycoff + co + j] = v[j];
    }
  }
}

}  // namespace

int plnr_conv2d_direct(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w,
                       const plnr_tensor* y, const plnr_epilogue* ep) {
  DirectParams p;
  memset(&p, 0, sizeof(p));
  p.x = x->ptr; p.w = w; p.y = y->ptr;
  p.N = x->n; p.H = x->h; p.W = x->w; p.xld = x->ld; p.xcoff = x->coff;
  p.OH = y->h; p.OW = y->w; p.yld = y->ld; p.ycoff = y->coff;
  p.Cg = x->c / d->groups; p.Cog = y->c / d->groups;
  p.kh = d->kh; p.kw = d->kw; p.pt = d->pad_t; p.pl = d->pad_l;
  p.sh = d->stride_h; p.sw = d->stride_w; p.dh = d->dil_h; p.dw = d->dil_w;
  p.Ktot = d->kh * d->kw * p.Cg;
  int64_t M = (int64_t)y->n * y->h * y->w;
  PLNR_REQUIRE(M < (1ll << 31), "conv2d(direct): more than 2^31 output pixels");
  p.M = (int)M;
  if (ep) {
    p.scale = ep->scale; p.shift = ep->shift; p.act = ep->act; p.alpha = ep->alpha;
    p.res_after = ep->res_after_act;
    if (ep->residual) { p.res = ep->residual->ptr; p.rld = ep->residual->ld; p.rcoff = ep->residual->coff; }
  }
  // fp32, one group, channel counts and pitches that allow 16-byte gathers: the 128 x 128 x 16 FFMA kernel
  {
    const bool al16 = ((reinterpret_cast<uintptr_t>(x->ptr) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y->ptr)) & 15) == 0;
    bool ok = d->dtype == PLNR_F32 && d->groups == 1 && p.Cg % 4 == 0 && x->ld % 4 == 0 && x->coff % 4 == 0 && y->ld % 4 == 0 &&
              y->coff % 4 == 0 && al16 && p.Cog >= 32;
    if (p.res) ok = ok && p.rld % 4 == 0 && p.rcoff % 4 == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0;
    if (const char* e = getenv("PLNR_DIRECT_SMALL")) { if (atoi(e)) ok = false; }
    if (ok) {
      dim3 grid((p.M + FM - 1) / FM, (p.Cog + FN - 1) / FN, 1);
      PLNR_REQUIRE(grid.y <= 65535, "conv2d(direct): too many channel tiles");
      conv_direct_f32x8_kernel<<<grid, 256, 0, ctx->stream>>>(p);
      return plnr_after_launch(ctx, "conv2d_direct_f32x8");
    }
  }
  dim3 grid((p.M + BM - 1) / BM, (p.Cog + BN - 1) / BN, d->groups);
  PLNR_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv2d(direct): too many channel tiles / groups");
  if (d->dtype == PLNR_F16) conv_direct_kernel<__half><<<grid, NT, 0, ctx->stream>>>(p);
  else conv_direct_kernel<float><<<grid, NT, 0, ctx->stream>>>(p);
  return plnr_after_launch(ctx, "conv2d_direct");
}

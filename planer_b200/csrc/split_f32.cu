// fp32 convolutions on the fp16 tensor pipe: the operand split.
//
// The reference computes Conv2d / Dense in the dtype of its arrays (planer/layer.py:22-26, :15-18); for float32 nets
// (BASELINE config 2) that is an fp32 GEMM.  tcgen05 has no fp32 kind, and kind::tf32 drops 13 mantissa bits (1e-3 on
// one product), so the fp32 path splits BOTH operands into two fp16 numbers each,
//     x * 2^ex = xh + xl,   w * 2^ew = wh + wl        (xh = fp16(x 2^ex), xl = fp16(x 2^ex - xh): 22 mantissa bits)
// and runs ONE fp16 implicit GEMM over 3C channels per filter tap
//     [xh | xh | xl] . [wh | wl | wh] = xh wh + xh wl + xl wh          (xl wl ~ 2^-22 of the product is dropped)
// whose fp32 accumulator (TMEM) the epilogue of conv_tcgen05.cu writes out as fp32 after multiplying by 2^-(ex+ew)
// (plnr_epilogue.out_f32 / acc_scale).  Every fp16 x fp16 product is exact in fp32; what remains is the fp32 accumulation
// of the tensor core -- the same order of error as an fp32 FFMA loop (tests/test_gpu_parity.py pins it against the
// CUDA-core kernel at 2e-5 and against the reference fixtures at the north star's 1e-3).
//
// The power-of-two pre-scales keep the LOW parts out of fp16's subnormal range whatever the magnitude of the tensors:
// weights are scaled once so that max|w| lands in [2^13, 2^14) (plnr_absmax_f32, the host picks the exponent);
// activations are scaled per call the same way ON THE DEVICE (`dyn`: one absmax pass, the split kernel derives 2^ex from
// it and leaves 2^-ex for the conv epilogue, plnr_epilogue.acc_scale_dev) -- elements more than 2^25 below the tensor's
// largest lose bits, as they would against the fp32 sum they are added into.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void split1(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// rows of C fp32 channels (pitch ld, first channel coff) -> rows of cs fp16 channels [hi C | hi C | lo C | untouched pad]
__device__ __forceinline__ float dyn_prescale(float amax) {
  // 2^e with amax * 2^e in [2^13, 2^14); 1 for an all-zero / non-finite tensor
  if (!(amax > 0.f) || amax > 3.0e38f) return 1.f;
  int e = 13 - ilogbf(amax);
  e = e > 120 ? 120 : (e < -120 ? -120 : e);
  return ldexpf(1.f, e);
}

template <int V>
__global__ void split_rows_kernel(const float* __restrict__ x, int ld, int coff, int C, __half* __restrict__ xs, int cs,
                                  long long rows, float prescale, float* __restrict__ dyn) {
  if (dyn) {
    prescale = dyn_prescale(__uint_as_float(*reinterpret_cast<const volatile unsigned int*>(dyn)));
    if (blockIdx.x == 0 && threadIdx.x == 0) dyn[1] = 1.f / prescale;        // exact: a power of two
  }
  const int per_row = C / V;
  const long long total = rows * per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / per_row;
    const int c = (int)(i - r * per_row) * V;
    const float* src = x + r * ld + coff + c;
    __half* dst = xs + r * cs + c;
    if (V == 4) {
      const float4 f = *reinterpret_cast<const float4*>(src);
      __half h[4], l[4];
      split1(f.x * prescale, h[0], l[0]); split1(f.y * prescale, h[1], l[1]);
      split1(f.z * prescale, h[2], l[2]); split1(f.w * prescale, h[3], l[3]);
      const uint2 hv = *reinterpret_cast<const uint2*>(h), lv = *reinterpret_cast<const uint2*>(l);
      *reinterpret_cast<uint2*>(dst) = hv;
      *reinterpret_cast<uint2*>(dst + C) = hv;
      *reinterpret_cast<uint2*>(dst + 2 * C) = lv;
    } else {
      __half h, l;
      split1(src[0] * prescale, h, l);
      dst[0] = h; dst[C] = h; dst[2 * C] = l;
    }
  }
}

// OIHW fp32 -> [Cout][kh][kw][cs] fp16 with [hi C | lo C | hi C | untouched pad] per filter tap
__global__ void pack_weight_split_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout, int cin, int kh,
                                         int kw, int cs, float prescale) {
  const long long total = (long long)cout * kh * kw * cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin);
    long long t = i / cin;
    const int s = (int)(t % kw); t /= kw;
    const int r = (int)(t % kh);
    const int o = (int)(t / kh);
    __half h, l;
    split1(w[(((long long)o * cin + c) * kh + r) * kw + s] * prescale, h, l);
    __half* dst = out + (((long long)o * kh + r) * kw + s) * cs + c;
    dst[0] = h; dst[cin] = l; dst[2 * cin] = h;
  }
}

// max |x| over `rows` rows of C channels (pitch ld, first channel coff); a flat array is one row per element run
__global__ void absmax_kernel(const float* __restrict__ x, int ld, int coff, int C, long long rows, unsigned int* out) {
  float m = 0.f;
  const long long n = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const float v = fabsf(x[r * ld + coff + (i - r * C)]);
    if (v == v) m = fmaxf(m, v);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));      // non-negative floats order like their bit patterns
}

int grid_of(long long total, int sm_count) {
  long long g = (total + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int plnr_split_f32(plnr_ctx* ctx, const plnr_tensor* x, const plnr_tensor* xs, float prescale, float* dyn) {
  PLNR_REQUIRE(ctx && x && xs && x->ptr && xs->ptr, "split_f32: NULL argument");
  PLNR_REQUIRE(xs->n == x->n && xs->h == x->h && xs->w == x->w, "split_f32: the two tensors differ in (n, h, w)");
  PLNR_REQUIRE(xs->c >= 3 * x->c && xs->coff == 0 && xs->ld == xs->c,
               "split_f32: the fp16 tensor must be dense with at least 3 x %d channels (has %d, ld %d, coff %d)", x->c, xs->c,
               xs->ld, xs->coff);
  const long long rows = (long long)x->n * x->h * x->w;
  if (rows == 0) return PLNR_OK;
  const bool v4 = x->c % 4 == 0 && x->ld % 4 == 0 && x->coff % 4 == 0 && (reinterpret_cast<uintptr_t>(x->ptr) & 15) == 0 &&
                  xs->c % 4 == 0 && (reinterpret_cast<uintptr_t>(xs->ptr) & 7) == 0;
  if (dyn) {
    PLNR_CHECK_CUDA(cudaMemsetAsync(dyn, 0, sizeof(float), ctx->stream));
    absmax_kernel<<<grid_of(rows * x->c, ctx->sm_count), kThreads, 0, ctx->stream>>>((const float*)x->ptr, x->ld, x->coff, x->c,
                                                                                   rows, (unsigned int*)dyn);
    int rc = plnr_after_launch(ctx, "split_f32(absmax)");
    if (rc != PLNR_OK) return rc;
  }
  if (v4)
    split_rows_kernel<4><<<grid_of(rows * (x->c / 4), ctx->sm_count), kThreads, 0, ctx->stream>>>(
        (const float*)x->ptr, x->ld, x->coff, x->c, (__half*)xs->ptr, xs->c, rows, prescale, dyn);
  else
    split_rows_kernel<1><<<grid_of(rows * x->c, ctx->sm_count), kThreads, 0, ctx->stream>>>(
        (const float*)x->ptr, x->ld, x->coff, x->c, (__half*)xs->ptr, xs->c, rows, prescale, dyn);
  return plnr_after_launch(ctx, "split_f32");
}

int plnr_pack_conv_weight_split(plnr_ctx* ctx, const void* w, void* out, int cout, int cin, int kh, int kw, int cs,
                                float prescale) {
  PLNR_REQUIRE(ctx && w && out, "pack_conv_weight_split: NULL argument");
  PLNR_REQUIRE(cout >= 1 && cin >= 1 && kh >= 1 && kw >= 1 && cs >= 3 * cin, "pack_conv_weight_split: bad extents");
  const long long total = (long long)cout * kh * kw * cin;
  pack_weight_split_kernel<<<grid_of(total, ctx->sm_count), kThreads, 0, ctx->stream>>>((const float*)w, (__half*)out, cout,
                                                                                      cin, kh, kw, cs, prescale);
  return plnr_after_launch(ctx, "pack_conv_weight_split");
}

int plnr_absmax_f32(plnr_ctx* ctx, const void* x, int64_t n, float* out) {
  PLNR_REQUIRE(ctx && x && out && n >= 0, "absmax_f32: bad argument");
  PLNR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), ctx->stream));
  if (n == 0) return PLNR_OK;
  absmax_kernel<<<grid_of(n, ctx->sm_count), kThreads, 0, ctx->stream>>>((const float*)x, 1, 0, 1, (long long)n, (unsigned int*)out);
  return plnr_after_launch(ctx, "absmax_f32");
}

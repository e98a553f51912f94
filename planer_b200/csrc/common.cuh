// Shared declarations of libplaner_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unordered_map>
#include <string>
#include <vector>

#include "../../include/planer_b200.h"

struct plnr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  int cc_major = 0, cc_minor = 0;
  int64_t l2_bytes = 0;
  int64_t launches = 0;            // kernels enqueued (eager) + graph nodes replayed
  bool capturing = false;
  int64_t capture_launches = 0;    // kernels recorded into the graph being captured
  // device-side error word written by kernels that time out on a barrier (debug aid)
  int* dev_error = nullptr;
  long long* prof = nullptr;       // debug: per-CTA role cycle counters of the last tcgen05 conv launch
  float* sk_ws = nullptr;          // stream-K workspace of the shift conv kernel: fp32 partial tiles, one slot per unit
  int* sk_flags = nullptr;         // ... and their ready flags (self-cleaning: the consumer resets them)
  std::unordered_map<long long, int> shift_max_units;   // (cta group, smem bytes) -> co-resident units of the shift conv kernel
  const char* last_kernel = "";   // name handed to plnr_after_launch by the most recent launch (diagnostics)
  bool shift_attr_set = false;
  bool igemm_attr_set = false;     // cudaFuncSetAttribute(max dynamic smem) done for this device
  // cache of TMA descriptors keyed by a byte string of their parameters
  std::unordered_map<std::string, CUtensorMap> tmap_cache;
};

struct plnr_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t nodes = 0;
};

struct plnr_event {
  cudaEvent_t ev = nullptr;
};

void plnr_set_error(const char* fmt, ...);

#define PLNR_CHECK_CUDA(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      plnr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PLNR_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

#define PLNR_REQUIRE(cond, ...)                  \
  do {                                           \
    if (!(cond)) {                               \
      plnr_set_error(__VA_ARGS__);               \
      return PLNR_ERR_INVALID;                   \
    }                                            \
  } while (0)

// Bookkeeping after every kernel launch: surfaces launch-configuration errors immediately.
static inline int plnr_after_launch(plnr_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    plnr_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return PLNR_ERR_CUDA;
  }
  if (ctx->capturing) ctx->capture_launches++;
  else ctx->launches++;
  ctx->last_kernel = what;
  return PLNR_OK;
}

// The tensor-core conv kernels are launched with programmatic dependent launch: the next grid is scheduled while the
// previous kernel of the stream drains, and everything it does before griddepcontrol.wait (barrier init, TMEM allocation,
// tensor-map prefetch) overlaps that kernel's tail.  Every CTA needs the whole SM's shared memory, so a dependent CTA starts
// when its predecessor on that SM exits: the gain is the launch latency, +0.7 % on the ResNet-18 batch-128 step
// (profiles/r02_kernel_experiments.md 5).  PLNR_PDL=0 turns it off.
static inline bool plnr_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PLNR_PDL"); v = (e && !atoi(e)) ? 0 : 1; }
  return v == 1;
}

static inline size_t plnr_dtype_size(int dt) { return dt == PLNR_F16 ? 2 : (dt == PLNR_U8 ? 1 : 4); }

static inline int plnr_out_size(int n_in, int pad_lo, int pad_hi, int k, int dil, int stride) {
  // planer/util.py:25-26
  return (n_in + pad_lo + pad_hi - (k - 1) * dil - 1 + stride) / stride;
}

// ---- element load/store helpers usable from templated kernels -------------------------------
template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ld_f<uint8_t>(const uint8_t* p) { return (float)*p; }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__half>(__half* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ float plnr_apply_act(float v, int act, float alpha) {
  // relu: x*(x>0) ; leaky: x*((x>0)*(1-a)+a) ; sigmoid: 1/(1+exp(-x))   (planer/layer.py:44-64)
  if (act == PLNR_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == PLNR_ACT_LEAKY) return v > 0.f ? v : v * alpha;
  if (act == PLNR_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

// exact floor(n / d) for n < 2^31 by multiply-high:  q = umulhi(n, mul) >> shr   (mul = ceil(2^(31+s) / d), 2^s >= d)
struct FastDiv { uint32_t mul, shr, d; };
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f; f.d = d;
  if (d <= 1) { f.mul = 0; f.shr = 0; return f; }          // d == 1 handled in fast_div
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.mul = (uint32_t)((((unsigned long long)1 << (31 + s)) + d - 1) / d);
  f.shr = s - 1;
  return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) {
  return f.d <= 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}

template <int A> struct ActTag { static constexpr int value = A; };   // compile-time activation switch of the conv epilogues

// internal entry points implemented in the other translation units
int plnr_conv2d_direct(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w,
                       const plnr_tensor* y, const plnr_epilogue* ep);
int plnr_conv2d_tcgen05(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w,
                        const plnr_tensor* y, const plnr_epilogue* ep);
bool plnr_conv2d_tcgen05_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y);
int plnr_conv2d_shift(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w,
                      const plnr_tensor* y, const plnr_epilogue* ep, const plnr_tensor* x2 = nullptr, int s2 = 1);
int plnr_conv2d_stack(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w, const plnr_tensor* y,
                      const plnr_epilogue* ep, const plnr_tensor* x2, int s2);
bool plnr_conv2d_stack_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y, const plnr_epilogue* ep,
                                 const plnr_tensor* x2, int s2);
bool plnr_conv2d_shift_shortcut_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* x2, int s2,
                                          const plnr_tensor* y);
bool plnr_conv2d_shift_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y);
int plnr_conv2d_shift_pool_parts(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y);
// pointwise convolutions with small resident filters (conv_pw.cu)
bool plnr_conv2d_pw_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y, const plnr_epilogue* ep);
int plnr_conv2d_pw(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w, const plnr_tensor* y,
                   const plnr_epilogue* ep);

// Context, memory, CUDA-graph and event entry points of the C ABI (include/planer_b200.h).
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[1024] = "";

void plnr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

int plnr_abi_version(void) { return PLNR_ABI_VERSION; }
const char* plnr_last_error(void) { return g_err; }

int plnr_create(int device, void* stream, plnr_ctx** out) {
  PLNR_REQUIRE(out != nullptr, "plnr_create: out is NULL");
  int ndev = 0;
  PLNR_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  PLNR_REQUIRE(device >= 0 && device < ndev, "plnr_create: device %d out of range (%d visible)", device, ndev);
  PLNR_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PLNR_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    plnr_set_error("plnr_create: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
                   prop.minor);
    return PLNR_ERR_UNSUPPORTED;
  }
  plnr_ctx* ctx = new plnr_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  ctx->l2_bytes = prop.l2CacheSize;
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    PLNR_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  PLNR_CHECK_CUDA(cudaMalloc(&ctx->dev_error, sizeof(int) * 4));
  PLNR_CHECK_CUDA(cudaMemset(ctx->dev_error, 0, sizeof(int) * 4));
  *out = ctx;
  return PLNR_OK;
}

int plnr_destroy(plnr_ctx* ctx) {
  if (!ctx) return PLNR_OK;
  cudaSetDevice(ctx->device);
  if (ctx->dev_error) cudaFree(ctx->dev_error);
  if (ctx->sk_ws) cudaFree(ctx->sk_ws);
  if (ctx->sk_flags) cudaFree(ctx->sk_flags);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return PLNR_OK;
}

int plnr_set_stream(plnr_ctx* ctx, void* stream) {
  PLNR_REQUIRE(ctx, "plnr_set_stream: ctx is NULL");
  PLNR_REQUIRE(!ctx->capturing, "plnr_set_stream: a graph capture is in progress");
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->own_stream = false;
  ctx->stream = (cudaStream_t)stream;
  return PLNR_OK;
}

int plnr_stream_sync(plnr_ctx* ctx) {
  PLNR_REQUIRE(ctx, "plnr_stream_sync: ctx is NULL");
  PLNR_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  int err[4] = {0, 0, 0, 0};
  PLNR_CHECK_CUDA(cudaMemcpy(err, ctx->dev_error, sizeof(err), cudaMemcpyDeviceToHost));
  if (err[0] != 0) {
    plnr_set_error("device-side watchdog fired: code %d (block %d, role %d, aux %d)", err[0], err[1], err[2], err[3]);
    cudaMemset(ctx->dev_error, 0, sizeof(err));
    if (ctx->sk_flags) cudaMemset(ctx->sk_flags, 0, sizeof(int) * 16 * 256);
    return PLNR_ERR_CUDA;
  }
  return PLNR_OK;
}

int plnr_debug_conv_profile(plnr_ctx* ctx, int enable, int64_t* out, int n) {
  PLNR_REQUIRE(ctx, "plnr_debug_conv_profile: ctx is NULL");
  if (enable && !ctx->prof) {
    PLNR_CHECK_CUDA(cudaMalloc(&ctx->prof, sizeof(long long) * 8 * 256));
    PLNR_CHECK_CUDA(cudaMemset(ctx->prof, 0, sizeof(long long) * 8 * 256));
  }
  if (out && ctx->prof) {
    PLNR_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    PLNR_CHECK_CUDA(cudaMemcpy(out, ctx->prof, sizeof(long long) * (n < 2048 ? n : 2048), cudaMemcpyDeviceToHost));
  }
  if (!enable && ctx->prof) { cudaFree(ctx->prof); ctx->prof = nullptr; }
  return PLNR_OK;
}

int plnr_last_kernel(plnr_ctx* ctx, char* out, int n) {
  PLNR_REQUIRE(ctx && out && n > 0, "plnr_last_kernel: bad argument");
  strncpy(out, ctx->last_kernel ? ctx->last_kernel : "", (size_t)n - 1);
  out[n - 1] = 0;
  return PLNR_OK;
}

int plnr_launch_count(plnr_ctx* ctx, int64_t* out) {
  PLNR_REQUIRE(ctx && out, "plnr_launch_count: NULL argument");
  *out = ctx->launches;
  return PLNR_OK;
}

int plnr_device_info(plnr_ctx* ctx, int64_t* out4) {
  PLNR_REQUIRE(ctx && out4, "plnr_device_info: NULL argument");
  out4[0] = ctx->sm_count;
  out4[1] = ctx->cc_major;
  out4[2] = ctx->cc_minor;
  out4[3] = ctx->l2_bytes;
  return PLNR_OK;
}

int plnr_malloc(plnr_ctx* ctx, size_t bytes, void** out) {
  PLNR_REQUIRE(ctx && out, "plnr_malloc: NULL argument");
  PLNR_CHECK_CUDA(cudaSetDevice(ctx->device));
  PLNR_CHECK_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return PLNR_OK;
}

int plnr_free(plnr_ctx* ctx, void* ptr) {
  PLNR_REQUIRE(ctx, "plnr_free: ctx is NULL");
  PLNR_CHECK_CUDA(cudaFree(ptr));
  return PLNR_OK;
}

int plnr_memcpy_h2d(plnr_ctx* ctx, void* dst, const void* src, size_t bytes) {
  PLNR_REQUIRE(ctx, "plnr_memcpy_h2d: ctx is NULL");
  PLNR_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return PLNR_OK;
}

int plnr_memcpy_d2h(plnr_ctx* ctx, void* dst, const void* src, size_t bytes) {
  PLNR_REQUIRE(ctx, "plnr_memcpy_d2h: ctx is NULL");
  PLNR_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PLNR_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return PLNR_OK;
}

int plnr_memset(plnr_ctx* ctx, void* dst, int value, size_t bytes) {
  PLNR_REQUIRE(ctx, "plnr_memset: ctx is NULL");
  PLNR_CHECK_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream));
  return PLNR_OK;
}

// ---- CUDA graph capture ---------------------------------------------------------------------

int plnr_graph_begin(plnr_ctx* ctx) {
  PLNR_REQUIRE(ctx, "plnr_graph_begin: ctx is NULL");
  PLNR_REQUIRE(!ctx->capturing, "plnr_graph_begin: capture already in progress");
  PLNR_CHECK_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  ctx->capturing = true;
  ctx->capture_launches = 0;
  return PLNR_OK;
}

int plnr_graph_end(plnr_ctx* ctx, plnr_graph** out) {
  PLNR_REQUIRE(ctx && out, "plnr_graph_end: NULL argument");
  PLNR_REQUIRE(ctx->capturing, "plnr_graph_end: no capture in progress");
  ctx->capturing = false;
  cudaGraph_t graph = nullptr;
  PLNR_CHECK_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
  plnr_graph* g = new plnr_graph();
  g->graph = graph;
  g->nodes = ctx->capture_launches;
  cudaError_t e = cudaGraphInstantiate(&g->exec, graph, 0);
  if (e != cudaSuccess) {
    plnr_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    cudaGraphDestroy(graph);
    delete g;
    return PLNR_ERR_CUDA;
  }
  *out = g;
  return PLNR_OK;
}

int plnr_graph_launch(plnr_ctx* ctx, plnr_graph* g) {
  PLNR_REQUIRE(ctx && g && g->exec, "plnr_graph_launch: NULL argument");
  PLNR_CHECK_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
  ctx->launches += g->nodes;
  return PLNR_OK;
}

int plnr_graph_destroy(plnr_graph* g) {
  if (!g) return PLNR_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return PLNR_OK;
}

// ---- events -----------------------------------------------------------------------------------

int plnr_event_create(plnr_event** out) {
  PLNR_REQUIRE(out, "plnr_event_create: out is NULL");
  plnr_event* e = new plnr_event();
  cudaError_t err = cudaEventCreate(&e->ev);
  if (err != cudaSuccess) {
    plnr_set_error("cudaEventCreate failed: %s", cudaGetErrorString(err));
    delete e;
    return PLNR_ERR_CUDA;
  }
  *out = e;
  return PLNR_OK;
}

int plnr_event_record(plnr_ctx* ctx, plnr_event* ev) {
  PLNR_REQUIRE(ctx && ev, "plnr_event_record: NULL argument");
  PLNR_CHECK_CUDA(cudaEventRecord(ev->ev, ctx->stream));
  return PLNR_OK;
}

int plnr_event_elapsed_ms(plnr_event* start, plnr_event* stop, float* ms) {
  PLNR_REQUIRE(start && stop && ms, "plnr_event_elapsed_ms: NULL argument");
  PLNR_CHECK_CUDA(cudaEventSynchronize(stop->ev));
  PLNR_CHECK_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
  return PLNR_OK;
}

int plnr_event_destroy(plnr_event* ev) {
  if (!ev) return PLNR_OK;
  if (ev->ev) cudaEventDestroy(ev->ev);
  delete ev;
  return PLNR_OK;
}

// ---- conv dispatch ------------------------------------------------------------------------------

static int validate_conv(const plnr_conv_desc* d, const plnr_tensor* x, const void* w, const plnr_tensor* y) {
  PLNR_REQUIRE(d && x && w && y, "conv2d: NULL argument");
  PLNR_REQUIRE(d->dtype == PLNR_F32 || d->dtype == PLNR_F16, "conv2d: unsupported dtype code %d", d->dtype);
  PLNR_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->stride_h >= 1 && d->stride_w >= 1 && d->dil_h >= 1 && d->dil_w >= 1,
               "conv2d: kernel/stride/dilation must be >= 1");
  PLNR_REQUIRE(d->groups >= 1 && x->c % d->groups == 0 && y->c % d->groups == 0,
               "conv2d: groups=%d must divide Cin=%d and Cout=%d", d->groups, x->c, y->c);
  PLNR_REQUIRE(d->pad_t >= 0 && d->pad_l >= 0 && d->pad_b >= 0 && d->pad_r >= 0, "conv2d: negative padding");
  // The reference's pad() allocates top/left on BOTH sides and ignores bottom/right (planer/util.py:4-10,
  // SURVEY App. D Q1): results are well defined only for bottom <= top and right <= left.
  PLNR_REQUIRE(d->pad_b <= d->pad_t && d->pad_r <= d->pad_l,
               "conv2d: pads (t=%d,l=%d,b=%d,r=%d) with bottom>top or right>left are undefined in the reference "
               "(planer/util.py:4-10) and rejected here",
               d->pad_t, d->pad_l, d->pad_b, d->pad_r);
  int oh = plnr_out_size(x->h, d->pad_t, d->pad_b, d->kh, d->dil_h, d->stride_h);
  int ow = plnr_out_size(x->w, d->pad_l, d->pad_r, d->kw, d->dil_w, d->stride_w);
  PLNR_REQUIRE(oh >= 1 && ow >= 1, "conv2d: empty output (%d x %d)", oh, ow);
  PLNR_REQUIRE(y->n == x->n && y->h == oh && y->w == ow, "conv2d: y is (%d,%d,%d) but the problem gives (%d,%d,%d)",
               y->n, y->h, y->w, x->n, oh, ow);
  PLNR_REQUIRE(x->ld >= x->coff + x->c && y->ld >= y->coff + y->c, "conv2d: view exceeds its row pitch");
  return PLNR_OK;
}

int plnr_conv2d_algo(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y) {
  if (!d || !x || !y) return PLNR_ERR_INVALID;
  return plnr_conv2d_tcgen05_supported(d, x, y) ? PLNR_ALGO_TCGEN05 : PLNR_ALGO_DIRECT;
}

int plnr_conv2d_fwd(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w, const plnr_tensor* y,
                    const plnr_epilogue* ep) {
  PLNR_REQUIRE(ctx, "conv2d: ctx is NULL");
  int rc = validate_conv(d, x, w, y);
  if (rc != PLNR_OK) return rc;
  if (ep && ep->residual) {
    const plnr_tensor* r = ep->residual;
    PLNR_REQUIRE(r->n == y->n && r->h == y->h && r->w == y->w && r->c == y->c,
                 "conv2d: residual shape differs from the output shape");
  }
  if (ep && ep->out_nchw) {
    PLNR_REQUIRE(!ep->residual, "conv2d: out_nchw cannot be combined with a residual operand");
    PLNR_REQUIRE(d->algo != PLNR_ALGO_DIRECT && plnr_conv2d_shift_supported(d, x, y),
                 "conv2d: out_nchw needs the shift-GEMM kernel (plnr_conv2d_out_nchw_supported)");
    return plnr_conv2d_shift(ctx, d, x, w, y, ep);
  }
  if (ep && ep->pool_sum) {
    PLNR_REQUIRE(d->algo != PLNR_ALGO_DIRECT && plnr_conv2d_pool_parts(d, x, y) > 0,
                 "conv2d: pool_sum needs the shift-GEMM kernel with whole 32-position parts per image (plnr_conv2d_pool_parts)");
    return plnr_conv2d_shift(ctx, d, x, w, y, ep);
  }
  bool tc_ok = plnr_conv2d_tcgen05_supported(d, x, y);
  if (ep && ep->out_f32) {
    PLNR_REQUIRE(d->dtype == PLNR_F16 && tc_ok && d->algo != PLNR_ALGO_DIRECT,
                 "conv2d: out_f32 needs fp16 split operands on the tensor-core path (groups==1, Cin%%16==0, 16B-aligned x)");
    return plnr_conv2d_tcgen05(ctx, d, x, w, y, ep);
  }
  if (d->algo == PLNR_ALGO_TCGEN05 && !tc_ok) {
    plnr_set_error("conv2d: PLNR_ALGO_TCGEN05 requested but the problem is not eligible "
                   "(needs fp16, groups==1, Cin%%16==0, 16B-aligned views)");
    return PLNR_ERR_UNSUPPORTED;
  }
  if (d->algo != PLNR_ALGO_DIRECT && tc_ok) return plnr_conv2d_tcgen05(ctx, d, x, w, y, ep);
  return plnr_conv2d_direct(ctx, d, x, w, y, ep);
}

int plnr_conv2d_out_nchw_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y) {
  if (!d || !x || !y || d->dtype != PLNR_F16 || d->groups != 1 || d->algo == PLNR_ALGO_DIRECT) return 0;
  return plnr_conv2d_shift_supported(d, x, y) ? 1 : 0;
}

int plnr_conv2d_pool_parts(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y) {
  if (!d || !x || !y || d->dtype != PLNR_F16 || d->groups != 1 || d->algo == PLNR_ALGO_DIRECT) return 0;
  return plnr_conv2d_shift_pool_parts(d, x, y);
}

int plnr_conv2d_shortcut_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* x2, int stride2,
                                   const plnr_tensor* y) {
  if (!d || !x || !x2 || !y) return 0;
  if (d->dtype != PLNR_F16 || d->groups != 1 || d->algo == PLNR_ALGO_DIRECT) return 0;
  return plnr_conv2d_shift_shortcut_supported(d, x, x2, stride2, y) ? 1 : 0;
}

int plnr_conv2d_shortcut_fwd(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w_cat,
                             const plnr_tensor* x2, int stride2, const plnr_tensor* y, const plnr_epilogue* ep) {
  PLNR_REQUIRE(ctx && x2 && x2->ptr, "conv2d_shortcut: NULL argument");
  int rc = validate_conv(d, x, w_cat, y);
  if (rc != PLNR_OK) return rc;
  PLNR_REQUIRE(!(ep && ep->residual), "conv2d_shortcut: the shortcut replaces the residual operand");
  PLNR_REQUIRE(plnr_conv2d_shortcut_supported(d, x, x2, stride2, y), "conv2d_shortcut: problem not eligible "
               "(needs the fp16 stride-1 shift kernel, shortcut channels %% 64 == 0, stride 1 or 2, matching output size)");
  return plnr_conv2d_shift(ctx, d, x, w_cat, y, ep, x2, stride2);
}

int plnr_dense_fwd(plnr_ctx* ctx, int dtype, const void* x, const void* w, void* y, int m, int n, int k,
                   const plnr_epilogue* ep, int algo) {
  PLNR_REQUIRE(ctx && x && w && y, "dense: NULL argument");
  PLNR_REQUIRE(m >= 1 && n >= 1 && k >= 1, "dense: empty problem %d x %d x %d", m, n, k);
  // x[M,K] is an image of M pixels (1 x 1 x M) with K channels; w[N,K] is a packed 1x1 filter.
  plnr_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.dtype = dtype;
  d.kh = d.kw = 1;
  d.stride_h = d.stride_w = d.dil_h = d.dil_w = 1;
  d.groups = 1;
  d.algo = algo;
  plnr_tensor tx = {const_cast<void*>(x), 1, 1, m, k, k, 0};
  plnr_tensor ty = {y, 1, 1, m, n, n, 0};
  return plnr_conv2d_fwd(ctx, &d, &tx, w, &ty, ep);
}

}  // extern "C"

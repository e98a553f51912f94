/* planer_b200.h -- C ABI of libplaner_b200.so: the B200 (sm_100a) forward hot path of Planer.
 *
 * The reference (Image-Py/planer @ 39174495) is pure Python and has NO foreign-function boundary;
 * its only native call site is the destination-passing cuDNN call in planer/util.py:74-76.  This
 * ABI is modelled on that call: every tensor argument is a raw device pointer plus explicit
 * extents, outputs are pre-allocated by the caller, the library never allocates or frees user
 * tensors.  Each entry point names the reference function it replaces (file:line under
 * /root/reference).  A host binds it with ctypes (planer_b200/_capi.py; INTEGRATION.md shows the
 * stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = error; plnr_last_error() gives a thread-local text.
 *   - no exceptions or aborts cross the boundary.
 *   - a plnr_ctx binds one device and one stream; it is not thread-safe; different ctxs are
 *     independent.  All work is enqueued on the ctx stream and is asynchronous.
 *   - INTERNAL ACTIVATION LAYOUT IS NHWC ("pixel-major"): a tensor of logical shape (N,C,H,W) is
 *     stored as N*H*W pixel rows, each `ld` elements long, the tensor's channels occupying
 *     [coff, coff+C) of the row.  (The reference itself hands non-contiguous CNHW views between
 *     layers -- planer/util.py:44 -- so inter-layer layout is not part of its contract; NCHW is
 *     restored at the graph boundary by plnr_nhwc_to_nchw.)
 *   - dtype codes: PLNR_F32 / PLNR_F16 for everything the path computes on; PLNR_U8 is accepted as the SOURCE dtype of
 *     the graph-entry functions (plnr_nchw_to_nhwc, plnr_stem_pack, plnr_stem_pool_fwd_u8, plnr_cast): the numpy reference
 *     promotes a uint8 image against float weights (planer/layer.py:22-26 on a uint8 x), i.e. computes on x.astype(w.dtype).
 */
#ifndef PLANER_B200_H
#define PLANER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLNR_ABI_VERSION 3

enum { PLNR_F32 = 0, PLNR_F16 = 1, PLNR_U8 = 2 /* graph INPUTS only: uint8 images, converted by the input-time kernel */ };
enum { PLNR_ACT_NONE = 0, PLNR_ACT_RELU = 1, PLNR_ACT_LEAKY = 2, PLNR_ACT_SIGMOID = 3 };
enum { PLNR_ALGO_AUTO = 0, PLNR_ALGO_TCGEN05 = 1, PLNR_ALGO_DIRECT = 2 };
enum {
  PLNR_OK = 0, PLNR_ERR_INVALID = -1, PLNR_ERR_UNSUPPORTED = -2, PLNR_ERR_CUDA = -3, PLNR_ERR_DRIVER = -4
};
/* elementwise op codes for plnr_eltwise */
enum {
  PLNR_EW_RELU = 0,        /* y = x * (x > 0)                    planer/layer.py:44-46  */
  PLNR_EW_LEAKY = 1,       /* y = x * ((x>0)*(1-a) + a)          planer/layer.py:48-51  */
  PLNR_EW_SIGMOID = 2,     /* y = 1 / (1 + exp(-x))              planer/layer.py:61-64  */
  PLNR_EW_ADD = 3,         /* y = x + x2                         planer/layer.py:93-95  */
  PLNR_EW_SCALE_SHIFT = 4, /* y = x * K[c] + B[c]  (folded BN)   planer/layer.py:125-127 */
  PLNR_EW_CLIP = 5,        /* y = max(min(x, b), a)              planer/layer.py:247-251 (plnr_unary2) */
  PLNR_EW_HARDSIGMOID = 6  /* y = max(min(x*a + b, 1), 0)        planer/layer.py:66-69   (plnr_unary2) */
};

typedef struct plnr_ctx plnr_ctx;
typedef struct plnr_graph plnr_graph;
typedef struct plnr_event plnr_event;

/* A pixel-major (NHWC) activation view: see "INTERNAL ACTIVATION LAYOUT" above. */
typedef struct {
  void* ptr;      /* device pointer to pixel row 0, channel 0 of the underlying buffer */
  int32_t n, h, w, c;
  int32_t ld;     /* elements per pixel row of the underlying buffer (>= coff + c) */
  int32_t coff;   /* first channel of this view inside the row */
} plnr_tensor;

/* Fused epilogue of conv/dense: y = act(acc * scale[c] + shift[c] + residual)  [or act(..) + residual].  The caller folds
 * bias and BatchNorm into (scale, shift) with plnr_fold_affine.  Order follows the reference graph
 * conv(+bias) -> batchnorm -> add -> relu  (planer/layer.py:26, :125-127, :93-95, :44-51). */
typedef struct {
  const float* scale;       /* [Cout] fp32 or NULL (=1) */
  const float* shift;       /* [Cout] fp32 or NULL (=0) */
  const plnr_tensor* residual; /* same (n,h,w,c) and dtype as y, or NULL */
  int32_t act;              /* PLNR_ACT_* */
  float alpha;              /* leaky slope */
  int32_t res_after_act;    /* 0: act(v + residual) (ResNet);  1: act(v) + residual (Darknet shortcut) */
  int32_t out_nchw;         /* 1: y->ptr is a DENSE NCHW array of y's logical shape -- the graph-exit transpose (planer/net.py:100
                             *    hands NCHW arrays back) folded into the epilogue.  No residual; only where
                             *    plnr_conv2d_out_nchw_supported says so (the shift-GEMM kernel). */
  int32_t out_f32;          /* 1: the fp32 path on the tensor pipe -- desc.dtype is PLNR_F16 for the OPERANDS (x from plnr_split_f32,
                             *    w_packed from plnr_pack_conv_weight_split), the fp32 accumulator is written as fp32: y and residual
                             *    are fp32 tensors.  Runs the TMA-im2col kernel. */
  float acc_scale;          /* accumulator multiplier applied before scale/shift (the inverse of the operands' power-of-two
                             *    pre-scales); 0 means 1 */
  const float* acc_scale_dev; /* NULL, or a DEVICE scalar multiplied in as well (the per-call activation pre-scale left by
                             *    plnr_split_f32 in dyn[1]) */
  float* pool_sum;          /* (ABI 3) NULL, or the GlobalAveragePool that consumes this conv's output folded into its epilogue
                             *    (planer/layer.py:77-78 after :22-26/:125-127/:93-95/:44-46): fp32 [n][parts][cout] with parts =
                             *    plnr_conv2d_pool_parts(...) > 0 -- the SUM of the finished (fp16-rounded) outputs over each
                             *    32-position part of an image, one writer per element (deterministic); y is NOT written.
                             *    plnr_pooled_dense_fwd adds the parts, divides by h*w and applies the Dense layer. */
} plnr_epilogue;

typedef struct {
  int32_t dtype;            /* PLNR_F32 | PLNR_F16: dtype of x, w_packed, y, residual */
  int32_t kh, kw;
  int32_t pad_t, pad_l, pad_b, pad_r;
  int32_t stride_h, stride_w, dil_h, dil_w;
  int32_t groups;
  int32_t algo;             /* PLNR_ALGO_* */
} plnr_conv_desc;

/* ---- library / context ------------------------------------------------------------------ */
int plnr_abi_version(void);
const char* plnr_last_error(void);
/* stream: a cudaStream_t to adopt (e.g. torch's current stream), or NULL to create an own one. */
int plnr_create(int device, void* stream, plnr_ctx** out);
int plnr_destroy(plnr_ctx* ctx);
int plnr_set_stream(plnr_ctx* ctx, void* stream);
int plnr_stream_sync(plnr_ctx* ctx);
/* number of kernels this ctx has enqueued so far (graph replays add the graph's node count). */
int plnr_launch_count(plnr_ctx* ctx, int64_t* out);
/* sm count, compute capability major/minor, L2 bytes -> out[4] */
int plnr_device_info(plnr_ctx* ctx, int64_t* out4);

/* ---- raw memory for hosts that do not bring their own allocator (replaces np.zeros / np.asarray /
 *      .get() of the backend module: planer/net.py:21,98,100) -------------------------------- */
int plnr_malloc(plnr_ctx* ctx, size_t bytes, void** out);
int plnr_free(plnr_ctx* ctx, void* ptr);
int plnr_memcpy_h2d(plnr_ctx* ctx, void* dst, const void* src, size_t bytes);
int plnr_memcpy_d2h(plnr_ctx* ctx, void* dst, const void* src, size_t bytes);
int plnr_memset(plnr_ctx* ctx, void* dst, int value, size_t bytes);

/* ---- layout / dtype --------------------------------------------------------------------- */
/* NCHW (dense, x_dtype) -> pixel-major view y (y_dtype); channels [c, y.c) of y are zero-filled
 * (channel padding for the tensor-core path).  Graph entry: planer/net.py:96-98. */
int plnr_nchw_to_nhwc(plnr_ctx* ctx, const void* x, int x_dtype, int c_src, const plnr_tensor* y, int y_dtype);
/* First-layer input packing (fp16 out): NCHW x (n,c,h,w; few channels) -> y[n, h2, ow, (ph*kw + sx)*c + ci] =
 * x[n, ci, stride*h2 + ph, stride*ow + sx - pad_l] (zero outside), so that a kh x kw / stride-s convolution of the
 * graph input (planer/layer.py:22-26 on the raw image) runs as a (T x 1) stride-1 convolution over y with the
 * filter taps re-ordered accordingly by the host (planer_b200/executor.py).  y->c >= stride*kw*c, multiple of 8. */
int plnr_stem_pack(plnr_ctx* ctx, const void* x, int x_dtype, int n, int c, int h, int w, const plnr_tensor* y, int kw,
                   int stride, int pad_l);
/* Fused first layer (fp16): NCHW x (n, 3, h, w) -> kh x kw / stride-2 convolution (planer/layer.py:22-26 +
 * planer/util.py:17-44) -> *scale + shift (bias / folded BatchNorm, planer/layer.py:125-127) -> ReLU
 * (planer/layer.py:44-46) -> 3x3 / stride-2 / pad-1 Maxpool (planer/layer.py:71-72 + planer/util.py:79-95; after a
 * ReLU its zero padding and -1e4 floor are neutral) -> pixel-major y (n, poh, pow, 64), in ONE kernel
 * (csrc/stem_pool.cu).  w_packed: [64][T][64] fp16 with W[co, e, (ph*3 + c)*8 + sx + col_shift] =
 * K[co, c, 2(e + e_min) + ph + pad_t, sx] (zero elsewhere); e_min, T and col_shift from plnr_stem_pool_geometry.
 * plnr_stem_pool_supported returns 1 when the fused kernel applies; otherwise the caller runs plnr_stem_pack /
 * plnr_conv2d_fwd / plnr_maxpool2d. */
int plnr_stem_pool_supported(int dtype, int c, int h, int w, int cout, int kh, int kw, int stride, int pad_t, int pad_l,
                             int pad_b, int pad_r, int act, int pool_k, int pool_stride, int pool_pad);
int plnr_stem_pool_geometry(int kh, int pad_t, int pad_l, int* e_min, int* taps, int* col_shift);
int plnr_stem_pool_fwd(plnr_ctx* ctx, const void* x, int n, int c, int h, int w, const void* w_packed,
                       const float* scale, const float* shift, int kh, int kw, int stride, int pad_t, int pad_l,
                       int pad_b, int pad_r, int act, int pool_k, int pool_stride, int pool_pad, const plnr_tensor* y);
/* The same with a uint8 image x (n, 3, h, w): the producer warps convert while staging (exact: 0..255 are fp16 numbers).
 * Replaces planer/net.py:96-98 + planer/layer.py:22-26 on a uint8 array (numpy promotes uint8 x float16 to float16). */
int plnr_stem_pool_fwd_u8(plnr_ctx* ctx, const void* x, int n, int c, int h, int w, const void* w_packed,
                          const float* scale, const float* shift, int kh, int kw, int stride, int pad_t, int pad_l,
                          int pad_b, int pad_r, int act, int pool_k, int pool_stride, int pool_pad, const plnr_tensor* y);
/* Small first layer on the CUDA cores (csrc/stem_direct.cu): 3x3 / stride-1 / pad-1 convolution of an NCHW image x (n, c <= 3,
 * h, w; x_dtype PLNR_F16 or PLNR_U8) to y->c <= 32 channels (multiple of 8) + *scale + shift + activation -> pixel-major y.
 * Replaces Conv2d (planer/layer.py:22-26 + planer/util.py:17-44) -> BatchNorm (:125-127) -> ReLU / LeakyReLU (:44-51) at
 * the head of YOLOv3-style networks, where K = 27 and N = 32 leave the tensor cores nothing to do.  w_oihw: HOST pointer
 * to the fp16 OIHW filter (y->c, c, 3, 3); scale / shift: HOST fp32 [y->c] or NULL -- 1.7 KB that travel in the kernel
 * parameters (the caller reads them back from the device once, at load time). */
int plnr_stem3x3_supported(int dtype, int c, int cout, int kh, int kw, int stride, int pad_t, int pad_l, int pad_b, int pad_r,
                           int dil);
int plnr_stem3x3_fwd(plnr_ctx* ctx, const void* x, int x_dtype, int n, int c, int h, int w, const void* w_oihw,
                     const float* scale, const float* shift, int act, float alpha, const plnr_tensor* y);
/* pixel-major view x -> NCHW dense y.  Graph exit: planer/net.py:100. */
int plnr_nhwc_to_nchw(plnr_ctx* ctx, const plnr_tensor* x, int x_dtype, void* y, int y_dtype);
/* flat cast, n elements (Net.half, planer/net.py:26-29). */
int plnr_cast(plnr_ctx* ctx, const void* x, int x_dtype, void* y, int y_dtype, int64_t n);
/* OIHW weight (w_dtype) -> packed [Cout][kh][kw][cin_pad] (out_dtype), zero-filling padded input
 * channels.  One-off at load (K = core.reshape(Co,-1) of planer/util.py:41, re-ordered so the
 * contraction index is (r,s,c) with c innermost). */
int plnr_pack_conv_weight(plnr_ctx* ctx, const void* w, int w_dtype, void* out, int out_dtype,
                          int cout, int cin_g, int kh, int kw, int cin_pad);
/* The float32 Conv2d / Dense of a float32 net (planer/layer.py:22-26, :15-18 on float32 arrays) on the fp16 tensor pipe:
 * both operands are split into two fp16 numbers, v * prescale = hi + lo, and ONE fp16 implicit GEMM over 3C channels per
 * filter tap computes xh.wh + xh.wl + xl.wh with an fp32 accumulator (csrc/split_f32.cu).
 *   plnr_split_f32:              x fp32 (n,h,w,C) view -> xs fp16 DENSE (n,h,w,cs), cs >= 3C: [hi C | hi C | lo C | pad];
 *                                pad channels are left untouched (zero them once).  dyn == NULL: x is multiplied by
 *                                `prescale`.  dyn = two device floats: the power-of-two pre-scale is derived from max|x| on the
 *                                device (dyn[0] = max|x|, dyn[1] = 1 / prescale for plnr_epilogue.acc_scale_dev).
 *   plnr_pack_conv_weight_split: OIHW fp32 -> [Cout][kh][kw][cs] fp16 with [hi C | lo C | hi C | pad] per tap (pad untouched).
 *   plnr_absmax_f32:             out[0] = max |x[i]| (device scalar; the caller picks the weights' power-of-two pre-scale).
 * Then plnr_conv2d_fwd(desc.dtype = PLNR_F16, x = xs, ep.out_f32 = 1, ep.acc_scale = 1 / prescale_w, ep.acc_scale_dev = dyn + 1). */
int plnr_split_f32(plnr_ctx* ctx, const plnr_tensor* x, const plnr_tensor* xs, float prescale, float* dyn);
int plnr_pack_conv_weight_split(plnr_ctx* ctx, const void* w, void* out, int cout, int cin, int kh, int kw, int cs,
                                float prescale);
int plnr_absmax_f32(plnr_ctx* ctx, const void* x, int64_t n, float* out);
/* scale[c] = bn_k ? bn_k[c] : 1 ; shift[c] = (bias ? bias[c] : 0) * scale[c] + (bn_b ? bn_b[c] : 0).
 * Inputs have dtype `dtype`; outputs are fp32.  Folds planer/layer.py:26 and :125-127. */
int plnr_fold_affine(plnr_ctx* ctx, const void* bias, const void* bn_k, const void* bn_b, int dtype,
                     float* scale, float* shift, int c);

/* ---- the FLOPs --------------------------------------------------------------------------- */
/* Conv2d forward.  Replaces planer/layer.py:22-26 (Conv2d) + planer/util.py:17-44 (conv_for:
 * zero-pad, im2col, one GEMM M=Co K=C*kh*kw N=N*oh*ow) and, through `ep`, the following
 * batchnorm / add / relu layers.  x: (n,h,w,c) view; w_packed from plnr_pack_conv_weight with
 * cin_pad == x.c / groups; y: (n,oh,ow,cout) view, oh/ow by the formula of planer/util.py:25-26.
 * fp16 + groups==1 + x.c%16==0 runs the TMA-im2col / tcgen05 implicit GEMM; everything else runs
 * the direct CUDA-core kernel (fp32 accumulate in both). */
int plnr_conv2d_fwd(plnr_ctx* ctx, const plnr_conv_desc* desc, const plnr_tensor* x, const void* w_packed,
                    const plnr_tensor* y, const plnr_epilogue* ep);
int plnr_conv2d_out_nchw_supported(const plnr_conv_desc* desc, const plnr_tensor* x, const plnr_tensor* y);
/* parts per image of plnr_epilogue.pool_sum for this problem, or 0 when the pooling fold does not apply (needs the fp16
 * stride-1 shift-GEMM kernel, cout % 32 == 0 and a padded image grid of a multiple of 32 positions, e.g. 7x7 + pad 1). */
int plnr_conv2d_pool_parts(const plnr_conv_desc* desc, const plnr_tensor* x, const plnr_tensor* y);
/* Conv2d + fused 1x1 shortcut convolution: y = act((conv(x, W) + conv1x1_stride(x2, W2)) * scale + shift).  Replaces the
 * tail of a down-sampling residual block -- Conv2d -> BatchNorm on the main path, Conv2d(1x1, stride) -> BatchNorm on
 * the shortcut, Add, ReLU (planer/layer.py:22-26, :125-127, :93-95, :44-46) -- by ONE launch: the shortcut's k-chunks
 * accumulate into the same tensor-memory tile.  w_cat: [Cout][kh*kw*Cin + C2] = plnr_pack_conv_weight(W) with
 * W2[co, c2] * (scale2[co] / scale[co]) appended along K (the caller folds the two BatchNorm scales; shift = shift +
 * shift2).  x2: (n, C2, h2, w2) with ceil(h2 / stride2) == y.h; needs the fp16 stride-1 tensor-core path. */
int plnr_conv2d_shortcut_supported(const plnr_conv_desc* desc, const plnr_tensor* x, const plnr_tensor* x2, int stride2,
                                   const plnr_tensor* y);
int plnr_conv2d_shortcut_fwd(plnr_ctx* ctx, const plnr_conv_desc* desc, const plnr_tensor* x, const void* w_cat,
                             const plnr_tensor* x2, int stride2, const plnr_tensor* y, const plnr_epilogue* ep);
/* Dense forward  y[M,N] = x[M,K] @ w[N,K]^T (+ epilogue).  Replaces planer/layer.py:15-18 (Dense);
 * w is the reference's (out,in) matrix as stored.  Runs as a 1x1 convolution over M pixels. */
int plnr_dense_fwd(plnr_ctx* ctx, int dtype, const void* x, const void* w, void* y, int m, int n, int k,
                   const plnr_epilogue* ep, int algo);

/* ---- HBM-bound companions ---------------------------------------------------------------- */
/* Maxpool with the reference's semantics: ZERO padding and a -1e4 floor (planer/util.py:79-95). */
int plnr_maxpool2d(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y,
                   int kh, int kw, int pad_t, int pad_l, int stride_h, int stride_w);
/* Bilinear resize to any output size (planer/util.py:194-210, upsample_size; fractional ONNX scales).  The per-row and
 * per-column tables (device pointers; y->h and y->w entries) hold the lower source index, the weight of index + 1 and the
 * weight of the index itself, computed by the caller in the image dtype as the reference does. */
int plnr_resize_linear(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, const int* row_lo, const float* row_w,
                       const float* row_w1, const int* col_lo, const float* col_w, const float* col_w1);
/* AveragePool with the reference's semantics: ZERO padding, divisor kh*kw whatever the window covers
 * (planer/layer.py:74-75 -> planer/util.py:97-100). */
int plnr_avgpool2d(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y,
                   int kh, int kw, int pad_t, int pad_l, int stride_h, int stride_w);
/* The two data movements of ConvTranspose2d (planer/layer.py:28-34), which then runs plnr_conv2d_fwd at stride 1:
 * zero_stuff: y[n, lo_h + i*stride_h, lo_w + j*stride_w, :] = x[n, i, j, :], zero elsewhere (planer/layer.py:32-33);
 * flip_weight: K (ci, co, kh, kw) -> K' (co, ci, kh, kw) = K.transpose(1,0,2,3)[:, :, ::-1, ::-1] (planer/layer.py:34). */
int plnr_zero_stuff(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, int lo_h, int lo_w,
                    int stride_h, int stride_w);
int plnr_flip_weight(plnr_ctx* ctx, int dtype, const void* w, void* out, int ci, int co, int kh, int kw);
/* Integer-factor nearest upsample, zero pixel shift (planer/util.py:184-192 with the default mode
 * strings of planer/util.py:212). */
int plnr_upsample_nearest(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, int fh, int fw);
/* Bilinear upsample by integer factors >= 2 (planer/util.py:121-153, upsample_blinear): edge-replicated input, 4-tap blend
 * with the (4, fh*fw) fp32 weight table `wmat` (device pointer; the reference's make_upmat, planer/util.py:121-131). */
int plnr_upsample_linear(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y, int fh, int fw,
                         const float* wmat);
/* Channel-slice copy x -> y (same n,h,w,c; different ld/coff): the building block of
 * np.concatenate(axis=1) (planer/layer.py:90-91). */
int plnr_copy_channels(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const plnr_tensor* y);
/* Elementwise family on dense pixel-major data of `npix` rows x `c` channels (ld == c).
 * p0/p1: second operand (ADD: x2) or per-channel K/B of dtype `dtype` (SCALE_SHIFT). */
int plnr_eltwise(plnr_ctx* ctx, int op, int dtype, const void* x, const void* p0, const void* p1, void* y,
                 int64_t npix, int c, float alpha);
/* Unary operators with two scalar parameters on `n` contiguous elements: PLNR_EW_CLIP (a = min, b = max),
 * PLNR_EW_HARDSIGMOID (a = alpha, b = beta). */
int plnr_unary2(plnr_ctx* ctx, int op, int dtype, const void* x, void* y, int64_t n, float a, float b);
/* Softmax over the last, contiguous axis of a dense (rows, c) array (planer/layer.py:141-146): the channel axis of
 * pixel-major activations, or the class axis of 2-D logits. */
int plnr_softmax(plnr_ctx* ctx, int dtype, const void* x, void* y, int64_t rows, int c);
/* Global average pool (n, hw, c) -> (n, c), fp32 accumulate (planer/layer.py:77-78). */
int plnr_global_avgpool(plnr_ctx* ctx, int dtype, const plnr_tensor* x, void* y);

/* GlobalAveragePool -> Flatten -> Dense fused (planer/layer.py:77-78, :59, :15-18): y[n, o] = act((mean_hw x[n, :, :] . w[o, :]) *
 * scale[o] + shift[o]); w is the reference's (out, in) matrix in dtype `dtype`; y is (n, out_features) of the same dtype. */
int plnr_gap_dense_fwd(plnr_ctx* ctx, int dtype, const plnr_tensor* x, const void* w, const float* scale,
                       const float* shift, void* y, int out_features, int act, float alpha);
/* The same tail when the pooling was folded into the producing convolution (plnr_epilogue.pool_sum): pool is fp32
 * [n][parts][c] of partial sums, hw the number of pooled positions (planer/layer.py:77-78 divides by h*w); fp16 w / y. */
int plnr_pooled_dense_fwd(plnr_ctx* ctx, const float* pool, int n, int parts, int c, int hw, const void* w, const float* scale,
                          const float* shift, void* y, int out_features, int act, float alpha);

/* ---- CUDA-graph capture of a planned forward (replaces the Python interpreter loop of
 *      planer/net.py:43-70 by one replayable launch) ------------------------------------- */
int plnr_graph_begin(plnr_ctx* ctx);
int plnr_graph_end(plnr_ctx* ctx, plnr_graph** out);
int plnr_graph_launch(plnr_ctx* ctx, plnr_graph* g);
int plnr_graph_destroy(plnr_graph* g);

/* ---- device timers (feeds Net.timer, planer/net.py:67-70, with real device time) -------- */
int plnr_event_create(plnr_event** out);
int plnr_event_record(plnr_ctx* ctx, plnr_event* ev);
int plnr_event_elapsed_ms(plnr_event* start, plnr_event* stop, float* ms);   /* syncs on stop */
int plnr_event_destroy(plnr_event* ev);

/* ---- diagnostics -------------------------------------------------------------------------- */
/* Debug aid: per-CTA cycle counters of the last tensor-core conv launch, 8 int64 per CTA:
 * [0] producer wait-for-empty, [1] producer total, [2] MMA wait-for-full, [3] MMA wait-for-accumulator,
 * [4] MMA total, [5] epilogue wait-for-accumulator, [6] epilogue total.  enable=1 arms it, out (host) may be NULL. */
int plnr_debug_conv_profile(plnr_ctx* ctx, int enable, int64_t* out, int n);
/* Family name of the kernel the most recent call on this ctx launched ("conv2d_shift", "conv2d_stack", "conv2d_tcgen05",
 * "conv2d_direct", "stem_pool", "gap_dense", ...): lets a host label a per-launch timing table (bench.py). */
int plnr_last_kernel(plnr_ctx* ctx, char* out, int n);
/* Test aid: virtual grid and per-plane tap tables the shift-GEMM kernel (csrc/conv_shift.cu) uses for a problem; out[0] = 0
 * when the kernel does not apply.  Layout in the source.  x->ptr only needs to be 16-byte aligned (not dereferenced). */
int plnr_debug_shift_geometry(const plnr_conv_desc* desc, const plnr_tensor* x, const plnr_tensor* y, int* out, int n);
/* Which kernel plnr_conv2d_fwd would pick for this problem: PLNR_ALGO_TCGEN05 or PLNR_ALGO_DIRECT. */
int plnr_conv2d_algo(const plnr_conv_desc* desc, const plnr_tensor* x, const plnr_tensor* y);

#ifdef __cplusplus
}
#endif
#endif /* PLANER_B200_H */

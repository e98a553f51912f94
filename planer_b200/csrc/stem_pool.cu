// Fused first layer of an image network for sm_100a:
//     NCHW fp16 image (3 channels) -> kh x kw / stride-2 convolution -> *scale + shift (folded bias / BatchNorm) -> ReLU
//     -> 3x3 / stride-2 / pad-1 max pooling -> pixel-major (NHWC) fp16
// in ONE kernel.  Replaces, for the ResNet stem, the reference chain  Conv2d (planer/layer.py:22-26 + planer/util.py:17-44)
// -> BatchNorm (planer/layer.py:125-127) -> ReLU (planer/layer.py:44-46) -> Maxpool (planer/layer.py:71-72 +
// planer/util.py:79-95) and the three separate launches this library used before (plnr_stem_pack -> conv -> maxpool):
// 38.5 MB read + 51 MB written per 128 images instead of ~920 MB of packed / conv-output round trips through HBM.
//
// Contraction.  Conv output row h, column ow, channel co:
//     y[h, ow, co] = sum_{r, sx, c} x[c, 2h + r - pad_t, 2ow + sx - pad_l] * K[co, c, r, sx]
// Input rows are taken in PAIRS t = (2t, 2t+1) ("packed rows"); a packed row is one A operand tile
//     A_t[ow, ph*24 + sx*3 + c] = x[c, 2t + ph, 2ow + sx - pad_l]          (112 pixels x 48 k, K-major, 128B swizzle)
// written to shared memory by producer warps (coalesced global loads -> channel-interleaved staging row -> one 48-byte
// run per (pixel, phase)); row h then is  D_h[128 pixels, 64 co] = sum_{e < T} A_{h + e + e_min} * W_e^T  with
// W_e[co, ph*24 + sx*3 + c] = K[co, c, 2(e + e_min) + ph + pad_t, sx]: T * 3 tcgen05.mma (M=128, N=64, K=16) per
// output row, every packed row being reused by T output rows straight from shared memory.
//
// Pooling.  Post-ReLU values are >= 0, so the reference's zero padding and -1e4 floor (planer/util.py:82,87-88) are
// neutral: pooled[p, q] = max over conv rows 2p-1..2p+1 and columns 2q-1..2q+1 that exist.  The epilogue warps keep the
// last three conv rows (fp16, swizzled) in shared memory and emit one pooled row per two conv rows with 16-byte stores.
//
// Work item = (image, band of PB pooled rows); persistent grid, items strided over CTAs.  16 warps:
//   warp 0       tcgen05.mma issuer (one elected lane)       warps 4-11   epilogue + pooling (TMEM lane quarter = warp % 4)
//   warp 1       TMEM allocator                              warps 12-15  producers, one packed row per warp at a time
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kAcc = 8;                 // TMEM accumulators (64 fp32 columns each)
constexpr int kMaxT = 4;                // packed-row taps per output row
constexpr int kMaxRing = 10;
constexpr int kProducerWarps = 4;
constexpr long long kWatchdogCycles = 4000000000ll;

struct StemParams {
  const __half* x;                      // NCHW
  int N, H, W, OH, OW, POH, POW;
  int pad_l, e_min, T;
  int PB, bands, items;
  int ring;                             // A tiles in the shared-memory ring
  uint32_t tile_bytes;                  // round_up(OW * 128, 1024)
  uint32_t srow_bytes;                  // one channel-interleaved staging row
  int npos;                             // staging positions per row
  const __half* w;                      // [64][T * 64] packed filter (K-major)
  const float* scale; const float* shift;
  __half* y; int yld, ycoff;
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
        if (atomicCAS(err, 0, 3) == 0) { err[1] = blockIdx.x; err[2] = role; err[3] = (int)parity; }
        __threadfence();
        return;
      }
    }
  }
}

// three K=16 steps (k = 0..47) of one packed row against one filter tap
__device__ __forceinline__ void umma_f16_x3(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                            uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred p, t;\n\t.reg .b64 da, db;\n\t.reg .b32 a1, b1;\n\t"
      "setp.ne.b32 p, %5, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "add.u32 a1, %1, 2;\n\tadd.u32 b1, %2, 2;\n\tmov.b64 da, {a1, %3};\n\tmov.b64 db, {b1, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, t;\n\t"
      "add.u32 a1, %1, 4;\n\tadd.u32 b1, %2, 4;\n\tmov.b64 da, {a1, %3};\n\tmov.b64 db, {b1, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, t;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
      : "memory");
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// rows of one work item: pooled rows [p0, p0 + np), conv rows [h0, h1] (the ones that exist), packed rows from t0
struct ItemGeom { int img, p0, np, h0, h1, nrows, npk; };
__device__ __forceinline__ ItemGeom item_geom(const StemParams& p, int item) {
  ItemGeom g;
  g.img = item / p.bands;
  const int band = item - g.img * p.bands;
  g.p0 = band * p.PB;
  g.np = min(p.PB, p.POH - g.p0);
  g.h0 = max(2 * g.p0 - 1, 0);
  g.h1 = min(2 * (g.p0 + g.np - 1) + 1, p.OH - 1);
  g.nrows = g.h1 - g.h0 + 1;
  g.npk = g.nrows + p.T - 1;
  return g;
}

__global__ void __launch_bounds__(kThreads, 1) stem_pool_kernel(const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  // layout: [A ring][filter taps T x 8 KB][conv-row ring 3 tiles][staging: 4 warps x 2 rows][scale|shift][barriers]
  const uint32_t ring = (uint32_t)p.ring, tile = p.tile_bytes;
  const uint32_t off_b = ring * tile;
  const uint32_t off_rows = off_b + (uint32_t)p.T * 8192u;
  const uint32_t off_stage = off_rows + 3u * tile;
  const uint32_t off_ss = off_stage + kProducerWarps * 2u * p.srow_bytes;
  const uint32_t off_bar = off_ss + 512u;
  const uint32_t bar_afull = base + off_bar, bar_aempty = bar_afull + 8 * kMaxRing;
  const uint32_t bar_tfull = bar_aempty + 8 * kMaxRing, bar_tempty = bar_tfull + 8 * kAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + off_bar + 16 * kMaxRing + 16 * kAcc);
  float* ss = reinterpret_cast<float*>(bp + off_ss);          // scale[64] | shift[64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (uint32_t i = 0; i < ring; ++i) { ptx::mbar_init(bar_afull + 8 * i, 32); ptx::mbar_init(bar_aempty + 8 * i, 1); }
    for (uint32_t i = 0; i < kAcc; ++i) { ptx::mbar_init(bar_tfull + 8 * i, 1); ptx::mbar_init(bar_tempty + 8 * i, 256); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish(); }
  // filter taps -> shared memory, K-major rows of 128 B with the 128B swizzle tcgen05 expects (16-byte chunk ^ row % 8)
  for (int i = threadIdx.x; i < p.T * 64 * 8; i += kThreads) {
    const int e = i >> 9, co = (i >> 3) & 63, ch = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(p.w + ((size_t)co * p.T + e) * 64 + ch * 8);
    *reinterpret_cast<uint4*>(bp + off_b + e * 8192 + co * 128 + ((ch ^ (co & 7)) << 4)) = v;
  }
  if (threadIdx.x < 128) {
    const int c = threadIdx.x & 63;
    ss[threadIdx.x] = threadIdx.x < 64 ? (p.scale ? __ldg(p.scale + c) : 1.f) : (p.shift ? __ldg(p.shift + c) : 0.f);
  }
  // staging rows: the positions left / right of the image stay zero for the whole kernel (zero padding)
  for (uint32_t i = threadIdx.x; i < kProducerWarps * 2u * p.srow_bytes / 4u; i += kThreads)
    reinterpret_cast<uint32_t*>(bp + off_stage)[i] = 0u;
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ===================================== MMA issuer =========================================
    const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t desc_hi = (uint32_t)(ptx::make_smem_desc(0, 1024, 2) >> 32);
    const uint32_t a_lo0 = (uint32_t)ptx::make_smem_desc(base, 1024, 2);
    const uint32_t b_lo0 = (uint32_t)ptx::make_smem_desc(base + off_b, 1024, 2);
    const uint32_t a_step = tile >> 4, b_step = 8192u >> 4;
    const int T = p.T;
    const bool elected = ptx::elect_one();
    uint32_t pk = 0;          // packed rows consumed so far (ring position of the item's first packed row)
    uint32_t crow = 0;        // conv rows issued so far (accumulator ring)
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemGeom g = item_geom(p, item);
      for (int j = 0; j < g.nrows; ++j, ++crow) {
        // packed rows j .. j+T-1 of this item; only the newest one can still be in flight
        for (int e = (j == 0 ? 0 : T - 1); e < T; ++e) {
          const uint32_t c = pk + (uint32_t)(j + e);
          mbar_wait(bar_afull + 8 * (c % ring), (c / ring) & 1u, p.err, 2);
        }
        const uint32_t acc = crow % kAcc;
        mbar_wait(bar_tempty + 8 * acc, ((crow / kAcc) & 1u) ^ 1u, p.err, 1);
        ptx::tc_fence_after();
        if (elected) {
          const uint32_t d = tmem_base + acc * 64u;
          for (int e = 0; e < T; ++e) {
            const uint32_t slot = (pk + (uint32_t)(j + e)) % ring;
            umma_f16_x3(d, a_lo0 + slot * a_step, b_lo0 + (uint32_t)e * b_step, desc_hi, idesc, e > 0 ? 1u : 0u);
          }
          ptx::umma_commit(bar_tfull + 8 * acc);
          ptx::umma_commit(bar_aempty + 8 * ((pk + (uint32_t)j) % ring));        // oldest packed row: last use
          if (j == g.nrows - 1)
            for (int e = 1; e < T; ++e) ptx::umma_commit(bar_aempty + 8 * ((pk + (uint32_t)(j + e)) % ring));
        }
      }
      pk += (uint32_t)g.npk;
    }
  } else if (warp >= 12) {
    // ===================================== producers ==========================================
    // Warp pw builds the packed rows whose running index is congruent to pw (mod 4): four rows are in flight at once,
    // which hides the global-load latency without any cross-warp synchronisation.
    const int pw = warp - 12;
    uint8_t* srow = bp + off_stage + (uint32_t)pw * 2u * p.srow_bytes;       // [phase][position][channel]
    const int W8 = p.W >> 3, OW = p.OW;
    const int nload = 2 * W8, ngather = 2 * OW;
    uint32_t pk = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemGeom g = item_geom(p, item);
      const int t0 = g.h0 + p.e_min;
      const __half* ximg = p.x + (size_t)g.img * 3 * p.H * p.W;
      for (int i = 0; i < g.npk; ++i) {
        const uint32_t c = pk + (uint32_t)i;
        if ((int)(c & 3u) != pw) continue;
        const int t = t0 + i;
        // 1. global -> staging: lane takes 8 columns of all three channels of one input row
        for (int it = lane; it < nload; it += 32) {
          const int ph = it / W8, v = it - ph * W8;
          const int ih = 2 * t + ph;
          uint4 q[3];
          if (ih >= 0 && ih < p.H) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
              q[ch] = __ldg(reinterpret_cast<const uint4*>(ximg + ((size_t)ch * p.H + ih) * p.W + v * 8));
          } else {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) q[ch] = make_uint4(0u, 0u, 0u, 0u);
          }
          __half* dst = reinterpret_cast<__half*>(srow + (uint32_t)ph * p.srow_bytes) + (size_t)(v * 8 + p.pad_l) * 3;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const __half* hv = reinterpret_cast<const __half*>(&q[ch]);
#pragma unroll
            for (int k = 0; k < 8; ++k) dst[k * 3 + ch] = hv[k];
          }
        }
        __syncwarp();
        // 2. staging -> A tile: the 8 taps x 3 channels of (pixel, phase) are 48 contiguous staging bytes at 12 * ow
        const uint32_t slot = c % ring;
        mbar_wait(bar_aempty + 8 * slot, ((c / ring) & 1u) ^ 1u, p.err, 0);
        uint8_t* atile = bp + slot * tile;
        for (int it = lane; it < ngather; it += 32) {
          const int ph = it / OW, ow = it - ph * OW;
          const uint32_t* src = reinterpret_cast<const uint32_t*>(srow + (uint32_t)ph * p.srow_bytes + 12u * (uint32_t)ow);
          uint32_t r[12];
#pragma unroll
          for (int k = 0; k < 12; ++k) r[k] = src[k];
          uint8_t* arow = atile + (uint32_t)ow * 128u;
          const uint32_t sw = (uint32_t)(ow & 7);
#pragma unroll
          for (int k = 0; k < 3; ++k)
            *reinterpret_cast<uint4*>(arow + ((((uint32_t)(ph * 3 + k)) ^ sw) << 4)) =
                make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        }
        ptx::fence_proxy_async_smem();           // generic-proxy writes -> visible to the tensor core's async proxy
        ptx::mbar_arrive(bar_afull + 8 * slot);
        __syncwarp();                            // the staging rows are rewritten by the next packed row of this warp
      }
      pk += (uint32_t)g.npk;
    }
  } else if (warp >= 4) {
    // ===================================== epilogue + pooling =================================
    const int ew = warp & 3;                 // TMEM lane quarter
    const int eg = (warp - 4) >> 2;          // channel half: channels [32 eg, 32 eg + 32)
    const int et = threadIdx.x - 128;        // 0..255
    const int px = ew * 32 + lane;           // conv column owned by this thread
    const bool pvalid = px < p.OW;
    uint8_t* rows = bp + off_rows;
    const float* sc = ss + eg * 32, *sf = ss + 64 + eg * 32;
    const int POW = p.POW, OW = p.OW;
    uint32_t crow = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemGeom g = item_geom(p, item);
      const int njj = 2 * g.np + 1;
      for (int jj = 0; jj < njj; ++jj) {
        const int hh = 2 * g.p0 - 1 + jj;
        uint8_t* rrow = rows + (uint32_t)(jj % 3) * tile + (uint32_t)px * 128u;
        const uint32_t sw = (uint32_t)(px & 7);
        if (hh >= 0 && hh < p.OH) {
          const uint32_t acc = crow % kAcc;
          mbar_wait(bar_tfull + 8 * acc, (crow / kAcc) & 1u, p.err, 3);
          ptx::tc_fence_after();
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 64u + (uint32_t)eg * 32u, v);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          ptx::mbar_arrive(bar_tempty + 8 * acc);
          ++crow;
          if (pvalid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int ch = q * 8 + 2 * k;
                const float a = fmaxf(fmaf(__uint_as_float(v[ch]), sc[ch], sf[ch]), 0.f);
                const float b = fmaxf(fmaf(__uint_as_float(v[ch + 1]), sc[ch + 1], sf[ch + 1]), 0.f);
                o[k] = pack_half2(a, b);
              }
              *reinterpret_cast<uint4*>(rrow + ((((uint32_t)(eg * 4 + q)) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        } else if (pvalid) {                 // row above / below the conv output: pooling pad (0 == ReLU floor)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(rrow + ((((uint32_t)(eg * 4 + q)) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (jj >= 2 && (jj & 1) == 0) {
          // conv rows jj-2, jj-1, jj are complete: pooled row p0 + jj/2 - 1
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int prow = g.p0 + (jj >> 1) - 1;
          __half* yrow = p.y + ((size_t)g.img * p.POH + prow) * POW * p.yld + p.ycoff;
          for (int idx = et; idx < POW * 8; idx += 256) {
            const int q = idx >> 3, c16 = idx & 7;
            uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int cx = 2 * q + dx;
              if (cx >= 0 && cx < OW) {
                const uint32_t off = (uint32_t)cx * 128u + ((((uint32_t)c16) ^ (uint32_t)(cx & 7)) << 4);
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                  const uint4 t = *reinterpret_cast<const uint4*>(rows + (uint32_t)s * tile + off);
                  m.x = hmax2_u32(m.x, t.x); m.y = hmax2_u32(m.y, t.y); m.z = hmax2_u32(m.z, t.z); m.w = hmax2_u32(m.w, t.w);
                }
              }
            }
            *reinterpret_cast<uint4*>(yrow + (size_t)q * p.yld + c16 * 8) = m;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct StemPlan {
  bool ok;
  int OH, OW, POH, POW, e_min, T, ring, npos;
  uint32_t tile_bytes, srow_bytes;
  size_t smem_bytes;
};

static StemPlan make_plan(int c, int h, int w, int cout, int kh, int kw, int stride, int pad_t, int pad_l, int pad_b,
                          int pad_r, int pool_k, int pool_s, int pool_p) {
  StemPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (c != 3 || cout != 64 || stride != 2 || kw < 1 || kw > 8 || kh < 1 || w % 8 != 0) return pl;
  if (pool_k != 3 || pool_s != 2 || pool_p != 1) return pl;
  if (pad_t < 0 || pad_l < 0 || pad_b > pad_t + 1 || pad_r > pad_l + 1) return pl;
  pl.OH = plnr_out_size(h, pad_t, pad_b, kh, 1, stride);
  pl.OW = plnr_out_size(w, pad_l, pad_r, kw, 1, stride);
  if (pl.OH < 2 || pl.OW < 2 || pl.OW > 128) return pl;
  pl.POH = plnr_out_size(pl.OH, 1, 1, 3, 1, 2);
  pl.POW = plnr_out_size(pl.OW, 1, 1, 3, 1, 2);
  // vertical tap r reads input row 2h + r - pad_t = 2(h + e) + ph:  e = floor((r - pad_t) / 2)
  auto fl2 = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };
  pl.e_min = fl2(-pad_t);
  const int e_max = fl2(kh - 1 - pad_t);
  pl.T = e_max - pl.e_min + 1;
  if (pl.T < 1 || pl.T > kMaxT) return pl;
  pl.tile_bytes = (uint32_t)round_up(pl.OW * 128, 1024);
  pl.npos = 2 * pl.OW + 8;                          // staging positions: column j of the image sits at j + pad_l
  if (w + pad_l > pl.npos) pl.npos = w + pad_l;
  pl.srow_bytes = (uint32_t)round_up(pl.npos * 6 + 48, 16);
  const size_t fixed = (size_t)pl.T * 8192 + 3 * (size_t)pl.tile_bytes + kProducerWarps * 2 * (size_t)pl.srow_bytes + 512 +
                       16 * kMaxRing + 16 * kAcc + 64 + 1024 + 2048;
  const size_t budget = 232448;
  if (fixed + 6 * (size_t)pl.tile_bytes > budget) return pl;
  pl.ring = (int)((budget - fixed) / pl.tile_bytes);
  if (pl.ring > kMaxRing) pl.ring = kMaxRing;
  if (pl.ring < pl.T + 2) return pl;
  pl.smem_bytes = fixed + (size_t)pl.ring * pl.tile_bytes;
  pl.ok = true;
  return pl;
}

}  // namespace

extern "C" int plnr_stem_pool_supported(int dtype, int c, int h, int w, int cout, int kh, int kw, int stride, int pad_t,
                                        int pad_l, int pad_b, int pad_r, int act, int pool_k, int pool_stride,
                                        int pool_pad) {
  if (dtype != PLNR_F16 || act != PLNR_ACT_RELU) return 0;
  return make_plan(c, h, w, cout, kh, kw, stride, pad_t, pad_l, pad_b, pad_r, pool_k, pool_stride, pool_pad).ok ? 1 : 0;
}

extern "C" int plnr_stem_pool_geometry(int h, int kh, int pad_t, int* e_min, int* taps) {
  auto fl2 = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };
  (void)h;
  *e_min = fl2(-pad_t);
  *taps = fl2(kh - 1 - pad_t) - *e_min + 1;
  return PLNR_OK;
}

extern "C" int plnr_stem_pool_fwd(plnr_ctx* ctx, const void* x, int n, int c, int h, int w, const void* w_packed,
                                  const float* scale, const float* shift, int kh, int kw, int stride, int pad_t,
                                  int pad_l, int pad_b, int pad_r, int act, int pool_k, int pool_stride, int pool_pad,
                                  const plnr_tensor* y) {
  PLNR_REQUIRE(ctx && x && w_packed && y && y->ptr, "stem_pool: NULL argument");
  PLNR_REQUIRE(act == PLNR_ACT_RELU, "stem_pool: only the ReLU epilogue makes the pooling pad neutral (act=%d)", act);
  const StemPlan pl = make_plan(c, h, w, y->c, kh, kw, stride, pad_t, pad_l, pad_b, pad_r, pool_k, pool_stride, pool_pad);
  PLNR_REQUIRE(pl.ok, "stem_pool: unsupported problem (c=%d h=%d w=%d cout=%d k=%dx%d s=%d pool=%d/%d/%d)", c, h, w, y->c,
               kh, kw, stride, pool_k, pool_stride, pool_pad);
  PLNR_REQUIRE(y->n == n && y->h == pl.POH && y->w == pl.POW, "stem_pool: output is (%d,%d,%d), expected (%d,%d,%d)",
               y->n, y->h, y->w, n, pl.POH, pl.POW);
  PLNR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0 && y->ld % 8 == 0 && y->coff % 8 == 0,
               "stem_pool: pointers must be 16-byte aligned, ld/coff multiples of 8");

  StemParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const __half*)x;
  p.N = n; p.H = h; p.W = w; p.OH = pl.OH; p.OW = pl.OW; p.POH = pl.POH; p.POW = pl.POW;
  p.pad_l = pad_l; p.e_min = pl.e_min; p.T = pl.T;
  // bands of PB pooled rows: enough items to balance the persistent grid, few enough that the one-row halo stays cheap
  int PB = 7;
  if (const char* e = getenv("PLNR_STEM_BAND")) { int v = atoi(e); if (v >= 1 && v <= 64) PB = v; }
  if (PB > pl.POH) PB = pl.POH;
  p.PB = PB;
  p.bands = (pl.POH + PB - 1) / PB;
  p.items = n * p.bands;
  p.ring = pl.ring; p.tile_bytes = pl.tile_bytes; p.srow_bytes = pl.srow_bytes; p.npos = pl.npos;
  p.w = (const __half*)w_packed; p.scale = scale; p.shift = shift;
  p.y = (__half*)y->ptr; p.yld = y->ld; p.ycoff = y->coff;
  p.err = ctx->dev_error;

  static bool attr_set = false;
  if (!attr_set) {
    PLNR_CHECK_CUDA(cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  int grid = ctx->sm_count < p.items ? ctx->sm_count : p.items;
  stem_pool_kernel<<<grid, kThreads, pl.smem_bytes, ctx->stream>>>(p);
  return plnr_after_launch(ctx, "stem_pool");
}

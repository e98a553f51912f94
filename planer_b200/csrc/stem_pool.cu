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
//     A_t[ow, (ph*3 + c)*8 + j] = x[c, 2t + ph, 2ow + j - P]      P = pad_l rounded up to even   (112 pixels x 48 k)
// written to shared memory (K-major, 128B swizzle) by producer warps: the six planar source rows of a packed row
// (3 channels x 2 phases, zero outside the image) are copied with 16-byte loads into a staging buffer (the loads of the
// NEXT row are already in flight while the current one is gathered), and every lane then turns 20 contiguous staging
// bytes per (pixel pair, phase, channel) into two 16-byte operand chunks -- no shuffling,
// because ANY fixed order of the contraction index works as long as the filter is packed the same way.  Row h then is
//     D_h[128 pixels, 64 co] = sum_{e < T} A_{h + e + e_min} * W_e^T,
//     W_e[co, (ph*3 + c)*8 + j] = K[co, c, 2(e + e_min) + ph + pad_t, j - (P - pad_l)]   (0 outside the filter):
// T * 3 tcgen05.mma (M=128, N=64, K=16) per output row, every packed row reused by T output rows from shared memory.
//
// Epilogue + pooling.  A thread owns one conv column and 32 channels: TMEM -> fp16 pairs -> HFMA2.RELU with the folded
// (scale, shift) held in registers (the reference applies BatchNorm in fp16 as well, planer/layer.py:125-127).  Post-ReLU
// values are >= 0, so the reference's zero padding and -1e4 floor (planer/util.py:82,87-88) are neutral:
// pooled[p, q] = max over conv rows 2p-1..2p+1 and columns 2q-1..2q+1 that exist.  The VERTICAL max of three conv rows
// is taken in registers (the two previous rows stay in 32 registers), only that row goes to shared memory (double
// buffered, one barrier per pooled row), and the horizontal max + 16-byte stores read it back three columns at a time.
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
constexpr int kRing = 8;                // A tiles in the shared-memory ring (power of two: index math by mask)
constexpr int kProducerWarps = 4;
constexpr long long kWatchdogCycles = 4000000000ll;

struct StemParams {
  const void* x;                        // NCHW, fp16 or (U8 instantiation) uint8
  int N, H, W, OH, OW, POH, POW;
  int e_min, T;
  int P;                                // staging position of image column 0 (pad_l rounded up to even)
  int PB, bands, items;
  int ring;                             // A tiles in the shared-memory ring
  uint32_t tile_bytes;                  // round_up(OW * 128, 1024)
  uint32_t srow_bytes;                  // one planar staging row: npos fp16
  uint32_t stg_bytes;                   // one staging buffer: 6 rows, rounded up to 128
  int npos;                             // staging positions per row
  const __half* w;                      // [64][T * 64] packed filter (K-major)
  const float* scale; const float* shift;
  __half* y; int yld, ycoff;
  int* err;
  long long* prof;                      // optional [grid][8] cycle counters per role (debug, plnr_debug_conv_profile)
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

__device__ __forceinline__ uint32_t h2_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) { return h2_u32(__hmax2(u32_h2(a), u32_h2(b))); }

// eight uint8 pixels -> eight fp16 values, exactly: byte b under exponent byte 0x64 is the fp16 number 1024 + b
__device__ __forceinline__ uint4 u8x8_to_h8(uint2 r) {
  const __half2 k1024 = u32_h2(0x64006400u);
  uint4 o;
  o.x = h2_u32(__hsub2(u32_h2(__byte_perm(r.x, 0x64646464u, 0x5140)), k1024));
  o.y = h2_u32(__hsub2(u32_h2(__byte_perm(r.x, 0x64646464u, 0x7362)), k1024));
  o.z = h2_u32(__hsub2(u32_h2(__byte_perm(r.y, 0x64646464u, 0x5140)), k1024));
  o.w = h2_u32(__hsub2(u32_h2(__byte_perm(r.y, 0x64646464u, 0x7362)), k1024));
  return o;
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

template <bool U8>
__global__ void __launch_bounds__(kThreads, 1) stem_pool_kernel(const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  // layout: [A ring][filter taps T x 8 KB][pooled-row stage 2 tiles][staging: 4 warps x 2 buffers][barriers]
  constexpr uint32_t ring = kRing;
  const uint32_t tile = p.tile_bytes;
  const uint32_t off_b = ring * tile;
  const uint32_t off_v = off_b + (uint32_t)p.T * 8192u;
  const uint32_t off_stage = off_v + 2u * tile;
  const uint32_t off_bar = off_stage + kProducerWarps * 2u * p.stg_bytes;
  const uint32_t bar_afull = base + off_bar, bar_aempty = bar_afull + 8 * kRing;
  const uint32_t bar_tfull = bar_aempty + 8 * kRing, bar_tempty = bar_tfull + 8 * kAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + off_bar + 16 * kRing + 16 * kAcc);

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
  // staging rows: the positions left / right of the image stay zero for the whole kernel (zero padding)
  for (uint32_t i = threadIdx.x; i < kProducerWarps * 2u * p.stg_bytes / 4u; i += kThreads)
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
    long long t_full = 0, t_tempty = 0;
    const long long t_all0 = clock64();
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemGeom g = item_geom(p, item);
      for (int j = 0; j < g.nrows; ++j, ++crow) {
        // packed rows j .. j+T-1 of this item; only the newest one can still be in flight
        const long long tf0 = clock64();
        for (int e = (j == 0 ? 0 : T - 1); e < T; ++e) {
          const uint32_t c = pk + (uint32_t)(j + e);
          mbar_wait(bar_afull + 8 * (c % ring), (c / ring) & 1u, p.err, 2);
        }
        const long long tf1 = clock64();
        const uint32_t acc = crow % kAcc;
        mbar_wait(bar_tempty + 8 * acc, ((crow / kAcc) & 1u) ^ 1u, p.err, 1);
        t_full += tf1 - tf0; t_tempty += clock64() - tf1;
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
    if (p.prof && lane == 0) { p.prof[blockIdx.x * 8 + 2] = t_full; p.prof[blockIdx.x * 8 + 3] = t_tempty; p.prof[blockIdx.x * 8 + 4] = clock64() - t_all0; }
  } else if (warp >= 12) {
    // ===================================== producers ==========================================
    // Warp pw builds the packed rows whose running index is congruent to pw (mod 4): four rows are in flight at once.
    // Per row: six planar source rows -> staging (their 16-byte loads were issued while the previous row was gathered);
    // then every lane turns (pixel pair, phase, channel) = 20 staging bytes into two swizzled 16-byte operand chunks.
    const int pw = warp - 12;
    const uint32_t stg0 = off_stage + (uint32_t)pw * 2u * p.stg_bytes;
    const int OW = p.OW, npairs = (OW + 1) >> 1, nitems = 2 * npairs;
    const int W8 = p.W >> 3, nload = 6 * W8;               // 16-byte pieces of the six source rows
    constexpr int kMaxLd = 6;                               // pieces per lane (W <= 256)
    struct Cur { int item, i; uint32_t pk; ItemGeom g; bool valid; };
    auto advance = [&](Cur& c) {
      for (;;) {
        if (++c.i >= c.g.npk) {
          c.pk += (uint32_t)c.g.npk;
          c.item += gridDim.x;
          if (c.item >= p.items) { c.valid = false; return; }
          c.g = item_geom(p, c.item);
          c.i = 0;
        }
        if ((int)((c.pk + (uint32_t)c.i) & 3u) == pw) return;
      }
    };
    // Everything that depends only on (lane, piece) is computed ONCE: the per-row loops below are loads / stores at
    // precomputed offsets (this warp's instruction stream, not bandwidth, is what paces the kernel otherwise).
    uint4 q[kMaxLd];
    int ld_goff[kMaxLd];            // element offset of the piece inside the image (phase row not yet added), -1 = none
    int ld_ph[kMaxLd];              // phase of the piece (0 / 1)
    uint32_t ld_soff[kMaxLd];       // byte offset of the piece in the staging buffer
#pragma unroll
    for (int k = 0; k < kMaxLd; ++k) {
      const int it = lane + 32 * k;
      const int row = it / W8, v = it - row * W8;           // row = channel * 2 + phase
      ld_goff[k] = it < nload ? ((row >> 1) * p.H + (row & 1)) * p.W + v * 8 : -1;
      ld_ph[k] = row & 1;
      ld_soff[k] = (uint32_t)row * p.srow_bytes + (uint32_t)(v * 8 + p.P) * 2u;
    }
    constexpr int kMaxG = 4;        // (pixel pair, phase) items per lane (OW <= 124)
    uint32_t g_src[kMaxG], g_dst[kMaxG], g_sw[kMaxG];
    bool g_ok[kMaxG], g_ok1[kMaxG];
#pragma unroll
    for (int k = 0; k < kMaxG; ++k) {
      const int it = lane + 32 * k;
      const int ph = it / npairs, m = it - ph * npairs;
      g_ok[k] = it < nitems;
      g_ok1[k] = 2 * m + 1 < OW;
      g_src[k] = (uint32_t)ph * p.srow_bytes + 8u * (uint32_t)m;      // + channel * 2 * srow_bytes
      g_dst[k] = (uint32_t)(2 * m) * 128u;
      g_sw[k] = ((uint32_t)(2 * m) & 7u) | ((uint32_t)ph << 8);       // swizzle phase of pixel 2m | phase
    }
    const int two_w = 2 * p.W;
    auto load_row = [&](const Cur& c) {                     // global -> registers
      const int t = c.g.h0 + p.e_min + c.i;
      const long long row0 = (long long)c.g.img * 3 * p.H * p.W + (long long)t * two_w;     // in elements
#pragma unroll
      for (int k = 0; k < kMaxLd; ++k) {
        const int ih = 2 * t + ld_ph[k];
        q[k] = make_uint4(0u, 0u, 0u, 0u);
        if (ld_goff[k] >= 0 && ih >= 0 && ih < p.H) {
          if (U8)       // uint8 image: 8-byte loads, converted here so that staging and gather are those of the fp16 path
            q[k] = u8x8_to_h8(__ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(p.x) + row0 + ld_goff[k])));
          else
            q[k] = __ldcs(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.x) + row0 + ld_goff[k]));
        }
      }
    };
    auto store_row = [&](uint8_t* stg) {                    // registers -> planar staging rows [channel][phase][position]
#pragma unroll
      for (int k = 0; k < kMaxLd; ++k) {
        if (ld_goff[k] >= 0) {
          uint32_t* dst = reinterpret_cast<uint32_t*>(stg + ld_soff[k]);
          dst[0] = q[k].x; dst[1] = q[k].y; dst[2] = q[k].z; dst[3] = q[k].w;
        }
      }
    };
    Cur is;
    is.item = blockIdx.x; is.i = -1; is.pk = 0; is.valid = blockIdx.x < p.items;
    if (is.valid) { is.g = item_geom(p, is.item); advance(is); }
    Cur cs = is;
    uint32_t n_cons = 0;
    long long t_wait = 0;
    const long long t_all0 = clock64();
    if (is.valid) { load_row(is); advance(is); }
    while (cs.valid) {
      uint8_t* stg = bp + stg0 + (n_cons & 1u) * p.stg_bytes;
      store_row(stg);
      __syncwarp();
      if (is.valid) { load_row(is); advance(is); }         // next row's loads fly during this gather
      const uint32_t c = cs.pk + (uint32_t)cs.i;
      const uint32_t slot = c % ring;
      const long long tw0 = clock64();
      mbar_wait(bar_aempty + 8 * slot, ((c / ring) & 1u) ^ 1u, p.err, 0);
      t_wait += clock64() - tw0;
      uint8_t* atile = bp + slot * tile;
      const uint32_t srow2 = 2u * p.srow_bytes;
#pragma unroll
      for (int k = 0; k < kMaxG; ++k) {
        if (g_ok[k]) {
          const uint8_t* src = stg + g_src[k];
          uint8_t* arow0 = atile + g_dst[k];
          const uint32_t sw0 = g_sw[k] & 7u, sw1 = sw0 + 1u, ph3 = (g_sw[k] >> 8) * 3u;    // pixel 2m is even: (2m+1)&7 = sw0+1
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const uint2 a = *reinterpret_cast<const uint2*>(src + ch * srow2), b = *reinterpret_cast<const uint2*>(src + ch * srow2 + 8);
            const uint32_t e = *reinterpret_cast<const uint32_t*>(src + ch * srow2 + 16);
            const uint32_t chunk = ph3 + (uint32_t)ch;
            *reinterpret_cast<uint4*>(arow0 + ((chunk ^ sw0) << 4)) = make_uint4(a.x, a.y, b.x, b.y);
            if (g_ok1[k]) *reinterpret_cast<uint4*>(arow0 + 128 + ((chunk ^ sw1) << 4)) = make_uint4(a.y, b.x, b.y, e);
          }
        }
      }
      ptx::fence_proxy_async_smem();           // generic-proxy writes -> visible to the tensor core's async proxy
      ptx::mbar_arrive(bar_afull + 8 * slot);
      __syncwarp();
      ++n_cons;
      advance(cs);
    }
    if (p.prof && pw == 0 && lane == 0) { p.prof[blockIdx.x * 8 + 0] = t_wait; p.prof[blockIdx.x * 8 + 1] = clock64() - t_all0; }
  } else if (warp >= 4) {
    // ===================================== epilogue + pooling =================================
    // The instruction stream of these warps paces the kernel, so everything that depends only on the thread is hoisted:
    // folded BatchNorm as fp16 pairs in registers, the staging address of the thread's column, the two pooled outputs
    // (q, 16-byte channel group) it produces per pooled row with the three staging offsets each of them reads.
    const int ew = warp & 3;                 // TMEM lane quarter
    const int eg = (warp - 4) >> 2;          // channel half: channels [32 eg, 32 eg + 32)
    const int et = threadIdx.x - 128;        // 0..255
    const int px = ew * 32 + lane;           // conv column owned by this thread
    const bool pvalid = px < p.OW;
    uint8_t* vbuf = bp + off_v;
    const int POW = p.POW, OW = p.OW;
    __half2 sc[16], sf[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int ch = eg * 32 + 2 * k;
      sc[k] = __floats2half2_rn(p.scale ? __ldg(p.scale + ch) : 1.f, p.scale ? __ldg(p.scale + ch + 1) : 1.f);
      sf[k] = __floats2half2_rn(p.shift ? __ldg(p.shift + ch) : 0.f, p.shift ? __ldg(p.shift + ch + 1) : 0.f);
    }
    uint32_t v_dst[4];                       // swizzled staging offsets of this thread's four 16-byte channel groups
#pragma unroll
    for (int q = 0; q < 4; ++q) v_dst[q] = (uint32_t)px * 128u + ((((uint32_t)(eg * 4 + q)) ^ (uint32_t)(px & 7)) << 4);
    constexpr int kPoolItems = 2;            // pooled outputs per thread and pooled row (POW * 8 <= 512)
    uint32_t h_src[kPoolItems][3];           // staging offsets of columns 2q-1, 2q, 2q+1 (0xffffffff = outside)
    uint32_t h_dst[kPoolItems];              // element offset in the output row, 0xffffffff = no item
#pragma unroll
    for (int i = 0; i < kPoolItems; ++i) {
      const int idx = et + 256 * i, q = idx >> 3, c16 = idx & 7;
      h_dst[i] = idx < POW * 8 ? (uint32_t)(q * p.yld + c16 * 8) : 0xffffffffu;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int cx = 2 * q + d - 1;
        h_src[i][d] = (cx >= 0 && cx < OW && idx < POW * 8)
                          ? (uint32_t)cx * 128u + ((((uint32_t)c16) ^ (uint32_t)(cx & 7)) << 4) : 0xffffffffu;
      }
    }
    const uint32_t taddr0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)eg * 32u;
    uint32_t crow = 0, prow_cnt = 0;
    long long t_tfull = 0, t_bar = 0;
    const long long t_all0 = clock64();

    // one conv row of the band: accumulator -> fp16 pairs, BatchNorm + ReLU (zero row when the row does not exist)
    auto conv_row = [&](int hh, uint32_t (&out)[16]) {
      if (hh >= 0 && hh < p.OH) {
        const uint32_t acc = crow & (kAcc - 1);
        const long long tt0 = clock64();
        mbar_wait(bar_tfull + 8 * acc, (crow / kAcc) & 1u, p.err, 3);
        t_tfull += clock64() - tt0;
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr0 + acc * 64u, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_tempty + 8 * acc);
        ++crow;
#pragma unroll
        for (int k = 0; k < 16; ++k)
          out[k] = h2_u32(__hfma2_relu(__floats2half2_rn(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1])), sc[k], sf[k]));
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) out[k] = 0u;
      }
    };

    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const ItemGeom g = item_geom(p, item);
      const int hh0 = 2 * g.p0 - 1;
      uint32_t ra[16];                       // conv row 2k of the band (top row of the pooling window)
      conv_row(hh0, ra);
      __half* yrow = p.y + ((size_t)g.img * p.POH + g.p0) * POW * p.yld + p.ycoff;
      for (int k = 0; k < g.np; ++k, yrow += (size_t)POW * p.yld) {
        uint32_t rb[16], rc[16];
        conv_row(hh0 + 2 * k + 1, rb);
        conv_row(hh0 + 2 * k + 2, rc);
        // vertical max of the three conv rows in registers -> staging (double buffered) -> horizontal max + store
        uint8_t* vb = vbuf + (prow_cnt & 1u) * tile;
        if (pvalid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = hmax2_u32(hmax2_u32(ra[4 * q], rb[4 * q]), rc[4 * q]);
            o.y = hmax2_u32(hmax2_u32(ra[4 * q + 1], rb[4 * q + 1]), rc[4 * q + 1]);
            o.z = hmax2_u32(hmax2_u32(ra[4 * q + 2], rb[4 * q + 2]), rc[4 * q + 2]);
            o.w = hmax2_u32(hmax2_u32(ra[4 * q + 3], rb[4 * q + 3]), rc[4 * q + 3]);
            *reinterpret_cast<uint4*>(vb + v_dst[q]) = o;
          }
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) ra[q] = rc[q];
        // ONE barrier per pooled row: the stage is double buffered, and a thread can only overwrite buffer b two
        // pooled rows later, i.e. after passing the next barrier, which every reader of b reaches after its reads
        const long long tb0 = clock64();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        t_bar += clock64() - tb0;
#pragma unroll
        for (int i = 0; i < kPoolItems; ++i) {
          if (h_dst[i] != 0xffffffffu) {
            uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              if (h_src[i][d] != 0xffffffffu) {
                const uint4 t = *reinterpret_cast<const uint4*>(vb + h_src[i][d]);
                m.x = hmax2_u32(m.x, t.x); m.y = hmax2_u32(m.y, t.y); m.z = hmax2_u32(m.z, t.z); m.w = hmax2_u32(m.w, t.w);
              }
            }
            *reinterpret_cast<uint4*>(yrow + h_dst[i]) = m;
          }
        }
        ++prow_cnt;
      }
    }
    if (p.prof && threadIdx.x == 128) { p.prof[blockIdx.x * 8 + 5] = t_tfull; p.prof[blockIdx.x * 8 + 6] = clock64() - t_all0; p.prof[blockIdx.x * 8 + 7] = t_bar; }
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
  int OH, OW, POH, POW, e_min, T, ring, npos, P;
  uint32_t tile_bytes, srow_bytes, stg_bytes;
  size_t smem_bytes;
};

static inline int floor_half(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

static StemPlan make_plan(int c, int h, int w, int cout, int kh, int kw, int stride, int pad_t, int pad_l, int pad_b,
                          int pad_r, int pool_k, int pool_s, int pool_p) {
  StemPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (c != 3 || cout != 64 || stride != 2 || kw < 1 || kh < 1 || w % 8 != 0) return pl;
  if (pool_k != 3 || pool_s != 2 || pool_p != 1) return pl;
  if (pad_t < 0 || pad_l < 0 || pad_b > pad_t + 1 || pad_r > pad_l + 1) return pl;
  pl.P = pad_l + (pad_l & 1);
  if (kw + (pl.P - pad_l) > 8) return pl;               // 8 horizontal slots per (phase, channel)
  pl.OH = plnr_out_size(h, pad_t, pad_b, kh, 1, stride);
  pl.OW = plnr_out_size(w, pad_l, pad_r, kw, 1, stride);
  if (pl.OH < 2 || pl.OW < 2 || pl.OW > 124) return pl;
  pl.POH = plnr_out_size(pl.OH, 1, 1, 3, 1, 2);
  pl.POW = plnr_out_size(pl.OW, 1, 1, 3, 1, 2);
  // vertical tap r reads input row 2h + r - pad_t = 2(h + e) + ph:  e = floor((r - pad_t) / 2)
  pl.e_min = floor_half(-pad_t);
  pl.T = floor_half(kh - 1 - pad_t) - pl.e_min + 1;
  if (pl.T < 1 || pl.T > kMaxT) return pl;
  pl.tile_bytes = (uint32_t)round_up(pl.OW * 128, 1024);
  pl.npos = round_up(2 * pl.OW + 8, 8);                 // pixel pair m reads staging positions 4m .. 4m + 9
  if (round_up(w + pl.P, 8) > pl.npos) pl.npos = round_up(w + pl.P, 8);   // image column j is stored at position j + P
  if (w > 256) return pl;                               // six source rows = at most 6 x 32 16-byte pieces per warp
  pl.srow_bytes = (uint32_t)pl.npos * 2u;
  pl.stg_bytes = (uint32_t)round_up(6 * (int)pl.srow_bytes, 128);
  const size_t fixed = (size_t)pl.T * 8192 + 2 * (size_t)pl.tile_bytes + kProducerWarps * 2 * (size_t)pl.stg_bytes +
                       16 * kRing + 16 * kAcc + 64 + 1024 + 2048;
  const size_t budget = 232448;
  if (fixed + kRing * (size_t)pl.tile_bytes > budget) return pl;
  pl.ring = kRing;
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

extern "C" int plnr_stem_pool_geometry(int kh, int pad_t, int pad_l, int* e_min, int* taps, int* col_shift) {
  *e_min = floor_half(-pad_t);
  *taps = floor_half(kh - 1 - pad_t) - *e_min + 1;
  *col_shift = pad_l & 1;
  return PLNR_OK;
}

static int stem_pool_launch(plnr_ctx* ctx, const void* x, bool u8, int n, int c, int h, int w, const void* w_packed,
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
  PLNR_REQUIRE((reinterpret_cast<uintptr_t>(x) & (u8 ? 7 : 15)) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0 && y->ld % 8 == 0 && y->coff % 8 == 0,
               "stem_pool: pointers must be 16-byte aligned, ld/coff multiples of 8");

  StemParams p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  p.N = n; p.H = h; p.W = w; p.OH = pl.OH; p.OW = pl.OW; p.POH = pl.POH; p.POW = pl.POW;
  p.e_min = pl.e_min; p.T = pl.T; p.P = pl.P;
  // bands of PB pooled rows: enough items to balance the persistent grid, few enough that the one-row halo stays cheap
  int PB = 7;
  if (const char* e = getenv("PLNR_STEM_BAND")) { int v = atoi(e); if (v >= 1 && v <= 64) PB = v; }
  if (PB > pl.POH) PB = pl.POH;
  p.PB = PB;
  p.bands = (pl.POH + PB - 1) / PB;
  p.items = n * p.bands;
  p.ring = pl.ring; p.tile_bytes = pl.tile_bytes; p.srow_bytes = pl.srow_bytes; p.stg_bytes = pl.stg_bytes; p.npos = pl.npos;
  p.w = (const __half*)w_packed; p.scale = scale; p.shift = shift;
  p.y = (__half*)y->ptr; p.yld = y->ld; p.ycoff = y->coff;
  p.err = ctx->dev_error;
  p.prof = ctx->prof;

  static bool attr_set = false;
  if (!attr_set) {
    PLNR_CHECK_CUDA(cudaFuncSetAttribute(stem_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    PLNR_CHECK_CUDA(cudaFuncSetAttribute(stem_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  int grid = ctx->sm_count < p.items ? ctx->sm_count : p.items;
  if (u8) stem_pool_kernel<true><<<grid, kThreads, pl.smem_bytes, ctx->stream>>>(p);
  else stem_pool_kernel<false><<<grid, kThreads, pl.smem_bytes, ctx->stream>>>(p);
  return plnr_after_launch(ctx, "stem_pool");
}

extern "C" int plnr_stem_pool_fwd(plnr_ctx* ctx, const void* x, int n, int c, int h, int w, const void* w_packed,
                                  const float* scale, const float* shift, int kh, int kw, int stride, int pad_t,
                                  int pad_l, int pad_b, int pad_r, int act, int pool_k, int pool_stride, int pool_pad,
                                  const plnr_tensor* y) {
  return stem_pool_launch(ctx, x, false, n, c, h, w, w_packed, scale, shift, kh, kw, stride, pad_t, pad_l, pad_b, pad_r, act,
                          pool_k, pool_stride, pool_pad, y);
}

extern "C" int plnr_stem_pool_fwd_u8(plnr_ctx* ctx, const void* x, int n, int c, int h, int w, const void* w_packed,
                                     const float* scale, const float* shift, int kh, int kw, int stride, int pad_t,
                                     int pad_l, int pad_b, int pad_r, int act, int pool_k, int pool_stride, int pool_pad,
                                     const plnr_tensor* y) {
  return stem_pool_launch(ctx, x, true, n, c, h, w, w_packed, scale, shift, kh, kw, stride, pad_t, pad_l, pad_b, pad_r, act,
                          pool_k, pool_stride, pool_pad, y);
}

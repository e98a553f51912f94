// Stride-1 fp16 convolution as a "shift GEMM": every input row is loaded into shared memory ONCE per 64-channel chunk
// and all kh*kw filter taps are fed to tcgen05.mma from row-shifted views of that one buffer.
//
// Replaces the same reference code as conv_tcgen05.cu (planer/layer.py:22-26 + planer/util.py:17-44 and the fused
// batchnorm / add / relu layers) for stride-1 convolutions with Cin % 64 == 0 -- 16 of ResNet-18's 20 convolutions and
// every 3x3/s1 and 1x1 convolution of YOLOv3.  The im2col kernel re-reads each input pixel kh*kw times through the
// TMA unit, whose im2col mode sustains only ~46 B/clk per SM (tools/tma_stream_probe.cu): its main loop is bound by
// that, not by the tensor pipe.  Here:
//
//   * the image is viewed as rows of Wv = W + pad_left pixels (left padding materialised by TMA out-of-bounds zero
//     fill; the right padding of one row IS the left padding of the next) and Hv = H + pad_top rows per image, all
//     flattened: output position o = (n*Hv + p)*Wv + q needs input position o + r*dil_h*Wv + s*dil_w for tap (r, s);
//   * a tile is 128 consecutive positions o (x2 for a CTA pair); its A buffer holds positions [o0 - Wv', o0 + 128 +
//     halo) as whole rows fetched by tiled TMA boxes (64 ch x Wv px, ~80-100 B/clk); positions with p >= OH or
//     q >= OW are computed and discarded (OH*OW/(Hv*Wv) efficiency: 96.5 % at 56x56, 87 % at 14x14);
//   * tcgen05 reads a 128B-swizzled K-major operand from ANY 128-byte-aligned start address (the swizzle is a
//     function of the absolute shared-memory address; tools/swizzle_probe.cu), so tap (r, s) is just the A descriptor
//     advanced by (r*dil_h*Wv + s*dil_w) rows;
//   * weights stream through their own ring, one [n_tile x 64] box per (chunk, tap) -- or stay resident in shared
//     memory for the whole kernel when they fit (Cout = 64 / 128 layers);
//   * CG = 2 runs a CTA pair per 256 positions with cta_group::2 MMAs (M = 256): half the B bytes per CTA and half
//     the MMA instructions per output row, which matters because an M=128 MMA costs ~85-100 clk even at N <= 128.
//
// Warp roles, TMEM double buffering, the fused epilogue and the watchdog are those of conv_tcgen05.cu.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 384;          // warps 0-3: producer / MMA / TMEM / table; warps 4-11: two epilogue groups
constexpr int kMaxA = 8, kMaxB = 40;
constexpr uint32_t kEpiBytes = 2 * 2 * 256 * 2;   // [buffer][scale | shift][256] fp16 (the epilogue math is packed fp16)
constexpr uint32_t kStageBytes = 8 * 1024;        // epilogue transposition stage (per epilogue warp: 32 rows x 32 B = 16 channels)
constexpr long long kWatchdogCycles = 4000000000ll;

struct AMaps { CUtensorMap m[4]; };     // the A operand's tensor map with boxes of 1, 2, 4, 8 rows

struct ShiftParams {
  FastDiv div_hvwv, div_wv, div_mt, div_hv;     // / (Hv*Wv), / Wv, / num_m_tiles, / Hv
  int N, OH, OW, Hv, Wv, HvWv;
  int pad_t, pad_l;
  long long Mv;             // N * Hv * Wv virtual output positions
  int halo;                 // (R-1)*dil_h*Wv + (S-1)*dil_w
  int R, S, dh, dw;
  // Phase planes.  Stride 1: ONE plane holding every tap.  Stride 2: the input splits into (row parity, column parity)
  // planes x[2i + a, 2j + b]; tap (r, s) reads plane ((r - pad_t) mod 2, (s - pad_l) mod 2) at the shifted position
  // (floor((r - pad_t) / 2), floor((s - pad_l) / 2)), so the strided convolution is a sum of up to four stride-1 shift
  // GEMMs, one per plane, over the SAME virtual output grid.  A plane is fetched by a tiled TMA box with element stride
  // `cs` along W (every other pixel: whole 128-byte channel rows, no wasted sectors) and row coordinate hrow * cs + h0.
  int nplanes, cs;
  int max_hlog;             // log2 of the tallest A box in use (0..3)
  int pl_w0[4], pl_h0[4];   // TMA start coordinates of a plane's virtual column 0 / virtual row 0
  int pl_first[5];          // taps [pl_first[i], pl_first[i + 1]) of the tables below belong to plane i
  uint32_t tap_aoff[kMaxB]; // A descriptor offset of a tap inside its plane buffer: (r' * Wv + s') * 8
  uint8_t tap_w[kMaxB];     // index r * S + s of the tap in the packed filter
  int C, cchunks;
  int c2chunks, s2;         // fused 1x1 shortcut convolution: extra 64-channel chunks read from a second tensor at stride s2
  uint32_t ctr;             // descriptor offset of the un-shifted (centre) view: (pad_t * dil_h * Wv + pad_l * dil_w) * 8
  int n_tile, num_m_tiles, num_tiles;
  int na, nb, b_resident;
  uint32_t a_buf_bytes, b_stage_bytes, idesc, tmem_cols;
  __half* y; int yld, ycoff, Cout;
  const float* scale; const float* shift;
  const __half* res; int rld, rcoff;
  int act; float alpha; int res_after;
  int vec_ok;
  int stage_wide;           // epilogue transposition stage: 2 KB per warp (a whole 32-column chunk per round trip) instead of 1 KB
  int y_nchw, ohw;          // 1: y is a dense NCHW array (graph exit folded into the epilogue); OH * OW
  // Global-average-pool fold (plnr_epilogue.pool_sum): instead of storing y, every epilogue warp writes the fp32 SUM of its
  // 32 accumulator rows (all of one image: HvWv % 32 == 0), finished values, rim rows excluded, to
  // pool[(image * pool_parts + (position in the image) / 32) * Cout + channel] -- one writer per element, no atomics
  float* pool; int pool_parts;
  int* err;
  long long* prof;
  // stream-K (see SegList): iterations per tile = weight boxes per tile; fp32 partial tiles and their ready flags
  int streamk, ipt;
  float* sk_ws;             // [unit][CG * 128 rows][n_tile] fp32: the partial accumulator a unit contributes to another unit's tile
  int* sk_flags;            // [unit][2 ranks][8 epilogue warps]: 1 = that warp's part of the partial is in sk_ws
};

// Work of one persistent unit (CTA or CTA pair) as a list of segments.  Without stream-K a segment is a whole tile
// (tiles unit, unit + nunits, ...).  With stream-K the (tile, k-iteration) space -- an iteration is one weight box = one
// (64-channel chunk, tap) = four tcgen05.mma -- is cut into nunits equal contiguous ranges, so that a last, partly filled
// wave (225 tiles on 148 SMs at 14x14) costs its share instead of a whole round.  A range covers: the tail of a tile
// another unit started (CONTRIB: the fp32 partial accumulator goes to sk_ws), whole tiles, and the head of a tile
// (OWNER: waits for the partials of the units that continue the tile, adds them in unit order -- deterministic -- and runs
// the epilogue).  Order inside a unit: CONTRIB first, OWNER second, whole tiles last: every partial is produced at the very
// start of its unit's timeline, so an owner never waits in practice and no wait can depend on another wait.
enum { SEG_FULL = 0, SEG_OWNER = 1, SEG_CONTRIB = 2 };
struct Seg { int tile, it0, it1, mode; };
struct SegList {
  int nseg, first_tile, it0_first, it1_last, ipt, unit, nunits, streamk;
  bool has_contrib, has_owner;
  __device__ __forceinline__ static void range(const ShiftParams& p, int u, int nunits, long long& g0, long long& g1) {
    const long long T = (long long)p.num_tiles * p.ipt;
    g0 = T * u / nunits;
    g1 = T * (u + 1) / nunits;
  }
  __device__ __forceinline__ void init(const ShiftParams& p, int unit_, int nunits_) {
    unit = unit_; nunits = nunits_; ipt = p.ipt; streamk = p.streamk;
    has_contrib = has_owner = false;
    first_tile = 0; it0_first = 0; it1_last = ipt;
    if (!streamk) {
      nseg = unit < p.num_tiles ? (p.num_tiles - unit + nunits - 1) / nunits : 0;
      return;
    }
    long long g0, g1;
    range(p, unit, nunits, g0, g1);
    if (g1 <= g0) { nseg = 0; return; }
    first_tile = (int)(g0 / ipt);
    const int last_tile = (int)((g1 - 1) / ipt);
    nseg = last_tile - first_tile + 1;
    it0_first = (int)(g0 - (long long)first_tile * ipt);
    it1_last = (int)(g1 - (long long)last_tile * ipt);
    has_contrib = it0_first > 0;
    has_owner = it1_last < ipt && (nseg > 1 || it0_first == 0);
  }
  __device__ __forceinline__ Seg at(int k) const {
    Seg sg;
    if (!streamk) { sg.tile = unit + k * nunits; sg.it0 = 0; sg.it1 = ipt; sg.mode = SEG_FULL; return sg; }
    int j = k;
    if (has_owner) {
      if (has_contrib) j = k == 0 ? 0 : (k == 1 ? nseg - 1 : k - 1);
      else j = k == 0 ? nseg - 1 : k - 1;
    }
    sg.tile = first_tile + j;
    sg.it0 = j == 0 ? it0_first : 0;
    sg.it1 = j == nseg - 1 ? it1_last : ipt;
    sg.mode = sg.it0 > 0 ? SEG_CONTRIB : (sg.it1 < ipt ? SEG_OWNER : SEG_FULL);
    return sg;
  }
};

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int role) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFF) == 0) {
      if (*reinterpret_cast<volatile int*>(err) != 0) return;
      if (clock64() - t0 > kWatchdogCycles) {
        if (atomicCAS(err, 0, 2) == 0) { err[1] = blockIdx.x; err[2] = role; err[3] = (int)parity; }
        __threadfence();
        return;
      }
    }
  }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// tiled 4-D box load (c, w, h, n); CG = 2 reports the bytes to the pair leader's mbarrier
template <int CG>
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n) {
  if (CG == 2) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & ptx::kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n),
          "l"(ptx::kTmaCacheHintDefault)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
        : "memory");
  }
}

// rows of the A buffer a CTA whose first position is o0 has to load: [v0, v0 + nrows)
__device__ __forceinline__ void tile_rows(long long o0, int Wv, int halo, const FastDiv& div_wv, long long& v0, int& nrows, int& off) {
  // every position (+ halo) is below 2^31 (make_plan), so the divisions are 32-bit multiply-highs: with 64-bit divisions
  // the producer warp spent ~700 clk per tile on address arithmetic and paced layers with few MMAs per tile (YOLOv3's
  // pointwise convolutions: 2300 clk per tile for four MMAs)
  const uint32_t u0 = (uint32_t)o0;
  const uint32_t q0 = fast_div(u0, div_wv);
  v0 = (long long)q0;
  off = (int)(u0 - q0 * (uint32_t)Wv);
  nrows = (int)(fast_div(u0 + (uint32_t)(kTileM - 1 + halo), div_wv) - q0) + 1;
}

// POOL: the instantiation whose epilogue folds GlobalAveragePool (ShiftParams::pool); a separate instantiation so that the
// register allocation of every other launch is exactly what it was without the fold (168 registers, no spills)
template <int CG, bool POOL = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_shift_f16_kernel(const __grid_constant__ AMaps mapsA, const __grid_constant__ CUtensorMap mapB,
                      const __grid_constant__ CUtensorMap mapA2, const ShiftParams p) {
  extern __shared__ uint8_t smem_raw[];
  const long long t_entry = clock64();
  const bool stamp = p.prof && blockIdx.x == 0;      // debug: phase time stamps of CTA 0 at prof[1400..]
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t na = (uint32_t)p.na, nb = (uint32_t)p.nb;
  const uint32_t sA = base;
  const uint32_t sB = sA + na * p.a_buf_bytes;
  const uint32_t epi_off = na * p.a_buf_bytes + nb * p.b_stage_bytes;
  float* epi = reinterpret_cast<float*>(base_ptr + epi_off);
  const uint32_t sBar = base + epi_off + kEpiBytes;
  const uint32_t bar_afull = sBar, bar_aempty = sBar + 8 * kMaxA;
  const uint32_t bar_bfull = sBar + 16 * kMaxA, bar_bempty = bar_bfull + 8 * kMaxB;
  const uint32_t bar_tfull = bar_bempty + 8 * kMaxB, bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + epi_off + kEpiBytes + 16 * kMaxA + 16 * kMaxB + 32);
  uint8_t* stage = base_ptr + epi_off + kEpiBytes + 16 * kMaxA + 16 * kMaxB + 64;   // 4 warps x 2 KB

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int Wv = p.Wv, RS = p.R * p.S;
  const uint32_t row_bytes = (uint32_t)Wv * 128u;       // one virtual row of 64 channels in shared memory

  if (warp == 0 && lane == 0) {
    for (int i = 0; i <= p.max_hlog; ++i) ptx::prefetch_tmap(&mapsA.m[i]);
    ptx::prefetch_tmap(&mapB);
    if (p.c2chunks) ptx::prefetch_tmap(&mapA2);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < na; ++i) { ptx::mbar_init(bar_afull + 8 * i, 1); ptx::mbar_init(bar_aempty + 8 * i, 1); }
    for (uint32_t i = 0; i < nb; ++i) { ptx::mbar_init(bar_bfull + 8 * i, 1); ptx::mbar_init(bar_bempty + 8 * i, 1); }
    for (uint32_t a = 0; a < 2; ++a) { ptx::mbar_init(bar_tfull + 8 * a, 1); ptx::mbar_init(bar_tempty + 8 * a, 256 * CG); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if (CG == 2) { ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), p.tmem_cols); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), p.tmem_cols); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all();      // the peer's barriers are initialised and its TMEM allocated before anything is sent to it
  __syncthreads();                           // (also under CG == 2: a CTA barrier is what compute-sanitizer's racecheck models)
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // programmatic dependent launch: let the next kernel of the stream start its own set-up, then wait until every
  // kernel before this one has completed and flushed its results (no-ops when launched without the attribute)
  ptx::grid_launch_dependents();
  if (warp != 0) ptx::grid_dependency_wait();      // the producer warp first requests weights (below), then waits
  if (stamp && threadIdx.x == 0) {
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[1400] = (long long)gt; p.prof[1401] = clock64() - t_entry;
  }

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    uint32_t ab = 0, aph = 0, bs = 0, bph = 0;
    bool first = true;
    long long t_wait = 0;
    const long long t_all0 = clock64();
    SegList segs;
    segs.init(p, unit, nunits);
    // (Requesting the first weight boxes BEFORE this wait -- weights do not depend on the previous kernel -- was tried: fine
    // in the single-CTA stacked kernel, intermittent launch failures in this kernel's CTA-pair configuration, removed:
    // profiles/r02_kernel_experiments.md 8.)
    ptx::grid_dependency_wait();
    for (int k = 0; k < segs.nseg; ++k) {
      const Seg sg = segs.at(k);
      const int tile = sg.tile;
      const int n_idx = (int)fast_div((uint32_t)tile, p.div_mt), m_idx = tile - n_idx * p.num_m_tiles;
      const long long o0 = ((long long)m_idx * CG + rank) * kTileM;
      long long v0; int nrows, off;
      tile_rows(o0, Wv, p.halo, p.div_wv, v0, nrows, off);
      uint32_t a_bytes = (uint32_t)nrows * row_bytes;
      if (CG == 2) {                                     // the leader arms the barrier for both CTAs' rows
        long long v0p; int nrows_p, off_p;
        tile_rows(o0 + (leader ? kTileM : -kTileM), Wv, p.halo, p.div_wv, v0p, nrows_p, off_p);
        a_bytes += (uint32_t)nrows_p * row_bytes;
      }
      const int n_row = n_idx * p.n_tile + (int)rank * (p.n_tile / CG);
      // position o0 always lands at row-offset Wv of the buffer, so the A descriptors are tile-independent and
      // identical in both CTAs of a pair (they share ONE descriptor per MMA)
      const uint32_t row0_off = (uint32_t)(Wv - off) * 128u;
      // chunks [0, cchunks): 64 input channels of the convolution, RS taps each; chunks [cchunks, +c2chunks): 64
      // channels of the fused 1x1 shortcut, read from the second tensor at stride s2 into the same row layout, ONE tap
      const int nmain = p.cchunks * p.nplanes;
      for (int lu = 0; lu < nmain + p.c2chunks; ++lu) {
        // load unit = (64-channel chunk, plane) of the convolution, then the chunks of the fused shortcut
        const bool sc = lu >= nmain;
        const int cc = sc ? lu - nmain : (p.nplanes == 1 ? lu : lu / p.nplanes);
        const int pi = sc ? 0 : lu - cc * p.nplanes;
        // taps of this unit inside the segment's iteration range (iteration of tap t: cc * RS + t; shortcut chunks follow)
        int t0 = sc ? 0 : p.pl_first[pi], t1 = sc ? 1 : p.pl_first[pi + 1];
        const int ibase = sc ? p.cchunks * RS + cc : cc * RS;
        t0 = max(t0, sg.it0 - ibase);
        t1 = min(t1, sg.it1 - ibase);
        if (t0 >= t1) continue;
        const long long tw0 = clock64();
        mbar_wait(bar_aempty + 8 * ab, aph ^ 1, p.err, 0);
        t_wait += clock64() - tw0;
        if (ptx::elect_one()) {
          const uint32_t full = bar_afull + 8 * ab;
          if (leader) ptx::mbar_arrive_expect_tx(full, a_bytes);
          uint32_t dst = sA + ab * p.a_buf_bytes + row0_off;
          long long v = v0;
          int img = (int)fast_div((uint32_t)v, p.div_hv), hrow = (int)v - img * p.Hv;
          const int cs = sc ? p.s2 : p.cs, c0 = cc * 64;
          const int w0 = sc ? -p.pad_l * p.s2 : p.pl_w0[pi], h0 = sc ? -p.pad_t * p.s2 : p.pl_h0[pi];
          // The rows of one image go out as boxes of 8 / 4 / 2 / 1 rows: the issuing thread pays ~100-150 clk per TMA
          // instruction whatever its size, and at 7x7 / 14x14 a tile is 12-21 rows per chunk (per plane at stride 2) --
          // issued row by row this warp, not the tensor pipe, paced those layers (profiles/r02_tma_issue.md)
          const int max_hlog = sc ? 0 : p.max_hlog;
          int left = nrows;
          while (left > 0) {
            int seg = min(left, p.Hv - hrow);                // rows of this image
            left -= seg;
            while (seg > 0) {
              int k = seg >= 8 ? 3 : (seg >= 4 ? 2 : (seg >= 2 ? 1 : 0));
              if (k > max_hlog) k = max_hlog;
              tma_load_4d<CG>(dst, sc ? &mapA2 : &mapsA.m[k], full, c0, w0, hrow * cs + h0, img);
              dst += row_bytes << k;
              hrow += 1 << k;
              seg -= 1 << k;
            }
            if (hrow == p.Hv) { hrow = 0; ++img; }
          }
        }
        __syncwarp();
        if (++ab == na) { ab = 0; aph ^= 1; }
        if (!p.b_resident || first) {
          for (int t = t0; t < t1; ++t) {
            if (!p.b_resident) {
              const long long tb0 = clock64();
              mbar_wait(bar_bempty + 8 * bs, bph ^ 1, p.err, 4);
              t_wait += clock64() - tb0;
            }
            if (ptx::elect_one()) {
              const uint32_t full = bar_bfull + 8 * bs;
              const int kcol = sc ? RS * p.C + cc * 64 : (int)p.tap_w[t] * p.C + cc * 64;
              if (leader) ptx::mbar_arrive_expect_tx(full, (uint32_t)CG * p.b_stage_bytes);
              if (CG == 2) ptx::tma_load_2d_pair(sB + bs * p.b_stage_bytes, &mapB, full, kcol, n_row);
              else ptx::tma_load_2d(sB + bs * p.b_stage_bytes, &mapB, full, kcol, n_row);
            }
            __syncwarp();
            if (++bs == nb) { bs = 0; bph ^= 1; }
          }
        }
      }
      first = false;
    }
    if (p.prof && lane == 0) { p.prof[blockIdx.x * 8 + 0] = t_wait; p.prof[blockIdx.x * 8 + 1] = clock64() - t_all0; }
  } else if (warp == 1 && leader) {
    // ===================================== MMA issuer =========================================
    // The single issuing thread paces every tile with N <= 128 unless its instruction stream is short (48 / 64 clk of
    // tensor-pipe time per MMA at N = 64 / 128): descriptors are 32-bit low words advanced by adds, four K=16 steps go
    // out in one asm block, resident weights are waited for during the first tile only, and nothing else sits between two taps.
    uint32_t ab = 0, aph = 0, bs = 0, bph = 0, it = 0;
    long long t_full = 0, t_tempty = 0;
    const long long t_all0 = clock64();
    const uint32_t idesc = p.idesc;
    const uint32_t desc_hi = (uint32_t)(ptx::make_smem_desc(0, 1024, 2) >> 32);
    // descriptor of position o0 (row offset Wv) in A buffer 0; +8 per pixel row (128 B >> 4)
    const uint32_t a_lo0 = (uint32_t)ptx::make_smem_desc(sA + (uint32_t)Wv * 128u, 1024, 2);
    const uint32_t b_lo0 = (uint32_t)ptx::make_smem_desc(sB, 1024, 2);
    const uint32_t a_step = p.a_buf_bytes >> 4, b_step = p.b_stage_bytes >> 4;
    const int R = p.R, S = p.S, cchunks = p.cchunks, n_tile = p.n_tile;
    const bool resident = p.b_resident != 0;
    const bool elected = ptx::elect_one();       // the same lane issues every MMA and every commit
    SegList segs;
    segs.init(p, unit, nunits);
    for (int k = 0; k < segs.nseg; ++k, ++it) {
      const Seg sg = segs.at(k);
      const uint32_t a = it & 1, tph = (it >> 1) & 1;
      const long long te0 = clock64();
      mbar_wait(bar_tempty + 8 * a, tph ^ 1, p.err, 1);
      t_tempty += clock64() - te0;
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * (uint32_t)n_tile;
      uint32_t acc = 0u;
      uint32_t b_lo = b_lo0;                     // resident: (chunk, tap) boxes in order from the start of sB
      const int nmain = cchunks * p.nplanes, nchunks_all = nmain + p.c2chunks;
      uint32_t ridx = 0;                         // resident: running index of the next weight box
      for (int cc = 0; cc < nchunks_all; ++cc) { // cc = load unit: (chunk, plane), then the fused-shortcut chunks
        const bool sc = cc >= nmain;             // chunk of the fused 1x1 shortcut: one tap, the un-shifted view
        const int pi = (sc || p.nplanes == 1) ? 0 : cc % p.nplanes;
        int t0 = sc ? 0 : p.pl_first[pi], t1 = sc ? 1 : p.pl_first[pi + 1];
        {                                        // the part of this unit inside the segment (see the producer)
          const int ibase = sc ? cchunks * R * S + (cc - nmain) : (p.nplanes == 1 ? cc : cc / p.nplanes) * R * S;
          const int nt = t1 - t0;
          t0 = max(t0, sg.it0 - ibase);
          t1 = min(t1, sg.it1 - ibase);
          if (t0 >= t1) { if (resident && !sc) { ridx += (uint32_t)nt; b_lo += (uint32_t)nt * b_step; } continue; }
        }
        const long long tf0 = clock64();
        mbar_wait(bar_afull + 8 * ab, aph, p.err, 2);
        t_full += clock64() - tf0;
        if (stamp && it == 0 && cc == 0 && lane == 0) p.prof[1402] = clock64() - t_entry;
        ptx::tc_fence_after();
        const uint32_t a_row = a_lo0 + ab * a_step;
        if (sc) {
          const uint32_t bi = resident ? (uint32_t)(cchunks * R * S + (cc - nmain)) : bs;
          if (!resident || it == 0) {
            mbar_wait(bar_bfull + 8 * bi, resident ? 0u : bph, p.err, 5);
            ptx::tc_fence_after();
          }
          if (elected) {
            ptx::umma_f16_x4<CG>(d_tmem, a_row + p.ctr, b_lo0 + bi * b_step, desc_hi, idesc, acc);
            if (!resident) {
              if (CG == 2) ptx::umma_commit_pair(bar_bempty + 8 * bs, 3);
              else ptx::umma_commit(bar_bempty + 8 * bs);
            }
          }
          acc = 1u;
          if (!resident) { if (++bs == nb) { bs = 0; bph ^= 1; } }
        } else if (resident) {
          if (it == 0) {                         // weights are loaded once per CTA, unit by unit behind the A rows
            for (int t = t0; t < t1; ++t) mbar_wait(bar_bfull + 8 * (ridx + (uint32_t)(t - t0)), 0, p.err, 5);
            ptx::tc_fence_after();
          }
          for (int t = t0; t < t1; ++t, b_lo += b_step) {
            if (elected) ptx::umma_f16_x4<CG>(d_tmem, a_row + p.tap_aoff[t], b_lo, desc_hi, idesc, acc);
            acc = 1u;
          }
          ridx += (uint32_t)(t1 - t0);
        } else {
          for (int t = t0; t < t1; ++t) {
            mbar_wait(bar_bfull + 8 * bs, bph, p.err, 5);
            ptx::tc_fence_after();
            if (elected) {
              ptx::umma_f16_x4<CG>(d_tmem, a_row + p.tap_aoff[t], b_lo0 + bs * b_step, desc_hi, idesc, acc);
              if (CG == 2) ptx::umma_commit_pair(bar_bempty + 8 * bs, 3);
              else ptx::umma_commit(bar_bempty + 8 * bs);
            }
            acc = 1u;
            if (++bs == nb) { bs = 0; bph ^= 1; }
          }
        }
        if (elected) {
          if (CG == 2) ptx::umma_commit_pair(bar_aempty + 8 * ab, 3);
          else ptx::umma_commit(bar_aempty + 8 * ab);
        }
        if (++ab == na) { ab = 0; aph ^= 1; }
      }
      if (elected) {                             // accumulator of this segment complete (commits are ordered)
        if (CG == 2) ptx::umma_commit_pair(bar_tfull + 8 * a, 3);
        else ptx::umma_commit(bar_tfull + 8 * a);
      }
    }
    if (p.prof && lane == 0) { p.prof[blockIdx.x * 8 + 2] = t_full; p.prof[blockIdx.x * 8 + 3] = t_tempty; p.prof[blockIdx.x * 8 + 4] = clock64() - t_all0; }
    if (stamp && lane == 0) p.prof[1403] = clock64() - t_entry;
  } else if (warp >= 4) {
    // ===================================== epilogue ===========================================
    // A thread owns one accumulator ROW (TMEM lane) and 32 channels per chunk; rows are transposed through a per-warp
    // shared-memory stage, 16 channels (32 B per row) at a time -- 1 KB per warp, which is what lets the weights of the
    // 128-channel layers stay resident next to the A buffers -- so that every global access of the warp moves 16 rows x
    // 32 contiguous bytes: lane l serves row 16i + l/2, 16-byte piece l%2 (i = 0, 1), twice per 32-channel chunk.  The residual operand is read in that same coalesced pattern and
    // added AFTER the transposition (the reference rounds to fp16 between batchnorm, add and relu as well:
    // planer/layer.py:125-127, :93-95, :44-46), so its loads need no row-owner gather and are issued one tile ahead.
    const int ew = warp & 3;                 // the TMEM lane quarter this warp may read (warp % 4)
    const int eg = (warp - 4) >> 2;          // epilogue group: 0 -> first half of the tile's columns, 1 -> second half
    const int et = threadIdx.x - 128;        // 0..255
    uint32_t it = 0, ebuf = 0;
    int staged_n = -1;
    long long t_tfull = 0;
    const long long t_all0 = clock64();
    const int nchunks = p.n_tile >> 5, half = (nchunks + 1) >> 1;
    // one 32-channel chunk per tile (Cout <= 32): the two column groups take the tiles in turn (group = accumulator buffer)
    // instead of one group idling -- such layers (YOLO's 3 -> 32 and 64 -> 32 convs) are epilogue-bound
    const bool alternate = nchunks == 1;
    const int c_begin = alternate ? 0 : eg * half * 32, c_end = alternate ? 32 : min(nchunks, (eg + 1) * half) * 32;
    const bool has_res = p.res != nullptr;
    const uint32_t HvWv = (uint32_t)p.HvWv, uWv = (uint32_t)Wv, uMv = (uint32_t)p.Mv;
    uint8_t* st_o = stage + (warp - 4) * (p.stage_wide ? 2048 : 1024);
    const uint32_t my_sw = (uint32_t)((lane >> 2) & 1);
    const int piece = lane & 1;

    // geometry of a tile for this thread: its own row (scalar tail path) and the four rows it serves in the coalesced
    // pattern; pixel index -1 = padded rim / beyond the tensor (computed and discarded)
    struct Geo { int own; int row[2]; };
    auto tile_geo = [&](int tile_) {
      Geo g;
      const uint32_t m_idx_ = (uint32_t)tile_ - fast_div((uint32_t)tile_, p.div_mt) * (uint32_t)p.num_m_tiles;
      const uint32_t o = (m_idx_ * CG + rank) * kTileM + (uint32_t)(ew * 32 + lane);
      g.own = -1;
      if (o < uMv) {
        const uint32_t img = fast_div(o, p.div_hvwv), rem = o - img * HvWv;
        const uint32_t pr = fast_div(rem, p.div_wv), q = rem - pr * uWv;
        if (pr < (uint32_t)p.OH && q < (uint32_t)p.OW) g.own = (int)((img * (uint32_t)p.OH + pr) * (uint32_t)p.OW + q);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) g.row[i] = __shfl_sync(0xffffffffu, g.own, 16 * i + (lane >> 1));
      return g;
    };
    // residual pieces of a tile's FIRST chunk, fetched one tile ahead: by the time an epilogue warp reaches a tile its
    // accumulator is usually complete, and a load issued then would expose the DRAM latency once per tile
    uint4 rvp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) rvp[i] = make_uint4(0u, 0u, 0u, 0u);
    auto fetch_res = [&](const Geo& g, int cb, uint4 (&dst)[4]) {      // [16-channel half h][row i] -> dst[2h + i]
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (g.row[i] >= 0)
            dst[2 * h + i] = *reinterpret_cast<const uint4*>(p.res + (size_t)g.row[i] * p.rld + p.rcoff + cb + h * 16 + piece * 8);
    };
    SegList segs;
    segs.init(p, unit, nunits);
    const Seg sg0 = segs.at(0);
    Geo gn = tile_geo(segs.nseg > 0 ? sg0.tile : 0);
    if (has_res && p.vec_ok && c_begin < c_end && segs.nseg > 0 && sg0.mode != SEG_CONTRIB && (!alternate || eg == 0)) {
      const int cb = (int)fast_div((uint32_t)sg0.tile, p.div_mt) * p.n_tile + c_begin;
      if (cb + 32 <= p.Cout) fetch_res(gn, cb, rvp);
    }

    for (int k = 0; k < segs.nseg; ++k, ++it) {
      const Seg sg = segs.at(k);
      const int tile = sg.tile;
      const uint32_t a = it & 1, tph = (it >> 1) & 1;
      const int n_idx = (int)fast_div((uint32_t)tile, p.div_mt);
      const int n0 = n_idx * p.n_tile;
      // per-channel scale/shift of this tile's channel block: re-staged only when the block changes
      if (n_idx != staged_n) {
        staged_n = n_idx;
        ebuf ^= 1u;
        __half* hsc = reinterpret_cast<__half*>(epi) + ebuf * 512, *hsf = hsc + 256;
        for (int i = et; i < p.n_tile; i += 256) {
          const int c = n0 + i;
          float sc = 0.f, sf = 0.f;
          if (c < p.Cout) { sc = p.scale ? __ldg(p.scale + c) : 1.f; sf = p.shift ? __ldg(p.shift + c) : 0.f; }
          hsc[i] = __float2half_rn(sc); hsf[i] = __float2half_rn(sf);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const __half* ep_hscale = reinterpret_cast<const __half*>(epi) + ebuf * 512, *ep_hshift = ep_hscale + 256;

      const Geo g = gn;
      uint4 rv[4], rvn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { rv[i] = rvp[i]; rvn[i] = make_uint4(0u, 0u, 0u, 0u); }
      // next segment: geometry + residual of its first chunk (in flight during this whole epilogue)
      if (k + 1 < segs.nseg) {
        const Seg sn = segs.at(k + 1);
        gn = tile_geo(sn.tile);
        if (has_res && p.vec_ok && c_begin < c_end && sn.mode != SEG_CONTRIB && (!alternate || ((it + 1) & 1u) == (uint32_t)eg)) {
          const int cb = (int)fast_div((uint32_t)sn.tile, p.div_mt) * p.n_tile + c_begin;
          if (cb + 32 <= p.Cout) fetch_res(gn, cb, rvp);
        }
      }

      const long long tt0 = clock64();
      mbar_wait(bar_tfull + 8 * a, tph, p.err, 3);
      t_tfull += clock64() - tt0;
      const bool estamp = stamp && threadIdx.x == 128 && k == segs.nseg - 1;     // debug: time stamps of the last epilogue
      int estamp_i = 1408;
      if (estamp) p.prof[estamp_i++] = clock64() - t_entry;
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + a * (uint32_t)p.n_tile;
      const int ce_ = (alternate && (it & 1u) != (uint32_t)eg) ? c_begin : c_end;      // not this group's tile
      if (sg.mode == SEG_CONTRIB) {
        // stream-K: this unit computed the tail of a tile another unit owns -- the raw fp32 accumulator goes to the
        // workspace in the order the registers hold it: [unit][rank][lane quarter][32-column chunk][4-column group][lane],
        // so that every store / load instruction of a warp moves 512 contiguous bytes (row-major rows cost 32 sectors each)
        float4* wq = reinterpret_cast<float4*>(p.sk_ws) +
                     (size_t)((unit * CG + (int)rank) * 4 + ew) * (size_t)(p.n_tile >> 5) * 256 + lane;
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          __syncwarp();
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(t_row + c0, v);
          ptx::tmem_ld_wait();
          if (c0 + 32 >= c_end) {
            ptx::tc_fence_before();
            if (CG == 2 && !leader) ptx::mbar_arrive_remote(bar_tempty + 8 * a, 0);
            else ptx::mbar_arrive(bar_tempty + 8 * a);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(wq + ((c0 >> 5) * 8 + j) * 32,
                   make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3])));
        }
        __threadfence();                  // the partial is visible device-wide before its flag
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile int*>(p.sk_flags + (unit * 2 + (int)rank) * 8 + (warp - 4)) = 1;
        continue;
      }
      int npieces = 0;                    // stream-K owner: units unit+1 .. unit+npieces hold the rest of this tile
      if (sg.mode == SEG_OWNER) {
        const long long tile_end = (long long)(tile + 1) * p.ipt;
        for (int v = unit + 1; v < nunits; ++v) {
          long long g0, g1;
          SegList::range(p, v, nunits, g0, g1);
          if (g0 >= tile_end) break;
          if (g1 > g0) npieces = v - unit;
        }
        if (lane == 0) {
          for (int pc = 1; pc <= npieces; ++pc) {
            volatile int* f = p.sk_flags + ((unit + pc) * 2 + (int)rank) * 8 + (warp - 4);
            const long long t0 = clock64();
            while (*f == 0) {
              if (*reinterpret_cast<volatile int*>(p.err) != 0) break;
              if (clock64() - t0 > kWatchdogCycles) {
                if (atomicCAS(p.err, 0, 4) == 0) { p.err[1] = blockIdx.x; p.err[2] = 6; p.err[3] = unit + pc; }
                break;
              }
            }
            *f = 0;                       // consumed: the next launch starts from clean flags
          }
          __threadfence();
        }
        __syncwarp();
      }
      if (c_begin >= ce_) {               // nothing to read for this group: release at once
        ptx::tc_fence_before();
        if (CG == 2 && !leader) ptx::mbar_arrive_remote(bar_tempty + 8 * a, 0);
        else ptx::mbar_arrive(bar_tempty + 8 * a);
      }
      for (int c0 = c_begin; c0 < ce_; c0 += 32) {
        __syncwarp();
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(t_row + c0, v);
        const int cb = n0 + c0;
        const bool fast = p.vec_ok && (cb + 32 <= p.Cout);
        if (has_res && p.vec_ok && c0 + 32 < ce_ && cb + 64 <= p.Cout) fetch_res(g, cb + 32, rvn);   // one chunk ahead
        ptx::tmem_ld_wait();
        if (estamp && estamp_i < 1424) p.prof[estamp_i++] = clock64() - t_entry;
        if (c0 + 32 >= ce_) {
          ptx::tc_fence_before();
          if (CG == 2 && !leader) ptx::mbar_arrive_remote(bar_tempty + 8 * a, 0);
          else ptx::mbar_arrive(bar_tempty + 8 * a);
        }
        for (int pc = 1; pc <= npieces; ++pc) {     // stream-K owner: add the other units' partial sums, in unit order
          const float4* pq = reinterpret_cast<const float4*>(p.sk_ws) +
                             (size_t)(((unit + pc) * CG + (int)rank) * 4 + ew) * (size_t)(p.n_tile >> 5) * 256 + lane;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 q4 = __ldcg(pq + ((c0 >> 5) * 8 + j) * 32);
            v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + q4.x);
            v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + q4.y);
            v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + q4.z);
            v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + q4.w);
          }
        }
        if (fast) {
          // The math of a chunk is compiled twice, once with the activation fixed to ReLU: with the generic runtime
          // switch this block is ~900 dependent instructions per chunk and the epilogue, not the tensor pipe, paces
          // Cout = 64 layers.  With a residual the activation (or the add, for the Darknet shortcut x + act(..))
          // happens after the transposition, in packed fp16 -- the reference rounds to fp16 between batchnorm, add
          // and relu too, and fp16 + fp16 rounded once is exactly what HADD2 computes.
          // The accumulator is rounded to fp16 FIRST (the reference's conv output is an fp16 array: planer/layer.py:22-26
          // on fp16 inputs), transposed, and batchnorm / bias, add and the activation run on the transposed pieces in
          // packed fp16 -- x*K+B as one HFMA2 (the reference: two fp16 roundings, planer/layer.py:125-127).  After the
          // transposition a lane owns the SAME 8 channels in each of its four rows, so scale/shift are two 16-byte
          // loads per chunk instead of sixteen broadcast loads per row, and the math is 4 HFMA2 per row instead of
          // 8 FFMA + 8 FMNMX: the shared-memory pipe these loads shared with the MMA operand reads is what bounds the
          // Cout = 64 / 128 layers (profiles/r01_smem_budget.md).
          auto chunk_math = [&](auto act_tag, auto wide_tag) {
            constexpr int kAct = decltype(act_tag)::value;      // 1 = ReLU, 2 = LeakyReLU (0 <= alpha <= 1), 0 = generic
            constexpr bool kWide = decltype(wide_tag)::value != 0;
            const bool act_first = !has_res || p.res_after;
            const __half2 zero2 = __float2half2_rn(0.f), alpha2 = __float2half2_rn(p.alpha);
            auto act2 = [&](__half2 x) {
              if (kAct == 1) return __hmax2(x, zero2);
              if (kAct == 2) return __hmax2(x, __hmul2(x, alpha2));
              const float2 f = __half22float2(x);
              return __floats2half2_rn(plnr_apply_act(f.x, p.act, p.alpha), plnr_apply_act(f.y, p.act, p.alpha));
            };
            auto finish = [&](uint4& val, const __half2* sch, const __half2* sfh, const uint4& res4) {
              __half2* vh = reinterpret_cast<__half2*>(&val);
              const __half2* rh = reinterpret_cast<const __half2*>(&res4);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __half2 x;
                if (kAct == 1 && act_first) x = __hfma2_relu(vh[e], sch[e], sfh[e]);
                else {
                  x = __hfma2(vh[e], sch[e], sfh[e]);
                  if (act_first) x = act2(x);
                }
                if (has_res) {
                  x = __hadd2(x, rh[e]);
                  if (!p.res_after) x = act2(x);
                }
                vh[e] = x;
              }
            };
            if (kWide) {
              // 2 KB stage per warp: the whole 32-column chunk is transposed in ONE round trip (four STS, one warp barrier,
              // four LDS) and the four 16-byte results are finished together -- twice the independent work per dependent
              // step of this latency-bound chain (12 % issue utilisation with the two-round version, profiles/r02_kernel_experiments.md 6).
              // Row of 64 B = four 16-byte slots, slot = piece ^ f(row), f = bit-reversed (row >> 1) & 3: conflict-free both ways.
              const uint32_t wsw = (((uint32_t)lane >> 1) & 1u) << 1 | (((uint32_t)lane >> 2) & 1u);
              const uint32_t rsw = (((uint32_t)lane >> 2) & 1u) << 1 | (((uint32_t)lane >> 3) & 1u);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 pk;
                pk.x = pack_half2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
                pk.y = pack_half2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
                pk.z = pack_half2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
                pk.w = pack_half2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
                *reinterpret_cast<uint4*>(st_o + lane * 64 + (((uint32_t)q ^ wsw) << 4)) = pk;
              }
              uint4 sc4[2], sf4[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                sc4[h] = *reinterpret_cast<const uint4*>(ep_hscale + c0 + h * 16 + piece * 8);
                sf4[h] = *reinterpret_cast<const uint4*>(ep_hshift + c0 + h * 16 + piece * 8);
              }
              __syncwarp();
              uint4 val[4];
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const int row = 16 * i + (lane >> 1);
                  val[2 * h + i] = *reinterpret_cast<const uint4*>(st_o + row * 64 + ((((uint32_t)(2 * h + piece)) ^ rsw) << 4));
                }
              if (POOL && p.pool) {
                // GlobalAveragePool folded into this epilogue (the layer's only consumer): the lane's two rows are added in
                // fp32, then a reduce-scatter over the sixteen lanes that hold the same channels (lane bits 1..4: 8 + 4 + 2 + 1
                // shuffles) leaves every lane with the 32-row sum of ONE channel of the chunk; nothing is stored to y.
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                  for (int i = 0; i < 2; ++i)
                    finish(val[2 * h + i], reinterpret_cast<const __half2*>(&sc4[h]), reinterpret_cast<const __half2*>(&sf4[h]), rv[2 * h + i]);
                float s16[16];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const __half2* r0 = reinterpret_cast<const __half2*>(&val[2 * h]);
                  const __half2* r1 = reinterpret_cast<const __half2*>(&val[2 * h + 1]);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 a0 = __half22float2(r0[e]), a1 = __half22float2(r1[e]);
                    s16[h * 8 + 2 * e] = (g.row[0] >= 0 ? a0.x : 0.f) + (g.row[1] >= 0 ? a1.x : 0.f);
                    s16[h * 8 + 2 * e + 1] = (g.row[0] >= 0 ? a0.y : 0.f) + (g.row[1] >= 0 ? a1.y : 0.f);
                  }
                }
                float s8[8], s4[4], s2[2];
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float got = __shfl_xor_sync(0xffffffffu, b4 ? s16[j] : s16[j + 8], 16);
                  s8[j] = (b4 ? s16[j + 8] : s16[j]) + got;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float got = __shfl_xor_sync(0xffffffffu, b3 ? s8[j] : s8[j + 4], 8);
                  s4[j] = (b3 ? s8[j + 4] : s8[j]) + got;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const float got = __shfl_xor_sync(0xffffffffu, b2 ? s4[j] : s4[j + 2], 4);
                  s2[j] = (b2 ? s4[j + 2] : s4[j]) + got;
                }
                const float tot = (b1 ? s2[1] : s2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? s2[0] : s2[1], 2);
                // the kept value: 16-channel half b4, element 4 b3 + 2 b2 + b1 of this lane's 8-channel piece
                const int ch = cb + (b4 ? 16 : 0) + piece * 8 + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
                // (image, 32-row part) this warp's rows belong to
                const uint32_t m_idx_ = (uint32_t)tile - (uint32_t)n_idx * (uint32_t)p.num_m_tiles;
                const uint32_t ob = (m_idx_ * CG + rank) * kTileM + (uint32_t)(ew * 32);
                if (ob < uMv) {
                  const uint32_t img = fast_div(ob, p.div_hvwv);
                  const uint32_t pool_row = img * (uint32_t)p.pool_parts + ((ob - img * HvWv) >> 5);
                  p.pool[(size_t)pool_row * p.Cout + ch] = tot;
                }
                __syncwarp();
                return;
              }
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  finish(val[2 * h + i], reinterpret_cast<const __half2*>(&sc4[h]), reinterpret_cast<const __half2*>(&sf4[h]), rv[2 * h + i]);
                  const bool ok = g.row[i] >= 0;                         // rim rows: computed, not stored (predicated store, no branch)
                  __half* dst = p.y + (size_t)(ok ? g.row[i] : 0) * p.yld + p.ycoff + cb + h * 16 + piece * 8;
                  if (ok) *reinterpret_cast<uint4*>(dst) = val[2 * h + i];
                }
              __syncwarp();                    // the stage is rewritten by the next chunk
              return;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int q = 2 * h; q < 2 * h + 2; ++q) {
                uint4 pk;
                pk.x = pack_half2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
                pk.y = pack_half2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
                pk.z = pack_half2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
                pk.w = pack_half2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
                *reinterpret_cast<uint4*>(st_o + lane * 32 + ((((uint32_t)q & 1u) ^ my_sw) << 4)) = pk;
              }
              const uint4 sc4 = *reinterpret_cast<const uint4*>(ep_hscale + c0 + h * 16 + piece * 8);
              const uint4 sf4 = *reinterpret_cast<const uint4*>(ep_hshift + c0 + h * 16 + piece * 8);
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int row = 16 * i + (lane >> 1);
                uint4 val = *reinterpret_cast<const uint4*>(st_o + row * 32 + ((((uint32_t)piece) ^ ((uint32_t)(row >> 2) & 1u)) << 4));
                if (g.row[i] >= 0) {
                  finish(val, reinterpret_cast<const __half2*>(&sc4), reinterpret_cast<const __half2*>(&sf4), rv[2 * h + i]);
                  *reinterpret_cast<uint4*>(p.y + (size_t)g.row[i] * p.yld + p.ycoff + cb + h * 16 + piece * 8) = val;
                }
              }
              __syncwarp();                    // the 1 KB stage is rewritten by the next half / chunk
            }
          };
          const bool leaky01 = p.act == PLNR_ACT_LEAKY && p.alpha >= 0.f && p.alpha <= 1.f;
          if (p.stage_wide) {
            if (p.act == PLNR_ACT_RELU) chunk_math(ActTag<1>{}, ActTag<1>{});
            else if (leaky01) chunk_math(ActTag<2>{}, ActTag<1>{});
            else chunk_math(ActTag<0>{}, ActTag<1>{});
          } else {
            if (p.act == PLNR_ACT_RELU) chunk_math(ActTag<1>{}, ActTag<0>{});
            else if (leaky01) chunk_math(ActTag<2>{}, ActTag<0>{});
            else chunk_math(ActTag<0>{}, ActTag<0>{});
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) rv[i] = rvn[i];
          if (estamp && estamp_i < 1424) p.prof[estamp_i++] = clock64() - t_entry;
        } else if (g.own >= 0) {
          // NCHW exit: pixel g.own = img * OH*OW + pix  ->  y[(img * Cout + c) * OH*OW + pix]
          const uint32_t img_ = p.y_nchw ? (uint32_t)g.own / (uint32_t)p.ohw : 0u;
          __half* yrow = p.y_nchw ? p.y + ((size_t)img_ * p.Cout) * p.ohw + ((uint32_t)g.own - img_ * (uint32_t)p.ohw)
                                  : p.y + (size_t)g.own * p.yld + p.ycoff;
          const size_t cstride = p.y_nchw ? (size_t)p.ohw : 1;
          const __half* rrow = has_res ? p.res + (size_t)g.own * p.rld + p.rcoff : nullptr;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int c = cb + e;
            if (c < p.Cout) {
              float o1 = fmaf(__half2float(__float2half_rn(__uint_as_float(v[e]))), __half2float(ep_hscale[c0 + e]),
                              __half2float(ep_hshift[c0 + e]));
              const float rf = rrow ? __half2float(rrow[c]) : 0.f;
              if (!p.res_after) o1 += rf;
              o1 = plnr_apply_act(o1, p.act, p.alpha);
              if (p.res_after) o1 += rf;
              yrow[(size_t)c * cstride] = __float2half_rn(o1);
            }
          }
        }
      }
    }
    if (p.prof && threadIdx.x == 128) { p.prof[blockIdx.x * 8 + 5] = t_tfull; p.prof[blockIdx.x * 8 + 6] = clock64() - t_all0; }
    if (stamp && threadIdx.x == 128) p.prof[1404] = clock64() - t_entry;
  }

  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (stamp && threadIdx.x == 0) p.prof[1405] = clock64() - t_entry;
  if (warp == 2) {
    ptx::tc_fence_after();
    if (CG == 2) ptx::tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else ptx::tmem_dealloc(tmem_base, p.tmem_cols);
    if (stamp && lane == 0) {
      p.prof[1406] = clock64() - t_entry;
      unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      p.prof[1407] = (long long)gt;
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static int g_driver_version = 0;

static int resolve_driver() {
  if (g_encode_tiled) return PLNR_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    plnr_set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return PLNR_ERR_DRIVER;
  }
  g_encode_tiled = (EncodeTiledFn)fn;
  cudaDriverGetVersion(&g_driver_version);
  return PLNR_OK;
}

static void small_tensor_fixup(CUtensorMap* m, uint64_t tensor_bytes) {
  if (g_driver_version <= 13010 && tensor_bytes < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

static inline int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

struct ShiftPlan {
  int stage_wide;
  int cg, n_tile, na, nb, b_resident, rows_max, Wv, Hv, halo;
  int cs, pv_t, pv_l;       // conv stride (1 | 2); virtual padding rows / columns of the (plane) grid
  uint32_t a_buf_bytes, b_stage_bytes;
  size_t smem_bytes;
  bool ok;
};

// Shared-memory plan for a problem; ok == false when the shift kernel does not apply / does not fit.
static ShiftPlan make_plan(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y, int sm_count,
                           int c2 = 0) {
  ShiftPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (d->dtype != PLNR_F16 || d->groups != 1 || d->stride_h != d->stride_w || d->stride_h < 1 || d->stride_h > 2) return pl;
  if (x->c % 64 != 0 || x->ld % 8 != 0 || x->coff % 8 != 0 || (reinterpret_cast<uintptr_t>(x->ptr) & 15)) return pl;
  if (d->pad_b > d->pad_t || d->pad_r > d->pad_l) return pl;
  if (d->kh * d->kw > kMaxB) return pl;
  pl.cs = d->stride_h;
  if (pl.cs == 1) {
    pl.pv_t = d->pad_t; pl.pv_l = d->pad_l;
    pl.Wv = x->w + d->pad_l;
    pl.Hv = x->h + d->pad_t;
    pl.halo = (d->kh - 1) * d->dil_h * pl.Wv + (d->kw - 1) * d->dil_w;
  } else {
    // stride 2 (phase planes, see ShiftParams): tap r reads plane row p + floor((r - pad_t) / 2); the grid needs
    // ceil(pad / 2) virtual padding rows / columns and holds the larger plane (ceil(H / 2) x ceil(W / 2))
    if (d->dil_h != 1 || d->dil_w != 1 || c2 != 0) return pl;
    // Opt-in (PLNR_SHIFT_S2=1): measured on ResNet-18 batch 128 the three 3x3/s2 convolutions take 31-35 us this way
    // against 23-29 us through TMA im2col (profiles/r02_kernel_experiments.md) -- the phase planes leave the tile
    // quantisation of the N=256 CTA-pair schedule (2 rounds for 1.5) where the im2col kernel picks N=128 tiles that fill
    // 2.65 of 3 rounds.  Correct (tests force it on), kept for shapes where the im2col TMA rate binds.
    { const char* e = getenv("PLNR_SHIFT_S2"); if (!(e && atoi(e))) return pl; }
    pl.pv_t = (d->pad_t + 1) / 2; pl.pv_l = (d->pad_l + 1) / 2;
    pl.Wv = (x->w + 1) / 2 + pl.pv_l;
    pl.Hv = (x->h + 1) / 2 + pl.pv_t;
    const int r_max = floor_div(d->kh - 1 - d->pad_t, 2) + pl.pv_t, s_max = floor_div(d->kw - 1 - d->pad_l, 2) + pl.pv_l;
    pl.halo = r_max * pl.Wv + s_max;
    // reads past the end of a virtual row / image must land in the zero padding of the next one
    if (y->w - 1 + s_max - pl.Wv >= pl.pv_l || y->h - 1 + r_max - pl.Hv >= pl.pv_t) return pl;
    if (y->w > pl.Wv || y->h > pl.Hv) return pl;
  }
  if (pl.Wv * pl.cs > 256) return pl;
  pl.rows_max = (pl.Wv - 1 + kTileM - 1 + pl.halo) / pl.Wv + 1;
  // position o0 sits at row offset Wv whatever its column `off` (see the producer): the rows of a tile start at
  // Wv - off and there are floor((off + 127 + halo) / Wv) + 1 of them -- the buffer is sized for the worst `off`
  int need = 0;
  for (int off = 0; off < pl.Wv; ++off) {
    const int end = pl.Wv - off + ((off + kTileM - 1 + pl.halo) / pl.Wv + 1) * pl.Wv;
    if (end > need) need = end;
  }
  pl.a_buf_bytes = (uint32_t)round_up(need * 128, 1024);
  const long long Mv = (long long)x->n * pl.Hv * pl.Wv;
  if (Mv + pl.halo + 1024 >= (1ll << 31)) return pl;       // positions (+ halo) stay below 2^31: 32-bit multiply-high divisions
  const double eff = (double)y->h * y->w / ((double)pl.Hv * pl.Wv);
  if (eff < 0.70) return pl;                              // small maps: too many discarded rim positions
  const int m_tiles_128 = (int)((Mv + kTileM - 1) / kTileM);
  // A CTA pair (cta_group::2: M = 256, each CTA supplies half of the B rows) is chosen when it is what makes the WEIGHTS
  // RESIDENT: a 128 -> 128 channel 3x3 filter is 288 KB -- too much for one CTA, whose shared-memory pipe then carries
  // 32*N B of streamed-weight TMA writes per MMA on top of the operand reads -- but 144 KB per CTA of a pair fits next to
  // the A buffers (measured, 128 -> 128 @14x14 x512: 50.1 -> 35.7 us, profiles/r01_cg2_resident.md).  With streamed
  // weights the pair was never faster (profiles/r01_shift_cg.md).  PLNR_SHIFT_CTA_GROUP=1|2 forces either.
  pl.cg = 1;
  int forced = 0;
  if (const char* e = getenv("PLNR_SHIFT_CTA_GROUP")) forced = atoi(e);
  if (forced == 2 && m_tiles_128 >= 2) pl.cg = 2;
  const int cout_r = round_up(y->c, 32);
  pl.n_tile = cout_r < 256 ? cout_r : 256;
  if (const char* e = getenv("PLNR_SHIFT_NTILE")) { int v = atoi(e); if (v >= 32 && v <= 256 && v % 32 == 0 && v < pl.n_tile) pl.n_tile = v; }
  pl.b_stage_bytes = (uint32_t)(pl.n_tile / pl.cg) * 128;
  const int kst = d->kh * d->kw * (x->c / 64) + c2 / 64;    // weight boxes per tile: (chunk, tap) + shortcut chunks
  const int num_n_tiles = (y->c + pl.n_tile - 1) / pl.n_tile;
  size_t fixed = kEpiBytes + 16 * kMaxA + 16 * kMaxB + 64 + kStageBytes + 1024;
  const size_t budget = 232448;
  // resident weights: all (chunk, tap) boxes stay in shared memory for the whole kernel
  pl.na = 2;
  auto resident_fits = [&](int cg_, int na_) {
    return num_n_tiles == 1 && kst <= kMaxB &&
           fixed + (size_t)na_ * pl.a_buf_bytes + (size_t)kst * (size_t)(pl.n_tile / cg_) * 128 <= budget;
  };
  // Epilogue stage: 2 KB per warp (one transposition round trip per 32-column chunk) unless those 8 KB are what keeps the
  // weights resident (128 -> 128 channels at 28x28 as a CTA pair: 144 KB of filter + two A buffers fill the SM).
  {
    const bool res_narrow = resident_fits(1, 2) || (m_tiles_128 >= 2 && resident_fits(2, 2));
    fixed += kStageBytes;
    const bool res_wide = resident_fits(1, 2) || (m_tiles_128 >= 2 && resident_fits(2, 2));
    pl.stage_wide = (res_wide || !res_narrow) ? 1 : 0;
    if (const char* e = getenv("PLNR_SHIFT_STAGE_WIDE")) pl.stage_wide = atoi(e) ? 1 : 0;
    if (!pl.stage_wide) fixed -= kStageBytes;
  }
  if (forced == 0 && !resident_fits(1, 2) && m_tiles_128 >= 2 && pl.n_tile % 32 == 0 && resident_fits(2, 2)) {
    pl.cg = 2;
    pl.b_stage_bytes = (uint32_t)(pl.n_tile / 2) * 128;
  }
  // streamed N = 256 weights: a pair halves the B bytes each CTA moves (TMA writes + MMA reads: 20 KB -> 12 KB per K=16 step,
  // under the 128-clk tensor time); measured +1.2 % on the ResNet-18 step (PLNR_SHIFT_CTA_GROUP A/B)
  if (forced == 0 && pl.cg == 1 && !resident_fits(1, 2) && pl.n_tile == 256 && m_tiles_128 >= 2) {
    pl.cg = 2;
    pl.b_stage_bytes = (uint32_t)(pl.n_tile / 2) * 128;
  }
  // stride 2: a plane buffer feeds only 1-4 taps (4-16 MMAs), so the A ring must be deep enough to cover the TMA latency
  // -- a CTA pair halves the weight bytes per CTA, which is what buys the extra A buffers
  int na_cap = kMaxA;
  if (const char* e = getenv("PLNR_SHIFT_S2_NA")) { int v = atoi(e); if (v >= 2 && v <= kMaxA) na_cap = v; }
  if (pl.cs == 2 && forced == 0 && pl.cg == 1 && m_tiles_128 >= 2 && pl.n_tile % 32 == 0 && pl.n_tile >= 64) {
    pl.cg = 2;
    pl.b_stage_bytes = (uint32_t)(pl.n_tile / 2) * 128;
  }
  if (resident_fits(pl.cg, 2)) {
    pl.b_resident = 1;
    pl.nb = kst;
    if (resident_fits(pl.cg, 3)) pl.na = 3;
    if (pl.cs == 2) for (int na = 4; na <= na_cap; ++na) if (resident_fits(pl.cg, na)) pl.na = na;
  } else if (pl.cs == 2) {
    if (fixed + 2 * (size_t)pl.a_buf_bytes + 3 * (size_t)pl.b_stage_bytes > budget) return pl;
    for (int na = 2; na <= na_cap; ++na)
      if (fixed + (size_t)na * pl.a_buf_bytes + 4 * (size_t)pl.b_stage_bytes <= budget) pl.na = na;
    pl.nb = (int)((budget - fixed - (size_t)pl.na * pl.a_buf_bytes) / pl.b_stage_bytes);
    if (pl.nb > 12) pl.nb = 12;
  } else {
    if (fixed + 2 * (size_t)pl.a_buf_bytes + 3 * (size_t)pl.b_stage_bytes > budget) return pl;
    size_t left = budget - fixed - 2 * (size_t)pl.a_buf_bytes;
    pl.nb = (int)(left / pl.b_stage_bytes);
    if (pl.nb > 12) pl.nb = 12;
    if (pl.nb >= 8 && fixed + 3 * (size_t)pl.a_buf_bytes + 6 * (size_t)pl.b_stage_bytes <= budget) {
      pl.na = 3;
      pl.nb = (int)((budget - fixed - 3 * (size_t)pl.a_buf_bytes) / pl.b_stage_bytes);
      if (pl.nb > 12) pl.nb = 12;
    }
  }
  pl.smem_bytes = fixed + (size_t)pl.na * pl.a_buf_bytes + (size_t)pl.nb * pl.b_stage_bytes;
  pl.ok = true;
  (void)sm_count;
  return pl;
}

// Tap tables of a problem, grouped by phase plane (see ShiftParams); stride 1 = one plane in filter order.
static bool fill_planes(ShiftParams& p, const ShiftPlan& pl, const plnr_conv_desc* d) {
  p.cs = pl.cs;
  int nt = 0;
  p.nplanes = 0;
  for (int a = 0; a < pl.cs; ++a) {
    for (int b = 0; b < pl.cs; ++b) {
      const int first = nt;
      for (int r = 0; r < d->kh; ++r) {
        for (int sx = 0; sx < d->kw; ++sx) {
          int rp, sp;
          if (pl.cs == 1) { rp = r * d->dil_h; sp = sx * d->dil_w; }
          else {
            const int dr = floor_div(r - d->pad_t, 2), dq = floor_div(sx - d->pad_l, 2);
            if (r - d->pad_t - 2 * dr != a || sx - d->pad_l - 2 * dq != b) continue;
            rp = dr + pl.pv_t; sp = dq + pl.pv_l;
          }
          p.tap_aoff[nt] = (uint32_t)(rp * pl.Wv + sp) * 8u;
          p.tap_w[nt] = (uint8_t)(r * d->kw + sx);
          ++nt;
        }
      }
      if (nt == first) continue;                         // no tap reads this plane (e.g. 1x1 / stride 2)
      p.pl_first[p.nplanes] = first;
      p.pl_w0[p.nplanes] = pl.cs == 1 ? -d->pad_l : -2 * pl.pv_l + b;
      p.pl_h0[p.nplanes] = pl.cs == 1 ? -d->pad_t : -2 * pl.pv_t + a;
      ++p.nplanes;
    }
  }
  p.pl_first[p.nplanes] = nt;
  return nt == d->kh * d->kw && p.nplanes >= 1;
}

}  // namespace

// Debug / test aid: the virtual grid and tap tables the shift kernel would use for a problem, as integers:
// out = [ok, cs, Hv, Wv, halo, nplanes, ntaps, (w0, h0, first) x 4, (aoff / 8, filter tap) x ntaps]; tests/test_host_logic.py
// replays them in numpy against the oracle convolution, so the plane algebra is pinned without a GPU.
extern "C" int plnr_debug_shift_geometry(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y, int* out, int n) {
  if (!d || !x || !y || !out || n < 19 + 2 * kMaxB) return PLNR_ERR_INVALID;
  memset(out, 0, sizeof(int) * (size_t)n);
  const ShiftPlan pl = make_plan(d, x, y, 148);
  if (!pl.ok) return PLNR_OK;
  ShiftParams p;
  memset(&p, 0, sizeof(p));
  if (!fill_planes(p, pl, d)) return PLNR_OK;
  out[0] = 1; out[1] = pl.cs; out[2] = pl.Hv; out[3] = pl.Wv; out[4] = pl.halo; out[5] = p.nplanes; out[6] = p.pl_first[p.nplanes];
  for (int i = 0; i < 4; ++i) { out[7 + 3 * i] = p.pl_w0[i]; out[8 + 3 * i] = p.pl_h0[i]; out[9 + 3 * i] = p.pl_first[i]; }
  for (int t = 0; t < out[6]; ++t) { out[19 + 2 * t] = (int)(p.tap_aoff[t] / 8u); out[20 + 2 * t] = p.tap_w[t]; }
  return PLNR_OK;
}

namespace {
}  // namespace

bool plnr_conv2d_shift_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y) {
  if (const char* e = getenv("PLNR_NO_SHIFT")) { if (atoi(e)) return false; }
  return make_plan(d, x, y, 148).ok;
}

// GlobalAveragePool fold: 32-row parts per image (> 0) when the shift kernel runs the problem with the wide epilogue stage and
// every TMEM lane quarter (32 consecutive virtual positions) lies inside ONE image; 0 otherwise.
int plnr_conv2d_shift_pool_parts(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y) {
  if (const char* e = getenv("PLNR_NO_POOL_FOLD")) { if (atoi(e)) return 0; }
  if (!plnr_conv2d_shift_supported(d, x, y)) return 0;
  const ShiftPlan pl = make_plan(d, x, y, 148);
  if (!pl.ok || !pl.stage_wide || pl.cs != 1 || y->c % 32 != 0 || (pl.Hv * pl.Wv) % 32 != 0) return 0;
  if (y->ld % 8 != 0 || y->coff % 8 != 0) return 0;
  return pl.Hv * pl.Wv / 32;
}

// The fused shortcut: x2 is the block input, (n, c2, >= (y.h-1)*s2+1, >= (y.w-1)*s2+1); its 1x1 / stride-s2 convolution is
// accumulated into the same TMEM tile as extra k-chunks (weights appended to the packed filter along K).
bool plnr_conv2d_shift_shortcut_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* x2, int s2,
                                          const plnr_tensor* y) {
  if (!plnr_conv2d_shift_supported(d, x, y)) return false;
  if (!x2 || x2->c % 64 != 0 || x2->ld % 8 != 0 || x2->coff % 8 != 0 || (reinterpret_cast<uintptr_t>(x2->ptr) & 15)) return false;
  if (s2 < 1 || s2 > 2 || x2->n != x->n) return false;
  if ((x2->h + s2 - 1) / s2 != y->h || (x2->w + s2 - 1) / s2 != y->w) return false;
  const ShiftPlan pl = make_plan(d, x, y, 148, x2->c);
  return pl.ok && pl.Wv * s2 <= 256;
}

int plnr_conv2d_shift(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w,
                      const plnr_tensor* y, const plnr_epilogue* ep, const plnr_tensor* x2, int s2) {
  // 3-wide filters over 64-channel output blocks whose taps fit in shared memory: the stacked variant (conv_stack.cu)
  if (!(ep && (ep->out_nchw || ep->pool_sum)) && plnr_conv2d_stack_supported(d, x, y, ep, x2, s2)) return plnr_conv2d_stack(ctx, d, x, w, y, ep, x2, s2);
  int rc = resolve_driver();
  if (rc != PLNR_OK) return rc;
  const int c2 = x2 ? x2->c : 0;
  const ShiftPlan pl = make_plan(d, x, y, ctx->sm_count, c2);
  PLNR_REQUIRE(pl.ok, "conv2d(shift): problem not eligible");
  PLNR_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv2d(shift): packed weights must be 16-byte aligned");
  const int cg = pl.cg;

  ShiftParams p;
  memset(&p, 0, sizeof(p));
  p.N = x->n; p.OH = y->h; p.OW = y->w; p.Hv = pl.Hv; p.Wv = pl.Wv; p.HvWv = pl.Hv * pl.Wv;
  p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  p.Mv = (long long)x->n * pl.Hv * pl.Wv;
  p.halo = pl.halo;
  p.R = d->kh; p.S = d->kw; p.dh = d->dil_h; p.dw = d->dil_w;
  PLNR_REQUIRE(fill_planes(p, pl, d), "conv2d(shift): tap table is inconsistent");
  p.C = x->c; p.cchunks = x->c / 64;
  p.c2chunks = c2 / 64; p.s2 = x2 ? s2 : 1;
  p.ctr = (uint32_t)(d->pad_t * pl.Wv + d->pad_l) * 8u;      // the view whose tap reads input pixel (p, q) itself
  p.n_tile = pl.n_tile;
  p.num_m_tiles = (int)((p.Mv + kTileM * cg - 1) / (kTileM * cg));
  const int num_n_tiles = (y->c + p.n_tile - 1) / p.n_tile;
  p.num_tiles = p.num_m_tiles * num_n_tiles;
  p.div_hvwv = make_fastdiv((uint32_t)p.HvWv);
  p.div_wv = make_fastdiv((uint32_t)p.Wv);
  p.div_mt = make_fastdiv((uint32_t)p.num_m_tiles);
  p.div_hv = make_fastdiv((uint32_t)p.Hv);
  p.na = pl.na; p.nb = pl.nb; p.b_resident = pl.b_resident;
  p.stage_wide = pl.stage_wide;
  p.a_buf_bytes = pl.a_buf_bytes; p.b_stage_bytes = pl.b_stage_bytes;
  p.idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)((kTileM * cg) >> 4) << 24);
  uint32_t cols = 32;
  while (cols < 2u * p.n_tile) cols <<= 1;
  p.tmem_cols = cols;
  p.y = (__half*)y->ptr; p.yld = y->ld; p.ycoff = y->coff; p.Cout = y->c;
  bool vec = (y->ld % 8 == 0) && (y->coff % 8 == 0) && ((reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0);
  if (ep) {
    p.scale = ep->scale; p.shift = ep->shift; p.act = ep->act; p.alpha = ep->alpha; p.res_after = ep->res_after_act;
    if (ep->residual) {
      const plnr_tensor* r = ep->residual;
      p.res = (const __half*)r->ptr; p.rld = r->ld; p.rcoff = r->coff;
      vec = vec && (r->ld % 8 == 0) && (r->coff % 8 == 0) && ((reinterpret_cast<uintptr_t>(r->ptr) & 15) == 0);
    }
  }
  p.vec_ok = vec ? 1 : 0;
  if (ep && ep->out_nchw) {          // every thread stores its own pixel, channel by channel: lanes = consecutive pixels of a plane
    PLNR_REQUIRE(!ep->residual, "conv2d(shift): out_nchw cannot be combined with a residual operand");
    p.y_nchw = 1; p.ohw = y->h * y->w; p.vec_ok = 0;
  }
  if (ep && ep->pool_sum) {          // GlobalAveragePool folded into the epilogue (plnr_conv2d_pool_parts)
    const int parts = plnr_conv2d_shift_pool_parts(d, x, y);
    PLNR_REQUIRE(parts > 0 && !ep->out_nchw && p.vec_ok && pl.stage_wide,
                 "conv2d(shift): pool_sum needs plnr_conv2d_pool_parts > 0, 16-byte-aligned views, no out_nchw");
    p.pool = ep->pool_sum; p.pool_parts = parts;
  }
  p.err = ctx->dev_error;
  p.prof = ctx->prof;

  AMaps mapsA;
  CUtensorMap mapB;
  memset(&mapsA, 0, sizeof(mapsA));
  // boxes of up to 256 positions (32 KB); PLNR_SHIFT_ROWBOX=0 keeps the row-by-row loads (A/B)
  p.max_hlog = 0;
  while (p.max_hlog < 3 && (2 << p.max_hlog) * pl.Wv <= 256 && (2 << p.max_hlog) <= pl.Hv) ++p.max_hlog;
  if (const char* e = getenv("PLNR_SHIFT_ROWBOX")) { int v = atoi(e); if (v >= 0 && v < p.max_hlog) p.max_hlog = v; }
  for (int k = 0; k <= p.max_hlog; ++k) {
    const cuuint64_t dims[4] = {(cuuint64_t)x->c, (cuuint64_t)x->w, (cuuint64_t)x->h, (cuuint64_t)x->n};
    const cuuint64_t strides[3] = {(cuuint64_t)x->ld * 2, (cuuint64_t)x->w * x->ld * 2,
                                   (cuuint64_t)x->h * x->w * x->ld * 2};
    // stride 2: Wv positions of a plane = every other pixel of a 2*Wv-wide box (element stride 2 along W), rows likewise
    const cuuint32_t box[4] = {64, (cuuint32_t)(pl.Wv * pl.cs), (cuuint32_t)((1 << k) * pl.cs), 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)pl.cs, (cuuint32_t)pl.cs, 1};
    void* gaddr = (void*)((__half*)x->ptr + x->coff);
    CUresult r = g_encode_tiled(&mapsA.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, gaddr, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      plnr_set_error("cuTensorMapEncodeTiled(A, %d rows) failed with CUresult %d (C=%d W=%d H=%d N=%d Wv=%d stride=%d)", 1 << k,
                     (int)r, x->c, x->w, x->h, x->n, pl.Wv, pl.cs);
      return PLNR_ERR_DRIVER;
    }
    small_tensor_fixup(&mapsA.m[k], (uint64_t)x->n * x->h * x->w * x->ld * 2);
  }
  {
    const int Ktot = d->kh * d->kw * x->c + c2;
    const cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)y->c};
    const cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)(p.n_tile / cg)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      plnr_set_error("cuTensorMapEncodeTiled(B) failed with CUresult %d (K=%d Cout=%d)", (int)r, Ktot, y->c);
      return PLNR_ERR_DRIVER;
    }
    small_tensor_fixup(&mapB, (uint64_t)Ktot * y->c * 2);
  }

  CUtensorMap mapA2 = mapsA.m[0];
  if (x2) {
    // x2 sampled at stride s2 in W (element stride) and H (row coordinate): one box = Wv positions of one virtual row
    const cuuint64_t dims[4] = {(cuuint64_t)x2->c, (cuuint64_t)x2->w, (cuuint64_t)x2->h, (cuuint64_t)x2->n};
    const cuuint64_t strides[3] = {(cuuint64_t)x2->ld * 2, (cuuint64_t)x2->w * x2->ld * 2,
                                   (cuuint64_t)x2->h * x2->w * x2->ld * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)(pl.Wv * s2), 1, 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)s2, 1, 1};
    void* gaddr = (void*)((__half*)x2->ptr + x2->coff);
    CUresult r = g_encode_tiled(&mapA2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, gaddr, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      plnr_set_error("cuTensorMapEncodeTiled(shortcut) failed with CUresult %d (C=%d W=%d H=%d N=%d Wv=%d s=%d)", (int)r,
                     x2->c, x2->w, x2->h, x2->n, pl.Wv, s2);
      return PLNR_ERR_DRIVER;
    }
    small_tensor_fixup(&mapA2, (uint64_t)x2->n * x2->h * x2->w * x2->ld * 2);
  }

  if (!ctx->shift_attr_set) {
    PLNR_CHECK_CUDA(cudaFuncSetAttribute(conv_shift_f16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    PLNR_CHECK_CUDA(cudaFuncSetAttribute(conv_shift_f16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    PLNR_CHECK_CUDA((cudaFuncSetAttribute(conv_shift_f16_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)));
    PLNR_CHECK_CUDA((cudaFuncSetAttribute(conv_shift_f16_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448)));
    ctx->shift_attr_set = true;
  }
  // Persistent grid = the number of units (CTAs / CTA pairs) that can be RESIDENT AT ONCE.  For pairs that is not
  // sm_count / 2: a cluster needs both SMs in one GPC, and GPCs with an odd number of usable SMs leave one over
  // (cudaOccupancyMaxActiveClusters; cached per shared-memory size).  Stream-K depends on it -- a unit that starts only
  // after another has finished delivers its partial tiles a whole unit-time late.
  int units = ctx->sm_count / cg;
  {
    const long long key = ((long long)cg << 32) | (long long)pl.smem_bytes;
    auto it = ctx->shift_max_units.find(key);
    if (it == ctx->shift_max_units.end()) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3((unsigned)(ctx->sm_count / cg * cg));
      q.blockDim = dim3(kThreads);
      q.dynamicSmemBytes = pl.smem_bytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = (unsigned)cg; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int n = 0;
      cudaError_t qe = cg == 2 ? cudaOccupancyMaxActiveClusters(&n, conv_shift_f16_kernel<2>, &q)
                               : cudaOccupancyMaxActiveClusters(&n, conv_shift_f16_kernel<1>, &q);
      if (qe != cudaSuccess || n < 1) { cudaGetLastError(); n = units; }
      it = ctx->shift_max_units.emplace(key, n).first;
    }
    if (it->second < units) units = it->second;
    if (const char* e = getenv("PLNR_SHIFT_UNITS")) { int v = atoi(e); if (v >= 1 && v < units) units = v; }
  }
  // stream-K (SegList): worth it when the last data-parallel round is poorly filled.  Streamed weights only (a resident
  // filter is loaded during a unit's first WHOLE tile), at least two 32-channel chunks per epilogue group, and enough
  // iterations that every unit gets a non-empty range.  PLNR_STREAMK=0 off (default), 1 auto, 2 whenever legal.
  p.ipt = d->kh * d->kw * p.cchunks + p.c2chunks;
  {
    // Default OFF: measured (profiles/r02_kernel_experiments.md) the balanced schedule shortens the MMA phase of the
    // 14x14 layers by 6k clk per CTA but costs as much again -- units at different K offsets stream DIFFERENT weight boxes
    // from L2 at the same time (the data-parallel schedule has all 148 SMs on the same box), 18 % more clk per iteration.
    int mode = 0;
    if (const char* e = getenv("PLNR_STREAMK")) mode = atoi(e);
    const double waves = (double)p.num_tiles / units;
    const double dp_rounds = (double)((p.num_tiles + units - 1) / units);
    const bool legal = !p.pool && !p.b_resident && p.n_tile >= 64 && (long long)p.num_tiles * p.ipt >= 4ll * units && p.num_tiles >= 2 &&
                       !(ctx->capturing && !ctx->sk_ws);
    if (legal && mode > 0 && (mode >= 2 || dp_rounds / (waves + 0.12) >= 1.06)) {
      const size_t ws_bytes = (size_t)units * cg * kTileM * 256 * sizeof(float);
      if (!ctx->sk_ws) {
        PLNR_CHECK_CUDA(cudaMalloc(&ctx->sk_ws, ws_bytes));
        PLNR_CHECK_CUDA(cudaMalloc(&ctx->sk_flags, sizeof(int) * 16 * 256));
        PLNR_CHECK_CUDA(cudaMemsetAsync(ctx->sk_flags, 0, sizeof(int) * 16 * 256, ctx->stream));
      }
      p.streamk = 1;
      p.sk_ws = ctx->sk_ws;
      p.sk_flags = ctx->sk_flags;
    }
  }
  if (!p.streamk && p.num_tiles < units) units = p.num_tiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(units * cg));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // Programmatic dependent launch: this grid may start while the previous kernel of the stream drains; everything it
  // does before griddepcontrol.wait (barrier init, TMEM allocation, tensor-map prefetch) overlaps that kernel's tail.
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = plnr_pdl_enabled() ? 2 : 1;
  cudaError_t le;
  if (p.pool) le = cg == 2 ? cudaLaunchKernelEx(&cfg, conv_shift_f16_kernel<2, true>, mapsA, mapB, mapA2, p)
                           : cudaLaunchKernelEx(&cfg, conv_shift_f16_kernel<1, true>, mapsA, mapB, mapA2, p);
  else le = cg == 2 ? cudaLaunchKernelEx(&cfg, conv_shift_f16_kernel<2>, mapsA, mapB, mapA2, p)
                    : cudaLaunchKernelEx(&cfg, conv_shift_f16_kernel<1>, mapsA, mapB, mapA2, p);
  if (le != cudaSuccess) {
    plnr_set_error("launch of conv_shift_f16_kernel<%d> failed: %s", cg, cudaGetErrorString(le));
    return PLNR_ERR_CUDA;
  }
  return plnr_after_launch(ctx, "conv2d_shift");
}

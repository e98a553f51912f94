// fp16 implicit-GEMM convolution for sm_100a:  TMA-im2col -> shared memory -> tcgen05.mma -> TMEM -> fused epilogue.
//
// Replaces planer/layer.py:22-26 (Conv2d) + planer/util.py:17-44 (conv_for) and the batchnorm / add / relu layers
// that follow it (planer/layer.py:125-127, :93-95, :44-51).  The reference materialises the im2col matrix
// (9-12x the input) and runs one GEMM  M=Co, K=C*kh*kw, N=N*oh*ow;  here the same contraction is tiled as
//
//     D[128 output pixels, n_tile output channels] += A[128 pixels, 64 k] * B[n_tile channels, 64 k]^T
//
// with the activation operand A fetched straight from the NHWC tensor by the TMA unit in im2col mode (zero fill
// = padding, element strides = conv stride, per-load filter-tap offsets = dilation), the packed weights B by a
// tiled TMA load, both landing in swizzled shared memory that tcgen05.mma reads through matrix descriptors;
// fp32 accumulators live in TMEM (double buffered: the epilogue of tile i overlaps the main loop of tile i+1).
//
// CTA = 8 warps, persistent over output tiles (grid = #SMs):
//   warp 0  lane 0 : TMA producer            (waits empty[s], arms full[s] with the stage's byte count)
//   warp 1  lane 0 : tcgen05.mma issuer      (waits full[s], issues 64/16 MMAs, commits -> empty[s] / tmem_full[a])
//   warp 2         : TMEM allocator / deallocator
//   warps 4-7      : epilogue, one TMEM lane quarter each: tcgen05.ld -> *scale+shift (+residual) -> act -> fp16
//
// K is walked in "units" of kc input channels of one filter tap (kc = 64/32/16 -> 128B/64B/32B swizzle); a
// pipeline stage holds 64/kc units, i.e. always 64 k-values = one 128-byte row per pixel / per output channel.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 384;          // warps 0-3: producer / MMA / TMEM / table; warps 4-11: two epilogue groups
constexpr int kMaxStages = 8;
constexpr uint32_t kAStageBytes = kTileM * 64 * 2;  // 16 KB: 128 pixels x 64 k x fp16
constexpr uint32_t kEpiBytes = 2 * 2 * 256 * 4 + 2 * 2 * 256 * 2;     // [buffer][scale|shift][256] fp32, then the same in fp16
constexpr uint32_t kStageBytes = 8 * 2048;          // epilogue transposition stage (per epilogue warp: 32 rows x 64 B)
constexpr long long kWatchdogCycles = 4000000000ll; // ~2 s: a stuck barrier becomes an error, not a hang

struct IgemmParams {
  FastDiv div_mt;   // / num_m_tiles
  int M, OW, OH;
  int pad_t, pad_l, sh, sw, dh, dw;
  int S;          // filter width
  int kc;         // input channels per unit
  int cchunks;    // Cin / kc
  int units;      // kh * kw * cchunks
  int tps;        // units per stage (64 / kc)
  int kstages;    // ceil(units / tps)
  int n_tile, num_m_tiles, num_tiles;
  int stages;
  uint32_t a_layout, a_sbo, a_unit_bytes, b_stage_bytes, idesc, tmem_cols;
  __half* y; int yld, ycoff, Cout;
  const float* scale; const float* shift;
  const __half* res; int rld, rcoff;
  int act; float alpha; int res_after;
  int vec_ok;
  int out_f32;       // plnr_epilogue.out_f32: y / residual are fp32 tensors (split-fp16 operands, split_f32.cu)
  float acc_scale;   // accumulator multiplier (2^-(ex+ew) of the operand pre-scales), folded into the staged scale
  const float* acc_scale_dev;   // ... times this device scalar (per-call activation pre-scale) when not NULL
  int* err;
  long long* prof;   // optional [grid][8] cycle counters per role (debug)
  int dbg;           // debug: bit0 = epilogue skips global loads/stores, bit1 = epilogue skips tcgen05.ld too
};

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int role) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFF) == 0) {
      if (*reinterpret_cast<volatile int*>(err) != 0) return;       // somebody else already gave up
      if (clock64() - t0 > kWatchdogCycles) {
        if (atomicCAS(err, 0, 1) == 0) { err[1] = blockIdx.x; err[2] = role; err[3] = (int)parity; }
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

// CG = 1: one CTA per 128-pixel tile (tcgen05 cta_group::1).
// CG = 2: a CTA PAIR (cluster of 2, one TPC) per 256-pixel tile (cta_group::2): each CTA loads its own 128 A rows and
//         HALF of the B tile; the leader CTA issues M=256 MMAs that read both CTAs' shared memory and write both CTAs'
//         TMEM.  Per CTA this halves the B bytes written to and read from shared memory -- the resource that bounds the
//         1-CTA kernel.
template <int CG>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_f16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                      const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const long long t_entry = clock64();
  const bool stamp = p.prof && blockIdx.x == 0;      // debug: phase time stamps of CTA 0 at prof[1400..] (tools/role_profile2.py)
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;           // 128B-swizzle atoms need 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t stages = (uint32_t)p.stages;
  const uint32_t sA = base;
  const uint32_t sB = sA + stages * kAStageBytes;
  const uint32_t epi_off = stages * (kAStageBytes + p.b_stage_bytes);
  float* epi = reinterpret_cast<float*>(base_ptr + epi_off);
  const uint32_t sBar = base + epi_off + kEpiBytes;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * stages;
  const uint32_t bar_tfull = sBar + 16 * stages, bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + epi_off + kEpiBytes + 16 * stages + 32);
  uint8_t* stage = base_ptr + epi_off + kEpiBytes + 16 * stages + 64;     // 4 warps x 2 KB epilogue transposition stage
  uint32_t* utab = reinterpret_cast<uint32_t*>(stage + kStageBytes);      // per k-unit: channel | tap offsets

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;     // position inside the CTA pair
  const bool leader = rank == 0;
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // persistent work-unit id
  const int nunits = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&mapA);
    ptx::prefetch_tmap(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < stages; ++i) {
      ptx::mbar_init(bar_full + 8 * i, 1);    // leader's arrive.expect_tx (+ TMA transaction bytes of the whole pair)
      ptx::mbar_init(bar_empty + 8 * i, 1);   // tcgen05.commit (multicast to both CTAs when CG == 2)
    }
    for (uint32_t a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_tfull + 8 * a, 1);          // tcgen05.commit after the last k-stage of a tile
      ptx::mbar_init(bar_tempty + 8 * a, 256 * CG);  // every epilogue thread of the pair (leader's copy is the live one)
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if (CG == 2) { ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), p.tmem_cols); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), p.tmem_cols); ptx::tmem_relinquish(); }
  }
  if (warp == 3) {
    // k-unit table: unit u = (filter tap, channel chunk) -> first channel and the im2col tap offsets, so that the
    // producer's inner loop is a table read + one TMA instruction
    for (int u = lane; u < p.units; u += 32) {
      const int tap = u / p.cchunks, cc = u - tap * p.cchunks;
      const int r = tap / p.S, sx = tap - r * p.S;
      utab[u] = (uint32_t)(cc * p.kc) | ((uint32_t)(sx * p.dw) << 16) | ((uint32_t)(r * p.dh) << 24);
    }
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // programmatic dependent launch: let the next kernel of the stream start its own set-up, then wait until every
  // kernel before this one has completed and flushed its results (no-ops when launched without the attribute)
  ptx::grid_launch_dependents();
  ptx::grid_dependency_wait();
  if (stamp && threadIdx.x == 0) {
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.prof[1400] = (long long)gt; p.prof[1401] = clock64() - t_entry;
  }

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // The whole warp walks the loop (converged, so address arithmetic stays on the uniform datapath); one elected
    // lane issues the TMA instructions.
    uint32_t s = 0, ph = 0;
    long long t_wait = 0;
    const long long t_all0 = clock64();
    const int tps = p.tps, units = p.units;
    const uint32_t a_unit_bytes = p.a_unit_bytes, b_stage_bytes = p.b_stage_bytes;
    for (int tile = unit; tile < p.num_tiles; tile += nunits) {
      const int m_idx = tile % p.num_m_tiles, n_idx = tile / p.num_m_tiles;
      const int m0 = (m_idx * CG + (int)rank) * kTileM;
      const int q0 = m0 % p.OW;
      const int t1 = m0 / p.OW;
      const int p0 = t1 % p.OH, img = t1 / p.OH;
      const int cw = q0 * p.sw - p.pad_l, chh = p0 * p.sh - p.pad_t;   // input coordinate of the tile's first pixel
      const int n_row = n_idx * p.n_tile + (int)rank * (p.n_tile / CG);
      int u = 0;
      for (int j = 0; j < p.kstages; ++j) {
        const long long tw0 = clock64();
        mbar_wait(bar_empty + 8 * s, ph ^ 1, p.err, 0);
        t_wait += clock64() - tw0;
        const int nu = min(tps, units - u);
        if (ptx::elect_one()) {
          const uint32_t full = bar_full + 8 * s;
          if (leader) ptx::mbar_arrive_expect_tx(full, (uint32_t)CG * ((uint32_t)nu * a_unit_bytes + b_stage_bytes));
          uint32_t a_dst = sA + s * kAStageBytes;
          for (int t = 0; t < nu; ++t) {
            const uint32_t e = utab[u + t];
            if (CG == 2)
              ptx::tma_load_im2col_4d_pair(a_dst, &mapA, full, (int)(e & 0xffffu), cw, chh, img, (uint16_t)((e >> 16) & 0xffu), (uint16_t)(e >> 24));
            else
              ptx::tma_load_im2col_4d(a_dst, &mapA, full, (int)(e & 0xffffu), cw, chh, img, (uint16_t)((e >> 16) & 0xffu), (uint16_t)(e >> 24));
            a_dst += a_unit_bytes;
          }
          if (CG == 2) ptx::tma_load_2d_pair(sB + s * b_stage_bytes, &mapB, full, j * 64, n_row);
          else ptx::tma_load_2d(sB + s * b_stage_bytes, &mapB, full, j * 64, n_row);
        }
        __syncwarp();
        u += nu;
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    if (p.prof && lane == 0) { p.prof[blockIdx.x * 8 + 0] = t_wait; p.prof[blockIdx.x * 8 + 1] = clock64() - t_all0; }
  } else if (warp == 1 && leader) {
    // ===================================== MMA issuer =========================================
    // Converged warp; one elected lane issues tcgen05.mma / tcgen05.commit.  Descriptors are advanced by adding
    // to their 14-bit (address >> 4) field: +2 per K=16 step (32 bytes), a constant per stage.
    uint32_t s = 0, ph = 0, it = 0;
    long long t_full = 0, t_tempty = 0;
    const long long t_all0 = clock64();
    const int kper = p.kc >> 4, tps = p.tps, units = p.units, kstages = p.kstages;
    const uint32_t idesc = p.idesc;
    const uint64_t adesc0 = ptx::make_smem_desc(sA, p.a_sbo, p.a_layout);
    const uint64_t bdesc0 = ptx::make_smem_desc(sB, 1024, 2);
    const uint32_t a_lo0 = (uint32_t)adesc0, b_lo0 = (uint32_t)bdesc0, desc_hi = (uint32_t)(bdesc0 >> 32);
    const uint32_t a_step = kAStageBytes >> 4, b_step = p.b_stage_bytes >> 4;
    const uint32_t a_unit_skip = (p.a_unit_bytes - (uint32_t)kper * 32u) >> 4;
    const bool elected = ptx::elect_one();       // the same lane issues every MMA and every commit
    const bool trace_on = p.prof && blockIdx.x == 0 && lane == 0;
    for (int tile = unit; tile < p.num_tiles; tile += nunits, ++it) {
      const uint32_t a = it & 1, aph = (it >> 1) & 1;
      const long long te0 = clock64();
      mbar_wait(bar_tempty + 8 * a, aph ^ 1, p.err, 1);   // epilogue(s) have drained this accumulator
      t_tempty += clock64() - te0;
      const bool trace = trace_on && it < 40;
      if (trace) { p.prof[1200 + it * 4 + 0] = te0 - t_all0; p.prof[1200 + it * 4 + 1] = clock64() - t_all0; }
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * (uint32_t)p.n_tile;
      if (kper == 4) {
        // Cin % 64 == 0: one k-unit per stage = four K=16 steps issued from one asm block; the issuing thread's
        // instruction stream between two stages is a barrier poll, two adds and two commits (see ptx::umma_f16_x4)
        uint32_t acc = 0u;
        for (int j = 0; j < kstages; ++j) {
          mbar_wait(bar_full + 8 * s, ph, p.err, 2);      // TMA bytes (of both CTAs) have landed
          if (stamp && it == 0 && j == 0 && lane == 0) p.prof[1402] = clock64() - t_entry;
          ptx::tc_fence_after();
          if (elected) {
            ptx::umma_f16_x4<CG>(d_tmem, a_lo0 + s * a_step, b_lo0 + s * b_step, desc_hi, idesc, acc);
            if (CG == 2) {
              ptx::umma_commit_pair(bar_empty + 8 * s, 3);
              if (j == kstages - 1) ptx::umma_commit_pair(bar_tfull + 8 * a, 3);
            } else {
              ptx::umma_commit(bar_empty + 8 * s);
              if (j == kstages - 1) ptx::umma_commit(bar_tfull + 8 * a);
            }
          }
          acc = 1u;
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      } else {
        int u = 0;
        for (int j = 0; j < kstages; ++j) {
          const long long tf0 = clock64();
          mbar_wait(bar_full + 8 * s, ph, p.err, 2);
          t_full += clock64() - tf0;
          ptx::tc_fence_after();
          const int nu = min(tps, units - u);
          if (elected) {
            uint64_t da = adesc0 + (uint64_t)(s * a_step), db = bdesc0 + (uint64_t)(s * b_step);
            uint32_t acc = j > 0 ? 1u : 0u;
            for (int t = 0; t < nu; ++t) {
              for (int k = 0; k < kper; ++k) {
                if (CG == 2) ptx::umma_f16_pair(d_tmem, da, db, idesc, acc);
                else ptx::umma_f16(d_tmem, da, db, idesc, acc);
                acc = 1u; da += 2; db += 2;
              }
              da += a_unit_skip;
            }
            if (CG == 2) {
              ptx::umma_commit_pair(bar_empty + 8 * s, 3);          // both CTAs' slots reusable once these MMAs retire
              if (j == kstages - 1) ptx::umma_commit_pair(bar_tfull + 8 * a, 3);
            } else {
              ptx::umma_commit(bar_empty + 8 * s);
              if (j == kstages - 1) ptx::umma_commit(bar_tfull + 8 * a);   // accumulator complete -> epilogue
            }
          }
          u += nu;
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
      if (trace) p.prof[1200 + it * 4 + 3] = clock64() - t_all0;
    }
    if (p.prof && lane == 0) { p.prof[blockIdx.x * 8 + 2] = t_full; p.prof[blockIdx.x * 8 + 3] = t_tempty; p.prof[blockIdx.x * 8 + 4] = clock64() - t_all0; }
    if (stamp && lane == 0) p.prof[1403] = clock64() - t_entry;
  } else if (warp >= 4) {
    // ===================================== epilogue ===========================================
    // A thread owns one accumulator ROW (TMEM lane) and 32 channels per chunk; rows are transposed through a per-warp
    // shared-memory stage so that every global access of the warp moves 8 rows x 64 contiguous bytes: lane l serves
    // row 8i + l/4, 16-byte piece l%4 (i = 0..3).  The residual operand is read in that same coalesced pattern and
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
    const bool res16 = has_res && !p.out_f32;          // fp16 residual: read in the coalesced pattern, one chunk ahead
    const uint32_t uM = (uint32_t)p.M;
    uint8_t* st_o = stage + (warp - 4) * 2048;
    const uint32_t my_sw = (uint32_t)((lane >> 1) & 3);
    const int piece = lane & 3;

    // geometry of a tile for this thread: its own row (scalar tail path) and the four rows it serves in the coalesced
    // pattern; pixel index -1 = beyond the tensor
    struct Geo { int own; int row[4]; };
    auto tile_geo = [&](int tile_) {
      Geo g;
      const uint32_t m_idx_ = (uint32_t)tile_ - fast_div((uint32_t)tile_, p.div_mt) * (uint32_t)p.num_m_tiles;
      const uint32_t o = (m_idx_ * CG + rank) * kTileM + (uint32_t)(ew * 32 + lane);
      g.own = o < uM ? (int)o : -1;          // output pixel index == GEMM row (no padded rim in the im2col kernel)
#pragma unroll
      for (int i = 0; i < 4; ++i) g.row[i] = __shfl_sync(0xffffffffu, g.own, 8 * i + (lane >> 2));
      return g;
    };
    // residual pieces of a tile's FIRST chunk, fetched one tile ahead: by the time an epilogue warp reaches a tile its
    // accumulator is usually complete, and a load issued then would expose the DRAM latency once per tile
    uint4 rvp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) rvp[i] = make_uint4(0u, 0u, 0u, 0u);
    auto fetch_res = [&](const Geo& g, int cb, uint4 (&dst)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (g.row[i] >= 0)
          dst[i] = *reinterpret_cast<const uint4*>(p.res + (size_t)g.row[i] * p.rld + p.rcoff + cb + piece * 8);
    };
    Geo gn = tile_geo(unit < p.num_tiles ? unit : 0);
    if (res16 && p.vec_ok && c_begin < c_end && unit < p.num_tiles && (!alternate || eg == 0)) {
      const int cb = (int)fast_div((uint32_t)unit, p.div_mt) * p.n_tile + c_begin;
      if (cb + 32 <= p.Cout) fetch_res(gn, cb, rvp);
    }

    for (int tile = unit; tile < p.num_tiles; tile += nunits, ++it) {
      const uint32_t a = it & 1, tph = (it >> 1) & 1;
      const int n_idx = (int)fast_div((uint32_t)tile, p.div_mt);
      const int n0 = n_idx * p.n_tile;
      // per-channel scale/shift of this tile's channel block: re-staged only when the block changes
      if (n_idx != staged_n) {
        staged_n = n_idx;
        ebuf ^= 1u;
        float* dsc = epi + ebuf * 512, *dsf = dsc + 256;
        __half* hsc = reinterpret_cast<__half*>(epi + 1024) + ebuf * 512, *hsf = hsc + 256;
        const float accs = p.acc_scale_dev ? p.acc_scale * __ldg(p.acc_scale_dev) : p.acc_scale;
        for (int i = et; i < p.n_tile; i += 256) {
          const int c = n0 + i;
          float sc = 0.f, sf = 0.f;
          if (c < p.Cout) { sc = (p.scale ? __ldg(p.scale + c) : 1.f) * accs; sf = p.shift ? __ldg(p.shift + c) : 0.f; }
          dsc[i] = sc; dsf[i] = sf;
          hsc[i] = __float2half_rn(sc); hsf[i] = __float2half_rn(sf);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const float* ep_scale = epi + ebuf * 512, *ep_shift = ep_scale + 256;
      const __half* ep_hscale = reinterpret_cast<const __half*>(epi + 1024) + ebuf * 512, *ep_hshift = ep_hscale + 256;

      const Geo g = gn;
      uint4 rv[4], rvn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { rv[i] = rvp[i]; rvn[i] = make_uint4(0u, 0u, 0u, 0u); }
      // next tile: geometry + residual of its first chunk (in flight during this whole epilogue)
      const int tile_n = tile + nunits;
      if (tile_n < p.num_tiles) {
        gn = tile_geo(tile_n);
        if (res16 && p.vec_ok && c_begin < c_end && (!alternate || ((it + 1) & 1u) == (uint32_t)eg)) {
          const int cb = (int)fast_div((uint32_t)tile_n, p.div_mt) * p.n_tile + c_begin;
          if (cb + 32 <= p.Cout) fetch_res(gn, cb, rvp);
        }
      }

      const long long tt0 = clock64();
      mbar_wait(bar_tfull + 8 * a, tph, p.err, 3);
      t_tfull += clock64() - tt0;
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + a * (uint32_t)p.n_tile;
      const int ce_ = (alternate && (it & 1u) != (uint32_t)eg) ? c_begin : c_end;      // not this group's tile
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
        const bool fast = !p.out_f32 && p.vec_ok && (cb + 32 <= p.Cout);
        if (res16 && p.vec_ok && c0 + 32 < ce_ && cb + 64 <= p.Cout) fetch_res(g, cb + 32, rvn);   // one chunk ahead
        ptx::tmem_ld_wait();
        if (c0 + 32 >= ce_) {
          ptx::tc_fence_before();
          if (CG == 2 && !leader) ptx::mbar_arrive_remote(bar_tempty + 8 * a, 0);
          else ptx::mbar_arrive(bar_tempty + 8 * a);
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
          auto chunk_math = [&](auto act_tag) {
            constexpr int kAct = decltype(act_tag)::value;      // 1 = ReLU, 2 = LeakyReLU (0 <= alpha <= 1), 0 = generic
            const bool act_first = !has_res || p.res_after;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 pk;
              pk.x = pack_half2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
              pk.y = pack_half2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
              pk.z = pack_half2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
              pk.w = pack_half2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
              *reinterpret_cast<uint4*>(st_o + lane * 64 + (((uint32_t)q ^ my_sw) << 4)) = pk;
            }
            const uint4 sc4 = *reinterpret_cast<const uint4*>(ep_hscale + c0 + piece * 8);
            const uint4 sf4 = *reinterpret_cast<const uint4*>(ep_hshift + c0 + piece * 8);
            const __half2* sch = reinterpret_cast<const __half2*>(&sc4);
            const __half2* sfh = reinterpret_cast<const __half2*>(&sf4);
            const __half2 zero2 = __float2half2_rn(0.f), alpha2 = __float2half2_rn(p.alpha);
            auto act2 = [&](__half2 x) {
              if (kAct == 1) return __hmax2(x, zero2);
              if (kAct == 2) return __hmax2(x, __hmul2(x, alpha2));
              const float2 f = __half22float2(x);
              return __floats2half2_rn(plnr_apply_act(f.x, p.act, p.alpha), plnr_apply_act(f.y, p.act, p.alpha));
            };
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = 8 * i + (lane >> 2);
              uint4 val = *reinterpret_cast<const uint4*>(st_o + row * 64 + ((piece ^ ((row >> 1) & 3)) << 4));
              if (g.row[i] >= 0) {
                __half2* vh = reinterpret_cast<__half2*>(&val);
                const __half2* rh = reinterpret_cast<const __half2*>(&rv[i]);
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
                *reinterpret_cast<uint4*>(p.y + (size_t)g.row[i] * p.yld + p.ycoff + cb + piece * 8) = val;
              }
            }
          };
          if (p.act == PLNR_ACT_RELU) chunk_math(ActTag<1>{});
          else if (p.act == PLNR_ACT_LEAKY && p.alpha >= 0.f && p.alpha <= 1.f) chunk_math(ActTag<2>{});
          else chunk_math(ActTag<0>{});
#pragma unroll
          for (int i = 0; i < 4; ++i) rv[i] = rvn[i];
        } else if (p.out_f32) {
          // fp32 result of the split-fp16 path: every thread finishes its own row (32 channels = 128 contiguous bytes),
          // scale / shift / residual / activation in fp32 like the reference's float32 layers
          if (g.own >= 0) {
            float* yrow = reinterpret_cast<float*>(p.y) + (size_t)g.own * p.yld + p.ycoff;
            const float* rrow = has_res ? reinterpret_cast<const float*>(p.res) + (size_t)g.own * p.rld + p.rcoff : nullptr;
            if (p.vec_ok && cb + 32 <= p.Cout) {
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                const float4 sc = *reinterpret_cast<const float4*>(ep_scale + c0 + e);
                const float4 sf = *reinterpret_cast<const float4*>(ep_shift + c0 + e);
                float4 rf = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rrow) rf = *reinterpret_cast<const float4*>(rrow + cb + e);
                float o[4] = {fmaf(__uint_as_float(v[e]), sc.x, sf.x), fmaf(__uint_as_float(v[e + 1]), sc.y, sf.y),
                              fmaf(__uint_as_float(v[e + 2]), sc.z, sf.z), fmaf(__uint_as_float(v[e + 3]), sc.w, sf.w)};
                const float r4[4] = {rf.x, rf.y, rf.z, rf.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (!p.res_after) o[j] += r4[j];
                  o[j] = plnr_apply_act(o[j], p.act, p.alpha);
                  if (p.res_after) o[j] += r4[j];
                }
                *reinterpret_cast<float4*>(yrow + cb + e) = make_float4(o[0], o[1], o[2], o[3]);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const int c = cb + e;
                if (c < p.Cout) {
                  float o1 = fmaf(__uint_as_float(v[e]), ep_scale[c0 + e], ep_shift[c0 + e]);
                  const float rf = rrow ? rrow[c] : 0.f;
                  if (!p.res_after) o1 += rf;
                  o1 = plnr_apply_act(o1, p.act, p.alpha);
                  if (p.res_after) o1 += rf;
                  yrow[c] = o1;
                }
              }
            }
          }
        } else if (g.own >= 0) {
          __half* yrow = p.y + (size_t)g.own * p.yld + p.ycoff;
          const __half* rrow = has_res ? p.res + (size_t)g.own * p.rld + p.rcoff : nullptr;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int c = cb + e;
            if (c < p.Cout) {
              float o1 = fmaf(__uint_as_float(v[e]), ep_scale[c0 + e], ep_shift[c0 + e]);
              const float rf = rrow ? __half2float(rrow[c]) : 0.f;
              if (!p.res_after) o1 += rf;
              o1 = plnr_apply_act(o1, p.act, p.alpha);
              if (p.res_after) o1 += rf;
              yrow[c] = __float2half_rn(o1);
            }
          }
        }
      }
    }
    if (p.prof && threadIdx.x == 128) { p.prof[blockIdx.x * 8 + 5] = t_tfull; p.prof[blockIdx.x * 8 + 6] = clock64() - t_all0; }
    if (stamp && threadIdx.x == 128) p.prof[1404] = clock64() - t_entry;
  }

  // ---- teardown: everyone (in both CTAs) is done with TMEM before the allocator warps free it ----
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (stamp && threadIdx.x == 0) p.prof[1405] = clock64() - t_entry;
  if (warp == 2) {
    ptx::tc_fence_after();
    if (CG == 2) ptx::tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else ptx::tmem_dealloc(tmem_base, p.tmem_cols);
    if (stamp && lane == 0) {
      unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      p.prof[1406] = clock64() - t_entry; p.prof[1407] = (long long)gt;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side: TMA descriptors (driver entry points resolved at run time; no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static int resolve_driver() {
  if (g_encode_tiled && g_encode_im2col) return PLNR_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    plnr_set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return PLNR_ERR_DRIVER;
  }
  g_encode_tiled = (EncodeTiledFn)fn;
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    plnr_set_error("cuTensorMapEncodeIm2col not available from the driver (%s)", cudaGetErrorString(e));
    return PLNR_ERR_DRIVER;
  }
  g_encode_im2col = (EncodeIm2colFn)fn;
  cudaDriverGetVersion(&g_driver_version);
  return PLNR_OK;
}

// Drivers up to CUDA 13.1 mis-encode one descriptor bit for tensors smaller than 128 KiB; CUTLASS clears it
// (cute/atom/copy_traits_sm90_tma.hpp, copy_traits_sm90_im2col.hpp) and so do we.
static void small_tensor_fixup(CUtensorMap* m, uint64_t tensor_bytes) {
  if (g_driver_version <= 13010 && tensor_bytes < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

static int pick_kc(int cin) { return cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16); }

// Estimated main-loop cost of a tile column width: MMA cycles per 64-k stage (N/2 per K=16 MMA at full rate;
// narrow tiles are shared-memory-read bound: ~48 cycles per MMA at N=64) times the number of waves.
static int pick_n_tile(int cout, int num_m_tiles, int sm_count) {
  const int cap = round_up(cout, 32) < 256 ? round_up(cout, 32) : 256;
  int best = cap;
  double best_cost = 1e30;
  for (int n = cap; n >= 32; n -= 32) {
    if (n != cap && n != 128 && n != 64) continue;
    const int n_tiles = (cout + n - 1) / n;
    const long long tiles = (long long)num_m_tiles * n_tiles;
    const long long waves = (tiles + sm_count - 1) / sm_count;
    // measured cycles per 64-k stage of this kernel (per-tile trace, round 1): the im2col A load alone is ~410
    const double per_stage = n > 192 ? 650.0 : (n > 128 ? 630.0 : (n > 64 ? 600.0 : 560.0));
    const double cost = (double)waves * per_stage;
    if (cost < best_cost * 0.97) { best_cost = cost; best = n; }
  }
  return best;
}

}  // namespace

bool plnr_conv2d_tcgen05_supported(const plnr_conv_desc* d, const plnr_tensor* x, const plnr_tensor* y) {
  if (d->dtype != PLNR_F16 || d->groups != 1) return false;
  if (x->c % 16 != 0 || x->ld % 8 != 0 || x->coff % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(x->ptr) & 15) != 0) return false;
  if (d->stride_h > 8 || d->stride_w > 8) return false;                       // TMA traversal stride limit
  if ((d->kh - 1) * d->dil_h > 255 || (d->kw - 1) * d->dil_w > 255) return false;   // im2col offsets are 8-bit
  const int up_w = d->pad_r - (d->kw - 1) * d->dil_w, up_h = d->pad_b - (d->kh - 1) * d->dil_h;
  if (d->pad_t > 128 || d->pad_l > 128 || up_w < -128 || up_h < -128 || up_w > 127 || up_h > 127) return false;
  if ((int64_t)y->n * y->h * y->w >= (1ll << 31) - 256) return false;
  return true;
}

int plnr_conv2d_tcgen05(plnr_ctx* ctx, const plnr_conv_desc* d, const plnr_tensor* x, const void* w,
                        const plnr_tensor* y, const plnr_epilogue* ep) {
  // stride-1 convolutions with Cin % 64 == 0 take the shift-GEMM kernel (conv_shift.cu): each input row is loaded once
  // (fp16 results only: the fp32 epilogue of the split-fp16 path lives in this kernel)
  // 1x1 / stride-1 convolutions with small filters: the lean resident-weight GEMM kernel (conv_pw.cu)
  if (plnr_conv2d_pw_supported(d, x, y, ep)) return plnr_conv2d_pw(ctx, d, x, w, y, ep);
  const bool out_f32 = ep && ep->out_f32;
  if (!out_f32 && plnr_conv2d_shift_supported(d, x, y)) return plnr_conv2d_shift(ctx, d, x, w, y, ep);
  int rc = resolve_driver();
  if (rc != PLNR_OK) return rc;
  PLNR_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv2d(tcgen05): packed weights must be 16-byte aligned");

  IgemmParams p;
  memset(&p, 0, sizeof(p));
  const int Cin = x->c, Cout = y->c;
  p.M = y->n * y->h * y->w; p.OW = y->w; p.OH = y->h;
  p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.sh = d->stride_h; p.sw = d->stride_w; p.dh = d->dil_h; p.dw = d->dil_w;
  p.S = d->kw;
  p.kc = pick_kc(Cin);
  p.cchunks = Cin / p.kc;
  p.units = d->kh * d->kw * p.cchunks;
  p.tps = 64 / p.kc;
  p.kstages = (p.units + p.tps - 1) / p.tps;
  const int m_tiles_128 = (p.M + kTileM - 1) / kTileM;
  // CTA pairs (cta_group::2, 256-pixel tiles) whenever there are at least two pixel tiles; PLNR_CTA_GROUP=1 forces
  // the single-CTA kernel (A/B comparisons, debugging)
  int cg = 1;   // measured: the im2col main loop is bound by the im2col TMA rate (46 B/clk), which a pair does not relieve
  if (const char* e = getenv("PLNR_CTA_GROUP")) { if (atoi(e) == 2 && m_tiles_128 >= 2) cg = 2; }
  p.num_m_tiles = (p.M + kTileM * cg - 1) / (kTileM * cg);
  p.n_tile = pick_n_tile(Cout, p.num_m_tiles, ctx->sm_count / cg);
  const int num_n_tiles = (Cout + p.n_tile - 1) / p.n_tile;
  p.num_tiles = p.num_m_tiles * num_n_tiles;
  p.div_mt = make_fastdiv((uint32_t)p.num_m_tiles);
  p.a_unit_bytes = (uint32_t)kTileM * p.kc * 2;
  p.b_stage_bytes = (uint32_t)(p.n_tile / cg) * 128;     // per CTA: a pair splits the B tile
  p.a_layout = p.kc == 64 ? 2u : (p.kc == 32 ? 4u : 6u);
  p.a_sbo = 8u * p.kc * 2;
  // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 @4, a/b format F16 (0) @7/@10,
  // both K-major (0) @15/@16, N>>3 @17, M>>4 @24
  p.idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)((kTileM * cg) >> 4) << 24);
  uint32_t cols = 32;
  while (cols < 2u * p.n_tile) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t stage_bytes = kAStageBytes + p.b_stage_bytes;
  PLNR_REQUIRE(p.units <= 4096, "conv2d(tcgen05): too many k-units (%d)", p.units);
  const uint32_t utab_bytes = (uint32_t)round_up(p.units * 4, 16);
  const uint32_t budget = 232448u - 1024u - kEpiBytes - 256u - kStageBytes - utab_bytes;
  int stages = (int)(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  PLNR_REQUIRE(stages >= 2, "conv2d(tcgen05): tile does not fit shared memory");
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * stage_bytes + kEpiBytes + 16 * stages + 64 + kStageBytes + utab_bytes + 1024;

  p.y = (__half*)y->ptr; p.yld = y->ld; p.ycoff = y->coff; p.Cout = Cout;
  p.out_f32 = out_f32 ? 1 : 0;
  p.acc_scale = (ep && ep->acc_scale != 0.f) ? ep->acc_scale : 1.f;
  p.acc_scale_dev = ep ? ep->acc_scale_dev : nullptr;
  const int va = out_f32 ? 4 : 8;          // elements per 16-byte vector of y / residual
  bool vec = (y->ld % va == 0) && (y->coff % va == 0) && ((reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0);
  if (ep) {
    p.scale = ep->scale; p.shift = ep->shift; p.act = ep->act; p.alpha = ep->alpha;
    p.res_after = ep->res_after_act;
    if (ep->residual) {
      const plnr_tensor* r = ep->residual;
      p.res = (const __half*)r->ptr; p.rld = r->ld; p.rcoff = r->coff;
      vec = vec && (r->ld % va == 0) && (r->coff % va == 0) && ((reinterpret_cast<uintptr_t>(r->ptr) & 15) == 0);
    }
  }
  p.vec_ok = vec ? 1 : 0;
  p.err = ctx->dev_error;
  p.prof = ctx->prof;
  if (const char* e = getenv("PLNR_DEBUG_EPI")) p.dbg = atoi(e);

  // ---- activation map: im2col over (C, W, H, N) ----
  CUtensorMap mapA, mapB;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)x->w, (cuuint64_t)x->h, (cuuint64_t)x->n};
    const cuuint64_t strides[3] = {(cuuint64_t)x->ld * 2, (cuuint64_t)x->w * x->ld * 2,
                                   (cuuint64_t)x->h * x->w * x->ld * 2};
    const int lower[2] = {-d->pad_l, -d->pad_t};
    const int upper[2] = {d->pad_r - (d->kw - 1) * d->dil_w, d->pad_b - (d->kh - 1) * d->dil_h};
    const cuuint32_t estr[4] = {1, (cuuint32_t)d->stride_w, (cuuint32_t)d->stride_h, 1};
    const CUtensorMapSwizzle sw = p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : (p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    void* gaddr = (void*)((__half*)x->ptr + x->coff);
    CUresult r = g_encode_im2col(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, gaddr, dims, strides, lower, upper,
                                 (cuuint32_t)p.kc, (cuuint32_t)kTileM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      plnr_set_error("cuTensorMapEncodeIm2col failed with CUresult %d (C=%d W=%d H=%d N=%d ld=%d kc=%d)", (int)r, Cin,
                     x->w, x->h, x->n, x->ld, p.kc);
      return PLNR_ERR_DRIVER;
    }
    small_tensor_fixup(&mapA, (uint64_t)x->n * x->h * x->w * x->ld * 2);
  }
  // ---- weight map: tiled over (K, Cout), box 64 x n_tile, 128B swizzle ----
  {
    const int Ktot = d->kh * d->kw * Cin;
    const cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)(p.n_tile / cg)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      plnr_set_error("cuTensorMapEncodeTiled failed with CUresult %d (K=%d Cout=%d n_tile=%d)", (int)r, Ktot, Cout,
                     p.n_tile);
      return PLNR_ERR_DRIVER;
    }
    small_tensor_fixup(&mapB, (uint64_t)Ktot * Cout * 2);
  }

  if (!ctx->igemm_attr_set) {
    PLNR_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_f16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    PLNR_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_f16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    ctx->igemm_attr_set = true;
  }
  int units = ctx->sm_count / cg;                          // persistent: one CTA (pair) per SM (pair)
  if (const char* cap = getenv("PLNR_MAX_CTAS")) {         // experiment knob: restrict the persistent grid
    int c = atoi(cap) / cg;
    if (c > 0 && c < units) units = c;
  }
  if (p.num_tiles < units) units = p.num_tiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(units * cg));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
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
  cudaError_t le = cg == 2 ? cudaLaunchKernelEx(&cfg, conv_igemm_f16_kernel<2>, mapA, mapB, p)
                           : cudaLaunchKernelEx(&cfg, conv_igemm_f16_kernel<1>, mapA, mapB, p);
  if (le != cudaSuccess) {
    plnr_set_error("launch of conv_igemm_f16_kernel<%d> failed: %s", cg, cudaGetErrorString(le));
    return PLNR_ERR_CUDA;
  }
  return plnr_after_launch(ctx, "conv2d_tcgen05");
}

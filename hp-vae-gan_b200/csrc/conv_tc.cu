// tcgen05 implicit-GEMM convolution (forward and data-gradient form) for the wide layers: Cin, Cout multiples of 64,
// NDHWC bf16 activations, fp32 accumulation in TMEM.
//
// Replaces aten::convolution / convolution_backward(grad_input) behind nn.Conv3d/Conv2d of ConvBlock3D/3DSN
// (modules/networks_3d.py:48-70) for the 64->64, 64->128 and 128->64 layers.
//
// Design (B200-first, see DESIGN.md "conv_tc"):
//  * A work unit is an output brick of NACC consecutive d-slices x 16 rows x 8 columns (NACC x 128 voxels).  The input
//    halo slab of every needed d-slice ((16+2) x (8+2) voxels x 64 channels = 23 KB) is fetched ONCE by one 5-D TMA
//    box (zero fill supplies the padding) and stays in shared memory for all 27 taps.
//  * Because tcgen05 applies the 128-byte swizzle to absolute shared-memory addresses (verified by
//    experiments/umma_desc_probe.cu), the A operand of tap (kd,kh,kw) is the SAME slab read through a descriptor whose
//    start address is moved by (kh*10+kw) rows and whose 8-row-group stride (SBO) is one slab row (10 voxels = 1280 B):
//    no im2col copy, no per-tap reload — input traffic from L2 is ~1.4x the tensor instead of 27x.
//  * Each 8 KB weight tap streams through a small TMA ring and is applied to all NACC accumulators (NACC x 64 TMEM
//    columns), so weights are read once per unit.
//  * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner, warps 2..5 = epilogue
//    (TMEM -> registers -> bias / LeakyReLU / LeakyReLU'-mask -> bf16 -> swizzled smem -> TMA store with hardware
//    clipping at the tensor edge, plus per-channel sum / sum-of-squares for BatchNorm).
#include "common.cuh"
#include <cstdlib>
#include <type_traits>

#ifndef HPVG_ROLL_STACK
#define HPVG_ROLL_STACK 1      // rolled (kh, kw) issue loops: see the MMA issuer
#endif
// The thin-output instantiations (NOUT = 16) keep the fully unrolled issue path: with the rolled loop the batched generation leg
// (two recorded forwards in flight on two streams, conv_col running next to the thin kernel) hit "illegal memory access" once in
// ~10 rounds of experiments/gen_stress.py, every time; with the thin kernels unrolled (this switch) 4 x 40 rounds on 2 and 3
// streams were clean, as were the fully unrolled builds.  The cause was not found (no sanitizer on this pool); the wide kernels
// are the ones that matter for time (110 of the ~125 tcgen05 convolution launches of an iteration).
#ifndef HPVG_ROLL_WIDE_ONLY
#define HPVG_ROLL_WIDE_ONLY 1
#endif
#ifndef HPVG_ROLL_PLAIN
#define HPVG_ROLL_PLAIN 1
#endif

namespace hpvg {

constexpr int BH = 16, BW = 8;                      // output brick rows x columns (M = 128)
constexpr int SLAB_H = BH + 2, SLAB_W = BW + 2;     // 18 x 10 halo
constexpr int SLAB_ROWS = SLAB_H * SLAB_W;          // 180 voxel rows of 128 B
constexpr int SLAB_BYTES = SLAB_ROWS * 128;         // 23040
constexpr int SLAB_STRIDE = SLAB_BYTES;             // 128-byte aligned is enough: TMA and tcgen05 swizzle on absolute address bits
constexpr int STG_BYTES = 128 * 128;                // one output tile: 128 voxels x 64 channels bf16
constexpr int MAX_COUT = 256;
constexpr int NEPI_WIDE = 8;                        // epilogue warps, wide (bf16 NDHWC) output: 2 per TMEM lane quadrant
constexpr int NEPI_THIN = 4;                        // epilogue warps, thin (fp32 NCDHW, Cout <= 16) output
constexpr int NMMA = 1;                             // MMA-issuing warps: one elected thread each, accumulators a % NMMA == index

// KCHUNKS: input channels / 64.  NACC: output d-slices (accumulators) per unit.  NGRP: accumulator groups per unit;
// the 27 weight taps are streamed once per group, so the epilogue of group i overlaps the MMAs of group i+1.
// NOUT: MMA N = output channels per accumulator: 64 (wide output) or 16 (thin output, Cout <= 16 zero-padded).
// STACK (3-D, Cin = 64 only): one weight-ring stage holds the three kd tiles of a (kh, kw) position, and one MMA with
// N = 3 * NOUT multiplies a slab view with all three of them at once — the accumulators are laid out in TMEM in
// descending slice order, so the three results land in the three accumulators that input slice feeds.  Reading the A
// operand (128 x 16 bf16 = 4 KB at 128 B/clk) is what bounds an N = 64 MMA; stacking amortises it over 3x the work.
template <int KCHUNKS, int NACC, int NGRP, int NOUT, bool STACK = false>
struct TcCfg {
  static constexpr bool THIN = NOUT < 64;
  static constexpr int NSLOTS = NACC + 2;                       // d-slices resident per unit (3-D, pad 1)
  static constexpr int NSLAB = NSLOTS * KCHUNKS;
  static constexpr int BTILE_BYTES = NOUT * 128;                // one tap: NOUT output channels x 64 input channels bf16
  static constexpr int STAGE_TILES = STACK ? 3 : 1;
  static constexpr int STAGE_BYTES = STAGE_TILES * BTILE_BYTES;
  static constexpr int NB = STACK ? (THIN ? 4 : (NGRP == 2 ? 3 : 2)) : (THIN ? 8 : ((KCHUNKS == 1) ? (NACC == 1 ? 3 : 6) : 3));   // weight ring depth (stages)
  static constexpr int NSTG = THIN ? 0 : ((KCHUNKS == 1 && NACC > 1 && !(STACK && NGRP == 2)) ? 2 : 1); // output staging buffers
  // NACC == 1 is the "two CTAs per SM" configuration for volumes with fewer 4-slice units than SMs: half the shared memory,
  // 4 epilogue warps (register budget), and the two co-resident CTAs overlap each other's prologue / MMA / epilogue phases
  static constexpr int NEPI = THIN ? NEPI_THIN : (NACC == 1 ? 4 : NEPI_WIDE);
  static constexpr int MIN_CTAS = (NACC == 1 && !THIN) ? 2 : 1;
  static constexpr int THREADS = 32 * (1 + NMMA) + 32 * NEPI;
  static constexpr int EPI_WARP0 = 1 + NMMA;
  static constexpr int ACC_COLS = NACC * NOUT;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512)));
  static constexpr int OFF_SLAB = 0;
  static constexpr int OFF_B = OFF_SLAB + NSLAB * SLAB_STRIDE;
  static constexpr int OFF_STG = OFF_B + NB * STAGE_BYTES;
  static constexpr int OFF_BIAS = OFF_STG + NSTG * STG_BYTES;          // float[MAX_COUT]
  static constexpr int OFF_BAR = OFF_BIAS + MAX_COUT * 4;
  static constexpr int NBARS = NSLAB + 2 * NB + 3;
  static constexpr int SMEM_BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;   // + tmem slot + alignment slack
  static_assert(NACC % NGRP == 0, "groups must divide the accumulators");
  static_assert(!STACK || KCHUNKS == 1, "kd stacking is implemented for Cin = 64");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct TcParams {
  ConvGeom g;
  int units_d, units_h, units_w, nblocks;   // unit grid: d-groups, h-tiles, w-tiles, 64-wide output-channel blocks
  long long num_units;
  int act;
  float slope;
  const float* bias;
  float* stats;
  const __nv_bfloat16* mask_src;
  float* y_thin;    // NOUT == 16: float32 NCDHW output
  const uint8_t* w_img;      // packed weight image (for the L2 prefetch; nullptr = off)
  unsigned w_img_bytes;
  long long* dbg;   // optional per-CTA phase clocks (development aid, hpvg_debug_set_clock_buffer)
  // FUSE (conv + BatchNorm(train) + LeakyReLU in one launch, see the fused epilogue)
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* nbt;
  float momentum, eps;
  long long count;            // voxels per channel: N * Do * Ho * Wo
  float* scale_shift;         // [2*Cout] saved for the backward pass
  float* mean_invstd;         // [2*Cout]
  unsigned* grid_counter;     // zero-initialised arrival counter of the grid barrier
  uint32_t* mask_bits;        // [voxel][Cout/32] LeakyReLU sign bits taken from the fp32 value (nullptr: not wanted)
  int store_y;                // also store the conv output y (bf16) for the BatchNorm backward
};

template <int KCHUNKS, int NACC, int KDT, int NGRP, int NOUT, bool STACK, bool FUSE>
__global__ void __launch_bounds__((TcCfg<KCHUNKS, NACC, NGRP, NOUT, STACK>::THREADS), (TcCfg<KCHUNKS, NACC, NGRP, NOUT, STACK>::MIN_CTAS))
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_o, const TcParams p) {
  using Cfg = TcCfg<KCHUNKS, NACC, NGRP, NOUT, STACK>;
  static_assert(!STACK || KDT == 3, "kd stacking needs a 3-deep kernel");
  static_assert(!FUSE || (STACK && NOUT == 64 && NGRP == 1 && KCHUNKS == 1), "the fused BatchNorm epilogue is built on the stacked 64 -> 64 form");
  static_assert(!FUSE || NACC * 2 * STG_BYTES <= Cfg::NSLAB * SLAB_STRIDE, "fused epilogue stages its tiles in the dead slab area");
  // accumulator a lives at TMEM columns (NACC - 1 - a) * NOUT: descending slice order (see STACK)
  auto acc_col = [](int a) { return (uint32_t)((NACC - 1 - a) * NOUT); };
  constexpr int GACC = NACC / NGRP;            // accumulators per group
  constexpr int NEPI_THREADS = 32 * Cfg::NEPI;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);

  const uint32_t s_slab = sbase + Cfg::OFF_SLAB;
  const uint32_t s_b = sbase + Cfg::OFF_B;
  const uint32_t s_stg = sbase + Cfg::OFF_STG;
  const uint32_t s_bar = sbase + Cfg::OFF_BAR;
  // barrier map
  auto bar_slab_full = [&](int i) { return s_bar + 8u * i; };
  auto bar_b_full = [&](int i) { return s_bar + 8u * (Cfg::NSLAB + i); };
  auto bar_b_empty = [&](int i) { return s_bar + 8u * (Cfg::NSLAB + Cfg::NB + i); };
  const uint32_t bar_acc_full = s_bar + 8u * (Cfg::NSLAB + 2 * Cfg::NB);
  const uint32_t bar_acc_empty = bar_acc_full + 8u;
  const uint32_t bar_slabs_free = bar_acc_full + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + Cfg::OFF_BAR + Cfg::NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;
  unsigned long long gt0 = 0;
  if (p.dbg && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));

  // unit -> coordinates
  // 32-bit arithmetic on purpose (the host checks num_units < 2^31): 64-bit division is a subroutine of dozens of instructions, and
  // the producer decodes its unit before it can issue the first load
  auto decode = [&](long long u64, int& nb, int& n, int& d0, int& h0, int& w0) {
    unsigned u = (unsigned)u64;
    unsigned q = u / (unsigned)p.units_w;
    w0 = (int)(u - q * (unsigned)p.units_w) * BW;
    u = q;
    q = u / (unsigned)p.units_h;
    h0 = (int)(u - q * (unsigned)p.units_h) * BH;
    u = q;
    q = u / (unsigned)p.units_d;
    d0 = (int)(u - q * (unsigned)p.units_d) * NACC;
    u = q;
    q = u / (unsigned)g.N;
    n = (int)(u - q * (unsigned)g.N);
    nb = (int)q;
  };

  // the loads a unit needs first: (kd-stacked form) the weight stage of position 0 into ring stage `wstage`, then the input slabs
  auto issue_front = [&](long long u, uint32_t wstage) {
    int nb, n, d0, h0, w0;
    decode(u, nb, n, d0, h0, w0);
    if (STACK) {
      // the first weight stage goes out BEFORE the slabs: it is small and L2-resident, and the first MMAs need it
      // together with slab 0 — behind 138 KB of slab traffic it arrived ~2 us late
      mbar_expect_tx(bar_b_full(wstage), Cfg::STAGE_BYTES);
#pragma unroll
      for (int kd = 0; kd < 3; ++kd)
        tma_load_2d(s_b + wstage * Cfg::STAGE_BYTES + kd * Cfg::BTILE_BYTES, &tmap_w, bar_b_full(wstage), 0, (kd * 9 + 0) * p.nblocks * NOUT + nb * NOUT);
    }
#pragma unroll
    for (int j = 0; j < NACC + KDT - 1; ++j) {
      const int d = d0 + j - g.pad_d;
      if (d < 0 || d >= g.Di) continue;
      // slice needed only if some valid accumulator reads it
      bool needed = false;
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        const int kd = j - a;
        if (kd >= 0 && kd < KDT && d0 + a < g.Do) needed = true;
      }
      if (!needed) continue;
#pragma unroll
      for (int kc = 0; kc < KCHUNKS; ++kc) {
        const int si = j * KCHUNKS + kc;
        mbar_expect_tx(bar_slab_full(si), SLAB_BYTES);
        tma_load_5d(s_slab + si * SLAB_STRIDE, &tmap_x, bar_slab_full(si), kc * 64, w0 - g.pad, h0 - g.pad, d, n);
      }
    }
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::NSLAB; ++i) mbar_init(bar_slab_full(i), 1);
    for (int i = 0; i < Cfg::NB; ++i) {
      mbar_init(bar_b_full(i), 1);
      mbar_init(bar_b_empty(i), NMMA);
    }
    mbar_init(bar_acc_full, NMMA);
    mbar_init(bar_acc_empty, NEPI_THREADS);
    mbar_init(bar_slabs_free, NMMA);
    mbar_fence_init();
    // (Tried: this thread issuing the first unit's loads right here, before the CTA-wide barrier — ~0.3 us earlier.  Parity tests were
    // green, but the recorded multi-stream iteration then died with "illegal memory access" on every run, with or without a
    // fence.proxy.async after the barrier initialisation; the cause was not found, the loads stay behind the barrier.)
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    if (!Cfg::THIN) tma_prefetch_desc(&tmap_y);
    if (FUSE) tma_prefetch_desc(&tmap_o);
  }
  if (threadIdx.x == 64 && p.w_img) {
    // Pull the whole packed weight image into L2, one slice per CTA.  The image of a frozen pyramid stage was last read an
    // iteration ago, so every ring stage would otherwise pay a DRAM round trip (about one stage time: the two-stage ring
    // cannot hide it), and all CTAs stream the same tiles in the same order.
    const unsigned chunk = ((p.w_img_bytes + gridDim.x - 1) / gridDim.x + 127u) & ~127u;
    const unsigned off = blockIdx.x * chunk;
    if (off < p.w_img_bytes) {
      const unsigned n = min(chunk, p.w_img_bytes - off);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.w_img + off), "r"(n) : "memory");
    }
  }
  // Programmatic dependent launch (when enabled): the barrier set-up above touches nothing the previous kernel of the stream
  // produces; from here on this CTA reads its output.
  pdl_trigger();
  pdl_wait();
  __syncthreads();                       // barriers initialised and visible to every warp
  // The producer warp goes straight to its TMA loads: it needs neither the TMEM address nor the bias.  TMEM allocation and the
  // bias fetch (a global round trip) of the other warps then overlap the first slabs' flight instead of preceding it
  // (measured with %globaltimer: ~3 us of a 17 us CTA lifetime were spent before the first load was issued).
  float* bias_s = reinterpret_cast<float*>(sgen + Cfg::OFF_BIAS);
  uint32_t tmem_base = 0;
  if (warp != 0) {
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
    for (int i = threadIdx.x - 32; i < MAX_COUT; i += Cfg::THREADS - 32) bias_s[i] = (p.bias && i < g.Cout) ? p.bias[i] : 0.f;
    tc_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(Cfg::THREADS - 32) : "memory");
    tc_fence_after();
    tmem_base = *tmem_slot;
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t bstage = 0, bphase = 0;
      int it = 0;
      for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
        int nb, n, d0, h0, w0;
        decode(u, nb, n, d0, h0, w0);
        const int q_first = STACK ? 1 : 0;
        if (it > 0) mbar_wait(bar_slabs_free, (uint32_t)((it - 1) & 1));
        if (STACK) mbar_wait(bar_b_empty(bstage), bphase ^ 1u);
        issue_front(u, bstage);
        if (STACK && ++bstage == Cfg::NB) { bstage = 0; bphase ^= 1u; }
        if (STACK) {
#pragma unroll 1
          for (int q = q_first; q < 9 * NGRP; ++q) {       // one stage = the kd = 0, 1, 2 tiles of position q % 9 = kh * 3 + kw
            mbar_wait(bar_b_empty(bstage), bphase ^ 1u);
            mbar_expect_tx(bar_b_full(bstage), Cfg::STAGE_BYTES);
#pragma unroll
            for (int kd = 0; kd < 3; ++kd)
              tma_load_2d(s_b + bstage * Cfg::STAGE_BYTES + kd * Cfg::BTILE_BYTES, &tmap_w, bar_b_full(bstage), 0,
                          (kd * 9 + q % 9) * p.nblocks * NOUT + nb * NOUT);
            if (++bstage == Cfg::NB) { bstage = 0; bphase ^= 1u; }
          }
        } else {
#pragma unroll 1
          for (int grp = 0; grp < NGRP; ++grp) {
#pragma unroll 1
            for (int t = 0; t < KDT * 9; ++t) {
#pragma unroll
              for (int kc = 0; kc < KCHUNKS; ++kc) {
                mbar_wait(bar_b_empty(bstage), bphase ^ 1u);
                mbar_expect_tx(bar_b_full(bstage), Cfg::BTILE_BYTES);
                tma_load_2d(s_b + bstage * Cfg::BTILE_BYTES, &tmap_w, bar_b_full(bstage), kc * 64, t * p.nblocks * NOUT + nb * NOUT);
                if (++bstage == Cfg::NB) { bstage = 0; bphase ^= 1u; }
              }
            }
          }
        }
      }
    }
  } else if (warp <= NMMA) {
    // ===================== MMA issuers: ONE elected thread per warp runs the whole loop =====================
    // Measured on B200 (experiments/bench_kernels.py clk): with both operands in shared memory the loop is bound by
    // the operand reads, (128 x 16 A + N x 16 B) x 2 B at 128 B/clk = 48 clk per MMA at N = 64 and 36 clk at N = 16,
    // not by issue — a second issuing warp (NMMA = 2, accumulators a % NMMA) or removing the weight waits changes
    // nothing.  That caps an N = 64 layer at 32/48 = 67 % of the tensor peak with this (SS) operand form.
    const int mma_id = warp - 1;
    // A tcgen05.mma of this shape occupies the tensor pipe for M*N/256 = 32 cycles (N = 64); the issuing thread must
    // stay below that per MMA, so descriptors are formed by adding constants to precomputed 64-bit bases, every inner
    // loop is unrolled and the region is guarded by elect.sync (a lane-id test makes ptxas broadcast each descriptor
    // through a loop: the first version of this kernel spent ~125 cycles per MMA on issue alone).
    if (elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_bf16(128, NOUT, 0, 0);
      const uint64_t a_base = umma_desc(s_slab, 16, SLAB_W * 128, 2);
      const uint64_t b_base = umma_desc(s_b, 16, 1024, 2);
      uint32_t bstage = 0, bphase = 0;
      // phase parity of every slab barrier: a slab that lies outside the volume (zero padding) is neither loaded nor waited
      // for, so its barrier does not complete in that unit — the parity is tracked per slab, not derived from the unit count
      uint32_t slab_phase = 0;
      int it = 0;
      long long t_start = clock64(), t_bwait = 0, t_swait = 0;
      for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
        int nb, n, d0, h0, w0;
        decode(u, nb, n, d0, h0, w0);
        if (it > 0) {
          mbar_wait(bar_acc_empty, (uint32_t)((it - 1) & 1));
          tc_fence_after();
        }
        uint32_t touched = 0, waited = 0;
        if (STACK) {
          // per input slot j, the kd range whose accumulators a = j - kd exist.  For a full unit (all NACC output slices
          // inside the volume) the ranges are compile-time constants and only the presence of a slab (zero padding at the
          // volume faces) is a run-time bit, so every descriptor below folds to base + constant; partial units take the
          // same code with run-time ranges (FULL = false).
          const int amax = min(NACC, g.Do - d0);
          uint32_t svalid = 0;
#pragma unroll
          for (int j = 0; j < NACC + 2; ++j) {
            const int d = d0 + j - g.pad_d;
            if (d >= 0 && d < g.Di) svalid |= 1u << j;
          }
          // The issue code is kept SMALL on purpose: the (kh, kw) loop is rolled (one peeled first position for the
          // "first touch" logic + one body of 4 k-steps x NACC+2 slabs executed eight times).  Fully unrolled it was ~50 KB of
          // straight-line SASS per instantiation that one thread walks through once per unit — inside the iteration, where other
          // kernels run in between, every launch then paid the instruction-cache misses of its whole issue path.
          auto run = [&](auto full_tag, auto grp_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            constexpr int GRP = decltype(grp_tag)::value;
            auto position = [&](auto first_tag, const uint32_t tapoff) {      // tapoff = ((kh * SLAB_W + kw) * 128) >> 4
              constexpr bool FIRST = decltype(first_tag)::value;
              if (p.dbg) t_bwait -= clock64();
              mbar_wait(bar_b_full(bstage), bphase);
              if (p.dbg) t_bwait += clock64();
              tc_fence_after();
              const uint64_t bs = b_base + (uint64_t)((bstage * Cfg::STAGE_BYTES) >> 4);
              const uint64_t aq = a_base + (uint64_t)tapoff;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                for (int j = 0; j < NACC + 2; ++j) {
                  const int g_lo = GRP * GACC, g_hi = (GRP + 1) * GACC - 1;     // accumulators of this group
                  const int a_lo = j - 2 > g_lo ? j - 2 : g_lo;
                  const int a_hi = FULL ? (j < g_hi ? j : g_hi) : min(min(j, g_hi), amax - 1);
                  const int klo = j - a_hi, n = a_hi - a_lo + 1;
                  if (n <= 0 || !((svalid >> j) & 1u)) continue;
                  const uint64_t ad = aq + (uint64_t)((j * SLAB_STRIDE + ks * 32) >> 4);
                  if (FIRST && ks == 0) {
                    // first touch: one MMA per accumulator so that each gets its own "overwrite" flag
                    if (!((waited >> j) & 1u)) {
                      if (p.dbg) t_swait -= clock64();
                      mbar_wait(bar_slab_full(j), (slab_phase >> j) & 1u);
                      slab_phase ^= 1u << j;
                      if (p.dbg) t_swait += clock64();
                      tc_fence_after();
                      waited |= 1u << j;
                    }
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                      if (kd < klo || kd >= klo + n) continue;
                      const int a = j - kd;
                      umma_bf16(tmem_base + acc_col(a), ad, bs + (uint64_t)((kd * Cfg::BTILE_BYTES) >> 4), IDESC, (touched >> a) & 1u);
                      touched |= 1u << a;
                    }
                  } else {
                    const uint64_t bd = bs + (uint64_t)((klo * Cfg::BTILE_BYTES + ks * 32) >> 4);
                    const uint32_t tacc = tmem_base + acc_col(j - klo);
                    if constexpr (!FULL) {
                      // run-time ranges: ONE instruction with the N field computed, not three predicated copies with three
                      // descriptor sets (the issue loop of a partial unit was instruction-bound, ~2x the time of a full unit)
                      umma_bf16_acc(tacc, ad, bd, umma_idesc_bf16(128, n * NOUT, 0, 0));
                    } else if (n == 3)
                      umma_bf16_acc(tacc, ad, bd, umma_idesc_bf16(128, 3 * NOUT, 0, 0));
                    else if (n == 2)
                      umma_bf16_acc(tacc, ad, bd, umma_idesc_bf16(128, 2 * NOUT, 0, 0));
                    else
                      umma_bf16_acc(tacc, ad, bd, IDESC);
                  }
                }
              }
              umma_commit(bar_b_empty(bstage));
              if (++bstage == Cfg::NB) { bstage = 0; bphase ^= 1u; }
            };
            position(std::true_type{}, 0u);
            uint32_t kh = 0, kw = 0;
            constexpr int QUNROLL = (HPVG_ROLL_STACK && !(HPVG_ROLL_WIDE_ONLY && Cfg::THIN)) ? 1 : 8;
#pragma unroll QUNROLL
            for (int q = 1; q < 9; ++q) {
              if (++kw == 3) { kw = 0; ++kh; }
              position(std::false_type{}, (kh * SLAB_W + kw) * 8u);
            }
          };
          if (amax == NACC) {
            run(std::true_type{}, std::integral_constant<int, 0>{});
            umma_commit(bar_acc_full);
            if (NGRP == 2) {
              run(std::true_type{}, std::integral_constant<int, NGRP - 1>{});
              umma_commit(bar_acc_full);
            }
          } else {
            run(std::false_type{}, std::integral_constant<int, 0>{});
            umma_commit(bar_acc_full);
            if (NGRP == 2) {
              run(std::false_type{}, std::integral_constant<int, NGRP - 1>{});
              umma_commit(bar_acc_full);
            }
          }
          umma_commit(bar_slabs_free);
          continue;
        }
#pragma unroll
        for (int grp = 0; grp < NGRP; ++grp) {
#pragma unroll
          for (int kd = 0; kd < KDT; ++kd) {
            // accumulators of this group that kd feeds
            uint32_t amask = 0;
#pragma unroll
            for (int a = grp * GACC; a < (grp + 1) * GACC; ++a) {
              const int d = d0 + a + kd - g.pad_d;
              if (a % NMMA == mma_id && d0 + a < g.Do && d >= 0 && d < g.Di) amask |= 1u << a;
            }
            if (amask == 0) {
              // nothing to do for this kd: the producer still streams its 9 weight taps, hand the slots back
#pragma unroll 1
              for (int q = 0; q < 9 * KCHUNKS; ++q) {
                mbar_wait(bar_b_full(bstage), bphase);
                umma_commit(bar_b_empty(bstage));
                if (++bstage == Cfg::NB) { bstage = 0; bphase ^= 1u; }
              }
              continue;
            }
            if (p.dbg) t_swait -= clock64();
#pragma unroll
            for (int a = grp * GACC; a < (grp + 1) * GACC; ++a) {
              if (!((amask >> a) & 1u)) continue;
              const int slot = a + kd;
              if ((waited >> slot) & 1u) continue;
              waited |= 1u << slot;
#pragma unroll
              for (int kc = 0; kc < KCHUNKS; ++kc) {
                const int si = slot * KCHUNKS + kc;
                mbar_wait(bar_slab_full(si), (slab_phase >> si) & 1u);
                slab_phase ^= 1u << si;
              }
            }
            if (p.dbg) t_swait += clock64();
            tc_fence_after();
            // rolled over the nine (kh, kw) positions: small issue code (see the stacked form above)
            uint32_t tapoff = 0, kwc = 0;      // ((kh * SLAB_W + kw) * 128) >> 4
#if HPVG_ROLL_PLAIN
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int q = 0; q < 9; ++q) {
              const uint64_t aq = a_base + (uint64_t)tapoff;
              if (++kwc == 3) { kwc = 0; tapoff += (SLAB_W - 2) * 8u; } else tapoff += 8u;
#pragma unroll
              for (int kc = 0; kc < KCHUNKS; ++kc) {
                if (p.dbg) t_bwait -= clock64();
                mbar_wait(bar_b_full(bstage), bphase);
                if (p.dbg) t_bwait += clock64();
                tc_fence_after();
                const uint64_t bd = b_base + (uint64_t)((bstage * Cfg::BTILE_BYTES) >> 4);
#pragma unroll
                for (int a = grp * GACC; a < (grp + 1) * GACC; ++a) {
                  if (!((amask >> a) & 1u)) continue;
                  const uint64_t ad = aq + (uint64_t)((((a + kd) * KCHUNKS + kc) * SLAB_STRIDE) >> 4);
                  const uint32_t tacc = tmem_base + acc_col(a);
                  umma_bf16(tacc, ad, bd, IDESC, (touched >> a) & 1u);
                  umma_bf16_acc(tacc, ad + 2, bd + 2, IDESC);
                  umma_bf16_acc(tacc, ad + 4, bd + 4, IDESC);
                  umma_bf16_acc(tacc, ad + 6, bd + 6, IDESC);
                }
                touched |= amask;
                umma_commit(bar_b_empty(bstage));
                if (++bstage == Cfg::NB) { bstage = 0; bphase ^= 1u; }
              }
            }
          }
          umma_commit(bar_acc_full);      // group `grp` of this unit is complete when everything issued so far is
        }
        umma_commit(bar_slabs_free);
      }
      if (p.dbg && mma_id == 0) {
        p.dbg[blockIdx.x * 8 + 0] = clock64() - t_start;   // MMA issue loop, all units
        p.dbg[blockIdx.x * 8 + 1] = t_bwait;               // waiting for weight tiles
        p.dbg[blockIdx.x * 8 + 2] = t_swait;               // waiting for input slabs
      }
    }
    __syncwarp();
  } else if constexpr (FUSE) {
    // ===================== fused epilogue: conv + BatchNorm(batch statistics) + LeakyReLU in ONE launch =====================
    // reference: modules/networks_3d.py:48-56 (nn.Conv3d -> nn.BatchNorm3d in training mode -> nn.LeakyReLU(0.2)).
    // The grid has exactly one unit per CTA and at most one CTA per SM (host-checked, cooperative launch), so the whole
    // layer output sits in the TMEM of the resident CTAs when the MMAs are done.  Pass 1 reads the accumulators and reduces
    // sum(y), sum(y^2) per channel FROM THE FP32 VALUES (warp-shuffle transpose reduction, one shared + one global atomic per
    // channel and CTA); a grid-wide arrive / spin barrier makes the totals visible; pass 2 reads the accumulators again,
    // normalises, applies the affine map and LeakyReLU in fp32 and stores the activated output, (optionally) the bf16 conv
    // output the BatchNorm backward needs, and one sign bit per element: the LeakyReLU derivative the backward pass uses is
    // then the fp32 one, not the sign of a bf16-rounded recomputation.
    constexpr int CPT = 32;
    const int q = warp & 3;
    const int ch = (warp - Cfg::EPI_WARP0) >> 2;              // column half: channels ch*32 .. ch*32+31
    const int m = q * 32 + lane;
    const int et = threadIdx.x - 32 * Cfg::EPI_WARP0;
    int nb, n, d0, h0, w0;
    decode((long long)blockIdx.x, nb, n, d0, h0, w0);
    const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
    const bool row_ok = (oh < g.Ho) && (ow < g.Wo);
    const int amax = min(NACC, g.Do - d0);
    float* csum = reinterpret_cast<float*>(sgen + Cfg::OFF_B);      // [128] channel sums of this CTA, then [128] scale | shift
    float* ss = csum + 128;
    // per-channel parameters the finalize step needs: fetched now, while the MMAs run, instead of as exposed global round trips
    // behind the grid barrier (where every CTA of the grid would wait for them)
    float gam = 0.f, bet = 0.f, rmean = 0.f, rvar = 0.f;
    if (et < 64) {
      gam = p.gamma[et];
      bet = p.beta[et];
      if (blockIdx.x == 0) {
        if (p.running_mean) rmean = p.running_mean[et];
        if (p.running_var) rvar = p.running_var[et];
      }
    }
    mbar_wait(bar_acc_full, 0);       // every MMA of the unit has completed: slabs and weight ring are dead from here on
    tc_fence_after();
    if (et < 128) csum[et] = 0.f;
    float s1[CPT], s2[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) s1[j] = s2[j] = 0.f;
    const float4* b4 = reinterpret_cast<const float4*>(bias_s + ch * CPT);
#pragma unroll 1
    for (int a = 0; a < amax; ++a) {
      uint32_t r[CPT];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc_col(a) + ch * CPT, r);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) {
          const float4 bq = b4[j];
          const float v0 = __uint_as_float(r[4 * j]) + bq.x, v1 = __uint_as_float(r[4 * j + 1]) + bq.y;
          const float v2 = __uint_as_float(r[4 * j + 2]) + bq.z, v3 = __uint_as_float(r[4 * j + 3]) + bq.w;
          s1[4 * j] += v0; s1[4 * j + 1] += v1; s1[4 * j + 2] += v2; s1[4 * j + 3] += v3;
          s2[4 * j] = fmaf(v0, v0, s2[4 * j]); s2[4 * j + 1] = fmaf(v1, v1, s2[4 * j + 1]);
          s2[4 * j + 2] = fmaf(v2, v2, s2[4 * j + 2]); s2[4 * j + 3] = fmaf(v3, v3, s2[4 * j + 3]);
        }
      }
    }
    // transpose reduction: 32 lanes x 32 columns -> lane l holds the sum of column l over the warp's 32 rows (31 shuffles per array)
#pragma unroll
    for (int wd = 16; wd >= 1; wd >>= 1) {
      const bool upper = (lane & wd) != 0;
#pragma unroll
      for (int j = 0; j < wd; ++j) {
        const float send1 = upper ? s1[j] : s1[j + wd], keep1 = upper ? s1[j + wd] : s1[j];
        const float send2 = upper ? s2[j] : s2[j + wd], keep2 = upper ? s2[j + wd] : s2[j];
        s1[j] = keep1 + __shfl_xor_sync(0xffffffffu, send1, wd);
        s2[j] = keep2 + __shfl_xor_sync(0xffffffffu, send2, wd);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");      // csum zeroed
    atomicAdd(csum + ch * CPT + lane, s1[0]);
    atomicAdd(csum + 64 + ch * CPT + lane, s2[0]);
    asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");
    if (et < 128) {
      atomicAdd(p.stats + et, csum[et]);           // stats = [sum y : 64][sum y^2 : 64]
      __threadfence();
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");
    if (et == 0) grid_barrier_arrive_and_wait(p.grid_counter, gridDim.x);      // all CTAs are co-resident (cooperative launch)
    asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");
    if (et < 64) {
      const int c = et;
      const double inv = 1.0 / (double)p.count;
      const double mean = (double)__ldcg(p.stats + c) * inv;
      double var = (double)__ldcg(p.stats + 64 + c) * inv - mean * mean;
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)p.eps));
      const float sc = gam * invstd;
      const float sh = bet - (float)mean * sc;
      ss[c] = sc;
      ss[64 + c] = sh;
      if (blockIdx.x == 0) {
        p.scale_shift[c] = sc;
        p.scale_shift[64 + c] = sh;
        p.mean_invstd[c] = (float)mean;
        p.mean_invstd[64 + c] = invstd;
        if (p.running_mean) p.running_mean[c] = (1.f - p.momentum) * rmean + p.momentum * (float)mean;
        if (p.running_var) {
          const double unbiased = p.count > 1 ? var * (double)p.count / (double)(p.count - 1) : var;
          p.running_var[c] = (1.f - p.momentum) * rvar + p.momentum * (float)unbiased;
        }
        if (c == 0 && p.nbt) p.nbt[0] += 1;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");
    // pass 2: normalise + LeakyReLU from the fp32 accumulators; every accumulator has its own staging tiles in the dead slab area
#pragma unroll 1
    for (int a = 0; a < amax; ++a) {
      const int od = d0 + a;
      uint32_t r[CPT];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc_col(a) + ch * CPT, r);
      tmem_ld_wait();
      const uint32_t sy = s_slab + (uint32_t)(2 * a) * STG_BYTES + m * 128;
      const uint32_t so = sy + STG_BYTES;
      const uint32_t sphase = (sy >> 7) & 7u;      // tiles are 16 KB apart: same swizzle phase for both
      uint32_t mbits = 0;
#pragma unroll
      for (int c = 0; c < CPT / 8; ++c) {
        float v[8], o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int col = ch * CPT + 8 * c + e;
          v[e] = __uint_as_float(r[8 * c + e]) + bias_s[col];
          const float z = fmaf(v[e], ss[col], ss[64 + col]);
          const bool pos = z > 0.f;
          mbits |= (pos ? 1u : 0u) << (8 * c + e);
          o[e] = pos ? z : z * p.slope;
        }
        const uint32_t chunk = ((uint32_t)((ch * (CPT / 8) + c) ^ sphase) << 4);
        if (p.store_y)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sy + chunk), "r"(pack_bf16x2(v[0], v[1])),
                       "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7]))
                       : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(so + chunk), "r"(pack_bf16x2(o[0], o[1])),
                     "r"(pack_bf16x2(o[2], o[3])), "r"(pack_bf16x2(o[4], o[5])), "r"(pack_bf16x2(o[6], o[7]))
                     : "memory");
      }
      if (p.mask_bits && row_ok)
        p.mask_bits[((((size_t)n * g.Do + od) * g.Ho + oh) * g.Wo + ow) * 2 + ch] = mbits;
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");
      if (et == 0) {
        if (p.store_y) tma_store_5d(&tmap_y, s_slab + (uint32_t)(2 * a) * STG_BYTES, 0, w0, h0, od, n);
        tma_store_5d(&tmap_o, s_slab + (uint32_t)(2 * a + 1) * STG_BYTES, 0, w0, h0, od, n);
        tma_store_commit();
      }
    }
#if HPVG_STORE_WAIT_READ
    if (et == 0) tma_store_wait_read<0>();      // the staging tiles have been read: the CTA may leave (the writes land by the end of the grid)
#else
    if (et == 0) tma_store_wait_all<0>();
#endif
  } else if (!Cfg::THIN) {
    // ===================== epilogue, wide output: NEPI warps, TMEM lane quadrant = warp & 3, CPT columns per thread ==========
    // Measured (bench_kernels.py clk): of the ~6.5 k cycles after the last MMA about half is this chain and half the
    // drain of the final TMA stores to HBM; 16 warps x 16 columns was not faster than 8 x 32.
    constexpr int CPT = 64 / (Cfg::NEPI / 4);                 // accumulator columns per thread (16 or 32)
    constexpr int RPT = 128 / (NEPI_THREADS / 64);            // rows per thread in the BatchNorm-sum pass
    const int q = warp & 3;                                   // TMEM lane quadrant this warp may read
    const int ch = (warp - Cfg::EPI_WARP0) >> 2;              // which CPT-wide column group
    const int m = q * 32 + lane;                              // accumulator row = brick voxel (hh = m / 8, ww = m % 8)
    const int et = threadIdx.x - 32 * Cfg::EPI_WARP0;         // index inside the epilogue group
    int it = 0;
    int stg = 0;
    uint32_t full_phase = 0;
    long long t_start = clock64(), t_accwait = 0, t_ld = 0, t_bar = 0, t_stat = 0;
    for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
      int nb, n, d0, h0, w0;
      decode(u, nb, n, d0, h0, w0);
      const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
      const bool row_ok = (oh < g.Ho) && (ow < g.Wo);
      const bool brick_full = (h0 + BH <= g.Ho) && (w0 + BW <= g.Wo);
      float st_s = 0.f, st_s2 = 0.f;     // BatchNorm sums of this unit: channel (et & 63), row group (et >> 6)
      // LeakyReLU'-mask epilogue (data gradients): the sign bits of this thread's row are fetched NOW, while the MMAs of the unit
      // run and the epilogue warps have nothing to do — one register per accumulator.  Read after the accumulators were complete,
      // each of the unit's accumulators paid an exposed global round trip (4 x ~1 k cycles of a 6 k-cycle epilogue).
      uint32_t mbits[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) mbits[a] = 0xffffffffu;
      if (p.mask_src && row_ok) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          const int od = d0 + a;
          if (od >= g.Do) continue;
          const uint4* mp = reinterpret_cast<const uint4*>(
              p.mask_src + ((((size_t)n * g.Do + od) * g.Ho + oh) * g.Wo + ow) * g.Cout + nb * 64 + ch * CPT);
          uint4 mv[CPT / 8];
#pragma unroll
          for (int c = 0; c < CPT / 8; ++c) mv[c] = __ldg(mp + c);
          uint32_t b = 0;
#pragma unroll
          for (int c = 0; c < CPT / 8; ++c) {
            const uint32_t wds[4] = {mv[c].x, mv[c].y, mv[c].z, mv[c].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = unpack_bf16x2(wds[k]);
              b |= (f.x > 0.f ? 1u : 0u) << (8 * c + 2 * k);
              b |= (f.y > 0.f ? 1u : 0u) << (8 * c + 2 * k + 1);
            }
          }
          mbits[a] = b;
        }
      }
#pragma unroll 1
      for (int grp = 0; grp < NGRP; ++grp) {
        long long tq = clock64();
        mbar_wait(bar_acc_full, full_phase);
        full_phase ^= 1u;
        t_accwait += clock64() - tq;
        tc_fence_after();
#pragma unroll 1
        for (int a = grp * GACC; a < (grp + 1) * GACC; ++a) {
          const int od = d0 + a;
          if (od >= g.Do) break;
          uint32_t r[CPT];
          if (p.dbg) t_ld -= clock64();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc_col(a) + ch * CPT;
          if (CPT == 64) {
            tmem_ld32(taddr, r);
            tmem_ld32(taddr + 32, r + (CPT == 64 ? 32 : 0));
          } else if (CPT == 32) {
            tmem_ld32(taddr, r);
          } else {
            tmem_ld16(taddr, r);
          }
          tmem_ld_wait();
          if (p.dbg) t_ld += clock64();
          float v[CPT];
#pragma unroll
          for (int j = 0; j < CPT; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + nb * 64 + ch * CPT);
#pragma unroll
            for (int j = 0; j < CPT / 4; ++j) {
              const float4 bq = b4[j];
              v[4 * j + 0] += bq.x; v[4 * j + 1] += bq.y; v[4 * j + 2] += bq.z; v[4 * j + 3] += bq.w;
            }
          }
          if (p.mask_src) {
            uint32_t mb = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < NACC; ++k)
              if (k == a) mb = mbits[k];
#pragma unroll
            for (int j = 0; j < CPT; ++j) v[j] *= ((mb >> j) & 1u) ? 1.f : p.slope;
          }
          if (p.act == HPVG_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) v[j] = fmaxf(v[j], v[j] * p.slope);     // slope in (0, 1): max(x, slope * x)
          }
          if (!brick_full && !row_ok) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) v[j] = 0.f;
          }
          // staging buffer must be free: the thread that issued its last TMA store waits for the read to finish
          if (p.dbg) t_bar -= clock64();
          if (et == 0) tma_store_wait_read<Cfg::NSTG - 1>();
          asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");
          if (p.dbg) t_bar += clock64();
          const uint32_t sdst = s_stg + stg * STG_BYTES + m * 128;
          const uint32_t sphase = (sdst >> 7) & 7u;      // 128-byte swizzle phase of this row (absolute address bits)
#pragma unroll
          for (int c = 0; c < CPT / 8; ++c) {
            const uint32_t addr = sdst + ((uint32_t)((ch * (CPT / 8) + c) ^ sphase) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[8 * c + 0], v[8 * c + 1])),
                         "r"(pack_bf16x2(v[8 * c + 2], v[8 * c + 3])), "r"(pack_bf16x2(v[8 * c + 4], v[8 * c + 5])),
                         "r"(pack_bf16x2(v[8 * c + 6], v[8 * c + 7]))
                         : "memory");
          }
          fence_proxy_async();
          if (p.dbg) t_bar -= clock64();
          asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");
          if (p.dbg) t_bar += clock64();
          if (et == 0) {
            tma_store_5d(&tmap_y, s_stg + stg * STG_BYTES, nb * 64, w0, h0, od, n);
            tma_store_commit();
          }
          if (p.dbg) t_stat -= clock64();
          if (p.stats) {
            // column sums over the staged (bf16-rounded, invalid rows zeroed) tile: thread = channel, RPT rows each
            const int c = et & 63, rg = et >> 6;
            const uint8_t* tile = sgen + Cfg::OFF_STG + stg * STG_BYTES;
            const uint32_t tile_row0 = (s_stg + stg * STG_BYTES) >> 7;
#pragma unroll
            for (int rr = 0; rr < RPT; ++rr) {
              const int row = rg * RPT + rr;
              const __nv_bfloat16 bv = *reinterpret_cast<const __nv_bfloat16*>(
                  tile + row * 128 + (((c >> 3) ^ ((tile_row0 + row) & 7)) << 4) + (c & 7) * 2);
              const float f = bf2f(bv);
              st_s += f;
              st_s2 = fmaf(f, f, st_s2);
            }
          }
          if (p.dbg) t_stat += clock64();
          stg = (stg + 1 == Cfg::NSTG) ? 0 : stg + 1;
        }
      }
      if (p.stats) {
        float* sp = p.stats + (size_t)n * g.stats_stride;      // per-sample statistics: a unit lies inside one sample
        atomicAdd(sp + nb * 64 + (et & 63), st_s);
        atomicAdd(sp + g.Cout + nb * 64 + (et & 63), st_s2);
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty);
    }
#if HPVG_STORE_WAIT_READ
    if (et == 0) tma_store_wait_read<0>();      // the staging tiles have been read: the CTA may leave (the writes land by the end of the grid)
#else
    if (et == 0) tma_store_wait_all<0>();
#endif
    if (p.dbg && et == 0) {
      p.dbg[blockIdx.x * 8 + 3] = clock64() - t_start;   // epilogue warps, all units
      p.dbg[blockIdx.x * 8 + 4] = t_accwait;             // of which waiting for the accumulators
      p.dbg[blockIdx.x * 8 + 5] = t_ld;                  // TMEM loads
      p.dbg[blockIdx.x * 8 + 6] = t_bar;                 // staging-buffer barriers
      p.dbg[blockIdx.x * 8 + 7] = t_stat;                // BatchNorm sums
    }
  } else {
    // ===================== epilogue, thin output: float32 NCDHW, Cout <= 16, bias only =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int it = 0;
    uint32_t full_phase = 0;
    const size_t out_sp = (size_t)g.Do * g.Ho * g.Wo;
    for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
      int nb, n, d0, h0, w0;
      decode(u, nb, n, d0, h0, w0);
      const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
      const bool row_ok = (oh < g.Ho) && (ow < g.Wo);
#pragma unroll 1
      for (int grp = 0; grp < NGRP; ++grp) {
        mbar_wait(bar_acc_full, full_phase);
        full_phase ^= 1u;
        tc_fence_after();
#pragma unroll 1
        for (int a = grp * GACC; a < (grp + 1) * GACC; ++a) {
          const int od = d0 + a;
          if (od >= g.Do) break;
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc_col(a), r);
          tmem_ld_wait();
          if (row_ok) {
            float* yp = p.y_thin + (size_t)n * g.Cout * out_sp + ((size_t)od * g.Ho + oh) * g.Wo + ow;
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c < g.Cout) yp[(size_t)c * out_sp] = __uint_as_float(r[c]) + bias_s[c];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  if (p.dbg && threadIdx.x == 0) {
    // wall-clock span of this CTA (ns): launch ramp and teardown show up as the difference to the event-timed duration
    unsigned long long gt1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
    p.dbg[blockIdx.x * 8 + 5] = (long long)gt0;
    p.dbg[blockIdx.x * 8 + 6] = (long long)gt1;
  }
}

template <int KCHUNKS, int NACC, int KDT, int NGRP, int NOUT, bool STACK = false>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<KCHUNKS, NACC, NGRP, NOUT, STACK>;
  static std::atomic<unsigned long long> attr_mask{0};
  if (attr_pending(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<KCHUNKS, NACC, KDT, NGRP, NOUT, STACK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("conv_tc: cannot opt in to %d bytes of shared memory: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return -2;
    }
    attr_set(attr_mask);
  }
  const ConvGeom& g = p.g;
  p.units_d = (int)cdiv(g.Do, NACC);
  p.units_h = (int)cdiv(g.Ho, BH);
  p.units_w = (int)cdiv(g.Wo, BW);
  p.nblocks = NOUT == 64 ? g.Cout / 64 : 1;
  p.num_units = (long long)p.nblocks * g.N * p.units_d * p.units_h * p.units_w;
  if (p.num_units >= (1LL << 31)) {
    set_error("conv_tc: %lld work items exceed the 32-bit unit index of the kernel", (long long)p.num_units);
    return -1;
  }
  const int grid = (int)min((long long)num_sms(), p.num_units);
  launch_k(conv_tc_kernel<KCHUNKS, NACC, KDT, NGRP, NOUT, STACK, false>, grid, Cfg::THREADS, Cfg::SMEM_BYTES, st, mx, mw, my, my, p);
  HPVG_CHECK_LAUNCH("conv_tc_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// conv + BatchNorm(train) + LeakyReLU in one launch (FUSE epilogue).  One unit per CTA, one CTA per SM, the whole grid
// co-resident (cooperative launch): eligible when the layer's units fit the SMs.
// ---------------------------------------------------------------------------------------------------------------
static int fused_nacc(const ConvGeom& g) {
  if (g.KD != 3 || g.Cin != 64 || g.Cout != 64 || g.pad_d != g.pad) return 0;
  const long long per_slice = (long long)g.N * cdiv(g.Ho, BH) * cdiv(g.Wo, BW);
  const long long sms = num_sms();
  // 2-slice units put twice as many SMs to work on the small pyramid levels; 4-slice units have the better MMA shapes
  if (g.Do % 2 == 0 && per_slice * (g.Do / 2) <= sms) return 2;
  if (per_slice * cdiv(g.Do, 4) <= sms) return 4;
  return 0;
}

bool conv_bn_fused_supported(const ConvGeom& g) { return fused_nacc(g) != 0; }

template <int NACC>
static int launch_fused(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const CUtensorMap& mo, TcParams& p,
                        cudaStream_t st) {
  using Cfg = TcCfg<1, NACC, 1, 64, true>;
  auto kernel = conv_tc_kernel<1, NACC, 3, 1, 64, true, true>;
  static std::atomic<unsigned long long> attr_mask{0};
  if (attr_pending(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("conv_bn_fused: cannot opt in to %d bytes of shared memory: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return -2;
    }
    attr_set(attr_mask);
  }
  const ConvGeom& g = p.g;
  p.units_d = (int)cdiv(g.Do, NACC);
  p.units_h = (int)cdiv(g.Ho, BH);
  p.units_w = (int)cdiv(g.Wo, BW);
  p.nblocks = 1;
  p.num_units = (long long)g.N * p.units_d * p.units_h * p.units_w;
  if (p.num_units >= (1LL << 31)) {
    set_error("conv_bn_lrelu_fused: %lld work items exceed the 32-bit unit index of the kernel", (long long)p.num_units);
    return -1;
  }
  if (p.num_units > num_sms()) {
    set_error("conv_bn_fused: %lld units do not fit %d SMs", p.num_units, num_sms());
    return -1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.num_units);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // co-residency of the whole grid is guaranteed by the driver (or the launch fails)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  static const bool coop = !(getenv("HPVG_FUSED_COOP") && atoi(getenv("HPVG_FUSED_COOP")) == 0);
  cfg.numAttrs = coop ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, mx, mw, my, mo, p);
  HPVG_CHECK_LAUNCH("conv_tc_kernel<fused BatchNorm>");
  return 0;
}

int conv_bn_fused(const void* x, const void* w_packed, const float* bias, void* y, void* out, const ConvGeom& g, float slope,
                  const float* gamma, const float* beta, float* running_mean, float* running_var, long long* nbt, float momentum,
                  float eps, float* stats, unsigned* grid_counter, float* scale_shift, float* mean_invstd, uint32_t* mask_bits,
                  cudaStream_t st) {
  const int nacc = fused_nacc(g);
  if (nacc == 0) {
    set_error("conv_bn_fused: layer not eligible (Cin=%d Cout=%d KD=%d %dx%dx%d)", g.Cin, g.Cout, g.KD, g.Do, g.Ho, g.Wo);
    return -1;
  }
  CUtensorMap mx, mw, my, mo;
  {
    uint64_t dims[5] = {(uint64_t)g.Cin, (uint64_t)g.Wi, (uint64_t)g.Hi, (uint64_t)g.Di, (uint64_t)g.N};
    uint32_t box[5] = {64, SLAB_W, SLAB_H, 1, 1};
    if (int rc = make_tmap_bf16(&mx, x, 5, dims, box)) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)g.Cin, (uint64_t)g.taps * g.Cout};
    uint32_t box[2] = {64, 64};
    if (int rc = make_tmap_bf16(&mw, w_packed, 2, dims, box)) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)g.Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)g.N};
    uint32_t box[5] = {64, BW, BH, 1, 1};
    if (int rc = make_tmap_bf16(&mo, out, 5, dims, box)) return rc;
    if (y) {
      if (int rc = make_tmap_bf16(&my, y, 5, dims, box)) return rc;
    } else {
      my = mo;
    }
  }
  TcParams p = {};
  p.g = g;
  p.act = HPVG_ACT_NONE;
  p.slope = slope;
  p.bias = bias;
  p.stats = stats;
  p.mask_src = nullptr;
  p.y_thin = nullptr;
  p.dbg = nullptr;
  p.w_img = reinterpret_cast<const uint8_t*>(w_packed);
  p.w_img_bytes = (unsigned)((size_t)g.taps * g.Cout * g.Cin * 2);
  p.gamma = gamma; p.beta = beta;
  p.running_mean = running_mean; p.running_var = running_var; p.nbt = nbt;
  p.momentum = momentum; p.eps = eps;
  p.count = (long long)g.N * g.Do * g.Ho * g.Wo;
  p.scale_shift = scale_shift; p.mean_invstd = mean_invstd;
  p.grid_counter = grid_counter;
  p.mask_bits = mask_bits;
  p.store_y = y != nullptr;
  if (nacc == 2) return launch_fused<2>(mx, mw, my, mo, p, st);
  return launch_fused<4>(mx, mw, my, mo, p, st);
}

// wide -> wide (Cout multiple of 64) or wide -> thin (Cout <= 16, packed weights zero-padded to 16 rows per tap)
bool conv_tc_supported(int x_fmt, int y_fmt, const ConvGeom& g, const void* w_packed) {
  if (x_fmt != HPVG_FMT_NDHWC_BF16 || w_packed == nullptr || !(g.Cin == 64 || g.Cin == 128) || g.Wi > 65535 || g.Hi > 65535) return false;
  if (y_fmt == HPVG_FMT_NDHWC_BF16) return (g.Cout % 64 == 0) && g.Cout >= 64 && g.Cout <= MAX_COUT;
  return g.Cout <= 16 && g.Cin == 64;
}

// conv_col.cu
bool conv_col_supported(int x_fmt, int y_fmt, const ConvGeom& g, const void* w_packed);
long long conv_col_brick_units(const ConvGeom& g);
int conv_col(const void* x, const void* w_packed, const float* bias, void* y, const ConvGeom& g, int act, float slope, float* stats,
             const void* mask_src, cudaStream_t st);

// Column-streaming kernel (conv_col.cu) or brick kernel?  The brick kernel is faster per unit of work (N = 192 MMAs) but
// deals work out in 4-slice x 128-voxel units, one round of at most num_sms() units at a time: a volume with 6 d-slices
// wastes half of every second unit, 54 x 54 wastes a fifth of its bricks, and 448 units on 148 SMs take four rounds.  The
// column kernel deals out single 128-voxel x 32-channel tiles in contiguous runs.  Measured (batch-8 generation, us per
// launch brick / column): 4x32x32 23.0 / 15.9, 4x39x39 23.7 / 21.4, 6x46x46 65.4 / 29.3, 6x54x54 85.3 / 41.0, 16x64x64
// 110.2 / 113.4; batch 1 at 16x64x64 22.6 / 24.6.  Rule (mode -1, the default): brick when the voxels it really computes
// fill at least HPVG_TC_COL_EFF (default 0.8) of the unit slots of its rounds.  (Host logic only: hpvg_conv_kernel_choice
// exposes it to the CPU tests.)
bool conv_tc_picks_column(int y_fmt, const ConvGeom& g, const void* w_packed) {
  const int col_mode = conv_col_mode();
  static const double col_eff = getenv("HPVG_TC_COL_EFF") ? atof(getenv("HPVG_TC_COL_EFF")) : 0.8;
  if (col_mode == 0 || !conv_col_supported(HPVG_FMT_NDHWC_BF16, y_fmt, g, w_packed)) return false;
  if (col_mode == 1) return true;
  const long long units = conv_col_brick_units(g) / (g.Cout / 64);     // per 64-channel block
  const long long rounds = cdiv(units * (g.Cout / 64), num_sms());
  const double eff = (double)g.N * g.Do * g.Ho * g.Wo * (g.Cout / 64) / ((double)rounds * num_sms() * 512.0);
  // ... or when its last d-unit would be partial (Do % 4 != 0, e.g. the 13 / 7 / 5-frame levels of the reference's default
  // sampling rates): a partial unit takes the run-time-range issue path, ~2x the time of a full unit, and with one unit
  // per CTA it sets the kernel time
  return eff < col_eff || (g.Do % 4 != 0);
}

int conv_tc(const void* x, const void* w_packed, const float* bias, void* y, int y_fmt, const ConvGeom& g, int act, float slope,
            float* stats, const void* mask_src, cudaStream_t st) {
  const bool thin = y_fmt == HPVG_FMT_NCDHW_F32;
  if (conv_tc_picks_column(y_fmt, g, w_packed)) return conv_col(x, w_packed, bias, y, g, act, slope, stats, mask_src, st);
  const int nout = thin ? 16 : 64;
  CUtensorMap mx, mw, my;
  {
    uint64_t dims[5] = {(uint64_t)g.Cin, (uint64_t)g.Wi, (uint64_t)g.Hi, (uint64_t)g.Di, (uint64_t)g.N};
    uint32_t box[5] = {64, SLAB_W, SLAB_H, 1, 1};
    if (int rc = make_tmap_bf16(&mx, x, 5, dims, box)) return rc;
  }
  {
    const int rows_per_tap = thin ? 16 : g.Cout;
    uint64_t dims[2] = {(uint64_t)g.Cin, (uint64_t)g.taps * rows_per_tap};
    uint32_t box[2] = {64, (uint32_t)nout};
    if (int rc = make_tmap_bf16(&mw, w_packed, 2, dims, box)) return rc;
  }
  if (!thin) {
    uint64_t dims[5] = {(uint64_t)g.Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)g.N};
    uint32_t box[5] = {64, BW, BH, 1, 1};
    if (int rc = make_tmap_bf16(&my, y, 5, dims, box)) return rc;
  } else {
    my = mx;
  }
  TcParams p = {};
  p.g = g;
  p.act = act;
  p.slope = slope;
  p.bias = bias;
  p.stats = stats;
  p.mask_src = reinterpret_cast<const __nv_bfloat16*>(mask_src);
  p.y_thin = thin ? reinterpret_cast<float*>(y) : nullptr;
  p.dbg = debug_clock_buffer();
  static const bool w_prefetch = !(getenv("HPVG_TC_WPREFETCH") && atoi(getenv("HPVG_TC_WPREFETCH")) == 0);
  p.w_img = w_prefetch ? reinterpret_cast<const uint8_t*>(w_packed) : nullptr;
  p.w_img_bytes = (unsigned)((size_t)g.taps * (thin ? 16 : g.Cout) * g.Cin * 2);
  static const int variant = getenv("HPVG_TC_VARIANT") ? atoi(getenv("HPVG_TC_VARIANT")) : 0;   // tuning knob: 1 = unstacked, 2 = unstacked two groups, 3 = stacked two groups (measured: 24.5 us vs 22.4 us for the default)
  if (thin) {
    // 2-slice units when the depth is even but not a multiple of 4 (the 6-slice pyramid levels): every unit is then full.
    // A 4-slice unit with only two valid output slices takes the run-time-range path of the issue loop — measured 39.5 us
    // per launch at 6 x 46 x 46 and 6 x 54 x 54 against 17.6 us for the (larger) 16 x 64 x 64 volume.
    if (g.KD == 3 && variant != 1 && g.Do % 4 != 0 && g.Do % 2 == 0) return launch_tc<1, 2, 3, 1, 16, true>(mx, mw, my, p, st);
    if (g.KD == 3) return variant == 1 ? launch_tc<1, 4, 3, 1, 16>(mx, mw, my, p, st) : launch_tc<1, 4, 3, 1, 16, true>(mx, mw, my, p, st);
    return launch_tc<1, 4, 1, 1, 16>(mx, mw, my, p, st);
  }
  if (g.KD == 3) {
    if (g.Cin == 64) {
      if (variant == 1) return launch_tc<1, 4, 3, 1, 64>(mx, mw, my, p, st);
      if (variant == 2) return launch_tc<1, 4, 3, 2, 64>(mx, mw, my, p, st);
      if (variant == 3) return launch_tc<1, 4, 3, 2, 64, true>(mx, mw, my, p, st);
      if (variant == 4) return launch_tc<1, 1, 3, 1, 64>(mx, mw, my, p, st);
      return launch_tc<1, 4, 3, 1, 64, true>(mx, mw, my, p, st);
    }
    return launch_tc<2, 2, 3, 2, 64>(mx, mw, my, p, st);
  }
  if (g.Cin == 64) return launch_tc<1, 4, 1, 2, 64>(mx, mw, my, p, st);
  return launch_tc<2, 2, 1, 2, 64>(mx, mw, my, p, st);
}

}  // namespace hpvg

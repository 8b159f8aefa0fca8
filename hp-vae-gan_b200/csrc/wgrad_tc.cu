// tcgen05 weight-gradient kernel for the wide layers (Cin, Cout multiples of 64, NDHWC bf16 operands).
//
// Replaces aten::convolution_backward(grad_weight) of nn.Conv3d/Conv2d (modules/networks_3d.py:51,63) and the
// "wgrad-as-conv" nodes of the WGAN-GP double backward (modules/utils.py:14-18).
//
//   dw[co][ci][kd,kh,kw] = sum_{n,od,oh,ow} gy[n,od,oh,ow,co] * x[n,od+kd-p,oh+kh-p,ow+kw-p,ci]
//
// GEMM view per tap: D[ci][co] += X_tap^T[ci][voxel] * GY[voxel][co], K = voxels.  Both operands are MN-major views of
// NDHWC rows (128 B = 64 channels per voxel), so TMA tiles feed tcgen05 directly:
//  * B = the gy brick (16 x 8 output voxels x 64 co), a plain 128-row tile.
//  * A = the x halo slab of the matching input slice ((16+2) x (8+2) voxels x 64 ci); tap (kh,kw) is the same slab read
//    with the descriptor start moved by kh*10+kw rows and K-groups (8 voxels of one brick row) 10 rows apart.
//  * Two taps share one M = 128 instruction: the second 64-row atom of A is the slab shifted by the row distance
//    between the taps (descriptor LBO), so the tensor core runs at full M (verified by experiments/umma_desc_probe.cu
//    tests 17-21).  9 taps of one kd -> 5 accumulators x 64 TMEM columns.
//  * grid = (voxel splits, 64x64 channel blocks, KD); each CTA reduces its bricks into TMEM, then writes one fp32
//    partial; a second kernel sums the splits in fixed order (deterministic) into PyTorch [Cout][Cin][taps] layout.
#include "common.cuh"
#include <cstdlib>

namespace hpvg {

constexpr int WG_BH = 16, WG_BW = 8;
constexpr int WG_SLAB_W = WG_BW + 2, WG_SLAB_H = WG_BH + 2;
constexpr int WG_SLAB_BYTES = WG_SLAB_H * WG_SLAB_W * 128;   // 23040
constexpr int WG_SLAB_STRIDE = 23 * 1024;
constexpr int WG_GY_BYTES = 128 * 128;                        // 16384
constexpr int WG_STAGE_BYTES = WG_SLAB_STRIDE + WG_GY_BYTES;  // 39936 (multiple of 1024)
constexpr int WG_STAGES = 5;
constexpr int WG_THREADS = 192;
constexpr int WG_NACC = 5;
constexpr int WG_TMEM_COLS = 512;
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + (2 * WG_STAGES + 1) * 8 + 16 + 1024;

struct WgParams {
  ConvGeom g;
  int bricks_h, bricks_w;
  long long num_bricks;       // N * Do * bricks_h * bricks_w  (over gy / output voxels)
  long long bricks_per_split;
  int splits;
  float* partial;             // [splits][taps][Cin][Cout] fp32
  long long* dbg;             // optional per-CTA phase clocks (development aid)
  // in-kernel split-K reduction (kd-stacked kernel): after a grid-wide barrier every CTA sums a slice of the partials over the
  // splits and writes dw in PyTorch layout — the separate reduction launch and its ~9 us on the stream disappear
  float* dw;                  // nullptr: the partials are reduced by wgrad_reduce_kernel
  unsigned* counter;          // zero-initialised arrival counter of the grid barrier
};

// STAGED (hpvg_set_wgrad_mode(2); measured: drain 9.8 k -> 5.6 k cycles per CTA): the drain writes its partials through a swizzled shared-memory tile so that every
// store instruction covers 512 contiguous bytes (see wgrad_tc_kdstack_kernel); <false> is the measured kernel, unchanged.
template <bool STAGED>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_gy, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t s_bar = sbase + WG_STAGES * WG_STAGE_BYTES;
  auto bar_full = [&](int i) { return s_bar + 8u * i; };
  auto bar_empty = [&](int i) { return s_bar + 8u * (WG_STAGES + i); };
  const uint32_t bar_acc = s_bar + 8u * (2 * WG_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + WG_STAGES * WG_STAGE_BYTES + (2 * WG_STAGES + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;
  const int split = blockIdx.x;
  const int cblocks_in = g.Cin / 64;
  const int ci_blk = blockIdx.y % cblocks_in, co_blk = blockIdx.y / cblocks_in;
  const int kd = blockIdx.z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    mbar_init(bar_acc, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_gy);
  }
  pdl_trigger();      // see conv_tc.cu: set-up above overlaps the previous kernel's tail, operands are read after the wait
  pdl_wait();
  __syncthreads();    // barriers initialised and visible
  // the producer warp starts its TMA loads at once; the TMEM allocation of warp 1 overlaps their flight (see conv_tc.cu)
  uint32_t tmem_base = 0;
  if (warp != 0) {
    if (warp == 1) tmem_alloc<WG_TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(WG_THREADS - 32) : "memory");
    tc_fence_after();
    tmem_base = *tmem_slot;
  }

  long long t_d0 = 0, t_d1 = 0;   // drain-phase clocks (development aid)
  const long long b_begin = (long long)split * p.bricks_per_split;
  const long long b_end = min(p.num_bricks, b_begin + p.bricks_per_split);

  // brick -> (n, od, h0, w0); returns false when the input slice of this kd is outside the tensor (zero contribution)
  auto decode = [&](long long b64, int& n, int& od, int& h0, int& w0) -> bool {      // 32-bit arithmetic: see conv_tc.cu
    unsigned b = (unsigned)b64;
    unsigned q = b / (unsigned)p.bricks_w;
    w0 = (int)(b - q * (unsigned)p.bricks_w) * WG_BW;
    b = q;
    q = b / (unsigned)p.bricks_h;
    h0 = (int)(b - q * (unsigned)p.bricks_h) * WG_BH;
    b = q;
    q = b / (unsigned)g.Do;
    od = (int)(b - q * (unsigned)g.Do);
    n = (int)q;
    const int d = od + kd - g.pad_d;
    return d >= 0 && d < g.Di;
  };

  if (warp == 0) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (long long b = b_begin; b < b_end; ++b) {
        int n, od, h0, w0;
        if (!decode(b, n, od, h0, w0)) continue;
        mbar_wait(bar_empty(stage), phase ^ 1u);
        mbar_expect_tx(bar_full(stage), WG_SLAB_BYTES + WG_GY_BYTES);
        const uint32_t sa = sbase + stage * WG_STAGE_BYTES;
        tma_load_5d(sa, &tmap_x, bar_full(stage), ci_blk * 64, w0 - g.pad, h0 - g.pad, od + kd - g.pad_d, n);
        tma_load_5d(sa + WG_SLAB_STRIDE, &tmap_gy, bar_full(stage), co_blk * 64, w0, h0, od, n);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // one elected thread issues every MMA; descriptors are precomputed bases + compile-time constants (see conv_tc.cu:
    // a lane-id guard makes ptxas wrap each tcgen05.mma in a descriptor-broadcast loop)
    if (elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_bf16(128, 64, 1, 1);   // A and B MN-major
      uint32_t stage = 0, phase = 0;
      uint32_t first = 1;
      long long t_start = clock64(), t_wait = 0;
      for (long long b = b_begin; b < b_end; ++b) {
        int n, od, h0, w0;
        if (!decode(b, n, od, h0, w0)) continue;
        if (p.dbg) t_wait -= clock64();
        mbar_wait(bar_full(stage), phase);
        if (p.dbg) t_wait += clock64();
        tc_fence_after();
        const uint32_t sa = sbase + stage * WG_STAGE_BYTES;
        const uint64_t b_base = umma_desc(sa + WG_SLAB_STRIDE, 16, 1024, 2);
#pragma unroll
        for (int pr = 0; pr < WG_NACC; ++pr) {
          // taps q0 = 2*pr and q1 = 2*pr+1 (q = kh*3 + kw); the last pair duplicates tap 8
          const int q0 = 2 * pr, q1 = (2 * pr + 1 <= 8) ? 2 * pr + 1 : 8;
          const int off0 = (q0 / 3) * WG_SLAB_W + (q0 % 3), off1 = (q1 / 3) * WG_SLAB_W + (q1 % 3);
          const uint32_t lbo = (uint32_t)(off1 - off0) * 128u;
          const uint64_t a_base = umma_desc(sa + (uint32_t)off0 * 128u, lbo, WG_SLAB_W * 128, 2);
          const uint32_t tacc = tmem_base + pr * 64;
          // K step = 16 voxels = brick rows 2ks, 2ks+1 (8 voxels each): A advances 2 slab rows, B 2048 bytes
          umma_bf16(tacc, a_base, b_base, IDESC, first ^ 1u);
#pragma unroll
          for (int ks = 1; ks < 8; ++ks)
            umma_bf16_acc(tacc, a_base + (uint64_t)((2 * ks * WG_SLAB_W * 128) >> 4), b_base + (uint64_t)((ks * 2048) >> 4), IDESC);
        }
        umma_commit(bar_empty(stage));
        first = 0;
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(bar_acc);
      if (p.dbg) {
        const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.dbg[cta * 8 + 0] = clock64() - t_start;    // MMA issue loop
        p.dbg[cta * 8 + 1] = t_wait;                 // waiting for TMA stages
      }
    }
    __syncwarp();
  } else {
    // ===================== drain: TMEM -> fp32 partial [tap][ci][co] =====================
    bool any = false;
    for (long long b = b_begin; b < b_end; ++b) {
      int n, od, h0, w0;
      if (decode(b, n, od, h0, w0)) { any = true; break; }
    }
    t_d0 = clock64();
    mbar_wait(bar_acc, 0);
    t_d1 = clock64();
    tc_fence_after();
    const int q = warp & 3;
    const int m = q * 32 + lane;           // TMEM lane: (m / 64) selects the tap of the pair, m % 64 = ci
    const int ci = ci_blk * 64 + (m & 63);
    for (int pr = 0; pr < WG_NACC; ++pr) {
      const int tap9 = 2 * pr + (m >> 6);
      uint32_t r[64];
      if (any) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + pr * 64;
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 32, r + 32);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 64; ++j) r[j] = 0u;
      }
      if (tap9 <= 8) {
        const int tap = kd * 9 + tap9;
        if constexpr (STAGED) {
          uint8_t* stg = sgen + (warp - 2) * 8192;      // pipeline stage 0 is dead once bar_acc has completed
#pragma unroll
          for (int j = 0; j < 16; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 256 + ((j ^ (lane & 15)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
          float* tile = p.partial + (((size_t)split * g.taps + tap) * g.Cin + (ci - lane)) * g.Cout + co_blk * 64;
#pragma unroll
          for (int it = 0; it < 16; ++it) {
            const int c = it * 32 + lane, row = c >> 4, col = c & 15;
            const uint4 v = *reinterpret_cast<const uint4*>(stg + row * 256 + ((col ^ (row & 15)) << 4));
            *reinterpret_cast<uint4*>(tile + (size_t)row * g.Cout + col * 4) = v;
          }
          __syncwarp();
        } else {
          float4* dst = reinterpret_cast<float4*>(p.partial + (((size_t)split * g.taps + tap) * g.Cin + ci) * g.Cout + co_blk * 64);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                 __uint_as_float(r[4 * j + 3]));
        }
      }
    }
  }
  if (p.dbg && threadIdx.x == 64) {
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    p.dbg[cta * 8 + 2] = t_d1 - t_d0;              // drain warps waiting for the accumulators
    p.dbg[cta * 8 + 3] = clock64() - t_d1;         // drain (TMEM -> global partial)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<WG_TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// kd-stacked form (KD == 3 only; the default: hpvg_set_wgrad_mode(1)).  Measured on B200 (experiments/check_wgrad_stack.py,
// profiles/r02a_wgrad_stack.txt): bit-for-bit the same partial sums as the one-kd-per-CTA kernel up to fp32 summation order
// (1.4e-6 relative), 18.8 k instead of 29.8 k cycles of MMA issue per CTA, 29.7 instead of 37.9 us per call at 16 x 64 x 64.
//
// Measured cost model: an M = 128, K = 16 MMA with both operands in shared memory costs max(64, N/2)
// cycles, so the N = 64 instructions above run the tensor pipe at half rate by construction (5 x 64 = 320 clk per
// k-step and brick for 9 taps).  Here the work item is one INPUT slab (slice d); the gy bricks of the three output
// slices it feeds, od = d + pad_d - kd, are loaded side by side and read as ONE MN-major B operand of N = 192 whose three
// 64-column N-atoms lie WG_GY_BYTES apart (descriptor LBO — the same mechanism the A operand uses for its two taps):
//     D[ci][kd*64 + co] += X_d(kh,kw)^T[ci][voxel] * [GY_{d+p} | GY_{d+p-1} | GY_{d+p-2}][voxel][kd*64 + co]
// One instruction serves the three kd taps of a (kh,kw) position.  A CTA owns one kh (blockIdx.z): taps kw = 0,1 share an
// M = 128 instruction, kw = 2 takes a second one (upper half duplicated and ignored): 2 x 96 = 192 clk per k-step and
// slab for the same 9 taps.  TMEM: 2 accumulators x 192 columns.  gy slices outside the volume are fetched as fully
// out-of-bounds TMA boxes (zero fill, full byte count; confirmed on the GPU).
// Partials keep the layout [split][tap][ci][co]; the reduction kernel is shared.
// ---------------------------------------------------------------------------------------------------------------
constexpr int WK_STAGE_BYTES = WG_SLAB_STRIDE + 3 * WG_GY_BYTES;   // 72704 (multiple of 1024)
constexpr int WK_STAGES = 3;
constexpr int WK_NACC = 2;
constexpr int WK_ACC_COLS = 192;
constexpr int WK_SMEM_BYTES = WK_STAGES * WK_STAGE_BYTES + (2 * WK_STAGES + 1) * 8 + 16 + 1024;
static_assert(WK_STAGE_BYTES % 1024 == 0, "stage bases must keep the 1024-byte swizzle alignment");
static_assert(WK_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(WK_NACC * WK_ACC_COLS <= WG_TMEM_COLS, "TMEM budget");

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kdstack_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_gy, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t s_bar = sbase + WK_STAGES * WK_STAGE_BYTES;
  auto bar_full = [&](int i) { return s_bar + 8u * i; };
  auto bar_empty = [&](int i) { return s_bar + 8u * (WK_STAGES + i); };
  const uint32_t bar_acc = s_bar + 8u * (2 * WK_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + WK_STAGES * WK_STAGE_BYTES + (2 * WK_STAGES + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;
  const int split = blockIdx.x;
  const int cblocks_in = g.Cin / 64;
  const int ci_blk = blockIdx.y % cblocks_in, co_blk = blockIdx.y / cblocks_in;
  const int kh = blockIdx.z;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WK_STAGES; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    mbar_init(bar_acc, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_gy);
  }
  pdl_trigger();
  pdl_wait();
  __syncthreads();
  uint32_t tmem_base = 0;
  if (warp != 0) {
    if (warp == 1) tmem_alloc<WG_TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(WG_THREADS - 32) : "memory");
    tc_fence_after();
    tmem_base = *tmem_slot;
  }

  // work items = input slabs: num_bricks = N * Di * bricks_h * bricks_w here
  const long long b_begin = (long long)split * p.bricks_per_split;
  const long long b_end = min(p.num_bricks, b_begin + p.bricks_per_split);
  auto decode = [&](long long b64, int& n, int& d, int& h0, int& w0) {      // 32-bit arithmetic: see conv_tc.cu
    unsigned b = (unsigned)b64;
    unsigned q = b / (unsigned)p.bricks_w;
    w0 = (int)(b - q * (unsigned)p.bricks_w) * WG_BW;
    b = q;
    q = b / (unsigned)p.bricks_h;
    h0 = (int)(b - q * (unsigned)p.bricks_h) * WG_BH;
    b = q;
    q = b / (unsigned)g.Di;
    d = (int)(b - q * (unsigned)g.Di);
    n = (int)q;
  };

  if (warp == 0) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (long long b = b_begin; b < b_end; ++b) {
        int n, d, h0, w0;
        decode(b, n, d, h0, w0);
        mbar_wait(bar_empty(stage), phase ^ 1u);
        mbar_expect_tx(bar_full(stage), WG_SLAB_BYTES + 3 * WG_GY_BYTES);
        const uint32_t sa = sbase + stage * WK_STAGE_BYTES;
        tma_load_5d(sa, &tmap_x, bar_full(stage), ci_blk * 64, w0 - g.pad, h0 - g.pad, d, n);
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)      // N-atom kd = the gy slice that meets this slab through tap kd
          tma_load_5d(sa + WG_SLAB_STRIDE + kd * WG_GY_BYTES, &tmap_gy, bar_full(stage), co_blk * 64, w0, h0, d + g.pad_d - kd, n);
        if (++stage == WK_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_bf16(128, WK_ACC_COLS, 1, 1);   // A and B MN-major, N = 3 atoms of 64
      const uint32_t a_kh = (uint32_t)kh * WG_SLAB_W * 128u;
      uint32_t stage = 0, phase = 0;
      uint32_t first = 1;
      long long t_start = clock64(), t_wait = 0;
      for (long long b = b_begin; b < b_end; ++b) {
        if (p.dbg) t_wait -= clock64();
        mbar_wait(bar_full(stage), phase);
        if (p.dbg) t_wait += clock64();
        tc_fence_after();
        const uint32_t sa = sbase + stage * WK_STAGE_BYTES;
        const uint64_t b_base = umma_desc(sa + WG_SLAB_STRIDE, WG_GY_BYTES, 1024, 2);
#pragma unroll
        for (int pr = 0; pr < WK_NACC; ++pr) {
          // pr 0: taps kw = 0 (lanes 0-63) and kw = 1 (lanes 64-127), one slab row apart; pr 1: kw = 2 twice
          const uint64_t a_base = umma_desc(sa + a_kh + (uint32_t)(2 * pr) * 128u, pr == 0 ? 128u : 0u, WG_SLAB_W * 128, 2);
          const uint32_t tacc = tmem_base + pr * WK_ACC_COLS;
          umma_bf16(tacc, a_base, b_base, IDESC, first ^ 1u);
#pragma unroll
          for (int ks = 1; ks < 8; ++ks)
            umma_bf16_acc(tacc, a_base + (uint64_t)((2 * ks * WG_SLAB_W * 128) >> 4), b_base + (uint64_t)((ks * 2048) >> 4), IDESC);
        }
        umma_commit(bar_empty(stage));
        first = 0;
        if (++stage == WK_STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(bar_acc);
      if (p.dbg) {
        const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.dbg[cta * 8 + 0] = clock64() - t_start;
        p.dbg[cta * 8 + 1] = t_wait;
      }
    }
    __syncwarp();
  } else {
    // drain: TMEM [lane = half*64 + ci][pr*192 + kd*64 + co] -> partial[split][kd*9 + kh*3 + kw][ci][co], kw = 2*pr + half
    const bool any = b_begin < b_end;
    const long long t_d0 = clock64();
    mbar_wait(bar_acc, 0);
    const long long t_d1 = clock64();
    tc_fence_after();
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ci = ci_blk * 64 + (m & 63);
    for (int pr = 0; pr < WK_NACC; ++pr) {
      const int kw = 2 * pr + (m >> 6);
      for (int kd = 0; kd < 3; ++kd) {
        uint32_t r[64];
        if (any) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + pr * WK_ACC_COLS + kd * 64;
          tmem_ld32(taddr, r);
          tmem_ld32(taddr + 32, r + 32);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j) r[j] = 0u;
        }
        if (kw <= 2) {      // warp-uniform: m >> 6 is the same for the 32 lanes of a quadrant
          // A thread holds one 256-byte row (its ci): written straight to global memory, every store instruction would touch
          // 32 different lines.  The 32 x 256-byte tile of the warp goes through a private staging tile instead (the pipeline
          // stages are dead once bar_acc has completed; 16-byte chunks XOR-swizzled by row: conflict-free both ways) and
          // leaves as 512 contiguous bytes per instruction.
          const int tap = kd * 9 + kh * 3 + kw;
          uint8_t* stg = sgen + (warp - 2) * 8192;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 256 + ((j ^ (lane & 15)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
          float* tile = p.partial + (((size_t)split * g.taps + tap) * g.Cin + (ci - lane)) * g.Cout + co_blk * 64;   // row 0 = lane 0's ci
#pragma unroll
          for (int it = 0; it < 16; ++it) {
            const int c = it * 32 + lane, row = c >> 4, col = c & 15;
            const uint4 v = *reinterpret_cast<const uint4*>(stg + row * 256 + ((col ^ (row & 15)) << 4));
            *reinterpret_cast<uint4*>(tile + (size_t)row * g.Cout + col * 4) = v;
          }
          __syncwarp();     // the next tile reuses the staging area
        }
      }
    }
    if (p.dbg && threadIdx.x == 64) {      // per-CTA phase clocks of the drain warps (development aid)
      const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      p.dbg[cta * 8 + 2] = t_d1 - t_d0;          // waiting for the accumulators
      p.dbg[cta * 8 + 3] = clock64() - t_d1;     // drain (TMEM -> staging -> global partial)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<WG_TMEM_COLS>(tmem_base);
  if (p.dw != nullptr) {
    // ===================== split-K reduction inside the launch =====================
    // every CTA of the grid is resident (at most one per SM, cooperative launch): publish the partials, cross the grid barrier,
    // then sum a contiguous slice of the [taps][Cin][Cout] result over the splits in split order (deterministic) and write it in
    // PyTorch's [Cout][Cin][taps] layout.  110 592 outputs over ~141 CTAs: ~1 float4 per thread, 47 L2-resident loads each.
    __threadfence();
    __syncthreads();
    const unsigned nctas = gridDim.x * gridDim.y * gridDim.z;
    if (threadIdx.x == 0) grid_barrier_arrive_and_wait(p.counter, nctas);
    __syncthreads();
    const long long total4 = (long long)g.taps * g.Cin * g.Cout / 4;
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const long long per = (total4 + nctas - 1) / nctas;
    const long long i_end = min(total4, (long long)(cta + 1) * per);
    const float4* part4 = reinterpret_cast<const float4*>(p.partial);
    for (long long i = (long long)cta * per + threadIdx.x; i < i_end; i += WG_THREADS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int k = 0; k < p.splits; ++k) {
        const float4 v = __ldcg(part4 + (size_t)k * total4 + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      const long long e = i * 4;
      const int co = (int)(e % g.Cout);
      const int ci = (int)((e / g.Cout) % g.Cin);
      const int tap = (int)(e / ((long long)g.Cout * g.Cin));
      float* o = p.dw + ((size_t)co * g.Cin + ci) * g.taps + tap;
      const size_t stride = (size_t)g.Cin * g.taps;
      o[0] = acc.x; o[stride] = acc.y; o[2 * stride] = acc.z; o[3 * stride] = acc.w;
    }
  }
}

// dw[co][ci][tap] = sum_s partial[s][tap][ci][co]: 4 consecutive co per thread (128-bit loads); the splits of one output quad are
// walked by WR_PARTS threads (at most 3 loads each for 47 splits: the loads of a block are all in flight at once) and combined
// through shared memory in a fixed order (deterministic).  ncu on the 4-part form: 9.2 us for 20.8 MB, 0.36 waves.
constexpr int WR_PARTS = 16, WR_QUADS = 16;
__global__ void __launch_bounds__(WR_PARTS * WR_QUADS) wgrad_reduce_kernel(const float4* __restrict__ partial, float* __restrict__ dw,
                                                                           int splits, int taps, int Cin, int Cout) {
  pdl_enter();
  __shared__ float4 red[WR_PARTS][WR_QUADS];
  const long long total4 = (long long)taps * Cin * Cout / 4;
  const int q = threadIdx.x % WR_QUADS, part = threadIdx.x / WR_QUADS;
  const long long i = (long long)blockIdx.x * WR_QUADS + q;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < total4) {
#pragma unroll 4
    for (int k = part; k < splits; k += WR_PARTS) {
      const float4 v = __ldcg(partial + (size_t)k * total4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  red[part][q] = s;
  __syncthreads();
  if (part != 0 || i >= total4) return;
#pragma unroll
  for (int k = 1; k < WR_PARTS; ++k) {
    const float4 v = red[k][q];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const long long e = i * 4;
  const int co = (int)(e % Cout);
  const int ci = (int)((e / Cout) % Cin);
  const int tap = (int)(e / ((long long)Cout * Cin));
  float* o = dw + ((size_t)co * Cin + ci) * taps + tap;
  const size_t stride = (size_t)Cin * taps;
  o[0] = s.x; o[stride] = s.y; o[2 * stride] = s.z; o[3 * stride] = s.w;
}

bool wgrad_tc_supported(int x_fmt, int gy_fmt, const ConvGeom& g) {
  return x_fmt == HPVG_FMT_NDHWC_BF16 && gy_fmt == HPVG_FMT_NDHWC_BF16 && g.Cin % 64 == 0 && g.Cout % 64 == 0 && g.Cin >= 64 &&
         g.Cout >= 64;
}

static void wgrad_plan(const ConvGeom& g, int& bricks_h, int& bricks_w, long long& num_bricks, int& splits, long long& per_split) {
  bricks_h = (int)cdiv(g.Ho, WG_BH);
  bricks_w = (int)cdiv(g.Wo, WG_BW);
  num_bricks = (long long)g.N * g.Do * bricks_h * bricks_w;
  const int others = g.KD * (g.Cin / 64) * (g.Cout / 64);
  long long want = max(1LL, (long long)num_sms() / others);
  want = min(want, num_bricks);
  per_split = cdiv(num_bricks, want);
  splits = (int)cdiv(num_bricks, per_split);
}

// kd-stacked form: work items are input slabs (N * Di * bricks), grid.z = kh
static bool wgrad_kdstack_ok(const ConvGeom& g) { return g.KD == 3 && g.taps == 27; }
static void wgrad_plan_kdstack(const ConvGeom& g, int& bricks_h, int& bricks_w, long long& num_items, int& splits, long long& per_split) {
  bricks_h = (int)cdiv(g.Ho, WG_BH);
  bricks_w = (int)cdiv(g.Wo, WG_BW);
  num_items = (long long)g.N * g.Di * bricks_h * bricks_w;
  const int others = 3 * (g.Cin / 64) * (g.Cout / 64);
  long long want = max(1LL, (long long)num_sms() / others);
  want = min(want, num_items);
  per_split = cdiv(num_items, want);
  splits = (int)cdiv(num_items, per_split);
}

// large enough for either form, so that the mode can change between the workspace query and the call
size_t wgrad_tc_workspace(const ConvGeom& g) {
  int bh, bw, splits;
  long long nb, per;
  wgrad_plan(g, bh, bw, nb, splits, per);
  if (wgrad_kdstack_ok(g)) {
    int s2;
    wgrad_plan_kdstack(g, bh, bw, nb, s2, per);
    splits = max(splits, s2);
  }
  return (size_t)splits * g.taps * g.Cin * g.Cout * sizeof(float) + 256;      // + the grid barrier's counter (fused reduction)
}

int wgrad_tc(const void* x, const void* gy, float* dw, const ConvGeom& g, void* workspace, size_t ws_bytes, cudaStream_t st) {
  WgParams p;
  p.g = g;
  const bool stacked = wgrad_mode() == 1 && wgrad_kdstack_ok(g);
  if (stacked)
    wgrad_plan_kdstack(g, p.bricks_h, p.bricks_w, p.num_bricks, p.splits, p.bricks_per_split);
  else
    wgrad_plan(g, p.bricks_h, p.bricks_w, p.num_bricks, p.splits, p.bricks_per_split);
  if (p.num_bricks >= (1LL << 31)) {
    set_error("wgrad_tc: %lld work items exceed the 32-bit brick index of the kernel", p.num_bricks);
    return -1;
  }
  const size_t need = (size_t)p.splits * g.taps * g.Cin * g.Cout * sizeof(float);
  if (ws_bytes < need || workspace == nullptr) {
    set_error("wgrad_tc: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return -1;
  }
  p.partial = reinterpret_cast<float*>(workspace);
  p.dbg = debug_clock_buffer();
  p.dw = nullptr;
  p.counter = nullptr;
  CUtensorMap mx, mg;
  {
    uint64_t dims[5] = {(uint64_t)g.Cin, (uint64_t)g.Wi, (uint64_t)g.Hi, (uint64_t)g.Di, (uint64_t)g.N};
    uint32_t box[5] = {64, WG_SLAB_W, WG_SLAB_H, 1, 1};
    if (int rc = make_tmap_bf16(&mx, x, 5, dims, box)) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)g.Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)g.N};
    uint32_t box[5] = {64, WG_BW, WG_BH, 1, 1};
    if (int rc = make_tmap_bf16(&mg, gy, 5, dims, box)) return rc;
  }
  static std::atomic<unsigned long long> attr_mask{0};
  if (attr_pending(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wgrad_tc_kdstack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WK_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("wgrad_tc: cannot opt in to %d bytes of shared memory: %s", max(WG_SMEM_BYTES, WK_SMEM_BYTES), cudaGetErrorString(e));
      return -2;
    }
    attr_set(attr_mask);
  }
  if (stacked) {
    dim3 grid((unsigned)p.splits, (unsigned)((g.Cin / 64) * (g.Cout / 64)), 3u);
    // OFF by default (HPVG_WGRAD_FUSED_REDUCE=1 turns it on).  Measured on B200 inside the recorded iteration: the weight
    // gradients run on side streams next to the data-gradient chain, and a kernel whose CTAs spin on a grid barrier holds
    // every SM it got until its LAST CTA has found a free SM — 26.0 us per call against 17.8 + 9.0 us for the two launches,
    // and the iteration went from 4.37 to 4.78 ms.  Alone on the GPU the one-launch form saves the second launch.
    static const bool fused_reduce = getenv("HPVG_WGRAD_FUSED_REDUCE") && atoi(getenv("HPVG_WGRAD_FUSED_REDUCE")) != 0;
    const size_t counter_off = (need + 127) & ~(size_t)127;
    if (fused_reduce && (long long)grid.x * grid.y * grid.z <= num_sms() && ws_bytes >= counter_off + 8) {
      // one launch: the kernel reduces its own partials after a grid barrier (cooperative: the whole grid is resident)
      p.dw = dw;
      p.counter = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(workspace) + counter_off);
      if (cudaMemsetAsync(p.counter, 0, 8, st) != cudaSuccess) {
        set_error("wgrad_tc: cannot clear the barrier counter: %s", cudaGetErrorString(cudaGetLastError()));
        return -2;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = grid;
      cfg.blockDim = dim3(WG_THREADS);
      cfg.dynamicSmemBytes = WK_SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeCooperative;
      attr[0].val.cooperative = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, wgrad_tc_kdstack_kernel, mx, mg, p);
      HPVG_CHECK_LAUNCH("wgrad_tc_kdstack_kernel (fused reduction)");
      return 0;
    }
    launch_k(wgrad_tc_kdstack_kernel, grid, WG_THREADS, WK_SMEM_BYTES, st, mx, mg, p);
    HPVG_CHECK_LAUNCH("wgrad_tc_kdstack_kernel");
  } else {
    dim3 grid((unsigned)p.splits, (unsigned)((g.Cin / 64) * (g.Cout / 64)), (unsigned)g.KD);
    if (wgrad_mode() == 2)
      launch_k(wgrad_tc_kernel<true>, grid, WG_THREADS, WG_SMEM_BYTES, st, mx, mg, p);
    else
      launch_k(wgrad_tc_kernel<false>, grid, WG_THREADS, WG_SMEM_BYTES, st, mx, mg, p);
    HPVG_CHECK_LAUNCH("wgrad_tc_kernel");
  }
  const long long total4 = (long long)g.taps * g.Cin * g.Cout / 4;
  launch_k(wgrad_reduce_kernel, (unsigned)cdiv(total4, WR_QUADS), WR_PARTS * WR_QUADS, 0, st, reinterpret_cast<const float4*>(p.partial), dw, p.splits, g.taps,
                                                                 g.Cin, g.Cout);
  HPVG_CHECK_LAUNCH("wgrad_reduce_kernel");
  return 0;
}

}  // namespace hpvg

// tcgen05 implicit-GEMM 3x3x3 convolution, "column streaming" form, for the 64-channel layers at volumes that give the
// brick kernel (conv_tc.cu) at most a unit or two per SM (BASELINE config 2: 16 x 64 x 64 = 128 units on 148 SMs).
//
// Replaces the same ATen calls as conv_tc.cu: aten::convolution / convolution_backward(grad_input) behind the 64 -> 64
// Conv3d of ConvBlock3D / ConvBlock3DSN (modules/networks_3d.py:48-70).
//
// Why a second kernel.  With one unit per CTA the three phases of conv_tc do not overlap inside a launch (measured, per-CTA
// clocks: ~4 k cycles waiting for six input slabs, ~16 k of MMA issue, ~7 k of epilogue and store drain), and 128 units leave
// 20 SMs idle.  Here
//  * a CTA owns 32 OUTPUT CHANNELS of a run of consecutive output d-slices of one 16 x 8 brick column.  All 27 taps of those
//    32 channels (27 x 32 x 64 bf16 = 108 KB) are loaded into shared memory ONCE per CTA and stay resident, so the loop order
//    can be slab-outer: input slab j (one d-slice of the 18 x 10 halo, 23 KB, a 4-deep TMA ring) is multiplied with all nine
//    (kh, kw) positions, each as one N = 96 MMA against the kd = 0, 1, 2 tiles, which lands in the three accumulators that
//    slice feeds (TMEM columns in descending slice order, as in conv_tc's kd stacking);
//  * accumulator a is therefore complete as soon as slab a + 2 is done: its epilogue (TMEM -> bias / LeakyReLU / LeakyReLU'
//    mask -> bf16 -> 64-byte-swizzled staging -> TMA store, BatchNorm sums) runs while the MMAs of the following slabs issue,
//    and only the last accumulator's epilogue is exposed;
//  * the first MMA needs one slab and one 12 KB weight group, not six slabs;
//  * work is dealt out in 128-voxel x 32-channel tiles: config 2 has 1 024 of them, 6.9 per SM on all 148 SMs.
// Cost: the A operand (128 x 16 bf16 = 4 KB per MMA at 128 B/clk) is amortised over N = 96 instead of 192, so a full slab
// position takes ~56 clk instead of the 48 of the tensor pipe, every slab is read from L2 by two CTAs, and a run of 7 tiles
// needs 9-11 slabs (2 halo slabs per segment) where the brick kernel's 4-slice unit needs 6.
//
// MEASURED on B200 (64 -> 64 at 16 x 64 x 64, experiments/bench_kernels.py): the MMA phase grows from 15.9 k to 20.9 k
// cycles per CTA, which eats what the overlapped epilogue (exposed: 1.4 k instead of 7 k cycles) and the 148-SM balance
// give back: 24.6 us per launch against 22.6 us for conv_tc (cold L2), 18.0 against 16.7 us (graph of 10, warm L2).  The
// kernel is therefore OFF by default (hpvg_set_conv_col_mode / HPVG_TC_COL=1 turn it on; tests run both kernels against the
// CUDA-core kernels).  The structure — resident weights, slab ring, progressive accumulators — is what a cta_group::2
// version would need to keep N = 192 per instruction with half of B in each CTA of the pair.
#include "common.cuh"
#include <cstdlib>

namespace hpvg {

namespace col {
constexpr int BH = 16, BW = 8;
constexpr int SLAB_H = BH + 2, SLAB_W = BW + 2;
constexpr int SLAB_BYTES = SLAB_H * SLAB_W * 128;     // 23040
constexpr int NC = 32;                                // output channels per column
constexpr int WTILE = NC * 128;                       // one tap: 32 rows (output channels) x 64 input channels bf16
constexpr int WQ = 3 * WTILE;                         // the kd = 0, 1, 2 tiles of one (kh, kw) position
constexpr int W_BYTES = 9 * WQ;                       // 110592
constexpr int NSLOT = 4;                              // input slab ring
constexpr int SEG = 8;                                // accumulators (output slices) per segment = per TMEM half
constexpr int STG_BYTES = 128 * NC * 2;               // one output tile: 128 voxels x 32 channels bf16
constexpr int NSTG = 2;
constexpr int NEPI = 4;                               // epilogue warps, one per TMEM lane quadrant
constexpr int NEPI_THREADS = 32 * NEPI;
constexpr int THREADS = 64 + NEPI_THREADS;
constexpr int TMEM_COLS = 512;
constexpr int MAX_COUT = 256;
constexpr int OFF_W = 0;
constexpr int OFF_SLAB = OFF_W + W_BYTES;
constexpr int OFF_STG = OFF_SLAB + NSLOT * SLAB_BYTES;
constexpr int OFF_BIAS = OFF_STG + NSTG * STG_BYTES;
constexpr int OFF_BAR = OFF_BIAS + MAX_COUT * 4;
// barriers: 9 weight groups, NSLOT full + NSLOT empty, 2 * SEG accumulator-full, 2 half-empty, 1 weights-free
constexpr int NBARS = 9 + 2 * NSLOT + 2 * SEG + 2 + 1;
constexpr int SMEM_BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;
static_assert(OFF_SLAB % 1024 == 0 && OFF_STG % 1024 == 0, "swizzle atoms need 1024-byte aligned regions");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
}  // namespace col

struct ColParams {
  ConvGeom g;
  int units_h, units_w, ncb;   // brick grid and 32-wide output-channel blocks
  long long tiles;             // ncb * N * units_h * units_w * Do
  int act;
  float slope;
  const float* bias;
  float* stats;
  const __nv_bfloat16* mask_src;
  long long* dbg;
};

struct ColSeg {
  int cb, n, h0, w0, od0, cnt;
};

// next run of at most SEG output slices of one (channel block, brick) column inside [t, t1)
__device__ __forceinline__ bool col_next_seg(long long& t, long long t1, const ColParams& p, ColSeg& s) {
  if (t >= t1) return false;
  const int Do = p.g.Do;
  long long c = t / Do;
  s.od0 = (int)(t - c * Do);
  long long left = t1 - t;
  int cnt = Do - s.od0;
  if (cnt > col::SEG) cnt = col::SEG;
  if ((long long)cnt > left) cnt = (int)left;
  s.cnt = cnt;
  s.w0 = (int)(c % p.units_w) * col::BW;
  c /= p.units_w;
  s.h0 = (int)(c % p.units_h) * col::BH;
  c /= p.units_h;
  s.n = (int)(c % p.g.N);
  s.cb = (int)(c / p.g.N);
  t += cnt;
  return true;
}

__global__ void __launch_bounds__(col::THREADS, 1)
conv_col_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_y, const ColParams p) {
  using namespace col;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t s_w = sbase + OFF_W, s_slab = sbase + OFF_SLAB, s_stg = sbase + OFF_STG, s_bar = sbase + OFF_BAR;
  auto bar_w = [&](int q) { return s_bar + 8u * q; };
  auto bar_sfull = [&](int i) { return s_bar + 8u * (9 + i); };
  auto bar_sempty = [&](int i) { return s_bar + 8u * (9 + NSLOT + i); };
  auto bar_accfull = [&](int i) { return s_bar + 8u * (9 + 2 * NSLOT + i); };
  auto bar_hempty = [&](int h) { return s_bar + 8u * (9 + 2 * NSLOT + 2 * SEG + h); };
  const uint32_t bar_wfree = s_bar + 8u * (9 + 2 * NSLOT + 2 * SEG + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + OFF_BAR + NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;

  if (threadIdx.x == 0) {
    for (int q = 0; q < 9; ++q) mbar_init(bar_w(q), 1);
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(bar_sfull(i), 1);
      mbar_init(bar_sempty(i), 1);
    }
    for (int i = 0; i < 2 * SEG; ++i) mbar_init(bar_accfull(i), 1);
    mbar_init(bar_hempty(0), NEPI_THREADS);
    mbar_init(bar_hempty(1), NEPI_THREADS);
    mbar_init(bar_wfree, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
  }
  pdl_trigger();
  pdl_wait();
  __syncthreads();                       // barriers initialised and visible to every warp
  // the producer warp starts its TMA loads at once; TMEM allocation and the bias fetch of the other warps overlap their flight
  float* bias_s = reinterpret_cast<float*>(sgen + OFF_BIAS);
  uint32_t tmem_base = 0;
  if (warp != 0) {
    if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
    for (int i = threadIdx.x - 32; i < MAX_COUT; i += THREADS - 32) bias_s[i] = (p.bias && i < g.Cout) ? p.bias[i] : 0.f;
    tc_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(THREADS - 32) : "memory");
    tc_fence_after();
    tmem_base = *tmem_slot;
  }

  // this CTA's run of tiles
  const long long t_begin = p.tiles * blockIdx.x / gridDim.x;
  const long long t_end = p.tiles * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t slot = 0, sphase = 0, wloads = 0;
      int cur_cb = -1;
      long long t = t_begin;
      ColSeg s;
      auto load_wq = [&](int q) {
        mbar_expect_tx(bar_w(q), WQ);
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)
          tma_load_2d(s_w + q * WQ + kd * WTILE, &tmap_w, bar_w(q), 0, (kd * 9 + q) * g.Cout + cur_cb * NC);
      };
      while (col_next_seg(t, t_end, p, s)) {
        const bool loadw = s.cb != cur_cb;
        if (loadw) {
          // the resident weights are replaced: every MMA that reads the old ones must have completed
          if (cur_cb >= 0) mbar_wait(bar_wfree, (wloads - 1u) & 1u);
          cur_cb = s.cb;
          load_wq(0);
        }
        bool rest = loadw;
        for (int j = 0; j < s.cnt + 2; ++j) {
          const int d = s.od0 + j - g.pad_d;
          if (d < 0 || d >= g.Di) continue;
          mbar_wait(bar_sempty(slot), sphase ^ 1u);
          mbar_expect_tx(bar_sfull(slot), SLAB_BYTES);
          tma_load_5d(s_slab + slot * SLAB_BYTES, &tmap_x, bar_sfull(slot), 0, s.w0 - g.pad, s.h0 - g.pad, d, s.n);
          if (++slot == NSLOT) { slot = 0; sphase ^= 1u; }
          if (rest) {
            // the first MMAs need slab 0 and group 0 only: the other 96 KB of weights queue behind them
#pragma unroll 1
            for (int q = 1; q < 9; ++q) load_wq(q);
            rest = false;
          }
        }
        if (rest) {
#pragma unroll 1
          for (int q = 1; q < 9; ++q) load_wq(q);
        }
        if (loadw) ++wloads;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one elected thread =====================
    if (elect_one()) {
      constexpr uint32_t IDESC1 = umma_idesc_bf16(128, NC, 0, 0);
      const uint64_t a_base = umma_desc(s_slab, 16, SLAB_W * 128, 2);
      const uint64_t b_base = umma_desc(s_w, 16, 1024, 2);
      uint32_t slot = 0, sphase = 0, wl = 0;
      int cur_cb = -1, si = 0;
      long long t = t_begin;
      ColSeg s;
      long long t_start = clock64(), t_wwait = 0, t_swait = 0;
      while (col_next_seg(t, t_end, p, s)) {
        const int half = si & 1;
        if (si >= 2) {
          mbar_wait(bar_hempty(half), (uint32_t)(((si >> 1) - 1) & 1));
          tc_fence_after();
        }
        bool need_w = s.cb != cur_cb;
        cur_cb = s.cb;
        uint32_t touched = 0;
        const uint32_t tcol0 = tmem_base + (uint32_t)(half * SEG * NC);
        for (int j = 0; j < s.cnt + 2; ++j) {
          const int d = s.od0 + j - g.pad_d;
          if (d >= 0 && d < g.Di) {
            const int a_hi = j < s.cnt - 1 ? j : s.cnt - 1;
            const int a_lo = j - 2 > 0 ? j - 2 : 0;
            const int nk = a_hi - a_lo + 1;          // accumulators this slab feeds (1..3)
            const int klo = j - a_hi;                // their kd range starts here
            if (p.dbg) t_swait -= clock64();
            mbar_wait(bar_sfull(slot), sphase);
            if (p.dbg) t_swait += clock64();
            tc_fence_after();
            const uint64_t ad0 = a_base + (uint64_t)((slot * SLAB_BYTES) >> 4);
            const uint64_t bd0 = b_base + (uint64_t)((klo * WTILE) >> 4);
            const uint32_t tacc = tcol0 + (uint32_t)((SEG - 1 - a_hi) * NC);
            const uint32_t idesc = umma_idesc_bf16(128, NC * nk, 0, 0);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
              const int kh = q / 3, kw = q % 3;
              if (need_w) {
                if (p.dbg) t_wwait -= clock64();
                mbar_wait(bar_w(q), wl & 1u);
                if (p.dbg) t_wwait += clock64();
                tc_fence_after();
              }
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = ad0 + (uint64_t)(((kh * SLAB_W + kw) * 128 + ks * 32) >> 4);
                if (q == 0 && ks == 0) {
                  // first MMA of the slab: one per accumulator, so that each gets its own overwrite / accumulate flag
#pragma unroll
                  for (int kd = 0; kd < 3; ++kd) {
                    if (kd < klo || kd >= klo + nk) continue;
                    const int a = j - kd;
                    umma_bf16(tcol0 + (uint32_t)((SEG - 1 - a) * NC), ad, b_base + (uint64_t)((kd * WTILE) >> 4), IDESC1,
                              (touched >> a) & 1u);
                    touched |= 1u << a;
                  }
                } else {
                  umma_bf16_acc(tacc, ad, bd0 + (uint64_t)((q * WQ + ks * 32) >> 4), idesc);
                }
              }
            }
            if (need_w) {
              need_w = false;
              ++wl;
            }
            umma_commit(bar_sempty(slot));
            if (++slot == NSLOT) { slot = 0; sphase ^= 1u; }
          }
          // output slice j - 2 has now seen its three input slices (or they lie outside the volume)
          if (j >= 2) umma_commit(bar_accfull(half * SEG + j - 2));
        }
        {
          long long tt = t;
          ColSeg nx;
          if (col_next_seg(tt, t_end, p, nx) && nx.cb != s.cb) umma_commit(bar_wfree);
        }
        ++si;
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 8 + 0] = clock64() - t_start;   // MMA issue loop
        p.dbg[blockIdx.x * 8 + 1] = t_wwait;               // waiting for weight groups
        p.dbg[blockIdx.x * 8 + 2] = t_swait;               // waiting for input slabs
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: warp = TMEM lane quadrant, thread = voxel row, 32 channels each =====================
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;                       // accumulator row = brick voxel (hh = m / 8, ww = m % 8)
    const int et = threadIdx.x - 64;
    uint32_t full_phase = 0;
    int si = 0, stg = 0;
    long long t = t_begin;
    ColSeg s;
    long long t_start = clock64(), t_accwait = 0;
    while (col_next_seg(t, t_end, p, s)) {
      const int half = si & 1;
      const int oh = s.h0 + (m >> 3), ow = s.w0 + (m & 7);
      const bool row_ok = (oh < g.Ho) && (ow < g.Wo);
      float st_s = 0.f, st_s2 = 0.f;                    // BatchNorm sums: channel et & 31, row group et >> 5
#pragma unroll 1
      for (int a = 0; a < s.cnt; ++a) {
        const int od = s.od0 + a;
        const int bi = half * SEG + a;
        long long tq = clock64();
        mbar_wait(bar_accfull(bi), (full_phase >> bi) & 1u);
        full_phase ^= 1u << bi;
        t_accwait += clock64() - tq;
        tc_fence_after();
        uint32_t r[NC];
        tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(half * SEG * NC + (SEG - 1 - a) * NC), r);
        tmem_ld_wait();
        float v[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + s.cb * NC);
#pragma unroll
          for (int j = 0; j < NC / 4; ++j) {
            const float4 bq = b4[j];
            v[4 * j + 0] += bq.x; v[4 * j + 1] += bq.y; v[4 * j + 2] += bq.z; v[4 * j + 3] += bq.w;
          }
        }
        if (p.mask_src && row_ok) {
          const uint4* mp = reinterpret_cast<const uint4*>(
              p.mask_src + ((((size_t)s.n * g.Do + od) * g.Ho + oh) * g.Wo + ow) * g.Cout + s.cb * NC);
#pragma unroll
          for (int c = 0; c < NC / 8; ++c) {
            const uint4 mv = __ldg(mp + c);
            float2 f;
            f = unpack_bf16x2(mv.x); v[8 * c + 0] *= f.x > 0.f ? 1.f : p.slope; v[8 * c + 1] *= f.y > 0.f ? 1.f : p.slope;
            f = unpack_bf16x2(mv.y); v[8 * c + 2] *= f.x > 0.f ? 1.f : p.slope; v[8 * c + 3] *= f.y > 0.f ? 1.f : p.slope;
            f = unpack_bf16x2(mv.z); v[8 * c + 4] *= f.x > 0.f ? 1.f : p.slope; v[8 * c + 5] *= f.y > 0.f ? 1.f : p.slope;
            f = unpack_bf16x2(mv.w); v[8 * c + 6] *= f.x > 0.f ? 1.f : p.slope; v[8 * c + 7] *= f.y > 0.f ? 1.f : p.slope;
          }
        }
        if (p.act == HPVG_ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < NC; ++j) v[j] = fmaxf(v[j], v[j] * p.slope);
        }
        if (!row_ok) {
#pragma unroll
          for (int j = 0; j < NC; ++j) v[j] = 0.f;
        }
        // staging buffer free: the thread that issued its last TMA store waits until that store has read it
        if (et == 0) tma_store_wait_read<NSTG - 1>();
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");
        // 64-byte rows, 64-byte swizzle (16-byte chunk index ^= address bits 7..8; the buffer is 1024-byte aligned)
        const uint32_t sdst = s_stg + stg * STG_BYTES + m * (NC * 2);
        const uint32_t sphase = (uint32_t)(m >> 1) & 3u;
#pragma unroll
        for (int c = 0; c < NC / 8; ++c) {
          const uint32_t addr = sdst + (((uint32_t)c ^ sphase) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[8 * c + 0], v[8 * c + 1])),
                       "r"(pack_bf16x2(v[8 * c + 2], v[8 * c + 3])), "r"(pack_bf16x2(v[8 * c + 4], v[8 * c + 5])),
                       "r"(pack_bf16x2(v[8 * c + 6], v[8 * c + 7]))
                       : "memory");
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(NEPI_THREADS) : "memory");
        if (et == 0) {
          tma_store_5d(&tmap_y, s_stg + stg * STG_BYTES, s.cb * NC, s.w0, s.h0, od, s.n);
          tma_store_commit();
        }
        if (p.stats) {
          // column sums over the staged (bf16-rounded, invalid rows zeroed) tile: thread = channel, 32 rows each
          const int c = et & 31, rg = et >> 5;
          const uint8_t* tile = sgen + OFF_STG + stg * STG_BYTES;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const int row = rg * 32 + rr;
            const __nv_bfloat16 bv =
                *reinterpret_cast<const __nv_bfloat16*>(tile + row * (NC * 2) + (((c >> 3) ^ ((row >> 1) & 3)) << 4) + (c & 7) * 2);
            const float f = bf2f(bv);
            st_s += f;
            st_s2 = fmaf(f, f, st_s2);
          }
        }
        stg ^= 1;
      }
      if (p.stats) {
        float* sp = p.stats + (size_t)s.n * g.stats_stride;
        atomicAdd(sp + s.cb * NC + (et & 31), st_s);
        atomicAdd(sp + g.Cout + s.cb * NC + (et & 31), st_s2);
      }
      tc_fence_before();
      mbar_arrive(bar_hempty(half));
      ++si;
    }
#if HPVG_STORE_WAIT_READ
    if (et == 0) tma_store_wait_read<0>();      // the staging tiles have been read: the CTA may leave (the writes land by the end of the grid)
#else
    if (et == 0) tma_store_wait_all<0>();
#endif
    if (p.dbg && et == 0) {
      p.dbg[blockIdx.x * 8 + 3] = clock64() - t_start;   // epilogue warps, whole run
      p.dbg[blockIdx.x * 8 + 4] = t_accwait;             // of which waiting for accumulators
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

bool conv_col_supported(int x_fmt, int y_fmt, const ConvGeom& g, const void* w_packed) {
  return x_fmt == HPVG_FMT_NDHWC_BF16 && y_fmt == HPVG_FMT_NDHWC_BF16 && w_packed != nullptr && g.KD == 3 && g.Cin == 64 &&
         g.Cout % 64 == 0 && g.Cout >= 64 && g.Cout <= col::MAX_COUT && g.Wi <= 65535 && g.Hi <= 65535;
}

// brick-kernel units of this geometry (4 d-slices x 16 x 8 voxels x 64 channels): the dispatcher's size criterion
long long conv_col_brick_units(const ConvGeom& g) {
  return (long long)(g.Cout / 64) * g.N * cdiv(g.Do, 4) * cdiv(g.Ho, col::BH) * cdiv(g.Wo, col::BW);
}

int conv_col(const void* x, const void* w_packed, const float* bias, void* y, const ConvGeom& g, int act, float slope, float* stats,
             const void* mask_src, cudaStream_t st) {
  using namespace col;
  static std::atomic<unsigned long long> attr_mask{0};
  if (attr_pending(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(conv_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("conv_col: cannot opt in to %d bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
      return -2;
    }
    attr_set(attr_mask);
  }
  CUtensorMap mx, mw, my;
  {
    uint64_t dims[5] = {(uint64_t)g.Cin, (uint64_t)g.Wi, (uint64_t)g.Hi, (uint64_t)g.Di, (uint64_t)g.N};
    uint32_t box[5] = {64, SLAB_W, SLAB_H, 1, 1};
    if (int rc = make_tmap_bf16(&mx, x, 5, dims, box)) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)g.Cin, (uint64_t)g.taps * g.Cout};
    uint32_t box[2] = {64, (uint32_t)NC};
    if (int rc = make_tmap_bf16(&mw, w_packed, 2, dims, box)) return rc;
  }
  {
    uint64_t dims[5] = {(uint64_t)g.Cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)g.N};
    uint32_t box[5] = {(uint32_t)NC, BW, BH, 1, 1};
    if (int rc = make_tmap_bf16(&my, y, 5, dims, box, 64)) return rc;
  }
  ColParams p;
  p.g = g;
  p.units_h = (int)cdiv(g.Ho, BH);
  p.units_w = (int)cdiv(g.Wo, BW);
  p.ncb = g.Cout / NC;
  p.tiles = (long long)p.ncb * g.N * p.units_h * p.units_w * g.Do;
  p.act = act;
  p.slope = slope;
  p.bias = bias;
  p.stats = stats;
  p.mask_src = reinterpret_cast<const __nv_bfloat16*>(mask_src);
  p.dbg = debug_clock_buffer();
  const int grid = (int)min((long long)num_sms(), p.tiles);
  launch_k(conv_col_kernel, grid, THREADS, SMEM_BYTES, st, mx, mw, my, p);
  HPVG_CHECK_LAUNCH("conv_col_kernel");
  return 0;
}

}  // namespace hpvg

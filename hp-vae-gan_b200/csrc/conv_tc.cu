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

namespace hpvg {

constexpr int BH = 16, BW = 8;                      // output brick rows x columns (M = 128)
constexpr int SLAB_H = BH + 2, SLAB_W = BW + 2;     // 18 x 10 halo
constexpr int SLAB_ROWS = SLAB_H * SLAB_W;          // 180 voxel rows of 128 B
constexpr int SLAB_BYTES = SLAB_ROWS * 128;         // 23040
constexpr int SLAB_STRIDE = 23 * 1024;              // keep every slab 1024-aligned
constexpr int BTILE_BYTES = 64 * 128;               // one tap: 64 output channels x 64 input channels bf16
constexpr int STG_BYTES = 128 * 128;                // one output tile: 128 voxels x 64 channels bf16
constexpr int TC_THREADS = 192;

template <int KCHUNKS, int NACC>
struct TcCfg {
  static constexpr int NSLOTS = NACC + 2;                       // d-slices resident per unit (3-D, pad 1)
  static constexpr int NSLAB = NSLOTS * KCHUNKS;
  static constexpr int NB = (KCHUNKS == 1) ? 4 : 3;             // weight ring depth
  static constexpr int NSTG = (KCHUNKS == 1) ? 2 : 1;           // output staging buffers
  static constexpr int TMEM_COLS = (NACC * 64 <= 32) ? 32 : (NACC * 64 <= 64 ? 64 : (NACC * 64 <= 128 ? 128 : (NACC * 64 <= 256 ? 256 : 512)));
  static constexpr int OFF_SLAB = 0;
  static constexpr int OFF_B = OFF_SLAB + NSLAB * SLAB_STRIDE;
  static constexpr int OFF_STG = OFF_B + NB * BTILE_BYTES;
  static constexpr int OFF_BAR = OFF_STG + NSTG * STG_BYTES;
  static constexpr int NBARS = NSLAB + 2 * NB + 3;
  static constexpr int SMEM_BYTES = OFF_BAR + NBARS * 8 + 16 + 1024;   // + tmem slot + alignment slack
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
};

template <int KCHUNKS, int NACC>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_y, const TcParams p) {
  using Cfg = TcCfg<KCHUNKS, NACC>;
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

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::NSLAB; ++i) mbar_init(bar_slab_full(i), 1);
    for (int i = 0; i < Cfg::NB; ++i) {
      mbar_init(bar_b_full(i), 1);
      mbar_init(bar_b_empty(i), 1);
    }
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_acc_empty, 128);
    mbar_init(bar_slabs_free, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_y);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nslots = NACC + g.KD - 1;   // d-slices a unit touches

  // unit -> coordinates
  auto decode = [&](long long u, int& nb, int& n, int& d0, int& h0, int& w0) {
    w0 = (int)(u % p.units_w) * BW;
    u /= p.units_w;
    h0 = (int)(u % p.units_h) * BH;
    u /= p.units_h;
    d0 = (int)(u % p.units_d) * NACC;
    u /= p.units_d;
    n = (int)(u % g.N);
    nb = (int)(u / g.N);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t bstage = 0, bphase = 0;
      int it = 0;
      for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
        int nb, n, d0, h0, w0;
        decode(u, nb, n, d0, h0, w0);
        if (it > 0) mbar_wait(bar_slabs_free, (uint32_t)((it - 1) & 1));
        for (int j = 0; j < nslots; ++j) {
          const int d = d0 + j - g.pad_d;
          if (d < 0 || d >= g.Di) continue;
          // slice needed only if some valid accumulator reads it
          bool needed = false;
          for (int a = 0; a < NACC; ++a) {
            const int kd = j - a;
            if (kd >= 0 && kd < g.KD && d0 + a < g.Do) needed = true;
          }
          if (!needed) continue;
#pragma unroll
          for (int kc = 0; kc < KCHUNKS; ++kc) {
            const int si = j * KCHUNKS + kc;
            mbar_expect_tx(bar_slab_full(si), SLAB_BYTES);
            tma_load_5d(s_slab + si * SLAB_STRIDE, &tmap_x, bar_slab_full(si), kc * 64, w0 - g.pad, h0 - g.pad, d, n);
          }
        }
        for (int t = 0; t < g.taps; ++t) {
#pragma unroll
          for (int kc = 0; kc < KCHUNKS; ++kc) {
            mbar_wait(bar_b_empty(bstage), bphase ^ 1u);
            mbar_expect_tx(bar_b_full(bstage), BTILE_BYTES);
            tma_load_2d(s_b + bstage * BTILE_BYTES, &tmap_w, bar_b_full(bstage), kc * 64, t * g.Cout + nb * 64);
            if (++bstage == Cfg::NB) { bstage = 0; bphase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t IDESC = umma_idesc_bf16(128, 64, 0, 0);
    uint32_t bstage = 0, bphase = 0;
    int it = 0;
    for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
      int nb, n, d0, h0, w0;
      decode(u, nb, n, d0, h0, w0);
      if (it > 0) {
        mbar_wait(bar_acc_empty, (uint32_t)((it - 1) & 1));
        tc_fence_after();
      }
      uint32_t touched = 0, slab_ready = 0;
      for (int t = 0; t < g.taps; ++t) {
        const int kd = t / 9, kh = (t % 9) / 3, kw = t % 3;
#pragma unroll
        for (int kc = 0; kc < KCHUNKS; ++kc) {
          mbar_wait(bar_b_full(bstage), bphase);
          tc_fence_after();
          const uint32_t b_addr = s_b + bstage * BTILE_BYTES;
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            if (d0 + a >= g.Do) continue;
            const int slot = a + kd;
            const int d = d0 + slot - g.pad_d;
            if (d < 0 || d >= g.Di) continue;
            const int si = slot * KCHUNKS + kc;
            if (!((slab_ready >> si) & 1u)) {
              mbar_wait(bar_slab_full(si), (uint32_t)(it & 1));
              tc_fence_after();
              slab_ready |= 1u << si;
            }
            if (lane == 0) {
              const uint32_t a_addr = s_slab + si * SLAB_STRIDE + (kh * SLAB_W + kw) * 128;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = umma_desc(a_addr + ks * 32, 16, SLAB_W * 128, 2);
                const uint64_t bd = umma_desc(b_addr + ks * 32, 16, 1024, 2);
                umma_bf16(tmem_base + a * 64, ad, bd, IDESC, (uint32_t)(((touched >> a) & 1u) | (ks > 0)));
              }
            }
            touched |= 1u << a;
          }
          __syncwarp();
          if (lane == 0) umma_commit(bar_b_empty(bstage));
          if (++bstage == Cfg::NB) { bstage = 0; bphase ^= 1u; }
        }
      }
      if (lane == 0) {
        umma_commit(bar_acc_full);
        umma_commit(bar_slabs_free);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (128 threads) =====================
    const int q = warp & 3;                  // TMEM lane quadrant this warp may read
    const int m = q * 32 + lane;             // accumulator row = brick voxel (hh = m / 8, ww = m % 8)
    const int et = threadIdx.x - 64;         // 0..127 index inside the epilogue group
    int it = 0;
    int stg = 0;
    for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
      int nb, n, d0, h0, w0;
      decode(u, nb, n, d0, h0, w0);
      mbar_wait(bar_acc_full, (uint32_t)(it & 1));
      tc_fence_after();
      const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
      const bool row_ok = (oh < g.Ho) && (ow < g.Wo);
      for (int a = 0; a < NACC; ++a) {
        const int od = d0 + a;
        if (od >= g.Do) break;
        uint32_t r[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * 64;
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 32, r + 32);
        tmem_ld_wait();
        float v[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] += __ldg(p.bias + nb * 64 + j);
        }
        if (p.mask_src && row_ok) {
          const uint4* mp = reinterpret_cast<const uint4*>(
              p.mask_src + ((((size_t)n * g.Do + od) * g.Ho + oh) * g.Wo + ow) * g.Cout + nb * 64);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 mv = __ldg(mp + c);
            float2 f;
            f = unpack_bf16x2(mv.x); v[8 * c + 0] *= f.x > 0.f ? 1.f : p.slope; v[8 * c + 1] *= f.y > 0.f ? 1.f : p.slope;
            f = unpack_bf16x2(mv.y); v[8 * c + 2] *= f.x > 0.f ? 1.f : p.slope; v[8 * c + 3] *= f.y > 0.f ? 1.f : p.slope;
            f = unpack_bf16x2(mv.z); v[8 * c + 4] *= f.x > 0.f ? 1.f : p.slope; v[8 * c + 5] *= f.y > 0.f ? 1.f : p.slope;
            f = unpack_bf16x2(mv.w); v[8 * c + 6] *= f.x > 0.f ? 1.f : p.slope; v[8 * c + 7] *= f.y > 0.f ? 1.f : p.slope;
          }
        }
        if (p.act == HPVG_ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
        }
        if (!row_ok) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = 0.f;
        }
        // staging buffer must be free: the thread that issued its last TMA store waits for the read to finish
        if (et == 0) tma_store_wait_read<Cfg::NSTG - 1>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint32_t sdst = s_stg + stg * STG_BYTES + m * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t addr = sdst + ((uint32_t)(c ^ (m & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[8 * c + 0], v[8 * c + 1])),
                       "r"(pack_bf16x2(v[8 * c + 2], v[8 * c + 3])), "r"(pack_bf16x2(v[8 * c + 4], v[8 * c + 5])),
                       "r"(pack_bf16x2(v[8 * c + 6], v[8 * c + 7]))
                       : "memory");
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          tma_store_5d(&tmap_y, s_stg + stg * STG_BYTES, nb * 64, w0, h0, od, n);
          tma_store_commit();
        }
        if (p.stats) {
          // column sums over the staged (bf16-rounded, invalid rows zeroed) tile: thread = channel, 64 rows each
          const int c = et & 63, half = et >> 6;
          const uint8_t* tile = sgen + Cfg::OFF_STG + stg * STG_BYTES;
          float s = 0.f, s2 = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 64; ++rr) {
            const int row = half * 64 + rr;
            const __nv_bfloat16 bv =
                *reinterpret_cast<const __nv_bfloat16*>(tile + row * 128 + (((c >> 3) ^ (row & 7)) << 4) + (c & 7) * 2);
            const float f = bf2f(bv);
            s += f;
            s2 = fmaf(f, f, s2);
          }
          atomicAdd(p.stats + nb * 64 + c, s);
          atomicAdd(p.stats + g.Cout + nb * 64 + c, s2);
        }
        stg = (stg + 1 == Cfg::NSTG) ? 0 : stg + 1;
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty);
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int KCHUNKS, int NACC>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<KCHUNKS, NACC>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<KCHUNKS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("conv_tc: cannot opt in to %d bytes of shared memory: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return -2;
    }
    attr_done = true;
  }
  const ConvGeom& g = p.g;
  p.units_d = (int)cdiv(g.Do, NACC);
  p.units_h = (int)cdiv(g.Ho, BH);
  p.units_w = (int)cdiv(g.Wo, BW);
  p.nblocks = g.Cout / 64;
  p.num_units = (long long)p.nblocks * g.N * p.units_d * p.units_h * p.units_w;
  const int grid = (int)min((long long)num_sms(), p.num_units);
  conv_tc_kernel<KCHUNKS, NACC><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(mx, mw, my, p);
  HPVG_CHECK_LAUNCH("conv_tc_kernel");
  return 0;
}

bool conv_tc_supported(int x_fmt, int y_fmt, const ConvGeom& g, const void* w_packed) {
  return x_fmt == HPVG_FMT_NDHWC_BF16 && y_fmt == HPVG_FMT_NDHWC_BF16 && w_packed != nullptr && (g.Cin == 64 || g.Cin == 128) &&
         (g.Cout % 64 == 0) && g.Cout >= 64 && g.Wi <= 65535 && g.Hi <= 65535;
}

int conv_tc(const void* x, const void* w_packed, const float* bias, void* y, const ConvGeom& g, int act, float slope, float* stats,
            const void* mask_src, cudaStream_t st) {
  CUtensorMap mx, mw, my;
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
    if (int rc = make_tmap_bf16(&my, y, 5, dims, box)) return rc;
  }
  TcParams p;
  p.g = g;
  p.act = act;
  p.slope = slope;
  p.bias = bias;
  p.stats = stats;
  p.mask_src = reinterpret_cast<const __nv_bfloat16*>(mask_src);
  if (g.Cin == 64) return launch_tc<1, 4>(mx, mw, my, p, st);
  return launch_tc<2, 2>(mx, mw, my, p, st);
}

}  // namespace hpvg

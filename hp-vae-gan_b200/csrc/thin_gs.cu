// Wide -> thin convolution (64 input channels, <= 4 output channels) as ONE GEMM per input slab + a shift-add gather.
//
// Replaces aten::convolution behind the generator tails nn.Conv3d(64, 3, 3) (modules/networks_3d.py:341,362), the critic tail
// nn.Conv3d(64, 1, 3) (:175) and the data gradient of the 3 -> 64 heads (convolution_backward grad_input of :51 with Cin = 3).
//
//   out[v][o] = sum_tap sum_c x[v + s_tap][c] * B[tap][c][o]
//
// With O <= 4 outputs an implicit-GEMM tile (M = 128 voxels, N = 16) is bound by streaming its A operand: an M = 128, K = 16
// tcgen05.mma costs 64 cycles whatever N is, so the thin layers cost as much as a 64 -> 64 layer (conv_tc_kernel<NOUT = 16>:
// 17-19 us at 16 x 64 x 64 for 0.7 GFLOP).  Here the taps move from the A operand to the N dimension:
//
//   P_d[v'][tap * O + o] = sum_c x_d[v'][c] * B[tap][c][o]        one GEMM per input slab d: M = slab voxels, N = 27 * O, K = 64
//   out[d + pad - kd][v][o] = sum_{kh,kw} P_d[v + (kh, kw)][(kd, kh, kw), o]        summed over the three slabs kd = 0, 1, 2
//
// 8 MMAs (two M = 128 tiles over the 18 x 10 halo slab x four K steps, N = 96) instead of 54 per slab; the partial products go
// TMEM -> shared memory (fp32, one slab at a time), every output voxel gathers its 27 * O values and keeps the sums of the
// output slices in flight in registers.  The weights (27 * O x 64, the packed bf16 image with O rows per tap) are fetched once per
// CTA by one TMA load and stay resident.
// Same arithmetic as the tcgen05 thin kernel it replaces (bf16 operands, fp32 accumulation), different summation order.
//
// STATUS: parity-green (tests/test_gpu_fullsize.py, tests/test_gpu_layers.py with HPVG_THIN_GS=1) and OFF by default.  Measured on
// B200 (experiments/thin_bench.py, 10 dependent launches in a CUDA graph, 64 -> 3 at 16 x 64 x 64): 14.6 us per launch against 12.3 us
// for conv_tc_kernel<NOUT = 16>.  The MMA work fell from 54 to 8 instructions per slab as designed, but the consumer side became the
// bound: 69 KB of partial products written to and 41 KB gathered from shared memory per slab (~1.5-2 k cycles per slab on four
// warps, against 512 cycles of MMA).  Halving it would need the gather to run out of TMEM-adjacent registers (a shuffle network
// over the brick's rows) instead of shared memory.
#include "common.cuh"
#include <cstdlib>

namespace hpvg {

constexpr int GS_BH = 16, GS_BW = 8;
constexpr int GS_SLAB_H = GS_BH + 2, GS_SLAB_W = GS_BW + 2;
constexpr int GS_SLAB_ROWS = GS_SLAB_H * GS_SLAB_W;          // 180 halo voxels
constexpr int GS_SLAB_BYTES = GS_SLAB_ROWS * 128;            // 23040
constexpr int GS_STAGES = 3;
constexpr int GS_NMAX = 112;                                 // 27 taps x 4 outputs rounded up to 16
constexpr int GS_PSTRIDE = GS_NMAX + 1;                      // odd row pitch (words) of the partial-product buffer: conflict-free both ways
constexpr int GS_THREADS = 192;                              // warp 0: TMA, warp 1: MMA, warps 2-5: drain + gather + output
constexpr int GS_OFF_SLAB = 0;
constexpr int GS_OFF_P = GS_STAGES * GS_SLAB_BYTES;          // 69120: an M tile that starts at slab row 128 reads 76 rows past the last stage: into P
constexpr int GS_P_BYTES = GS_SLAB_ROWS * GS_PSTRIDE * 4;    // 81360
constexpr int GS_OFF_B = ((GS_OFF_P + GS_P_BYTES + 1023) / 1024) * 1024;     // weights tile, 1024-byte aligned (128-byte swizzle atoms)
constexpr int GS_B_BYTES = GS_NMAX * 128;
constexpr int GS_OFF_BAR = GS_OFF_B + GS_B_BYTES;
constexpr int GS_NBARS = 2 * GS_STAGES + 5;
constexpr int GS_SMEM_BYTES = GS_OFF_BAR + GS_NBARS * 8 + 16 + 1024;
static_assert(GS_SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct GsParams {
  ConvGeom g;                 // Cin = 64 (channels of x), Cout = O
  int units_d, units_h, units_w;
  long long num_units;
  int ncols;                  // 27 * O (or 9 * O) rounded up to a multiple of 16
  const float* bias;          // [O] or nullptr
  float* y;                   // float32 [N][O][Do][Ho][Wo]
};

template <int R>      // output d-slices per unit
__global__ void __launch_bounds__(GS_THREADS, 1) thin_conv_gs_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                     const __grid_constant__ CUtensorMap tmap_w, const GsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t s_slab = sbase + GS_OFF_SLAB;
  float* s_p = reinterpret_cast<float*>(sgen + GS_OFF_P);
  const uint32_t s_b = sbase + GS_OFF_B;
  const uint32_t s_bar = sbase + GS_OFF_BAR;
  auto bar_full = [&](int i) { return s_bar + 8u * i; };
  auto bar_empty = [&](int i) { return s_bar + 8u * (GS_STAGES + i); };
  auto bar_acc_full = [&](int i) { return s_bar + 8u * (2 * GS_STAGES + i); };
  auto bar_acc_empty = [&](int i) { return s_bar + 8u * (2 * GS_STAGES + 2 + i); };
  const uint32_t bar_w = s_bar + 8u * (2 * GS_STAGES + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + GS_OFF_BAR + GS_NBARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;
  const int O = g.Cout, KD = g.KD, taps = g.taps, NC = p.ncols;
  constexpr int NSLAB = R + 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GS_STAGES; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full(i), 1);
      mbar_init(bar_acc_empty(i), 128);
    }
    mbar_init(bar_w, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  pdl_trigger();
  pdl_wait();
  __syncthreads();
  uint32_t tmem_base = 0;
  if (warp != 0) {
    if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
    tc_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(GS_THREADS - 32) : "memory");
    tc_fence_after();
    tmem_base = *tmem_slot;
  }

  auto decode = [&](long long u, int& n, int& d0, int& h0, int& w0) {
    w0 = (int)(u % p.units_w) * GS_BW;
    u /= p.units_w;
    h0 = (int)(u % p.units_h) * GS_BH;
    u /= p.units_h;
    d0 = (int)(u % p.units_d) * R;
    n = (int)(u / p.units_d);
  };
  // input slab j of a unit = input slice d0 - pad_d + j; needed when some valid output slice a = j - kd (0 <= a < R, d0 + a < Do) reads it
  auto slab_needed = [&](int d0, int j) -> bool {
    const int d = d0 - g.pad_d + j;
    if (d < 0 || d >= g.Di) return false;
    for (int kd = 0; kd < KD; ++kd) {
      const int a = j - kd;
      if (a >= 0 && a < R && d0 + a < g.Do) return true;
    }
    return false;
  };

  if (warp == 0) {
    if (elect_one()) {
      // the weights first: the packed bf16 image [taps * O rows][64 channels] arrives through TMA with the 128-byte swizzle of a
      // K-major operand tile (rows past taps * O are out of bounds: zero fill) and stays resident for the whole kernel
      mbar_expect_tx(bar_w, (uint32_t)NC * 128u);
      tma_load_2d(s_b, &tmap_w, bar_w, 0, 0);
      uint32_t stage = 0, phase = 0;
      for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        int n, d0, h0, w0;
        decode(u, n, d0, h0, w0);
        for (int j = 0; j < NSLAB; ++j) {
          if (j >= R + KD - 1 || !slab_needed(d0, j)) continue;
          mbar_wait(bar_empty(stage), phase ^ 1u);
          mbar_expect_tx(bar_full(stage), GS_SLAB_BYTES);
          tma_load_5d(s_slab + stage * GS_SLAB_BYTES, &tmap_x, bar_full(stage), 0, w0 - g.pad, h0 - g.pad, d0 - g.pad_d + j, n);
          if (++stage == GS_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, NC, 0, 0);
      const uint64_t b_base = umma_desc(s_b, 16, 1024, 2);
      uint32_t stage = 0, phase = 0, buf = 0, acc_phase = 0;
      mbar_wait(bar_w, 0);
      for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        int n, d0, h0, w0;
        decode(u, n, d0, h0, w0);
        for (int j = 0; j < NSLAB; ++j) {
          if (j >= R + KD - 1 || !slab_needed(d0, j)) continue;
          mbar_wait(bar_acc_empty(buf), acc_phase ^ 1u);          // the consumers have drained this TMEM buffer
          mbar_wait(bar_full(stage), phase);
          tc_fence_after();
          const uint64_t a_base = umma_desc(s_slab + stage * GS_SLAB_BYTES, 16, 1024, 2);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint32_t tacc = tmem_base + (buf * 2 + t) * GS_NMAX;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tacc, a_base + (uint64_t)((t * 128 * 128 + ks * 32) >> 4), b_base + (uint64_t)((ks * 32) >> 4), idesc, ks != 0);
          }
          umma_commit(bar_empty(stage));        // the slab may be overwritten once these MMAs have read it
          umma_commit(bar_acc_full(buf));
          if (++stage == GS_STAGES) { stage = 0; phase ^= 1u; }
          buf ^= 1u;
          if (buf == 0) acc_phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== consumers: TMEM -> P (shared, fp32) -> gather into the output slices in flight =====================
    const int t = threadIdx.x - 64;              // 0..127: TMEM lane of the drain AND output voxel of the gather
    const int q = warp & 3;                      // TMEM lane quadrant this warp may read
    const int row_in_tile = q * 32 + lane;
    const int hh = t >> 3, ww = t & 7;
    uint32_t buf = 0, acc_phase = 0;
    const size_t out_sp = (size_t)g.Do * g.Ho * g.Wo;
    for (long long u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      int n, d0, h0, w0;
      decode(u, n, d0, h0, w0);
      float acc[R][4];
#pragma unroll
      for (int a = 0; a < R; ++a)
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[a][o] = 0.f;
#pragma unroll
      for (int j = 0; j < NSLAB; ++j) {
        if (j >= R + KD - 1 || !slab_needed(d0, j)) continue;
        mbar_wait(bar_acc_full(buf), acc_phase);
        tc_fence_after();
        // drain both M tiles: thread = slab row (tile * 128 + lane index), NC columns
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const int row = tile * 128 + row_in_tile;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2 + tile) * GS_NMAX;
          float* prow = s_p + (size_t)row * GS_PSTRIDE;
          for (int c0 = 0; c0 < NC; c0 += 32) {
            uint32_t r[32];
            if (NC - c0 >= 32) {
              tmem_ld32(taddr + c0, r);
              tmem_ld_wait();
              if (row < GS_SLAB_ROWS) {
#pragma unroll
                for (int e = 0; e < 32; ++e) prow[c0 + e] = __uint_as_float(r[e]);
              }
            } else {
              tmem_ld16(taddr + c0, r);
              tmem_ld_wait();
              if (row < GS_SLAB_ROWS) {
#pragma unroll
                for (int e = 0; e < 16; ++e) prow[c0 + e] = __uint_as_float(r[e]);
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar_acc_empty(buf));
        buf ^= 1u;
        if (buf == 0) acc_phase ^= 1u;
        asm volatile("bar.sync 1, 128;" ::: "memory");          // P complete
        // gather: output voxel (hh, ww) of output slice a = j - kd reads P[(hh + kh) * 10 + ww + kw][(kd, kh, kw), o]
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const int a = j - kd;
          if (kd >= KD || a < 0 || a >= R) continue;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float* pp = s_p + (size_t)((hh + kh) * GS_SLAB_W + ww + kw) * GS_PSTRIDE + ((kd * 3 + kh) * 3 + kw) * O;
#pragma unroll
              for (int o = 0; o < 4; ++o)
                if (o < O) acc[a][o] += pp[o];
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");          // everyone has read P: the next slab's drain may overwrite it
      }
      const int oh = h0 + hh, ow = w0 + ww;
      if (oh < g.Ho && ow < g.Wo) {
#pragma unroll
        for (int a = 0; a < R; ++a) {
          const int od = d0 + a;
          if (od >= g.Do) break;
          float* yp = p.y + (size_t)n * O * out_sp + ((size_t)od * g.Ho + oh) * g.Wo + ow;
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < O) yp[(size_t)o * out_sp] = acc[a][o] + (p.bias ? p.bias[o] : 0.f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

bool thin_gs_supported(int x_fmt, int y_fmt, const ConvGeom& g) {
  static const bool off = getenv("HPVG_THIN_GS") && atoi(getenv("HPVG_THIN_GS")) == 0;
  return !off && x_fmt == HPVG_FMT_NDHWC_BF16 && y_fmt == HPVG_FMT_NCDHW_F32 && g.Cin == 64 && g.Cout >= 1 && g.Cout <= 4 &&
         (g.KD == 1 || g.KD == 3) && g.Wi <= 65535 && g.Hi <= 65535;
}

template <int R>
static int launch_gs(const CUtensorMap& mx, const CUtensorMap& mw, GsParams& p, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_mask{0};
  if (attr_pending(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(thin_conv_gs_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("thin_conv_gs: cannot opt in to %d bytes of shared memory: %s", GS_SMEM_BYTES, cudaGetErrorString(e));
      return -2;
    }
    attr_set(attr_mask);
  }
  const ConvGeom& g = p.g;
  p.units_d = (int)cdiv(g.Do, R);
  p.units_h = (int)cdiv(g.Ho, GS_BH);
  p.units_w = (int)cdiv(g.Wo, GS_BW);
  p.num_units = (long long)g.N * p.units_d * p.units_h * p.units_w;
  const int grid = (int)min((long long)num_sms(), p.num_units);
  launch_k(thin_conv_gs_kernel<R>, grid, GS_THREADS, GS_SMEM_BYTES, st, mx, mw, p);
  HPVG_CHECK_LAUNCH("thin_conv_gs_kernel");
  return 0;
}

// x: NDHWC bf16 with 64 channels; w_packed: bf16 image [taps][O][64] of hpvg_pack_weights(rows = O) (forward or data-gradient form);
// y: float32 NCDHW with O channels.  g.Cin = 64, g.Cout = O.
int thin_conv_gs(const void* x, const void* w_packed, const float* bias, float* y, const ConvGeom& g, cudaStream_t st) {
  CUtensorMap mx, mw;
  {
    uint64_t dims[5] = {64, (uint64_t)g.Wi, (uint64_t)g.Hi, (uint64_t)g.Di, (uint64_t)g.N};
    uint32_t box[5] = {64, GS_SLAB_W, GS_SLAB_H, 1, 1};
    if (int rc = make_tmap_bf16(&mx, x, 5, dims, box)) return rc;
  }
  GsParams p;
  p.g = g;
  p.ncols = ((g.taps * g.Cout + 15) / 16) * 16;
  {
    uint64_t dims[2] = {64, (uint64_t)g.taps * g.Cout};
    uint32_t box[2] = {64, (uint32_t)p.ncols};
    if (int rc = make_tmap_bf16(&mw, w_packed, 2, dims, box)) return rc;
  }
  p.bias = bias;
  p.y = y;
  // 2-slice units when 4-slice units would leave SMs idle or waste half a unit (the 6-slice pyramid levels)
  const long long per_slice = (long long)g.N * cdiv(g.Ho, GS_BH) * cdiv(g.Wo, GS_BW);
  const bool two = g.KD == 3 && ((g.Do % 4 != 0 && g.Do % 2 == 0) || per_slice * cdiv(g.Do, 4) * 2 <= num_sms());
  if (g.KD == 1 || two) return launch_gs<2>(mx, mw, p, st);
  return launch_gs<4>(mx, mw, p, st);
}

}  // namespace hpvg

// tcgen05 kernels for the narrow ends of the networks (one side of the convolution has <= 3 channels, the other 64):
//
//   expand_tc_kernel   thin (float32 NCDHW, Cin in {1, 3}) -> wide (bf16 NDHWC, Cout = 64): the 3 -> 64 head convolutions
//                      (modules/networks_3d.py:51,63 with in_channel = nc_im) and the data gradient of the 64 -> 3 / 64 -> 1 tails.
//
// K = taps * Cin = 81 is tiny, but the warp-level mma.sync form of narrow.cu was instruction bound (11.7 us per launch at
// 16 x 64 x 64 against an HBM floor of ~1.5 us for the 8.4 MB bf16 output).  Here the im2col operand is BUILT in shared memory:
// a tile of 128 output voxels x K (padded to a multiple of 16) bf16, in the canonical 128-byte-swizzled K-major layout
// (two 64-wide K atoms), gathered from a float32 halo tile with compile-time offsets, and multiplied with the 64 x K filter by
// 2 - 6 tcgen05.mma (M = 128, N = 64, K = 16) into a TMEM accumulator.  Persistent CTAs (one per SM) pipeline three roles over
// their tiles through mbarriers: 4 builder warps (halo of tile i+1 in flight in registers while tile i is built), one MMA
// thread, 4 epilogue warps (TMEM -> bias / LeakyReLU -> bf16 -> swizzled staging -> TMA store, BatchNorm sums from the staged
// values).  The same operand tile read MN-major is the A operand of the narrow weight gradients (narrow_wgrad_tc below).
#include "common.cuh"
#include <cstdlib>

namespace hpvg {

constexpr int NT_BH = 8, NT_BW = 16;                       // output tile: 8 x 16 voxels = the 128 rows of the GEMM
constexpr int NT_HH = NT_BH + 2, NT_HS = NT_BW + 2;        // float32 halo tile rows x row length
constexpr int NT_HP = 48;                                  // halo row pitch in shared memory: the two tile rows a warp gathers from
                                                           // are 16 banks apart (pitch 18: 2-way conflicts on every gather)
constexpr int NT_ATOM_BYTES = 128 * 128;                   // [128 rows][64 k] bf16
constexpr int NT_A_BYTES = 2 * NT_ATOM_BYTES;              // K padded to 128: two atoms
constexpr int NT_W_ATOM_BYTES = 64 * 128;                  // [64 co][64 k] bf16
constexpr int NT_W_BYTES = 2 * NT_W_ATOM_BYTES;
constexpr int NT_STG_BYTES = 128 * 128;                    // one output tile: 128 voxels x 64 channels bf16
constexpr int NT_HALO_BYTES = ((3 * 3 * NT_HH * NT_HP * 4 + 127) / 128) * 128;      // 17280
// 8 builder warps (thread = half an operand row), 1 MMA warp, 8 epilogue warps (thread = 32 columns of an accumulator row).  With 4 + 4
// warps every role had ONE warp per scheduler and paid each dependent latency in full: ~1 us per tile in the builders and in the
// epilogue (per-CTA time stamps, experiments/head_clk.py).
constexpr int NT_BUILD = 256, NT_EPI = 256;
constexpr int NT_THREADS = NT_BUILD + 32 + NT_EPI;
constexpr int NT_EPI_WARP0 = NT_BUILD / 32 + 1;
constexpr int NT_OFF_A = 0;
constexpr int NT_OFF_W = NT_OFF_A + 2 * NT_A_BYTES;
constexpr int NT_OFF_STG = NT_OFF_W + NT_W_BYTES;
constexpr int NT_OFF_HALO = NT_OFF_STG + 2 * NT_STG_BYTES;
constexpr int NT_OFF_BIAS = NT_OFF_HALO + 2 * NT_HALO_BYTES;
constexpr int NT_OFF_BAR = NT_OFF_BIAS + (64 + 128) * 4;      // bias [64], BatchNorm-sum scratch [128]
constexpr int NT_NBARS = 9;                                // a_full[2], a_empty[2], acc_full[2], acc_empty[2], w_ready
constexpr int NT_SMEM_BYTES = NT_OFF_BAR + NT_NBARS * 8 + 16 + 1024;
constexpr int NT_TMEM_COLS = 128;                          // two accumulators of 64 columns
static_assert(NT_OFF_W % 1024 == 0 && NT_OFF_STG % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(NT_SMEM_BYTES <= 227 * 1024, "shared memory budget");

// halo offset (floats) of GEMM column k = tap * J + j, tap = (kd * 3 + kh) * 3 + kw — the order of the [tap][ci][co] filter image.
// FLIP: the thin tensor is read at "voxel - tap" (the tail's weight gradient) instead of "voxel + tap": mirrored offsets.
template <int KDT, int J, bool FLIP>
__host__ __device__ constexpr int nt_koff(int k) {
  const int j = k % J, tap = k / J, kd = tap / 9, kh = (tap % 9) / 3, kw = tap % 3;
  return ((j * KDT + (FLIP ? KDT - 1 - kd : kd)) * NT_HH + (FLIP ? 2 - kh : kh)) * NT_HP + (FLIP ? 2 - kw : kw);
}

// the thin tensor a tile's operand rows are gathered from: float32 [N][J][D][H][W]; element (d, h, w) of the halo of the tile at
// (d0, h0, w0) is the tensor at (d0 + sd + d, h0 + sh + h, w0 + sh + w), zero outside
struct NtThin {
  const float* ptr;
  int D, H, W;
  int sd, sh;
};

// tile index -> (n, d, h0, w0) without a division per tile: a mixed-radix counter advanced by a constant step (the divisions that
// split the first index and the step happen once).  Three dependent integer divisions per tile cost ~0.5 us in the builders.
struct NtTileIter {
  int w, h, d, n;          // current tile (w, h in tiles)
  int sw, sh, sd, sn;      // the step, same radices
  int tw, th, td;          // tiles per row, tile rows per slice, slices per sample
  long long left, stepsz;  // tiles from the current one to the end (<= 0: past the end)
  __device__ __forceinline__ void init(long long first, long long step, long long total, int tiles_w, int tiles_h, int D) {
    tw = tiles_w; th = tiles_h; td = D;
    unsigned t = (unsigned)first, q = t / (unsigned)tw;
    w = (int)(t - q * (unsigned)tw); t = q; q = t / (unsigned)th;
    h = (int)(t - q * (unsigned)th); t = q; q = t / (unsigned)td;
    d = (int)(t - q * (unsigned)td); n = (int)q;
    t = (unsigned)step; q = t / (unsigned)tw;
    sw = (int)(t - q * (unsigned)tw); t = q; q = t / (unsigned)th;
    sh = (int)(t - q * (unsigned)th); t = q; q = t / (unsigned)td;
    sd = (int)(t - q * (unsigned)td); sn = (int)q;
    left = total - first;
    stepsz = step;
  }
  __device__ __forceinline__ bool valid() const { return left > 0; }
  __device__ __forceinline__ void next() {
    w += sw; h += sh; d += sd; n += sn;
    if (w >= tw) { w -= tw; ++h; }
    if (h >= th) { h -= th; ++d; }
    if (d >= td) { d -= td; ++n; }
    left -= stepsz;
  }
};

template <int KDT, int J>
struct NtShape {
  static constexpr int K = KDT * 9 * J;                // 81, 27 or 9
  static constexpr int KSTEPS = (K + 1 + 15) / 16;     // K-major use: MMAs per tile (the column after the last tap is a column of ones)
  static constexpr int NCHUNK = 2 * KSTEPS;            // 16-byte chunks (8 k values) per operand row
  static constexpr int HALO_N = J * KDT * NT_HH * NT_HS;
  static constexpr int HREGS = (HALO_N + NT_BUILD - 1) / NT_BUILD;
};

// per-thread, tile-independent part of the halo gather: element i = tid + 128 r of the halo tile is (j, kd, yy, xx)
template <int KDT, int J>
__device__ __forceinline__ void nt_halo_prepare(int tid, const NtThin& th, int (&hoff)[NtShape<KDT, J>::HREGS], int (&hco)[NtShape<KDT, J>::HREGS]) {
  using S = NtShape<KDT, J>;
  const long long sp = (long long)th.D * th.H * th.W;
#pragma unroll
  for (int r = 0; r < S::HREGS; ++r) {
    const int i = tid + NT_BUILD * r;
    const int xx = i % NT_HS, yy = (i / NT_HS) % NT_HH, kd = (i / (NT_HS * NT_HH)) % KDT, j = i / (NT_HS * NT_HH * KDT);
    hoff[r] = (int)(j * sp) + (kd * th.H + yy) * th.W + xx;
    hco[r] = i < S::HALO_N ? (xx | (yy << 8) | (kd << 16)) : -1;
  }
}

// global -> registers: the halo of the tile whose first voxel is (n, d0, h0, w0)
template <int KDT, int J>
__device__ __forceinline__ void nt_halo_load(const NtThin& th, int n, int d0, int h0, int w0, const int (&hoff)[NtShape<KDT, J>::HREGS],
                                             const int (&hco)[NtShape<KDT, J>::HREGS], float (&hv)[NtShape<KDT, J>::HREGS]) {
  using S = NtShape<KDT, J>;
  const long long sp = (long long)th.D * th.H * th.W;
  const int dd = d0 + th.sd, hh = h0 + th.sh, ww = w0 + th.sh;
  const float* src = th.ptr + (long long)n * J * sp + ((long long)dd * th.H + hh) * (long long)th.W + ww;
#pragma unroll
  for (int r = 0; r < S::HREGS; ++r) {
    hv[r] = 0.f;
    if (hco[r] >= 0) {
      const int iw = ww + (hco[r] & 255), ih = hh + ((hco[r] >> 8) & 255), id = dd + (hco[r] >> 16);
      if ((unsigned)id < (unsigned)th.D && (unsigned)ih < (unsigned)th.H && (unsigned)iw < (unsigned)th.W) hv[r] = __ldg(src + hoff[r]);
    }
  }
}

template <int KDT, int J>
__device__ __forceinline__ void nt_halo_store(float* halo, int tid, const float (&hv)[NtShape<KDT, J>::HREGS]) {
  using S = NtShape<KDT, J>;
#pragma unroll
  for (int r = 0; r < S::HREGS; ++r) {
    const int i = tid + NT_BUILD * r;
    if (i < S::HALO_N) halo[(i / NT_HS) * NT_HP + i % NT_HS] = hv[r];
  }
}

// operand row m (= tile voxel (m / 16, m % 16)) of the im2col tile: K gathered values, a one, zeros; bf16, 128-byte swizzle
template <int KDT, int J, bool FLIP, int HALF>
__device__ __forceinline__ void nt_build_row(const float* halo, int m, uint32_t arow) {
  using S = NtShape<KDT, J>;
  const float* hb = halo + (m >> 4) * NT_HP + (m & 15);
#pragma unroll
  for (int c = HALF * (S::NCHUNK / 2); c < (HALF + 1) * (S::NCHUNK / 2); ++c) {
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k0 = 8 * c + 2 * e, k1 = k0 + 1;
      const float v0 = k0 < S::K ? hb[nt_koff<KDT, J, FLIP>(k0 < S::K ? k0 : 0)] : (k0 == S::K ? 1.f : 0.f);
      const float v1 = k1 < S::K ? hb[nt_koff<KDT, J, FLIP>(k1 < S::K ? k1 : 0)] : (k1 == S::K ? 1.f : 0.f);
      pk[e] = pack_bf16x2(v0, v1);
    }
    const uint32_t addr = arow + (c >> 3) * NT_ATOM_BYTES + (((c & 7) ^ (m & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
  }
}

struct NtParams {
  ConvGeom g;
  NtThin thin;          // x (float32 NCDHW)
  const float* w_tco;   // [taps][Cin][64] float32 (hpvg_pack_weights_expand)
  const float* bias;
  float* stats;         // BatchNorm sums [sum : 64][sum of squares : 64] per sample block, or nullptr
  int act;
  float slope;
  int tiles_h, tiles_w;
  long long num_tiles;
  long long* dbg;       // optional per-CTA time stamps (development aid, hpvg_debug_set_clock_buffer): 16 slots of %globaltimer ns
};

__device__ __forceinline__ long long nt_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}

template <int KDT, int CIN>
__global__ void __launch_bounds__(NT_THREADS, 1) expand_tc_kernel(const __grid_constant__ CUtensorMap tmap_y, const NtParams p) {
  using S = NtShape<KDT, CIN>;
  constexpr int K = S::K, KSTEPS = (K + 15) / 16, NCHUNK = S::NCHUNK, HREGS = S::HREGS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t s_bar = sbase + NT_OFF_BAR;
  auto bar_a_full = [&](int i) { return s_bar + 8u * i; };
  auto bar_a_empty = [&](int i) { return s_bar + 8u * (2 + i); };
  auto bar_acc_full = [&](int i) { return s_bar + 8u * (4 + i); };
  auto bar_acc_empty = [&](int i) { return s_bar + 8u * (6 + i); };
  const uint32_t bar_w_ready = s_bar + 8u * 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + NT_OFF_BAR + NT_NBARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;
  long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 16 : nullptr;
  if (dbg && threadIdx.x == 0) dbg[0] = nt_now();

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_a_full(i), NT_BUILD);
      mbar_init(bar_a_empty(i), 1);
      mbar_init(bar_acc_full(i), 1);
      mbar_init(bar_acc_empty(i), NT_EPI);
    }
    mbar_init(bar_w_ready, NT_EPI);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_y);
  }
  pdl_trigger();
  pdl_wait();

  auto decode = [&](long long t64, int& n, int& od, int& h0, int& w0) {      // 32-bit arithmetic (64-bit division is a subroutine)
    unsigned t = (unsigned)t64;
    unsigned q = t / (unsigned)p.tiles_w;
    w0 = (int)(t - q * (unsigned)p.tiles_w) * NT_BW;
    t = q;
    q = t / (unsigned)p.tiles_h;
    h0 = (int)(t - q * (unsigned)p.tiles_h) * NT_BH;
    t = q;
    q = t / (unsigned)g.Do;
    od = (int)(t - q * (unsigned)g.Do);
    n = (int)q;
  };
  // builders: the halo of a tile travels global -> registers -> shared memory, one tile ahead of the tile being built
  // (two register sets: the loads of tiles i + 1 and i + 2 are in flight while tile i is built — with one set the global
  // round trip of every tile was exposed, ~0.7 us per tile)
  float hv0[HREGS], hv1[HREGS];
  int hoff[HREGS], hco[HREGS];
  NtTileIter lit;        // the tile whose halo is loaded next (loads are issued in tile order, two tiles ahead of the build)
  auto load_halo = [&](float (&hv)[HREGS]) {
    if (!lit.valid()) return;
    nt_halo_load<KDT, CIN>(p.thin, lit.n, lit.d, lit.h * NT_BH, lit.w * NT_BW, hoff, hco, hv);
    lit.next();
  };
  if (warp < NT_BUILD / 32) {
    nt_halo_prepare<KDT, CIN>(threadIdx.x, p.thin, hoff, hco);
    lit.init(blockIdx.x, gridDim.x, p.num_tiles, p.tiles_w, p.tiles_h, g.Do);
    load_halo(hv0);
    load_halo(hv1);
  }

  __syncthreads();                 // barriers initialised
  uint32_t tmem_base = 0;
  if (warp >= NT_BUILD / 32) {
    // TMEM allocation concerns the MMA and epilogue warps only: the builders go straight to their first tile
    if (warp == NT_BUILD / 32) tmem_alloc<NT_TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    asm volatile("bar.sync 3, %0;" ::"n"(32 + NT_EPI) : "memory");
    tc_fence_after();
    tmem_base = *tmem_slot;
  }

  if (warp < NT_BUILD / 32) {
    // ===================== builders: thread = half of operand row m = output voxel (m / 16, m % 16) of the tile =====================
    const int m = threadIdx.x & 127, half = threadIdx.x >> 7;
    int it = 0;
    auto tile_step = [&](long long t, float (&hv)[HREGS]) {
      const int buf = it & 1;
      float* halo = reinterpret_cast<float*>(sgen + NT_OFF_HALO + buf * NT_HALO_BYTES);
      nt_halo_store<KDT, CIN>(halo, threadIdx.x, hv);
      asm volatile("bar.sync 1, %0;" ::"n"(NT_BUILD) : "memory");      // halo[buf] complete; everyone is done with halo[buf] of tile it - 2
      if (dbg && threadIdx.x == 0 && it == 0) dbg[1] = nt_now();
      load_halo(hv);
      if (it >= 2) mbar_wait(bar_a_empty(buf), (uint32_t)(((it >> 1) - 1) & 1));      // the MMAs of tile it - 2 have read A[buf]
      if (half == 0)
        nt_build_row<KDT, CIN, false, 0>(halo, m, sbase + NT_OFF_A + buf * NT_A_BYTES + m * 128);
      else
        nt_build_row<KDT, CIN, false, 1>(halo, m, sbase + NT_OFF_A + buf * NT_A_BYTES + m * 128);
      fence_proxy_async();
      mbar_arrive(bar_a_full(buf));
      if (dbg && threadIdx.x == 0 && it < 4) dbg[2 + it] = nt_now();
      ++it;
    };
    for (long long t = blockIdx.x; t < p.num_tiles; t += 2 * (long long)gridDim.x) {
      tile_step(t, hv0);
      if (t + gridDim.x < p.num_tiles) tile_step(t + gridDim.x, hv1);
    }
  } else if (warp == NT_BUILD / 32) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_bf16(128, 64, 0, 0);
      const uint64_t a_base = umma_desc(sbase + NT_OFF_A, 16, 1024, 2);
      const uint64_t b_base = umma_desc(sbase + NT_OFF_W, 16, 1024, 2);
      mbar_wait(bar_w_ready, 0);
      tc_fence_after();
      int it = 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        if (it >= 2) {
          mbar_wait(bar_acc_empty(buf), (uint32_t)(((it >> 1) - 1) & 1));
          tc_fence_after();
        }
        mbar_wait(bar_a_full(buf), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)buf * 64u;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint64_t ad = a_base + (uint64_t)((buf * NT_A_BYTES + (ks >> 2) * NT_ATOM_BYTES + (ks & 3) * 32) >> 4);
          const uint64_t bd = b_base + (uint64_t)(((ks >> 2) * NT_W_ATOM_BYTES + (ks & 3) * 32) >> 4);
          umma_bf16(tacc, ad, bd, IDESC, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_a_empty(buf));
        umma_commit(bar_acc_full(buf));
        if (dbg && it < 4) dbg[7 + it] = nt_now();
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: TMEM lane quadrant = warp & 3, column half = (warp - first) / 4 =====================
    const int q = warp & 3;
    const int ch = (warp - NT_EPI_WARP0) >> 2;
    const int m = q * 32 + lane;                       // accumulator row = tile voxel (m / 16, m % 16)
    const int et = threadIdx.x - (NT_BUILD + 32);
    float* bias_s = reinterpret_cast<float*>(sgen + NT_OFF_BIAS);
    {
      // filter image [k][co] float32 -> bf16 B operand [co][k], K-major, 128-byte swizzle, two 64-wide K atoms; k >= K is zero.
      // Every load of the thread is issued before the first one is consumed (one round trip, not one per chunk).
      const int co = et & 63, part = et >> 6;
      constexpr int WCH = (NCHUNK + 3) / 4;
      float wv[WCH][8];
#pragma unroll
      for (int cc = 0; cc < WCH; ++cc)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = 8 * (part + 4 * cc) + e;
          wv[cc][e] = k < K ? __ldg(p.w_tco + (size_t)k * 64 + co) : 0.f;
        }
      const float bias_v = (et < 64 && p.bias) ? __ldg(p.bias + et) : 0.f;
#pragma unroll
      for (int cc = 0; cc < WCH; ++cc) {
        const int c = part + 4 * cc;
        if (c < NCHUNK) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) pk[e] = pack_bf16x2(wv[cc][2 * e], wv[cc][2 * e + 1]);
          const uint32_t addr = sbase + NT_OFF_W + (c >> 3) * NT_W_ATOM_BYTES + co * 128 + (((c & 7) ^ (co & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
      }
      if (et < 64) bias_s[et] = bias_v;
      fence_proxy_async();
      mbar_arrive(bar_w_ready);
      asm volatile("bar.sync 2, %0;" ::"n"(NT_EPI) : "memory");
    }
    float st_s[2] = {0.f, 0.f}, st_q[2] = {0.f, 0.f};
    int stat_n = -1;
    // flush: the eight row groups combine in shared memory first (one global atomic per channel, sum and CTA — issued by every CTA
    // at the same moment, same-address atomics serialise in L2)
    float* red = reinterpret_cast<float*>(sgen + NT_OFF_BIAS) + 64;      // [128]
    auto flush_stats = [&]() {
      if (p.stats && stat_n >= 0) {
        if (et < 128) red[et] = 0.f;
        asm volatile("bar.sync 2, %0;" ::"n"(NT_EPI) : "memory");
        const int c = 2 * (et & 31);
        atomicAdd(red + c, st_s[0]);
        atomicAdd(red + c + 1, st_s[1]);
        atomicAdd(red + 64 + c, st_q[0]);
        atomicAdd(red + 64 + c + 1, st_q[1]);
        asm volatile("bar.sync 2, %0;" ::"n"(NT_EPI) : "memory");
        if (et < 128) atomicAdd(p.stats + (size_t)stat_n * g.stats_stride + et, red[et]);
      }
      st_s[0] = st_s[1] = st_q[0] = st_q[1] = 0.f;
    };
    int it = 0;
    NtTileIter eit;
    eit.init(blockIdx.x, gridDim.x, p.num_tiles, p.tiles_w, p.tiles_h, g.Do);
    for (; eit.valid(); eit.next(), ++it) {
      const int buf = it & 1;
      const int n = eit.n, od = eit.d, h0 = eit.h * NT_BH, w0 = eit.w * NT_BW;
      const int oh = h0 + (m >> 4), ow = w0 + (m & 15);
      const bool ok = oh < g.Ho && ow < g.Wo;
      mbar_wait(bar_acc_full(buf), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 64u + (uint32_t)ch * 32u, r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_acc_empty(buf));                 // the accumulator is in registers: the MMAs of tile it + 2 may overwrite it
      if (et == 0) tma_store_wait_read<1>();           // the staging tile used two tiles ago has been read by its TMA store
      asm volatile("bar.sync 2, %0;" ::"n"(NT_EPI) : "memory");
      const uint32_t sdst = sbase + NT_OFF_STG + buf * NT_STG_BYTES + m * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[8];
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch * 32 + 8 * c), b1 = *reinterpret_cast<const float4*>(bias_s + ch * 32 + 8 * c + 4);
        const float bq[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float a = __uint_as_float(r[8 * c + e]) + bq[e];
          if (p.act == HPVG_ACT_LRELU) a = fmaxf(a, a * p.slope);
          v[e] = ok ? a : 0.f;
        }
        const uint32_t addr = sdst + ((uint32_t)((ch * 4 + c) ^ (m & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[0], v[1])), "r"(pack_bf16x2(v[2], v[3])),
                     "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7]))
                     : "memory");
      }
      fence_proxy_async();
      asm volatile("bar.sync 2, %0;" ::"n"(NT_EPI) : "memory");
      if (et == 0) {
        tma_store_5d(&tmap_y, sbase + NT_OFF_STG + buf * NT_STG_BYTES, 0, w0, h0, od, n);
        tma_store_commit();
        if (dbg && it < 4) dbg[11 + it] = nt_now();
      }
      if (p.stats) {
        // column sums over the staged (bf16-rounded, invalid rows zeroed) tile: thread = channel pair, 16 rows each; the sums stay
        // in registers over the tiles of this CTA and are flushed when the sample changes (per-sample statistics) and at the end
        if (n != stat_n) {
          flush_stats();
          stat_n = n;
        }
        const int cp = et & 31, rg = et >> 5;
        const uint8_t* tile = sgen + NT_OFF_STG + buf * NT_STG_BYTES;
#pragma unroll 8
        for (int rr = 0; rr < 16; ++rr) {
          const int row = rg * 16 + rr;
          const float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(tile + row * 128 + (((cp >> 2) ^ (row & 7)) << 4) + (cp & 3) * 4));
          st_s[0] += f.x;
          st_s[1] += f.y;
          st_q[0] = fmaf(f.x, f.x, st_q[0]);
          st_q[1] = fmaf(f.y, f.y, st_q[1]);
        }
      }
    }
    flush_stats();
#if HPVG_STORE_WAIT_READ
    if (et == 0) tma_store_wait_read<0>();      // the staging tiles have been read: the CTA may leave (the writes land by the end of the grid)
#else
    if (et == 0) tma_store_wait_all<0>();
#endif
    if (dbg && et == 0) dbg[15] = nt_now();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NT_BUILD / 32) tmem_dealloc<NT_TMEM_COLS>(tmem_base);
}

bool expand_tc_supported(const ConvGeom& g) {
  static const bool off = getenv("HPVG_EXPAND_TC") && atoi(getenv("HPVG_EXPAND_TC")) == 0;
  return !off && (g.Cin == 1 || g.Cin == 3) && g.Cout == 64 && (g.KD == 1 || g.KD == 3) && g.taps == g.KD * 9;
}

template <int KDT, int CIN>
static int launch_expand_tc(const CUtensorMap& my, NtParams& p, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_mask{0};
  if (attr_pending(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(expand_tc_kernel<KDT, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, NT_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("expand_tc: cannot opt in to %d bytes of shared memory: %s", NT_SMEM_BYTES, cudaGetErrorString(e));
      return -2;
    }
    attr_set(attr_mask);
  }
  const int grid = (int)min((long long)num_sms(), p.num_tiles);
  launch_k(expand_tc_kernel<KDT, CIN>, grid, NT_THREADS, NT_SMEM_BYTES, st, my, p);
  HPVG_CHECK_LAUNCH("expand_tc_kernel");
  return 0;
}

// x: float32 NCDHW (Cin in {1, 3}); w_tco: float32 [taps][Cin][64]; y: bf16 NDHWC with 64 channels
int expand_tc(const void* x, const float* w_tco, const float* bias, void* y, const ConvGeom& g, int act, float slope, float* stats,
              cudaStream_t st) {
  CUtensorMap my;
  {
    uint64_t dims[5] = {64, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)g.N};
    uint32_t box[5] = {64, NT_BW, NT_BH, 1, 1};
    if (int rc = make_tmap_bf16(&my, y, 5, dims, box)) return rc;
  }
  NtParams p = {};
  p.g = g;
  p.thin.ptr = reinterpret_cast<const float*>(x);
  p.thin.D = g.Di; p.thin.H = g.Hi; p.thin.W = g.Wi;
  p.thin.sd = -g.pad_d; p.thin.sh = -g.pad;
  p.w_tco = w_tco;
  p.bias = bias;
  p.stats = stats;
  p.act = act;
  p.slope = slope;
  p.tiles_h = (int)cdiv(g.Ho, NT_BH);
  p.tiles_w = (int)cdiv(g.Wo, NT_BW);
  p.num_tiles = (long long)g.N * g.Do * p.tiles_h * p.tiles_w;
  if (p.num_tiles >= (1LL << 31)) {
    set_error("expand_tc: %lld work items exceed the 32-bit unit index of the kernel", (long long)p.num_tiles);
    return -1;
  }
  p.dbg = debug_clock_buffer();
  if (g.KD == 3) return g.Cin == 3 ? launch_expand_tc<3, 3>(my, p, st) : launch_expand_tc<3, 1>(my, p, st);
  return g.Cin == 3 ? launch_expand_tc<1, 3>(my, p, st) : launch_expand_tc<1, 1>(my, p, st);
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradients of the narrow layers (aten::convolution_backward grad_weight, modules/networks_3d.py:51,63,175,341,362):
//     OUT[(t, j)][k] = sum_p THIN[j][p (+/-) t + shift] * WIDE[p][k]          j < J <= 3 (float32 NCDHW), k < 64 (bf16 NDHWC)
//   head  (x thin, gy wide):  dw[co = k][ci = j][t] = sum_v x[ci][v + t - pad] * gy[v][co]
//   tail  (x wide, gy thin):  dw[c = j][ci = k][t]  = sum_u gy[c][u - t + pad] * x[u][ci]
// GEMM view: D[M = (t, j) : 128 rows, K_ + 1 used][N = 64] += IM2COL^T[(t, j)][voxel] * WIDE[voxel][k], K = voxels.  The A operand is
// the SAME im2col tile expand_tc_kernel builds (rows = voxels, 128 bytes = 64 (t, j) entries per row and atom), read MN-major; the
// B operand is the wide tile as TMA delivers it (NDHWC rows), MN-major as in wgrad_tc.cu.  The column of ones after the last tap
// makes row K_ of D the channel sum of WIDE: the bias gradient of a head layer comes for free.  One accumulator per CTA over all
// its tiles; per-CTA partials + the fixed-order reduction of narrow.cu (deterministic).
// ---------------------------------------------------------------------------------------------------------------
constexpr int NW_WIDE_BYTES = 128 * 128;
constexpr int NW_OFF_A = 0;
constexpr int NW_OFF_B = NW_OFF_A + 2 * NT_A_BYTES;
constexpr int NW_OFF_HALO = NW_OFF_B + 2 * NW_WIDE_BYTES;
constexpr int NW_OFF_BAR = NW_OFF_HALO + 2 * NT_HALO_BYTES;
constexpr int NW_NBARS = 7;                                // a_full[2], b_full[2], ab_empty[2], acc_full
constexpr int NW_SMEM_BYTES = NW_OFF_BAR + NW_NBARS * 8 + 16 + 1024;
constexpr int NW_TMEM_COLS = 64;
constexpr int NW_EPI = 128;                                // 4 drain warps (the accumulator is read once, at the end)
constexpr int NW_THREADS = NT_BUILD + 32 + NW_EPI;
constexpr int NW_ROWS = 96;                                // rows of a partial: narrow.cu's OM_ROWS
static_assert(NW_OFF_B % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(NW_SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct NwParams {
  NtThin thin;
  int N, Dw, Hw, Ww;          // extents of the wide tensor (the voxels the sum runs over)
  int tiles_h, tiles_w;
  long long num_tiles;
  float* partial;             // [grid][NW_ROWS][64] float32
};

template <int KDT, int J, bool FLIP>
__global__ void __launch_bounds__(NW_THREADS, 1) narrow_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_wide, const NwParams p) {
  using S = NtShape<KDT, J>;
  constexpr int HREGS = S::HREGS;
  static_assert(S::K + 1 <= NW_ROWS, "rows of a partial");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const uint32_t s_bar = sbase + NW_OFF_BAR;
  auto bar_a_full = [&](int i) { return s_bar + 8u * i; };
  auto bar_b_full = [&](int i) { return s_bar + 8u * (2 + i); };
  auto bar_ab_empty = [&](int i) { return s_bar + 8u * (4 + i); };
  const uint32_t bar_acc_full = s_bar + 8u * 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + NW_OFF_BAR + NW_NBARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_a_full(i), NT_BUILD);
      mbar_init(bar_b_full(i), 1);
      mbar_init(bar_ab_empty(i), 1);
    }
    mbar_init(bar_acc_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap_wide);
  }
  pdl_trigger();
  pdl_wait();

  auto decode = [&](long long t64, int& n, int& d, int& h0, int& w0) {
    unsigned t = (unsigned)t64;
    unsigned q = t / (unsigned)p.tiles_w;
    w0 = (int)(t - q * (unsigned)p.tiles_w) * NT_BW;
    t = q;
    q = t / (unsigned)p.tiles_h;
    h0 = (int)(t - q * (unsigned)p.tiles_h) * NT_BH;
    t = q;
    q = t / (unsigned)p.Dw;
    d = (int)(t - q * (unsigned)p.Dw);
    n = (int)q;
  };
  float hv0[HREGS], hv1[HREGS];
  int hoff[HREGS], hco[HREGS];
  NtTileIter lit;
  auto load_halo = [&](float (&hv)[HREGS]) {
    if (!lit.valid()) return;
    nt_halo_load<KDT, J>(p.thin, lit.n, lit.d, lit.h * NT_BH, lit.w * NT_BW, hoff, hco, hv);
    lit.next();
  };
  if (warp < NT_BUILD / 32) {
    nt_halo_prepare<KDT, J>(threadIdx.x, p.thin, hoff, hco);
    lit.init(blockIdx.x, gridDim.x, p.num_tiles, p.tiles_w, p.tiles_h, p.Dw);
    load_halo(hv0);
    load_halo(hv1);
    // the (t, j) entries behind the ones column are never written by nt_build_row: zero them once (rows K_+1 .. 127 of D are
    // not used, but they must not turn into NaN patterns that cost denormal / exception handling in the tensor pipe)
    const int m = threadIdx.x & 127;
    const int buf = threadIdx.x >> 7;
    {
#pragma unroll
      for (int c = S::NCHUNK; c < 16; ++c) {
        const uint32_t addr = sbase + NW_OFF_A + buf * NT_A_BYTES + m * 128 + (c >> 3) * NT_ATOM_BYTES + (((c & 7) ^ (m & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
      }
    }
  }
  __syncthreads();                 // barriers initialised
  uint32_t tmem_base = 0;
  if (warp >= NT_BUILD / 32) {
    if (warp == NT_BUILD / 32) tmem_alloc<NW_TMEM_COLS>(smem_u32(tmem_slot));
    tc_fence_before();
    asm volatile("bar.sync 3, %0;" ::"n"(32 + NW_EPI) : "memory");
    tc_fence_after();
    tmem_base = *tmem_slot;
  }

  if (warp < NT_BUILD / 32) {
    // ===================== builders (as in expand_tc_kernel) =====================
    const int m = threadIdx.x & 127, half = threadIdx.x >> 7;
    int it = 0;
    auto tile_step = [&](long long t, float (&hv)[HREGS]) {
      const int buf = it & 1;
      float* halo = reinterpret_cast<float*>(sgen + NW_OFF_HALO + buf * NT_HALO_BYTES);
      nt_halo_store<KDT, J>(halo, threadIdx.x, hv);
      asm volatile("bar.sync 1, %0;" ::"n"(NT_BUILD) : "memory");
      load_halo(hv);
      if (it >= 2) mbar_wait(bar_ab_empty(buf), (uint32_t)(((it >> 1) - 1) & 1));
      if (half == 0)
        nt_build_row<KDT, J, FLIP, 0>(halo, m, sbase + NW_OFF_A + buf * NT_A_BYTES + m * 128);
      else
        nt_build_row<KDT, J, FLIP, 1>(halo, m, sbase + NW_OFF_A + buf * NT_A_BYTES + m * 128);
      fence_proxy_async();
      mbar_arrive(bar_a_full(buf));
      ++it;
    };
    for (long long t = blockIdx.x; t < p.num_tiles; t += 2 * (long long)gridDim.x) {
      tile_step(t, hv0);
      if (t + gridDim.x < p.num_tiles) tile_step(t + gridDim.x, hv1);
    }
  } else if (warp == NT_BUILD / 32) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_bf16(128, 64, 1, 1);      // A and B MN-major
      int it = 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        mbar_wait(bar_a_full(buf), ph);
        mbar_wait(bar_b_full(buf), ph);
        tc_fence_after();
        // A: two 64-wide M atoms NT_ATOM_BYTES apart (LBO), K groups of 8 voxels 1024 bytes apart (SBO); B: one N atom
        const uint64_t a_base = umma_desc(sbase + NW_OFF_A + buf * NT_A_BYTES, NT_ATOM_BYTES, 1024, 2);
        const uint64_t b_base = umma_desc(sbase + NW_OFF_B + buf * NW_WIDE_BYTES, NW_WIDE_BYTES, 1024, 2);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16(tmem_base, a_base + (uint64_t)((ks * 2048) >> 4), b_base + (uint64_t)((ks * 2048) >> 4), IDESC, (it > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar_ab_empty(buf));
      }
      umma_commit(bar_acc_full);
    }
    __syncwarp();
  } else {
    // ===================== first drain warp: TMA producer of the wide tiles; then all four warps drain the accumulator =====================
    if (warp == NT_EPI_WARP0 && elect_one()) {
      int it = 0;
      NtTileIter pit;
      pit.init(blockIdx.x, gridDim.x, p.num_tiles, p.tiles_w, p.tiles_h, p.Dw);
      for (; pit.valid(); pit.next(), ++it) {
        const int buf = it & 1;
        if (it >= 2) mbar_wait(bar_ab_empty(buf), (uint32_t)(((it >> 1) - 1) & 1));
        mbar_expect_tx(bar_b_full(buf), NW_WIDE_BYTES);
        tma_load_5d(sbase + NW_OFF_B + buf * NW_WIDE_BYTES, &tmap_wide, bar_b_full(buf), 0, pit.w * NT_BW, pit.h * NT_BH, pit.d, pit.n);     // out-of-range rows: zeros
      }
    }
    __syncwarp();
    const int q = warp & 3;
    const int m = q * 32 + lane;                       // accumulator row = (t, j) entry m
    mbar_wait(bar_acc_full, 0);
    tc_fence_after();
    uint32_t r[64];
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    tmem_ld32(taddr, r);
    tmem_ld32(taddr + 32, r + 32);
    tmem_ld_wait();
    if (m < NW_ROWS) {
      float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.x * NW_ROWS + m) * 64);
#pragma unroll
      for (int c = 0; c < 16; ++c)
        dst[c] = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NT_BUILD / 32) tmem_dealloc<NW_TMEM_COLS>(tmem_base);
}

bool narrow_wgrad_tc_supported(const ConvGeom& g, bool head) {
  static const bool off = getenv("HPVG_NARROW_WGRAD_TC") && atoi(getenv("HPVG_NARROW_WGRAD_TC")) == 0;
  const int J = head ? g.Cin : g.Cout;
  return !off && (J == 1 || J == 3) && (g.KD == 1 || g.KD == 3) && g.taps == g.KD * 9;
}

int narrow_wgrad_tc_grid(const ConvGeom& g, bool head) {
  const long long tiles = head ? (long long)g.N * g.Do * cdiv(g.Ho, NT_BH) * cdiv(g.Wo, NT_BW)
                               : (long long)g.N * g.Di * cdiv(g.Hi, NT_BH) * cdiv(g.Wi, NT_BW);
  return (int)min((long long)num_sms(), tiles);
}

template <int KDT, int J, bool FLIP>
static int launch_narrow_wgrad_tc(const CUtensorMap& mw, const NwParams& p, int grid, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_mask{0};
  if (attr_pending(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(narrow_wgrad_tc_kernel<KDT, J, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, NW_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("narrow_wgrad_tc: cannot opt in to %d bytes of shared memory: %s", NW_SMEM_BYTES, cudaGetErrorString(e));
      return -2;
    }
    attr_set(attr_mask);
  }
  launch_k(narrow_wgrad_tc_kernel<KDT, J, FLIP>, grid, NW_THREADS, NW_SMEM_BYTES, st, mw, p);
  HPVG_CHECK_LAUNCH("narrow_wgrad_tc_kernel");
  return 0;
}

// per-CTA partials [grid][NW_ROWS][64]: row t * J + j = the (t, j) entry, row taps * J = channel sums of the wide tensor
int narrow_wgrad_tc(const void* wide, const float* thin, bool head, const ConvGeom& g, float* partial, int grid, cudaStream_t st) {
  NwParams p = {};
  p.thin.ptr = thin;
  p.N = g.N;
  if (head) {   // WIDE = gy over output voxels, THIN = x at v + t - pad
    p.Dw = g.Do; p.Hw = g.Ho; p.Ww = g.Wo;
    p.thin.D = g.Di; p.thin.H = g.Hi; p.thin.W = g.Wi;
    p.thin.sd = -g.pad_d; p.thin.sh = -g.pad;
  } else {      // WIDE = x over input voxels, THIN = gy at u - t + pad: halo origin u + pad - 2, mirrored tap offsets
    p.Dw = g.Di; p.Hw = g.Hi; p.Ww = g.Wi;
    p.thin.D = g.Do; p.thin.H = g.Ho; p.thin.W = g.Wo;
    p.thin.sd = g.pad_d - (g.KD - 1); p.thin.sh = g.pad - 2;
  }
  p.tiles_h = (int)cdiv(p.Hw, NT_BH);
  p.tiles_w = (int)cdiv(p.Ww, NT_BW);
  p.num_tiles = (long long)p.N * p.Dw * p.tiles_h * p.tiles_w;
  if (p.num_tiles >= (1LL << 31)) {
    set_error("narrow_wgrad_tc: %lld work items exceed the 32-bit unit index of the kernel", (long long)p.num_tiles);
    return -1;
  }
  p.partial = partial;
  CUtensorMap mw;
  {
    uint64_t dims[5] = {64, (uint64_t)p.Ww, (uint64_t)p.Hw, (uint64_t)p.Dw, (uint64_t)p.N};
    uint32_t box[5] = {64, NT_BW, NT_BH, 1, 1};
    if (int rc = make_tmap_bf16(&mw, wide, 5, dims, box)) return rc;
  }
  const int J = head ? g.Cin : g.Cout;
  if (head) {
    if (g.KD == 3) return J == 3 ? launch_narrow_wgrad_tc<3, 3, false>(mw, p, grid, st) : launch_narrow_wgrad_tc<3, 1, false>(mw, p, grid, st);
    return J == 3 ? launch_narrow_wgrad_tc<1, 3, false>(mw, p, grid, st) : launch_narrow_wgrad_tc<1, 1, false>(mw, p, grid, st);
  }
  if (g.KD == 3) return J == 3 ? launch_narrow_wgrad_tc<3, 3, true>(mw, p, grid, st) : launch_narrow_wgrad_tc<3, 1, true>(mw, p, grid, st);
  return J == 3 ? launch_narrow_wgrad_tc<1, 3, true>(mw, p, grid, st) : launch_narrow_wgrad_tc<1, 1, true>(mw, p, grid, st);
}


}  // namespace hpvg

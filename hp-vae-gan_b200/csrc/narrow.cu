// CUDA-core kernels for the narrow ends of the networks, where one side of the convolution has <= 4 channels (the
// 3-channel video / image, the 1-channel critic map) and the other side has 64:
//
//   expand_conv_kernel   thin (float32 NCDHW, Cin <= 4) -> wide (bf16 NDHWC, Cout = 64): the head convolutions
//                        (modules/networks_3d.py:51,63 with in_channel = nc_im) and the data gradient of the tails
//   outer_corr_kernel    weight gradients of both narrow layer types (aten::convolution_backward grad_weight):
//                          head  dw[co][ci][t] = sum_v gy[v][co] * x[ci][v + t - pad]      (x thin, gy wide)
//                          tail  dw[c][ci][t]  = sum_u x[u][ci]  * gy[c][u - t + pad]      (x wide, gy thin)
//                        both are   OUT[k][j][t] = sum_p WIDE[p][k] * THIN[j][p + s_t]   with k < 64, j <= 4
//
// K = 27 * 3 = 81 per output is far below a tensor-core tile and the 64-channel side is read exactly once, so these are
// FMA / HBM bound; the kernels stage a halo tile of the thin tensor and the filter in shared memory and keep
// 24-128 accumulators per thread.  (The wide -> thin direction runs on tcgen05: conv_tc.cu, NOUT = 16.)
#include "common.cuh"
#include <cstdlib>

namespace hpvg {

bool narrow_wgrad_tc_supported(const ConvGeom& g, bool head);
int narrow_wgrad_tc_grid(const ConvGeom& g, bool head);
int narrow_wgrad_tc(const void* wide, const float* thin, bool head, const ConvGeom& g, float* partial, int grid, cudaStream_t st);
bool expand_tc_supported(const ConvGeom& g);
int expand_tc(const void* x, const float* w_tco, const float* bias, void* y, const ConvGeom& g, int act, float slope, float* stats,
              cudaStream_t st);

constexpr int EX_TH = 8, EX_TW = 32;             // output tile of the expand kernel: 1 x 8 x 32 voxels; thread = 4 voxels x 32 channels
constexpr int EX_HW = EX_TW + 4, EX_HH = EX_TH + 2;   // halo row padded to an even length (64-bit loads)

// ---------------------------------------------------------------------------------------------------------------
// thin -> wide convolution, Cout == 64
// ---------------------------------------------------------------------------------------------------------------
template <int KDT>
__global__ void __launch_bounds__(128, 3) expand_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ w_tco, const float* __restrict__ bias,
                                                             __nv_bfloat16* __restrict__ y, ConvGeom g, int transposed, int act,
                                                             float slope, float* __restrict__ stats) {
  pdl_enter();
  extern __shared__ float sm[];
  constexpr int TAPS = KDT * 9;
  float* ws = sm;                                   // [TAPS][Cin][64]
  float* xs = sm + TAPS * g.Cin * 64;               // [Cin][KDT][EX_HH][EX_HW]
  const int tiles_w = (g.Wo + EX_TW - 1) / EX_TW, tiles_h = (g.Ho + EX_TH - 1) / EX_TH;
  int b = blockIdx.x;
  const int w0 = (b % tiles_w) * EX_TW;
  b /= tiles_w;
  const int h0 = (b % tiles_h) * EX_TH;
  b /= tiles_h;
  const int od = b % g.Do;
  const int n = b / g.Do;
  const int tid = threadIdx.x;

  if (w_tco) {
    // filter already arranged [tap][ci][co] by hpvg_pack_weights_expand: a straight 128-bit copy.  (Gathering it from
    // the PyTorch layout in every block — 40 rounds of 324-byte-strided loads — cost more than the convolution itself.)
    const float4* src = reinterpret_cast<const float4*>(w_tco);
    float4* dst = reinterpret_cast<float4*>(ws);
    for (int i = tid; i < TAPS * g.Cin * 16; i += 128) dst[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < TAPS * g.Cin * 64; i += 128) {
      const int co = i & 63, ci = (i >> 6) % g.Cin, t = i / (64 * g.Cin);
      ws[i] = transposed ? w[((size_t)ci * 64 + co) * TAPS + (TAPS - 1 - t)] : w[((size_t)co * g.Cin + ci) * TAPS + t];
    }
  }
  const size_t in_sp = (size_t)g.Di * g.Hi * g.Wi;
  for (int i = tid; i < g.Cin * KDT * EX_HH * EX_HW; i += 128) {
    const int ww = i % EX_HW, hh = (i / EX_HW) % EX_HH, kd = (i / (EX_HW * EX_HH)) % KDT, ci = i / (EX_HW * EX_HH * KDT);
    const int id = od + kd - g.pad_d, ih = h0 + hh - g.pad, iw = w0 + ww - g.pad;
    float v = 0.f;
    if (id >= 0 && id < g.Di && ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi)
      v = __ldg(x + ((size_t)n * g.Cin + ci) * in_sp + ((size_t)id * g.Hi + ih) * g.Wi + iw);
    xs[i] = v;
  }
  __syncthreads();

  // thread = 4 consecutive output voxels (along w) x 32 output channels.  The loop is bound by the shared-memory pipe,
  // not by FMA issue: a (broadcast) 128-bit filter load occupies it for 4 cycles, so each one has to feed 16 FFMA (4 voxels
  // x 4 channels) to keep the 128 FMA lanes of the SM busy; 2 voxels x 64 channels fed 8 and ran at a quarter of peak.
  const int half = tid & 1, quad = tid >> 1;        // channel half, voxel quad 0..63
  const int ty = quad >> 3, tx = (quad & 7) * 4;    // tile row, first tile column
  float acc[4][32];
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[v][j] = 0.f;
  // the (ci, kd, kh) loops stay rolled: a fully unrolled kernel stalled on instruction fetch (ncu: stall_no_inst on top)
#pragma unroll 1
  for (int ci = 0; ci < g.Cin; ++ci) {
#pragma unroll 1
    for (int kd = 0; kd < KDT; ++kd) {
#pragma unroll 1
      for (int kh = 0; kh < 3; ++kh) {
        const float* xr = xs + ((ci * KDT + kd) * EX_HH + ty + kh) * EX_HW + tx;
        const float2 x01 = *reinterpret_cast<const float2*>(xr), x23 = *reinterpret_cast<const float2*>(xr + 2),
                     x45 = *reinterpret_cast<const float2*>(xr + 4);
        const float xv[6] = {x01.x, x01.y, x23.x, x23.y, x45.x, x45.y};
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4* w4 = reinterpret_cast<const float4*>(ws + (((kd * 3 + kh) * 3 + kw) * g.Cin + ci) * 64 + half * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 q = w4[j];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              acc[v][4 * j + 0] = fmaf(xv[kw + v], q.x, acc[v][4 * j + 0]);
              acc[v][4 * j + 1] = fmaf(xv[kw + v], q.y, acc[v][4 * j + 1]);
              acc[v][4 * j + 2] = fmaf(xv[kw + v], q.z, acc[v][4 * j + 2]);
              acc[v][4 * j + 3] = fmaf(xv[kw + v], q.w, acc[v][4 * j + 3]);
            }
          }
        }
      }
    }
  }

  // epilogue: bias, activation, bf16 rounding, store, BatchNorm sums of the stored values
  const int oh = h0 + ty;
  float ssum[32], ssq[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) ssum[j] = ssq[j] = 0.f;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int ow = w0 + tx + v;
    const bool ok = oh < g.Ho && ow < g.Wo;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float a = acc[v][j] + (bias ? __ldg(bias + half * 32 + j) : 0.f);
      if (act == HPVG_ACT_LRELU) a = a > 0.f ? a : a * slope;
      a = ok ? bf2f(f2bf(a)) : 0.f;
      acc[v][j] = a;
      ssum[j] += a;
      ssq[j] = fmaf(a, a, ssq[j]);
    }
    if (ok) {
      __nv_bfloat16* yp = y + ((((size_t)n * g.Do + od) * g.Ho + oh) * g.Wo + ow) * 64 + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 q;
        q.x = pack_bf16x2(acc[v][j + 0], acc[v][j + 1]); q.y = pack_bf16x2(acc[v][j + 2], acc[v][j + 3]);
        q.z = pack_bf16x2(acc[v][j + 4], acc[v][j + 5]); q.w = pack_bf16x2(acc[v][j + 6], acc[v][j + 7]);
        *reinterpret_cast<uint4*>(yp + j) = q;
      }
    }
  }
  if (stats) {
    __syncthreads();                 // the filter stage is dead: reuse it for the block reduction [2][64]
    float* red = sm;
    red[tid] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      // lanes of equal parity hold the same channel half: reduce over the 16 lanes that share it
      float s = ssum[j], s2 = ssq[j];
#pragma unroll
      for (int o = 16; o >= 2; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if ((tid & 31) < 2) {
        atomicAdd(red + half * 32 + j, s);
        atomicAdd(red + 64 + half * 32 + j, s2);
      }
    }
    __syncthreads();
    atomicAdd(stats + (size_t)n * g.stats_stride + tid, red[tid]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// thin -> wide convolution on the tensor cores: implicit GEMM [voxels x K] x [K x 64] with K = taps * Cin (81 -> 88),
// TF32 mma.sync.m16n8k8 (fp32 operands rounded to 10-bit mantissas, fp32 accumulation).  K = 81 is far below a tcgen05
// tile and the layer is bound by the 8.4 MB bf16 output, so the legacy warp-level MMA is the right size here: it removes
// the 5184 FFMA per voxel of the CUDA-core kernel above (kept as the fallback for Cin = 4 / odd shapes).
// Block = 8 warps, tile = 8 x 32 output voxels; warp = one tile row = 2 m16 tiles x 8 n8 tiles (64 accumulators).
// ---------------------------------------------------------------------------------------------------------------
constexpr int EM_KMAX = 88;                      // K padded to a multiple of 8 (Cin <= 3: 27 * 3 = 81)
constexpr int EM_BSTRIDE = 72;                   // floats per filter row in smem: 72 = 64 + 8 keeps the B loads conflict free
constexpr int EM_HW = EX_TW + 4;                 // halo row length

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KDT>
__global__ void __launch_bounds__(256, 2) expand_conv_mma_kernel(const float* __restrict__ x, const float* __restrict__ w_tco,
                                                                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                                 ConvGeom g, int act, float slope, float* __restrict__ stats) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t em_smem[];
  constexpr int TAPS = KDT * 9;
  const int K = TAPS * g.Cin;                                   // <= 81
  const int ksteps = (K + 7) >> 3;
  uint32_t* bs = reinterpret_cast<uint32_t*>(em_smem);          // [EM_KMAX][EM_BSTRIDE] tf32 filter, k = tap * Cin + ci
  uint32_t* xs = bs + EM_KMAX * EM_BSTRIDE;                     // [Cin][KDT][EX_HH][EM_HW] tf32 halo tile
  constexpr int MAIN_WORDS = EM_KMAX * EM_BSTRIDE + 3 * KDT * EX_HH * EM_HW;
  constexpr int STAGE_WORDS = 8 * 32 * 128 / 4;                 // the epilogue staging tile (8 warps x 32 voxels) overlays bs / xs
  int* koff = reinterpret_cast<int*>(em_smem) + (MAIN_WORDS > STAGE_WORDS ? MAIN_WORDS : STAGE_WORDS);   // [EM_KMAX]
  float* red = reinterpret_cast<float*>(koff + EM_KMAX);        // [128] BatchNorm partial sums
  float* bias_s = red + 128;                                    // [64]
  const int tiles_w = (g.Wo + EX_TW - 1) / EX_TW, tiles_h = (g.Ho + EX_TH - 1) / EX_TH;
  int b = blockIdx.x;
  const int w0 = (b % tiles_w) * EX_TW;
  b /= tiles_w;
  const int h0 = (b % tiles_h) * EX_TH;
  b /= tiles_h;
  const int od = b % g.Do;
  const int n = b / g.Do;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;

  // filter: [K][64] fp32 image (hpvg_pack_weights_expand), 128-bit copies into rows of EM_BSTRIDE words; rows >= K are zero
  // Every global load of the prologue is issued before the first one is consumed (fully unrolled register staging): the
  // block is otherwise a chain of dependent DRAM / L2 round trips (ncu: 41 % of the warp stalls were long-scoreboard waits).
  constexpr int F_ITERS = (EM_KMAX * 16 + 255) / 256;                  // filter: 6 x 128-bit loads per thread
  constexpr int H_ITERS = (3 * KDT * EX_HH + 7) / 8;                   // halo rows per warp (Cin <= 3)
  const size_t in_sp = (size_t)g.Di * g.Hi * g.Wi;
  float4 fv[F_ITERS];
#pragma unroll
  for (int it = 0; it < F_ITERS; ++it) {
    const int i = tid + it * 256;
    fv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < EM_KMAX * 16 && (i >> 4) < K) fv[it] = __ldg(reinterpret_cast<const float4*>(w_tco) + i);
  }
  // halo tile: one warp per (ci, kd, row), lanes along w (no per-element index arithmetic)
  float hv[H_ITERS][2];
  const int hrows = g.Cin * KDT * EX_HH;
#pragma unroll
  for (int it = 0; it < H_ITERS; ++it) {
    const int r = warp + 8 * it;
    hv[it][0] = hv[it][1] = 0.f;
    if (r < hrows) {
      const int hh = r % EX_HH, kd = (r / EX_HH) % KDT, ci = r / (EX_HH * KDT);
      const int id = od + kd - g.pad_d, ih = h0 + hh - g.pad;
      const bool row_ok = id >= 0 && id < g.Di && ih >= 0 && ih < g.Hi;
      const float* src = x + ((size_t)n * g.Cin + ci) * in_sp + ((size_t)(row_ok ? id : 0) * g.Hi + (row_ok ? ih : 0)) * g.Wi;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int ww = lane + 32 * half;
        const int iw = w0 + ww - g.pad;
        if (ww < EM_HW && row_ok && iw >= 0 && iw < g.Wi) hv[it][half] = __ldg(src + iw);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < F_ITERS; ++it) {
    const int i = tid + it * 256;
    if (i < EM_KMAX * 16) {
      const int k = i >> 4, c4 = i & 15;
      *reinterpret_cast<uint4*>(bs + k * EM_BSTRIDE + c4 * 4) =
          make_uint4(to_tf32(fv[it].x), to_tf32(fv[it].y), to_tf32(fv[it].z), to_tf32(fv[it].w));
    }
  }
#pragma unroll
  for (int it = 0; it < H_ITERS; ++it) {
    const int r = warp + 8 * it;
    if (r < hrows) {
      xs[r * EM_HW + lane] = to_tf32(hv[it][0]);
      if (lane + 32 < EM_HW) xs[r * EM_HW + lane + 32] = to_tf32(hv[it][1]);
    }
  }
  if (tid < EM_KMAX) {
    int o = 0;
    if (tid < K) {
      const int tap = tid / g.Cin, ci = tid % g.Cin;
      const int kd = tap / 9, kh = (tap % 9) / 3, kw = tap % 3;
      o = ((ci * KDT + kd) * EX_HH + kh) * EM_HW + kw;
    }
    koff[tid] = o;
  }
  if (tid < 128) red[tid] = 0.f;
  if (tid < 64) bias_s[tid] = bias ? __ldg(bias + tid) : 0.f;
  __syncthreads();

  // warp = tile row `warp`: 2 m16 tiles (columns 0-15, 16-31) x 8 n8 tiles = 64 accumulators per thread
  float acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
  const int rowbase = warp * EM_HW + gq;
#pragma unroll 1
  for (int s = 0; s < ksteps; ++s) {
    const int k0 = 8 * s + tq;
    const int o0 = koff[k0], o1 = koff[k0 + 4];
    uint32_t bf[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      bf[nt][0] = bs[k0 * EM_BSTRIDE + nt * 8 + gq];
      bf[nt][1] = bs[(k0 + 4) * EM_BSTRIDE + nt * 8 + gq];
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      uint32_t a[4];
      a[0] = xs[rowbase + mt * 16 + o0];
      a[1] = xs[rowbase + mt * 16 + 8 + o0];
      a[2] = xs[rowbase + mt * 16 + o1];
      a[3] = xs[rowbase + mt * 16 + 8 + o1];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) mma_tf32(acc[mt][nt], a, bf[nt][0], bf[nt][1]);
    }
  }
  __syncthreads();                       // filter / halo stages are dead: reuse them as the bf16 staging tile

  // epilogue: bias, activation, bf16 -> swizzled staging [32 voxels][64] per warp -> coalesced stores; BatchNorm sums
  uint8_t* stg = em_smem + warp * (32 * 128);
  const int oh = h0 + warp;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int vox = mt * 16 + gq + 8 * hrow;                  // voxel = column inside the warp's tile row
      const bool ok = oh < g.Ho && w0 + vox < g.Wo;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float a0 = acc[mt][nt][2 * hrow] + bias_s[nt * 8 + 2 * tq], a1 = acc[mt][nt][2 * hrow + 1] + bias_s[nt * 8 + 2 * tq + 1];
        if (act == HPVG_ACT_LRELU) {
          a0 = a0 > 0.f ? a0 : a0 * slope;
          a1 = a1 > 0.f ? a1 : a1 * slope;
        }
        if (!ok) a0 = a1 = 0.f;
        *reinterpret_cast<uint32_t*>(stg + vox * 128 + ((nt ^ (vox & 7)) << 4) + 4 * tq) = pack_bf16x2(a0, a1);
      }
    }
  }
  __syncwarp();
  if (oh < g.Ho) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int vox = (lane >> 3) + 4 * i, chunk = lane & 7;
      if (w0 + vox < g.Wo) {
        const uint4 v = *reinterpret_cast<const uint4*>(stg + vox * 128 + ((chunk ^ (vox & 7)) << 4));
        *reinterpret_cast<uint4*>(y + ((((size_t)n * g.Do + od) * g.Ho + oh) * g.Wo + w0 + vox) * 64 + chunk * 8) = v;
      }
    }
  }
  if (stats) {
    // lane -> channels 2*lane, 2*lane+1 over the warp's 32 staged voxels (invalid voxels were staged as zeros)
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    const int chunk = lane >> 2, within = (lane & 3) * 4;
#pragma unroll 8
    for (int vox = 0; vox < 32; ++vox) {
      const uint32_t pr = *reinterpret_cast<const uint32_t*>(stg + vox * 128 + ((chunk ^ (vox & 7)) << 4) + within);
      const float2 f = unpack_bf16x2(pr);
      s0 += f.x; s1 += f.y;
      q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
    }
    atomicAdd(red + 2 * lane, s0);
    atomicAdd(red + 2 * lane + 1, s1);
    atomicAdd(red + 64 + 2 * lane, q0);
    atomicAdd(red + 64 + 2 * lane + 1, q1);
    __syncthreads();
    if (tid < 128) atomicAdd(stats + (size_t)n * g.stats_stride + tid, red[tid]);
  }
}

static size_t expand_mma_smem(int KD) {
  const size_t main_bytes = (size_t)(EM_KMAX * EM_BSTRIDE + 3 * KD * EX_HH * EM_HW) * 4;
  const size_t stage_bytes = 8 * 32 * 128;                       // staging overlays the filter/halo
  return (main_bytes > stage_bytes ? main_bytes : stage_bytes) + EM_KMAX * 4 + 128 * 4 + 64 * 4;
}

bool expand_conv_supported(int x_fmt, int y_fmt, const ConvGeom& g, const void* mask_src) {
  return x_fmt == HPVG_FMT_NCDHW_F32 && y_fmt == HPVG_FMT_NDHWC_BF16 && g.Cin <= 4 && g.Cout == 64 && mask_src == nullptr;
}

// [tap][ci][co] float32 image of a narrow layer's filter (co = 64), `transposed` as in hpvg_conv_forward
__global__ void pack_expand_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int taps, int transposed) {
  pdl_enter();
  const int total = taps * Cin * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i & 63, ci = (i >> 6) % Cin, t = i / (64 * Cin);
    out[i] = transposed ? w[((size_t)ci * 64 + co) * taps + (taps - 1 - t)] : w[((size_t)co * Cin + ci) * taps + t];
  }
}

int pack_expand(const float* w, float* out, int Cin, int taps, int transposed, cudaStream_t st) {
  launch_k(pack_expand_kernel, (unsigned)cdiv(taps * Cin * 64, 256), 256, 0, st, w, out, Cin, taps, transposed);
  HPVG_CHECK_LAUNCH("pack_expand_kernel");
  return 0;
}

int expand_conv(const void* x, const float* w, const float* w_tco, const float* bias, void* y, const ConvGeom& g, int transposed,
                int act, float slope, float* stats, cudaStream_t st) {
  const int tiles_w = (int)cdiv(g.Wo, EX_TW), tiles_h = (int)cdiv(g.Ho, EX_TH);
  const long long blocks = (long long)g.N * g.Do * tiles_h * tiles_w;
  static const bool no_mma = getenv("HPVG_EXPAND_FMA") != nullptr;     // development aid: force the CUDA-core kernel
  // tcgen05 form (narrow_tc.cu): im2col operand built in shared memory, bf16 operands; HPVG_EXPAND_TC=0 keeps the TF32 mma.sync kernel
  if (w_tco != nullptr && !no_mma && expand_tc_supported(g)) return expand_tc(x, w_tco, bias, y, g, act, slope, stats, st);
  if (w_tco != nullptr && g.Cin <= 3 && !no_mma) {
    const size_t smem_mma = expand_mma_smem(g.KD);
    static std::atomic<unsigned long long> attr_mask{0};
    if (attr_pending(attr_mask)) {
      cudaFuncSetAttribute(expand_conv_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      cudaFuncSetAttribute(expand_conv_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      attr_set(attr_mask);
    }
    if (g.KD == 3)
      launch_k(expand_conv_mma_kernel<3>, (unsigned)blocks, 256, smem_mma, st, reinterpret_cast<const float*>(x), w_tco, bias,
                                                                        reinterpret_cast<__nv_bfloat16*>(y), g, act, slope, stats);
    else
      launch_k(expand_conv_mma_kernel<1>, (unsigned)blocks, 256, smem_mma, st, reinterpret_cast<const float*>(x), w_tco, bias,
                                                                        reinterpret_cast<__nv_bfloat16*>(y), g, act, slope, stats);
    HPVG_CHECK_LAUNCH("expand_conv_mma_kernel");
    return 0;
  }
  const size_t smem = ((size_t)g.taps * g.Cin * 64 + (size_t)g.Cin * g.KD * EX_HH * EX_HW) * sizeof(float);
  if (g.KD == 3)
    launch_k(expand_conv_kernel<3>, (unsigned)blocks, 128, smem, st, reinterpret_cast<const float*>(x), w, w_tco, bias, reinterpret_cast<__nv_bfloat16*>(y),
                                                              g, transposed, act, slope, stats);
  else
    launch_k(expand_conv_kernel<1>, (unsigned)blocks, 128, smem, st, reinterpret_cast<const float*>(x), w, w_tco, bias, reinterpret_cast<__nv_bfloat16*>(y),
                                                              g, transposed, act, slope, stats);
  HPVG_CHECK_LAUNCH("expand_conv_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// OUT[k][j][t] = sum_p WIDE[p][k] * THIN[j][p + s_t]        k < 64 (bf16 NDHWC), j < J <= 4 (float32 NCDHW)
// ---------------------------------------------------------------------------------------------------------------
constexpr int OC_TH = 8, OC_TW = 32;               // tile of WIDE voxels per block iteration: 1 x 8 x 32
constexpr int OC_HH = OC_TH + 2, OC_HW = OC_TW + 2;
constexpr int OC_ONES = OC_TH * OC_HW + OC_TW;     // a run of 1.0f long enough for every voxel base offset
constexpr int OC_THREADS = 256;

struct OuterCorrParams {
  const __nv_bfloat16* wide;   // [N][Dw][Hw][Ww][64]
  const float* thin;           // [N][J][Dt][Ht][Wt]
  float* partial;              // [grid][NC][4][256] per-block partial sums (thread-major: coalesced), reduced by a second kernel
  int N, J, KD, taps;
  int Dw, Hw, Ww, Dt, Ht, Wt;
  int sign, shift;             // THIN coordinate = p + sign * (tap offset) + shift   on every filtered axis
  int shift_d;                 // the same for the depth axis (0 when KD == 1)
  int want_sum;                // also produce sum_p WIDE[p][k] (bias gradient of a head layer) in the spare slot of slice 15
  long long tiles;
  int tiles_h, tiles_w;
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// thread = (kq = tid & 15 -> channels 4kq..4kq+3 of WIDE, sl = tid >> 4 -> (j, t) combinations sl, sl+16, ...); NC = combinations
// per thread.  Unused slots point at offset 0 and are dropped by the reduction; with want_sum the last slot of slice 15
// points into a run of ones, so that it accumulates the plain channel sum.
template <int KDT, int NC>
__global__ void __launch_bounds__(OC_THREADS, 2) outer_corr_kernel(const OuterCorrParams p) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t osm[];
  constexpr int TAPS = KDT * 9;
  constexpr int WT_BYTES = OC_TH * OC_TW * 128;
  float* th = reinterpret_cast<float*>(osm + WT_BYTES);                 // [J][KDT][OC_HH][OC_HW] then OC_ONES ones
  const int th_elems = p.J * KDT * OC_HH * OC_HW;
  const int tid = threadIdx.x;
  const int kq = tid & 15, sl = tid >> 4;
  const int ncomb = TAPS * p.J;
  uint32_t off[NC];                                                      // byte offsets into the THIN halo tile
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = sl + 16 * i;
    off[i] = 0;
    if (c < ncomb) {
      const int j = c / TAPS, t = c % TAPS;
      const int kd = t / 9, kh = (t % 9) / 3, kw = t % 3;
      const int dd = (KDT == 3) ? (p.sign * kd + (p.sign > 0 ? 0 : 2)) : 0;
      const int hh = p.sign * kh + (p.sign > 0 ? 0 : 2);
      const int ww = p.sign * kw + (p.sign > 0 ? 0 : 2);
      off[i] = 4u * (uint32_t)(((j * KDT + dd) * OC_HH + hh) * OC_HW + ww);
    } else if (p.want_sum && sl == 15 && i == NC - 1) {
      off[i] = 4u * (uint32_t)th_elems;
    }
  }
  float acc[NC][4];
#pragma unroll
  for (int i = 0; i < NC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  for (int i = tid; i < OC_ONES; i += OC_THREADS) th[th_elems + i] = 1.0f;

  const uint32_t s_wt = smem_u32(osm) + 8u * kq, s_th = smem_u32(th);
  const size_t thin_sp = (size_t)p.Dt * p.Ht * p.Wt;
  for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    long long b = tile;
    const int w0 = (int)(b % p.tiles_w) * OC_TW;
    b /= p.tiles_w;
    const int h0 = (int)(b % p.tiles_h) * OC_TH;
    b /= p.tiles_h;
    const int d0 = (int)(b % p.Dw);
    const int n = (int)(b / p.Dw);
    __syncthreads();
    // WIDE tile: 256 voxels x 128 B, zero outside the tensor
    for (int i = tid; i < OC_TH * OC_TW * 8; i += OC_THREADS) {
      const int chunk = i & 7, vox = i >> 3;
      const int hh = vox / OC_TW, ww = vox % OC_TW;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (h0 + hh < p.Hw && w0 + ww < p.Ww)
        v = __ldg(reinterpret_cast<const uint4*>(p.wide + ((((size_t)n * p.Dw + d0) * p.Hw + h0 + hh) * p.Ww + w0 + ww) * 64) + chunk);
      reinterpret_cast<uint4*>(osm)[i] = v;
    }
    // THIN halo tile: origin chosen so that voxel (hh, ww) with tap offset k in {0,1,2} reads [hh + k'][ww + k'], k' >= 0
    const int od = d0 + p.shift_d - (KDT == 3 ? (p.sign > 0 ? 0 : 2) : 0);
    const int oh = h0 + p.shift - (p.sign > 0 ? 0 : 2), ow = w0 + p.shift - (p.sign > 0 ? 0 : 2);
    for (int i = tid; i < th_elems; i += OC_THREADS) {
      const int ww = i % OC_HW, hh = (i / OC_HW) % OC_HH, dd = (i / (OC_HW * OC_HH)) % KDT, j = i / (OC_HW * OC_HH * KDT);
      const int id = od + dd, ih = oh + hh, iw = ow + ww;
      float v = 0.f;
      if (id >= 0 && id < p.Dt && ih >= 0 && ih < p.Ht && iw >= 0 && iw < p.Wt)
        v = __ldg(p.thin + ((size_t)n * p.J + j) * thin_sp + ((size_t)id * p.Ht + ih) * p.Wt + iw);
      th[i] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int hh = 0; hh < OC_TH; ++hh) {
      uint32_t a_wt = s_wt + (uint32_t)(hh * OC_TW) * 128u;
      uint32_t a_th = s_th + 4u * (uint32_t)(hh * OC_HW);
#pragma unroll 4
      for (int ww = 0; ww < OC_TW; ++ww, a_wt += 128u, a_th += 4u) {
        const uint2 raw = lds_u64(a_wt);
        const float2 a = unpack_bf16x2(raw.x), c = unpack_bf16x2(raw.y);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const float t = lds_f32(a_th + off[i]);
          acc[i][0] = fmaf(a.x, t, acc[i][0]);
          acc[i][1] = fmaf(a.y, t, acc[i][1]);
          acc[i][2] = fmaf(c.x, t, acc[i][2]);
          acc[i][3] = fmaf(c.y, t, acc[i][3]);
        }
      }
    }
  }
  float* dst = p.partial + (size_t)blockIdx.x * (NC * 4 * OC_THREADS) + tid;
#pragma unroll
  for (int i = 0; i < NC; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dst[(i * 4 + e) * OC_THREADS] = acc[i][e];
}

// dw[...] = sum over blocks of the partials.  A block reduces 32 consecutive outputs (one 128-byte line of every partial
// row) with 8 threads per output walking the rows 8 apart, so the ~300 dependent loads of a serial sum become ~37.
__global__ void __launch_bounds__(256) outer_corr_reduce_kernel(const float* __restrict__ partial, int blocks, int nc, int taps, int J,
                                                                int wide_is_gy, float* __restrict__ dw, float* __restrict__ wide_sum) {
  pdl_enter();
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int total = nc * 4 * OC_THREADS;
  const int idx = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (idx < total)
    for (int b = slice; b < blocks; b += 8) s += partial[(size_t)b * total + idx];
  red[slice][lane] = s;
  __syncthreads();
  if (slice != 0 || idx >= total) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) s += red[k][lane];
  const int tid = idx % OC_THREADS, e = (idx / OC_THREADS) & 3, i = idx / (4 * OC_THREADS);
  const int kq = tid & 15, sl = tid >> 4;
  const int c = sl + 16 * i, k = 4 * kq + e;
  const int ncomb = taps * J;
  const bool is_sum = wide_sum != nullptr && sl == 15 && i == nc - 1 && c >= ncomb;
  if (c >= ncomb && !is_sum) return;
  if (is_sum) {
    wide_sum[k] = s;
    return;
  }
  const int j = c / taps, t = c % taps;
  if (wide_is_gy)
    dw[((size_t)k * J + j) * taps + t] = s;
  else
    dw[((size_t)j * 64 + k) * taps + t] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// The same reduction on the tensor cores:  OUT[c][k] = sum_p A[c][p] * B[p][k]  with A[c][p] = THIN[j_c][p + s_c]
// (fp32 halo tile, rounded to bf16 as it is loaded into the fragment) and B = the bf16 WIDE tile, read transposed with
// ldmatrix.  mma.sync.m16n8k16 bf16, fp32 accumulation; 6 warps = 6 m16 tiles of (j, t) combinations (96 >= 27 * 3 + 1:
// the extra row multiplies by 1.0 and yields the channel sum of WIDE), each warp all 8 n8 tiles of the 64 channels.
// ---------------------------------------------------------------------------------------------------------------
constexpr int OM_WARPS = 6, OM_THREADS = 32 * OM_WARPS, OM_ROWS = 16 * OM_WARPS;

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <int KDT>
__global__ void __launch_bounds__(OM_THREADS, 2) outer_corr_mma_kernel(const OuterCorrParams p) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t osm[];
  constexpr int TAPS = KDT * 9;
  constexpr int WT_BYTES = OC_TH * OC_TW * 128;
  float* th = reinterpret_cast<float*>(osm + WT_BYTES);                 // [J][KDT][OC_HH][OC_HW] then OC_ONES ones
  const int th_elems = p.J * KDT * OC_HH * OC_HW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int ncomb = TAPS * p.J;
  // halo offsets of this thread's two combination rows (gq and gq + 8 of m-tile `warp`)
  int off[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = warp * 16 + gq + 8 * h;
    off[h] = 0;
    if (c < ncomb) {
      const int j = c / TAPS, t = c % TAPS;
      const int kd = t / 9, kh = (t % 9) / 3, kw = t % 3;
      const int dd = (KDT == 3) ? (p.sign * kd + (p.sign > 0 ? 0 : 2)) : 0;
      const int hh = p.sign * kh + (p.sign > 0 ? 0 : 2);
      const int ww = p.sign * kw + (p.sign > 0 ? 0 : 2);
      off[h] = ((j * KDT + dd) * OC_HH + hh) * OC_HW + ww;
    } else if (c == ncomb && p.want_sum) {
      off[h] = th_elems;                                               // a run of ones: OUT row = channel sum of WIDE
    }
  }
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  for (int i = tid; i < OC_ONES; i += OM_THREADS) th[th_elems + i] = 1.0f;

  const uint32_t s_wt = smem_u32(osm);
  // ldmatrix row address of this lane: voxel (lane & 7) + 8 * ((lane >> 3) & 1) of the k16 step, channels of n-tile (lane >> 4)
  const uint32_t ld_row = (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * 128);
  const uint32_t ld_swz = (uint32_t)(lane & 7), ld_hi = (uint32_t)(lane >> 4);   // swizzle phase = voxel & 7 = lane & 7 at every k-step
  const size_t thin_sp = (size_t)p.Dt * p.Ht * p.Wt;
  for (long long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    long long b = tile;
    const int w0 = (int)(b % p.tiles_w) * OC_TW;
    b /= p.tiles_w;
    const int h0 = (int)(b % p.tiles_h) * OC_TH;
    b /= p.tiles_h;
    const int d0 = (int)(b % p.Dw);
    const int n = (int)(b / p.Dw);
    __syncthreads();
    // all global loads of the tile are issued before the first shared-memory store (fully unrolled register staging): the
    // tile used to be a chain of ~25 dependent global round trips per warp (ncu: 25 us per launch at 9.6 % issue activity)
    constexpr int W_ITERS = (OC_TH * OC_TW * 8 + OM_THREADS - 1) / OM_THREADS;     // 11 x 128-bit loads per thread
    constexpr int T_ITERS = (3 * KDT * OC_HH + OM_WARPS - 1) / OM_WARPS;           // thin halo rows per warp (J <= 3)
    uint4 wv[W_ITERS];
#pragma unroll
    for (int it = 0; it < W_ITERS; ++it) {
      const int i = tid + it * OM_THREADS;
      const int chunk = i & 7, vox = i >> 3;
      const int hh = vox / OC_TW, ww = vox % OC_TW;
      wv[it] = make_uint4(0, 0, 0, 0);
      if (i < OC_TH * OC_TW * 8 && h0 + hh < p.Hw && w0 + ww < p.Ww)
        wv[it] = __ldg(reinterpret_cast<const uint4*>(p.wide + ((((size_t)n * p.Dw + d0) * p.Hw + h0 + hh) * p.Ww + w0 + ww) * 64) + chunk);
    }
    const int od = d0 + p.shift_d - (KDT == 3 ? (p.sign > 0 ? 0 : 2) : 0);
    const int oh = h0 + p.shift - (p.sign > 0 ? 0 : 2), ow = w0 + p.shift - (p.sign > 0 ? 0 : 2);
    const int trows = p.J * KDT * OC_HH;
    float tv[T_ITERS][2];
#pragma unroll
    for (int it = 0; it < T_ITERS; ++it) {
      const int r = warp + OM_WARPS * it;
      tv[it][0] = tv[it][1] = 0.f;
      if (r < trows) {
        const int hh = r % OC_HH, dd = (r / OC_HH) % KDT, j = r / (OC_HH * KDT);
        const int id = od + dd, ih = oh + hh;
        const bool row_ok = id >= 0 && id < p.Dt && ih >= 0 && ih < p.Ht;
        const float* src = p.thin + ((size_t)n * p.J + j) * thin_sp + ((size_t)(row_ok ? id : 0) * p.Ht + (row_ok ? ih : 0)) * p.Wt;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int ww = lane + 32 * half;
          const int iw = ow + ww;
          if (ww < OC_HW && row_ok && iw >= 0 && iw < p.Wt) tv[it][half] = __ldg(src + iw);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < W_ITERS; ++it) {
      const int i = tid + it * OM_THREADS;
      // 16-byte chunk index XOR (voxel & 7): the eight rows of an ldmatrix 8x8 tile (same chunk, consecutive voxels, 128 bytes
      // apart) then fall into eight different bank groups instead of one (ncu counted 2.8 M shared-memory bank conflicts per launch)
      if (i < OC_TH * OC_TW * 8) reinterpret_cast<uint4*>(osm)[(i & ~7) | ((i & 7) ^ ((i >> 3) & 7))] = wv[it];
    }
#pragma unroll
    for (int it = 0; it < T_ITERS; ++it) {
      const int r = warp + OM_WARPS * it;
      if (r < trows) {
        th[r * OC_HW + lane] = tv[it][0];
        if (lane + 32 < OC_HW) th[r * OC_HW + lane + 32] = tv[it][1];
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int ks = 0; ks < OC_TH * OC_TW / 16; ++ks) {              // 16 voxels per step: half a tile row
      const int hh = ks >> 1, ww = (ks & 1) * 16;
      const int base = hh * OC_HW + ww + 2 * tq;
      uint32_t a[4];
      {
        const float* r0 = th + base + off[0];
        const float* r1 = th + base + off[1];
        a[0] = pack_bf16x2(r0[0], r0[1]);
        a[1] = pack_bf16x2(r1[0], r1[1]);
        a[2] = pack_bf16x2(r0[8], r0[9]);
        a[3] = pack_bf16x2(r1[8], r1[9]);
      }
      const uint32_t wrow = s_wt + (uint32_t)(ks * 16) * 128u + ld_row;
#pragma unroll
      for (int np = 0; np < 4; ++np) {                              // n-tile pairs (2 np, 2 np + 1)
        uint32_t bfrag[4];
        ldmatrix_x4_trans(bfrag, wrow + (((uint32_t)(2 * np) + ld_hi) ^ ld_swz) * 16u);
        mma_bf16(acc[2 * np], a, bfrag[0], bfrag[1]);
        mma_bf16(acc[2 * np + 1], a, bfrag[2], bfrag[3]);
      }
    }
  }
  // per-block partial [OM_ROWS][64]
  float* dst = p.partial + (size_t)blockIdx.x * (OM_ROWS * 64);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<float2*>(dst + (warp * 16 + gq) * 64 + nt * 8 + 2 * tq) = make_float2(acc[nt][0], acc[nt][1]);
    *reinterpret_cast<float2*>(dst + (warp * 16 + gq + 8) * 64 + nt * 8 + 2 * tq) = make_float2(acc[nt][2], acc[nt][3]);
  }
}

// dw[...] = sum over blocks of partial[b][c][k]; a block owns 32 consecutive (c, k) entries (one 128-byte line of every
// partial row) and walks the ~300 rows with 32 threads per entry, so each thread has ~10 independent loads in flight
__global__ void __launch_bounds__(1024) outer_corr_mma_reduce_kernel(const float* __restrict__ partial, int blocks, int taps, int J,
                                                                     int wide_is_gy, float* __restrict__ dw, float* __restrict__ wide_sum,
                                                                     int tap_major) {
  pdl_enter();
  __shared__ float red[32][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int total = OM_ROWS * 64;
  const int idx = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (idx < total) {
#pragma unroll 4
    for (int b = slice; b < blocks; b += 32) s += __ldg(partial + (size_t)b * total + idx);
  }
  red[slice][lane] = s;
  __syncthreads();
  if (slice != 0 || idx >= total) return;
#pragma unroll
  for (int k = 1; k < 32; ++k) s += red[k][lane];
  const int c = idx >> 6, k = idx & 63;
  const int ncomb = taps * J;
  if (c == ncomb) {
    if (wide_sum) wide_sum[k] = s;
    return;
  }
  if (c > ncomb) return;
  // rows of a partial: (j, t) with t fastest (outer_corr_mma_kernel) or (t, j) with j fastest (narrow_wgrad_tc_kernel)
  const int j = tap_major ? c % J : c / taps, t = tap_major ? c / J : c % taps;
  if (wide_is_gy)
    dw[((size_t)k * J + j) * taps + t] = s;
  else
    dw[((size_t)j * 64 + k) * taps + t] = s;
}

bool narrow_wgrad_supported(int x_fmt, int gy_fmt, const ConvGeom& g) {
  if (x_fmt == HPVG_FMT_NCDHW_F32 && gy_fmt == HPVG_FMT_NDHWC_BF16) return g.Cin <= 4 && g.Cout == 64;
  if (x_fmt == HPVG_FMT_NDHWC_BF16 && gy_fmt == HPVG_FMT_NCDHW_F32) return g.Cin == 64 && g.Cout <= 4;
  return false;
}

static int oc_slots(const ConvGeom& g, bool head, bool want_sum) {
  const int ncomb = g.taps * (head ? g.Cin : g.Cout);
  int need = (int)cdiv(ncomb, 16);
  if (want_sum && 15 + 16 * (need - 1) < ncomb) need += 1;      // slice 15 has no spare slot: add one
  return need <= 2 ? 2 : (need <= 4 ? 4 : (need <= 6 ? 6 : 7));
}
static int oc_grid(const ConvGeom& g, bool head) {
  const long long tiles = head ? (long long)g.N * g.Do * cdiv(g.Ho, OC_TH) * cdiv(g.Wo, OC_TW)
                               : (long long)g.N * g.Di * cdiv(g.Hi, OC_TH) * cdiv(g.Wi, OC_TW);
  return (int)min(tiles, (long long)num_sms() * 2);
}

size_t narrow_wgrad_workspace(int x_fmt, const ConvGeom& g) {
  const bool head = x_fmt == HPVG_FMT_NCDHW_F32;
  const size_t tc = (size_t)narrow_wgrad_tc_grid(g, head) * 96 * 64 * sizeof(float);      // narrow_wgrad_tc_kernel: [grid][96][64]
  return max((size_t)oc_grid(g, head) * 7 * 4 * OC_THREADS * sizeof(float), tc);
}

template <int KDT, int NC>
static void oc_launch(const OuterCorrParams& p, int grid, size_t smem, cudaStream_t st) {
  static std::atomic<unsigned long long> attr_mask{0};
  if (attr_pending(attr_mask)) {
    cudaFuncSetAttribute(outer_corr_kernel<KDT, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_set(attr_mask);
  }
  launch_k(outer_corr_kernel<KDT, NC>, grid, OC_THREADS, smem, st, p);
}

// Two launches: per-block partial sums into `workspace` (deterministic, no atomics), then a fixed-order reduction into
// dw (and dbias_wide = channel sum of the wide gy for a head layer).
int narrow_wgrad(const void* x, int x_fmt, const void* gy, float* dw, float* dbias_wide, const ConvGeom& g, void* workspace,
                 size_t ws_bytes, cudaStream_t st) {
  OuterCorrParams p;
  const bool head = x_fmt == HPVG_FMT_NCDHW_F32;
  p.N = g.N;
  p.KD = g.KD;
  p.taps = g.taps;
  if (head) {   // WIDE = gy over output voxels, THIN = x at v + t - pad
    p.wide = reinterpret_cast<const __nv_bfloat16*>(gy);
    p.thin = reinterpret_cast<const float*>(x);
    p.J = g.Cin;
    p.Dw = g.Do; p.Hw = g.Ho; p.Ww = g.Wo;
    p.Dt = g.Di; p.Ht = g.Hi; p.Wt = g.Wi;
    p.sign = 1;
    p.shift = -g.pad;
    p.shift_d = -g.pad_d;
  } else {      // WIDE = x over input voxels, THIN = gy at u - t + pad
    p.wide = reinterpret_cast<const __nv_bfloat16*>(x);
    p.thin = reinterpret_cast<const float*>(gy);
    p.J = g.Cout;
    p.Dw = g.Di; p.Hw = g.Hi; p.Ww = g.Wi;
    p.Dt = g.Do; p.Ht = g.Ho; p.Wt = g.Wo;
    p.sign = -1;
    p.shift = g.pad;
    p.shift_d = g.pad_d;
  }
  p.want_sum = (head && dbias_wide) ? 1 : 0;
  p.tiles_h = (int)cdiv(p.Hw, OC_TH);
  p.tiles_w = (int)cdiv(p.Ww, OC_TW);
  p.tiles = (long long)p.N * p.Dw * p.tiles_h * p.tiles_w;
  const int grid = oc_grid(g, head);
  static const bool no_mma = getenv("HPVG_WGRAD_FMA") != nullptr;      // development aid: force the CUDA-core kernel
  if (!no_mma && narrow_wgrad_tc_supported(g, head)) {
    // tcgen05 form (narrow_tc.cu): the im2col operand tile of the head convolution read MN-major against the wide tile
    const int tgrid = narrow_wgrad_tc_grid(g, head);
    const size_t need_tc = (size_t)tgrid * OM_ROWS * 64 * sizeof(float);
    if (workspace == nullptr || ws_bytes < need_tc) {
      set_error("narrow_wgrad: workspace too small (%zu < %zu bytes)", ws_bytes, need_tc);
      return -1;
    }
    if (int rc = narrow_wgrad_tc(p.wide, p.thin, head, g, reinterpret_cast<float*>(workspace), tgrid, st)) return rc;
    launch_k(outer_corr_mma_reduce_kernel, (unsigned)cdiv(OM_ROWS * 64, 32), 1024, 0, st, reinterpret_cast<const float*>(workspace), tgrid, g.taps,
                                                                                p.J, head ? 1 : 0, dw, p.want_sum ? dbias_wide : nullptr, 1);
    HPVG_CHECK_LAUNCH("outer_corr_mma_reduce_kernel");
    return 0;
  }
  if (g.taps * p.J + 1 <= OM_ROWS && !no_mma) {
    const size_t need_mma = (size_t)grid * OM_ROWS * 64 * sizeof(float);
    if (workspace == nullptr || ws_bytes < need_mma) {
      set_error("narrow_wgrad: workspace too small (%zu < %zu bytes)", ws_bytes, need_mma);
      return -1;
    }
    p.partial = reinterpret_cast<float*>(workspace);
    const size_t smem_mma = (size_t)OC_TH * OC_TW * 128 + ((size_t)p.J * g.KD * OC_HH * OC_HW + OC_ONES) * sizeof(float);
    static std::atomic<unsigned long long> attr_mma{0};
    if (attr_pending(attr_mma)) {
      cudaFuncSetAttribute(outer_corr_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      cudaFuncSetAttribute(outer_corr_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      attr_set(attr_mma);
    }
    if (g.KD == 3)
      launch_k(outer_corr_mma_kernel<3>, grid, OM_THREADS, smem_mma, st, p);
    else
      launch_k(outer_corr_mma_kernel<1>, grid, OM_THREADS, smem_mma, st, p);
    HPVG_CHECK_LAUNCH("outer_corr_mma_kernel");
    launch_k(outer_corr_mma_reduce_kernel, (unsigned)cdiv(OM_ROWS * 64, 32), 1024, 0, st, p.partial, grid, g.taps, p.J, head ? 1 : 0, dw,
                                                                                p.want_sum ? dbias_wide : nullptr, 0);
    HPVG_CHECK_LAUNCH("outer_corr_mma_reduce_kernel");
    return 0;
  }
  const int nc = oc_slots(g, head, p.want_sum != 0);
  const size_t need = (size_t)grid * nc * 4 * OC_THREADS * sizeof(float);
  if (workspace == nullptr || ws_bytes < need) {
    set_error("narrow_wgrad: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return -1;
  }
  p.partial = reinterpret_cast<float*>(workspace);
  const size_t smem = (size_t)OC_TH * OC_TW * 128 + ((size_t)p.J * g.KD * OC_HH * OC_HW + OC_ONES) * sizeof(float);
  if (g.KD == 3) {
    if (nc == 2) oc_launch<3, 2>(p, grid, smem, st);
    else if (nc == 4) oc_launch<3, 4>(p, grid, smem, st);
    else if (nc == 6) oc_launch<3, 6>(p, grid, smem, st);
    else oc_launch<3, 7>(p, grid, smem, st);
  } else {
    if (nc == 2) oc_launch<1, 2>(p, grid, smem, st);
    else if (nc == 4) oc_launch<1, 4>(p, grid, smem, st);
    else if (nc == 6) oc_launch<1, 6>(p, grid, smem, st);
    else oc_launch<1, 7>(p, grid, smem, st);
  }
  HPVG_CHECK_LAUNCH("outer_corr_kernel");
  const int total = nc * 4 * OC_THREADS;
  launch_k(outer_corr_reduce_kernel, (unsigned)cdiv(total, 32), 256, 0, st, p.partial, grid, nc, g.taps, p.J, head ? 1 : 0, dw,
                                                                      p.want_sum ? dbias_wide : nullptr);
  HPVG_CHECK_LAUNCH("outer_corr_reduce_kernel");
  return 0;
}

}  // namespace hpvg

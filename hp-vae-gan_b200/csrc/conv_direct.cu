// CUDA-core convolution kernels: the narrow layers of the path (Cin = 3 image heads, Cout in {1,3} tails, and their
// data/weight gradients) are HBM/FMA bound with K = 81 or N <= 3, far below a tensor-core tile; they run here.
// The same kernels accept any channel count and serve as the in-library cross-check for the tcgen05 kernels
// (hpvg_set_conv_backend(HPVG_BACKEND_DIRECT)).
//
// Reference semantics: nn.Conv3d/Conv2d forward (modules/networks_3d.py:51,63,175,341,362) and
// aten::convolution_backward (grad_input as a flipped/transposed forward, grad_weight as a voxel reduction).
#include "common.cuh"

namespace hpvg {

template <int FMT>
struct Acc;
template <>
struct Acc<HPVG_FMT_NCDHW_F32> {
  typedef float T;
};
template <>
struct Acc<HPVG_FMT_NDHWC_BF16> {
  typedef __nv_bfloat16 T;
};

// ---------------------------------------------------------------------------------------------------------------
// forward-type convolution: one thread = one output voxel x CO_TILE output channels
// ---------------------------------------------------------------------------------------------------------------
template <int XFMT, int YFMT, int CO_TILE, int CI_CHUNK>
__global__ void __launch_bounds__(128) conv_direct_kernel(const void* __restrict__ xv, const float* __restrict__ w,
                                                          const float* __restrict__ bias, void* __restrict__ yv, ConvGeom g,
                                                          int transposed, int act, float slope, float* __restrict__ stats,
                                                          const __nv_bfloat16* __restrict__ mask_src) {
  pdl_enter();
  extern __shared__ float ws[];  // [taps][CI_CHUNK][CO_TILE]
  const int co0 = blockIdx.y * CO_TILE;
  const long long out_vox = (long long)g.N * g.Do * g.Ho * g.Wo;
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = v < out_vox;
  int n = 0, od = 0, oh = 0, ow = 0;
  if (live) {
    long long t = v;
    ow = (int)(t % g.Wo);
    t /= g.Wo;
    oh = (int)(t % g.Ho);
    t /= g.Ho;
    od = (int)(t % g.Do);
    n = (int)(t / g.Do);
  }
  float acc[CO_TILE];
#pragma unroll
  for (int j = 0; j < CO_TILE; ++j) acc[j] = 0.f;

  const size_t in_sp = (size_t)g.Di * g.Hi * g.Wi;
  for (int ci0 = 0; ci0 < g.Cin; ci0 += CI_CHUNK) {
    const int cin_here = min(CI_CHUNK, g.Cin - ci0);
    __syncthreads();
    for (int i = threadIdx.x; i < g.taps * CI_CHUNK * CO_TILE; i += blockDim.x) {
      int j = i % CO_TILE;
      int c = (i / CO_TILE) % CI_CHUNK;
      int t = i / (CO_TILE * CI_CHUNK);
      float val = 0.f;
      int co = co0 + j, ci = ci0 + c;
      if (co < g.Cout && c < cin_here) {
        val = transposed ? w[((size_t)ci * g.Cout + co) * g.taps + (g.taps - 1 - t)] : w[((size_t)co * g.Cin + ci) * g.taps + t];
      }
      ws[i] = val;
    }
    __syncthreads();
    if (!live) continue;
    for (int kd = 0; kd < g.KD; ++kd) {
      const int id = od + kd - g.pad_d;
      if (id < 0 || id >= g.Di) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh + kh - g.pad;
        if (ih < 0 || ih >= g.Hi) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = ow + kw - g.pad;
          if (iw < 0 || iw >= g.Wi) continue;
          const int t = (kd * 3 + kh) * 3 + kw;
          const float* wt = ws + (size_t)t * CI_CHUNK * CO_TILE;
          const size_t sp = ((size_t)id * g.Hi + ih) * g.Wi + iw;
          if (XFMT == HPVG_FMT_NCDHW_F32) {
            const float* xp = reinterpret_cast<const float*>(xv) + ((size_t)n * g.Cin + ci0) * in_sp + sp;
            for (int c = 0; c < cin_here; ++c) {
              const float x = __ldg(xp + (size_t)c * in_sp);
              const float4* w4 = reinterpret_cast<const float4*>(wt + c * CO_TILE);
#pragma unroll
              for (int j = 0; j < CO_TILE / 4; ++j) {
                float4 q = w4[j];
                acc[4 * j + 0] = fmaf(x, q.x, acc[4 * j + 0]);
                acc[4 * j + 1] = fmaf(x, q.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(x, q.z, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(x, q.w, acc[4 * j + 3]);
              }
            }
          } else {
            const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(xv) + ((size_t)n * in_sp + sp) * g.Cin + ci0;
            if ((g.Cin & 7) == 0 && (CI_CHUNK & 7) == 0) {
              for (int c8 = 0; c8 < cin_here; c8 += 8) {
                uint4 raw = __ldg(reinterpret_cast<const uint4*>(xp + c8));
                float xs[8];
                float2 p;
                p = unpack_bf16x2(raw.x); xs[0] = p.x; xs[1] = p.y;
                p = unpack_bf16x2(raw.y); xs[2] = p.x; xs[3] = p.y;
                p = unpack_bf16x2(raw.z); xs[4] = p.x; xs[5] = p.y;
                p = unpack_bf16x2(raw.w); xs[6] = p.x; xs[7] = p.y;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const float4* w4 = reinterpret_cast<const float4*>(wt + (c8 + u) * CO_TILE);
#pragma unroll
                  for (int j = 0; j < CO_TILE / 4; ++j) {
                    float4 q = w4[j];
                    acc[4 * j + 0] = fmaf(xs[u], q.x, acc[4 * j + 0]);
                    acc[4 * j + 1] = fmaf(xs[u], q.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(xs[u], q.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(xs[u], q.w, acc[4 * j + 3]);
                  }
                }
              }
            } else {
              for (int c = 0; c < cin_here; ++c) {
                const float x = bf2f(xp[c]);
                const float4* w4 = reinterpret_cast<const float4*>(wt + c * CO_TILE);
#pragma unroll
                for (int j = 0; j < CO_TILE / 4; ++j) {
                  float4 q = w4[j];
                  acc[4 * j + 0] = fmaf(x, q.x, acc[4 * j + 0]);
                  acc[4 * j + 1] = fmaf(x, q.y, acc[4 * j + 1]);
                  acc[4 * j + 2] = fmaf(x, q.z, acc[4 * j + 2]);
                  acc[4 * j + 3] = fmaf(x, q.w, acc[4 * j + 3]);
                }
              }
            }
          }
        }
      }
    }
  }

  // ---- epilogue: bias, LeakyReLU-derivative mask, activation, store, BatchNorm sums ----
  const size_t out_sp = (size_t)g.Do * g.Ho * g.Wo;
  const size_t osp = ((size_t)od * g.Ho + oh) * g.Wo + ow;
#pragma unroll
  for (int j = 0; j < CO_TILE; ++j) {
    const int co = co0 + j;
    float r = acc[j];
    if (live && co < g.Cout) {
      if (bias) r += bias[co];
      if (mask_src) {
        float m = bf2f(mask_src[((size_t)n * out_sp + osp) * g.Cout + co]);
        r *= (m > 0.f) ? 1.f : slope;
      }
      if (act == HPVG_ACT_LRELU) r = r > 0.f ? r : r * slope;
      if (YFMT == HPVG_FMT_NDHWC_BF16) r = bf2f(f2bf(r));
    } else {
      r = 0.f;
    }
    acc[j] = r;
  }
  if (live) {
    if (YFMT == HPVG_FMT_NCDHW_F32) {
      float* yp = reinterpret_cast<float*>(yv) + (size_t)n * g.Cout * out_sp + osp;
#pragma unroll
      for (int j = 0; j < CO_TILE; ++j)
        if (co0 + j < g.Cout) yp[(size_t)(co0 + j) * out_sp] = acc[j];
    } else {
      __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(yv) + ((size_t)n * out_sp + osp) * g.Cout + co0;
      if ((g.Cout & 7) == 0 && CO_TILE >= 8) {
#pragma unroll
        for (int j = 0; j < CO_TILE; j += 8) {
          if (co0 + j < g.Cout) {
            uint4 q;
            q.x = pack_bf16x2(acc[j + 0], acc[j + 1]);
            q.y = pack_bf16x2(acc[j + 2], acc[j + 3]);
            q.z = pack_bf16x2(acc[j + 4], acc[j + 5]);
            q.w = pack_bf16x2(acc[j + 6], acc[j + 7]);
            *reinterpret_cast<uint4*>(yp + j) = q;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < CO_TILE; ++j)
          if (co0 + j < g.Cout) yp[j] = f2bf(acc[j]);
      }
    }
  }
  if (stats) {
    // warp-shuffle reduction per channel, one atomic per warp and channel (dead threads contribute zeros)
#pragma unroll
    for (int j = 0; j < CO_TILE; ++j) {
      const int co = co0 + j;
      if (co >= g.Cout) break;
      float s = warp_sum(acc[j]);
      float s2 = warp_sum(acc[j] * acc[j]);
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + co, s);
        atomicAdd(stats + g.Cout + co, s2);
      }
    }
  }
}

template <int XFMT, int YFMT, int CO_TILE, int CI_CHUNK>
static int launch_direct(const void* x, const float* w, const float* bias, void* y, const ConvGeom& g, int transposed, int act,
                         float slope, float* stats, const void* mask_src, cudaStream_t st) {
  const long long out_vox = (long long)g.N * g.Do * g.Ho * g.Wo;
  dim3 grid((unsigned)cdiv(out_vox, 128), (unsigned)cdiv(g.Cout, CO_TILE));
  size_t smem = (size_t)g.taps * CI_CHUNK * CO_TILE * sizeof(float);
  launch_k(conv_direct_kernel<XFMT, YFMT, CO_TILE, CI_CHUNK>, grid, 128, smem, st, 
      x, w, bias, y, g, transposed, act, slope, stats, reinterpret_cast<const __nv_bfloat16*>(mask_src));
  HPVG_CHECK_LAUNCH("conv_direct_kernel");
  return 0;
}

template <int XFMT, int YFMT>
static int dispatch_direct(const void* x, const float* w, const float* bias, void* y, const ConvGeom& g, int transposed, int act,
                           float slope, float* stats, const void* mask_src, cudaStream_t st) {
  // weight stage stays under the 48 KB default: taps*CI_CHUNK*CO_TILE*4 <= 27*{4*64, 16*16, 64*4}*4 = 27.6 KB
  if (g.Cout <= 4) return launch_direct<XFMT, YFMT, 4, 64>(x, w, bias, y, g, transposed, act, slope, stats, mask_src, st);
  if (g.Cin <= 4) return launch_direct<XFMT, YFMT, 64, 4>(x, w, bias, y, g, transposed, act, slope, stats, mask_src, st);
  return launch_direct<XFMT, YFMT, 16, 16>(x, w, bias, y, g, transposed, act, slope, stats, mask_src, st);
}

int conv_direct(const void* x, int x_fmt, const float* w, const float* bias, void* y, int y_fmt, const ConvGeom& g, int transposed,
                int act, float slope, float* stats, const void* mask_src, cudaStream_t st) {
  if (x_fmt == HPVG_FMT_NCDHW_F32 && y_fmt == HPVG_FMT_NCDHW_F32)
    return dispatch_direct<HPVG_FMT_NCDHW_F32, HPVG_FMT_NCDHW_F32>(x, w, bias, y, g, transposed, act, slope, stats, mask_src, st);
  if (x_fmt == HPVG_FMT_NCDHW_F32 && y_fmt == HPVG_FMT_NDHWC_BF16)
    return dispatch_direct<HPVG_FMT_NCDHW_F32, HPVG_FMT_NDHWC_BF16>(x, w, bias, y, g, transposed, act, slope, stats, mask_src, st);
  if (x_fmt == HPVG_FMT_NDHWC_BF16 && y_fmt == HPVG_FMT_NCDHW_F32)
    return dispatch_direct<HPVG_FMT_NDHWC_BF16, HPVG_FMT_NCDHW_F32>(x, w, bias, y, g, transposed, act, slope, stats, mask_src, st);
  return dispatch_direct<HPVG_FMT_NDHWC_BF16, HPVG_FMT_NDHWC_BF16>(x, w, bias, y, g, transposed, act, slope, stats, mask_src, st);
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient: thread = one (co, ci) pair with all taps in registers, block = one chunk of output voxels
// ---------------------------------------------------------------------------------------------------------------
template <int XFMT, int GFMT, int TAPS>
__global__ void __launch_bounds__(256) wgrad_direct_kernel(const void* __restrict__ xv, const void* __restrict__ gv,
                                                           float* __restrict__ dw, ConvGeom g, long long vox_per_block) {
  pdl_enter();
  const int pairs = g.Cin * g.Cout;
  const int p = blockIdx.y * blockDim.x + threadIdx.x;
  const bool live = p < pairs;
  int co, ci;
  if (GFMT == HPVG_FMT_NDHWC_BF16) {  // co contiguous in gy: lanes walk co
    co = live ? p % g.Cout : 0;
    ci = live ? p / g.Cout : 0;
  } else {
    ci = live ? p % g.Cin : 0;
    co = live ? p / g.Cin : 0;
  }
  float acc[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) acc[t] = 0.f;
  const long long out_vox = (long long)g.N * g.Do * g.Ho * g.Wo;
  const long long v0 = (long long)blockIdx.x * vox_per_block;
  const long long v1 = min(out_vox, v0 + vox_per_block);
  const size_t in_sp = (size_t)g.Di * g.Hi * g.Wi, out_sp = (size_t)g.Do * g.Ho * g.Wo;
  constexpr int KD = TAPS / 9;
  for (long long v = v0; v < v1; ++v) {
    long long t = v;
    const int ow = (int)(t % g.Wo);
    t /= g.Wo;
    const int oh = (int)(t % g.Ho);
    t /= g.Ho;
    const int od = (int)(t % g.Do);
    const int n = (int)(t / g.Do);
    const size_t osp = ((size_t)od * g.Ho + oh) * g.Wo + ow;
    float gy;
    if (GFMT == HPVG_FMT_NDHWC_BF16)
      gy = bf2f(reinterpret_cast<const __nv_bfloat16*>(gv)[((size_t)n * out_sp + osp) * g.Cout + co]);
    else
      gy = reinterpret_cast<const float*>(gv)[((size_t)n * g.Cout + co) * out_sp + osp];
#pragma unroll
    for (int kd = 0; kd < KD; ++kd) {
      const int id = od + kd - g.pad_d;
      if (id < 0 || id >= g.Di) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh + kh - g.pad;
        if (ih < 0 || ih >= g.Hi) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = ow + kw - g.pad;
          if (iw < 0 || iw >= g.Wi) continue;
          const size_t sp = ((size_t)id * g.Hi + ih) * g.Wi + iw;
          float x;
          if (XFMT == HPVG_FMT_NDHWC_BF16)
            x = bf2f(reinterpret_cast<const __nv_bfloat16*>(xv)[((size_t)n * in_sp + sp) * g.Cin + ci]);
          else
            x = reinterpret_cast<const float*>(xv)[((size_t)n * g.Cin + ci) * in_sp + sp];
          acc[(kd * 3 + kh) * 3 + kw] = fmaf(gy, x, acc[(kd * 3 + kh) * 3 + kw]);
        }
      }
    }
  }
  if (live) {
    float* o = dw + ((size_t)co * g.Cin + ci) * TAPS;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) atomicAdd(o + t, acc[t]);
  }
}

template <int XFMT, int GFMT>
static int launch_wgrad_direct(const void* x, const void* gy, float* dw, const ConvGeom& g, cudaStream_t st) {
  const int pairs = g.Cin * g.Cout;
  const int pair_blocks = (int)cdiv(pairs, 256);
  const long long out_vox = (long long)g.N * g.Do * g.Ho * g.Wo;
  long long chunks = cdiv((long long)num_sms() * 4, pair_blocks);
  chunks = max(1LL, min(chunks, cdiv(out_vox, 16)));
  const long long vpb = cdiv(out_vox, chunks);
  chunks = cdiv(out_vox, vpb);
  dim3 grid((unsigned)chunks, (unsigned)pair_blocks);
  if (g.taps == 27)
    launch_k(wgrad_direct_kernel<XFMT, GFMT, 27>, grid, 256, 0, st, x, gy, dw, g, vpb);
  else
    launch_k(wgrad_direct_kernel<XFMT, GFMT, 9>, grid, 256, 0, st, x, gy, dw, g, vpb);
  HPVG_CHECK_LAUNCH("wgrad_direct_kernel");
  return 0;
}

int wgrad_direct(const void* x, int x_fmt, const void* gy, int gy_fmt, float* dw, const ConvGeom& g, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)g.Cout * g.Cin * g.taps * sizeof(float), st);
  if (e != cudaSuccess) {
    set_error("wgrad_direct: memset failed: %s", cudaGetErrorString(e));
    return -2;
  }
  if (x_fmt == HPVG_FMT_NCDHW_F32 && gy_fmt == HPVG_FMT_NCDHW_F32)
    return launch_wgrad_direct<HPVG_FMT_NCDHW_F32, HPVG_FMT_NCDHW_F32>(x, gy, dw, g, st);
  if (x_fmt == HPVG_FMT_NCDHW_F32 && gy_fmt == HPVG_FMT_NDHWC_BF16)
    return launch_wgrad_direct<HPVG_FMT_NCDHW_F32, HPVG_FMT_NDHWC_BF16>(x, gy, dw, g, st);
  if (x_fmt == HPVG_FMT_NDHWC_BF16 && gy_fmt == HPVG_FMT_NCDHW_F32)
    return launch_wgrad_direct<HPVG_FMT_NDHWC_BF16, HPVG_FMT_NCDHW_F32>(x, gy, dw, g, st);
  return launch_wgrad_direct<HPVG_FMT_NDHWC_BF16, HPVG_FMT_NDHWC_BF16>(x, gy, dw, g, st);
}

}  // namespace hpvg

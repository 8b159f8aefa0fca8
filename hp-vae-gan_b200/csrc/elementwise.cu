// HBM-bound kernels of the path: BatchNorm(+LeakyReLU) forward/backward, LeakyReLU backward, trilinear/bilinear
// resize (+noise) and its adjoint, tanh(+residual), the VAE head (reparameterisation, KL), the WGAN-GP penalty,
// layout conversion, weight packing and spectral normalisation.  All are one pass over their operands with
// 128-bit accesses where the layout allows; reductions use warp shuffles and one atomic per warp or block.
#include "common.cuh"
#include <cstdlib>

namespace hpvg {

static inline int ew_blocks(long long work_items, int threads) {
  long long b = cdiv(work_items, threads);
  long long cap = (long long)num_sms() * 16;
  return (int)max(1LL, min(b, cap));
}

// ===============================================================================================================
// BatchNorm (training) + LeakyReLU        reference: modules/networks_3d.py:54-56 (nn.BatchNorm3d, nn.LeakyReLU(0.2))
// ===============================================================================================================
__global__ void bn_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ nbt, float momentum, float eps, long long count,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_invstd, int C) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const double inv = 1.0 / (double)count;
    const double mean = (double)stats[c] * inv;
    double var = (double)stats[C + c] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale_shift[c] = sc;
    scale_shift[C + c] = beta[c] - (float)mean * sc;
    mean_invstd[c] = (float)mean;
    mean_invstd[C + c] = invstd;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
      const double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
  if (c == 0 && nbt) nbt[0] += 1;
}

__global__ void __launch_bounds__(256) bn_apply_lrelu_kernel(const uint4* __restrict__ y, const float* __restrict__ scale_shift,
                                                             uint4* __restrict__ out, long long nvec, int C, float slope) {
  pdl_enter();
  extern __shared__ float ss[];  // scale[C], shift[C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) ss[i] = scale_shift[i];
  __syncthreads();
  const int cvec = C >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cvec) << 3;
    uint4 v = __ldg(y + i);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = unpack_bf16x2(w[k]);
      float a = fmaf(f.x, ss[c0 + 2 * k], ss[C + c0 + 2 * k]);
      float b = fmaf(f.y, ss[c0 + 2 * k + 1], ss[C + c0 + 2 * k + 1]);
      a = a > 0.f ? a : a * slope;
      b = b > 0.f ? b : b * slope;
      w[k] = pack_bf16x2(a, b);
    }
    out[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// bn_finalize + bn_apply_lrelu in one launch: every block derives scale/shift for all channels from the sums (C <= 256
// channels: trivial), block 0 also publishes them for the backward pass and updates the running statistics.
__global__ void __launch_bounds__(256) bn_finalize_apply_lrelu_kernel(const uint4* __restrict__ y, const float* __restrict__ stats,
                                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                      float* __restrict__ running_mean, float* __restrict__ running_var,
                                                                      long long* __restrict__ nbt, float momentum, float eps,
                                                                      long long count, float* __restrict__ scale_shift,
                                                                      float* __restrict__ mean_invstd, uint4* __restrict__ out,
                                                                      long long nvec, int C, float slope) {
  pdl_enter();
  extern __shared__ float ss[];  // scale[C], shift[C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double inv = 1.0 / (double)count;
    const double mean = (double)stats[c] * inv;
    double var = (double)stats[C + c] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    const float sh = beta[c] - (float)mean * sc;
    ss[c] = sc;
    ss[C + c] = sh;
    if (blockIdx.x == 0) {
      scale_shift[c] = sc;
      scale_shift[C + c] = sh;
      mean_invstd[c] = (float)mean;
      mean_invstd[C + c] = invstd;
      if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      if (running_var) {
        const double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
      if (c == 0 && nbt) nbt[0] += 1;
    }
  }
  __syncthreads();
  const int cvec = C >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cvec) << 3;
    uint4 v = __ldg(y + i);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = unpack_bf16x2(w[k]);
      float a = fmaf(f.x, ss[c0 + 2 * k], ss[C + c0 + 2 * k]);
      float b = fmaf(f.y, ss[c0 + 2 * k + 1], ss[C + c0 + 2 * k + 1]);
      a = a > 0.f ? a : a * slope;
      b = b > 0.f ? b : b * slope;
      w[k] = pack_bf16x2(a, b);
    }
    out[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// per-channel sums of bf16 vectors a thread produced (its 8-channel group is fixed because the grid stride is a multiple
// of C/8): block reduction through shared memory, then one atomic per channel and block
__device__ __forceinline__ void block_channel_sum(const float (&acc)[8], int c0, int C, float* red, float* __restrict__ out) {
  for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    // lanes of a warp cover 32 / (C/8) rows x (C/8) channel groups: combine the rows that share a group first
    float v = acc[k];
    for (int o = 16; o >= (C >> 3); o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) < (C >> 3)) atomicAdd(red + c0 + k, v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, red[i]);
}

// thread = (row lane, 8-channel group); per-thread partial sums in registers, block reduction through shared memory
__global__ void __launch_bounds__(256) bn_lrelu_bwd_reduce_kernel(const uint4* __restrict__ y, const uint4* __restrict__ gout,
                                                                  const float* __restrict__ scale_shift,
                                                                  const float* __restrict__ mean_invstd, float* __restrict__ sums,
                                                                  long long nvox, int C, float slope,
                                                                  const uint8_t* __restrict__ mask) {
  pdl_enter();
  extern __shared__ float sm[];  // [4*C] params, then [rows][2*C] partials
  float* prm = sm;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    prm[i] = scale_shift[i];
    prm[2 * C + i] = mean_invstd[i];
  }
  __syncthreads();
  const int cvec = C >> 3;
  const int rows = blockDim.x / cvec;
  const int cg = threadIdx.x % cvec, rl = threadIdx.x / cvec;
  const int c0 = cg << 3;
  float s0[8], s1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s0[k] = s1[k] = 0.f;
  if (rl < rows) {
    for (long long r = (long long)blockIdx.x * rows + rl; r < nvox; r += (long long)gridDim.x * rows) {
      const uint4 yv = __ldg(y + r * cvec + cg);
      const uint4 gv = __ldg(gout + r * cvec + cg);
      // mask (fused forward): bit b of byte [voxel][channel / 8] = the fp32 pre-activation of channel 8*(c/8) + b was positive
      const uint32_t mb = mask ? (uint32_t)__ldg(mask + r * cvec + cg) : 0u;
      const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 yf = unpack_bf16x2(yw[k]), gf = unpack_bf16x2(gw[k]);
        const float ye[2] = {yf.x, yf.y}, ge[2] = {gf.x, gf.y};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = c0 + 2 * k + e;
          const float z = fmaf(ye[e], prm[c], prm[C + c]);
          const bool pos = mask ? ((mb >> (2 * k + e)) & 1u) != 0u : z > 0.f;
          const float dz = pos ? ge[e] : ge[e] * slope;
          const float xh = (ye[e] - prm[2 * C + c]) * prm[3 * C + c];
          s0[2 * k + e] += dz;
          s1[2 * k + e] = fmaf(dz, xh, s1[2 * k + e]);
        }
      }
    }
  }
  float* part = sm + 4 * C;  // [rows][2*C]
  if (rl < rows) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      part[rl * 2 * C + c0 + k] = s0[k];
      part[rl * 2 * C + C + c0 + k] = s1[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += part[r * 2 * C + i];
    atomicAdd(sums + i, s);
  }
}

__global__ void __launch_bounds__(256) bn_lrelu_bwd_apply_kernel(const uint4* __restrict__ y, const uint4* __restrict__ gout,
                                                                 const float* __restrict__ scale_shift,
                                                                 const float* __restrict__ mean_invstd, const float* __restrict__ sums,
                                                                 uint4* __restrict__ gy, float* __restrict__ dgamma,
                                                                 float* __restrict__ dbeta, long long nvox, int C, float slope,
                                                                 float* __restrict__ chsum, const uint8_t* __restrict__ mask) {
  pdl_enter();
  extern __shared__ float prm[];  // scale, shift, mean, invstd, m0 = sum dz / M, m1 = sum dz*xhat / M, [C] reduction scratch
  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int my_c0 = 0;
  const float invM = 1.f / (float)nvox;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    prm[i] = scale_shift[i];
    prm[2 * C + i] = mean_invstd[i];
    prm[4 * C + i] = sums[i] * invM;
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      if (dbeta) dbeta[i] = sums[i];
      if (dgamma) dgamma[i] = sums[C + i];
    }
  }
  __syncthreads();
  const int cvec = C >> 3;
  const long long nvec = nvox * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cvec) << 3;
    const uint4 yv = __ldg(y + i), gv = __ldg(gout + i);
    const uint32_t mb = mask ? (uint32_t)__ldg(mask + i) : 0u;
    const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 yf = unpack_bf16x2(yw[k]), gf = unpack_bf16x2(gw[k]);
      const float ye[2] = {yf.x, yf.y}, ge[2] = {gf.x, gf.y};
      float res[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = c0 + 2 * k + e;
        const float z = fmaf(ye[e], prm[c], prm[C + c]);
        const bool pos = mask ? ((mb >> (2 * k + e)) & 1u) != 0u : z > 0.f;
        const float dz = pos ? ge[e] : ge[e] * slope;
        const float xh = (ye[e] - prm[2 * C + c]) * prm[3 * C + c];
        res[e] = prm[c] * (dz - prm[4 * C + c] - xh * prm[5 * C + c]);
      }
      ow[k] = pack_bf16x2(res[0], res[1]);
      if (chsum) {   // sum of the STORED (bf16) gradient: the bias gradient of the convolution in front of this BatchNorm
        const float2 st = unpack_bf16x2(ow[k]);
        csum[2 * k] += st.x;
        csum[2 * k + 1] += st.y;
      }
    }
    my_c0 = c0;
    gy[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
  if (chsum) {
    __syncthreads();
    block_channel_sum(csum, ((threadIdx.x % cvec) << 3), C, prm + 6 * C, chsum);
    (void)my_c0;
  }
}


// BatchNorm running statistics applied AFTER the fact, in order: entry e updates running_mean / running_var / num_batches_tracked
// of its layer from the batch mean / invstd that the layer's forward saved (mean_invstd), exactly as bn_finalize would have:
//   running <- (1 - momentum) * running + momentum * {mean, unbiased variance}.
// The generator's reconstruction and sampling passes run concurrently on two streams and go through the SAME BatchNorm layers;
// their updates are logged and applied by one launch in the reference's order (all of 'rec', then all of 'rand') instead of
// racing on the buffers.  One thread per channel walks the entries sequentially: entries of the same layer stay ordered.
struct BnRunBatch {
  int n;
  float* rm[HPVG_BN_LOG_MAX];
  float* rv[HPVG_BN_LOG_MAX];
  long long* nbt[HPVG_BN_LOG_MAX];
  const float* mean_invstd[HPVG_BN_LOG_MAX];
  long long count[HPVG_BN_LOG_MAX];
  int C[HPVG_BN_LOG_MAX];
  float momentum[HPVG_BN_LOG_MAX];
  float eps[HPVG_BN_LOG_MAX];
};

__global__ void __launch_bounds__(256) bn_running_update_kernel(const BnRunBatch b) {
  pdl_enter();
  // block = one LAYER: the first entry that names a buffer owns it and applies every later entry of the same buffer, in order;
  // blocks whose entry is not the first of its layer leave at once.  Layers are independent, so they update in parallel.
  const int first = blockIdx.x;
  for (int e = 0; e < first; ++e)
    if (b.mean_invstd[e] != nullptr && b.rm[e] == b.rm[first] && b.rv[e] == b.rv[first] && b.nbt[e] == b.nbt[first]) return;
  const int c = threadIdx.x;
  for (int e = first; e < b.n; ++e) {
    if (e != first && !(b.rm[e] == b.rm[first] && b.rv[e] == b.rv[first] && b.nbt[e] == b.nbt[first])) continue;
    if (c < b.C[e]) {
      const float mean = b.mean_invstd[e][c];
      const float invstd = b.mean_invstd[e][b.C[e] + c];
      double var = 1.0 / ((double)invstd * (double)invstd) - (double)b.eps[e];
      if (var < 0.0) var = 0.0;
      const long long n = b.count[e];
      const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
      const float m = b.momentum[e];
      if (b.rm[e]) b.rm[e][c] = (1.f - m) * b.rm[e][c] + m * mean;
      if (b.rv[e]) b.rv[e][c] = (1.f - m) * b.rv[e][c] + m * (float)unbiased;
    }
    if (c == 0 && b.nbt[e]) b.nbt[e][0] += 1;
  }
}

// BatchNorm + LeakyReLU backward in ONE launch (reduce + apply): every CTA loads its share of y and gout ONCE into shared
// memory, adds its partial sums of dz and dz * xhat to the global totals, crosses a grid-wide barrier, and applies
// gy = scale * (dz - mean(dz) - xhat * mean(dz * xhat)) from shared memory.  HBM traffic: y + gout read once and gy written once
// (25 MB per 64-channel layer at 16 x 64 x 64) instead of y + gout twice (42 MB) in two launches.  Persistent grid of at most one CTA
// per SM, cooperative launch; eligible when the tensor fits the SMs' shared memory (host check), else the two-launch path runs.
// `sums`: float32 [3C + 1], zeroed by the caller: [0,C) sum dz, [C,2C) sum dz*xhat, [2C,3C) bias-gradient sums, [3C] barrier counter.
constexpr int BNB_THREADS = 512;
__global__ void __launch_bounds__(BNB_THREADS, 1) bn_lrelu_bwd_fused_kernel(const uint4* __restrict__ y, const uint4* __restrict__ gout,
                                                                            const float* __restrict__ scale_shift,
                                                                            const float* __restrict__ mean_invstd, float* __restrict__ sums,
                                                                            uint4* __restrict__ gy, float* __restrict__ dgamma,
                                                                            float* __restrict__ dbeta, long long nvox, int C, float slope,
                                                                            int want_chsum, const uint8_t* __restrict__ mask,
                                                                            int rows_per_cta) {
  extern __shared__ __align__(16) uint8_t bnb_smem[];
  const int cvec = C >> 3;
  const long long row0 = (long long)blockIdx.x * rows_per_cta;
  const int rows = (int)max(0LL, min((long long)rows_per_cta, nvox - row0));
  const int nvec = rows * cvec;
  const uint32_t mbar = smem_u32(bnb_smem);                        // 16 bytes: the bulk copies' completion barrier
  float* prm = reinterpret_cast<float*>(bnb_smem + 16);            // scale, shift, mean, invstd [4C]; m0, m1 [2C]; red [3C]
  float* red = prm + 6 * C;
  uint4* sy = reinterpret_cast<uint4*>(bnb_smem + 16 + 9 * C * sizeof(float));
  uint4* sg = sy + (size_t)rows_per_cta * cvec;
  uint8_t* sm = reinterpret_cast<uint8_t*>(sg + (size_t)rows_per_cta * cvec);
  const uint4* yb = y + row0 * cvec;
  const uint4* gb = gout + row0 * cvec;
  const uint8_t* mb_base = mask ? mask + row0 * cvec : nullptr;
  // The CTA's rows are contiguous in memory: y and gout arrive as a few 1-D bulk asynchronous copies (cp.async.bulk, at most
  // 32 KB each) that complete on one mbarrier — every byte of the CTA's 2 x 113 KB is in flight at once.  (Per-thread 16-byte
  // loads in a loop left one round trip per iteration exposed: 36 k active cycles per launch in ncu for 25 MB of traffic.)
  if (threadIdx.x == 0) {
    mbar_init(mbar, 1);
    mbar_fence_init();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t bytes = (uint32_t)nvec * 16u;
    mbar_expect_tx(mbar, 2u * bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t n = min(32768u, bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sy) + off),
                   "l"(reinterpret_cast<const uint8_t*>(yb) + off), "r"(n), "r"(mbar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sg) + off),
                   "l"(reinterpret_cast<const uint8_t*>(gb) + off), "r"(n), "r"(mbar)
                   : "memory");
    }
  }
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    prm[i] = scale_shift[i];
    prm[2 * C + i] = mean_invstd[i];
  }
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) red[i] = 0.f;
  if (mb_base)
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) sm[v] = __ldg(mb_base + v);
  __syncthreads();          // barrier initialised (thread 0) before anyone waits on it; parameters and mask bytes in place
  mbar_wait(mbar, 0);
  const int cg = threadIdx.x % cvec;      // blockDim.x is a multiple of cvec: a thread's 8-channel group is fixed
  const int c0 = cg << 3;
  // a thread's 8 channels never change: their parameters live in registers (read from shared memory per element they were
  // 4 - 6 loads per element and pass — ncu: 7.4 M warp instructions per launch, MIO-throttle / short-scoreboard stalls)
  float p_sc[8], p_sh[8], p_mu[8], p_is[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    p_sc[k] = prm[c0 + k];
    p_sh[k] = prm[C + c0 + k];
    p_mu[k] = prm[2 * C + c0 + k];
    p_is[k] = prm[3 * C + c0 + k];
  }
  float s0[8], s1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s0[k] = s1[k] = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint4 yv = sy[v], gv = sg[v];
    const uint32_t mb = mb_base ? (uint32_t)sm[v] : 0u;
    const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 yf = unpack_bf16x2(yw[k]), gf = unpack_bf16x2(gw[k]);
      const float ye[2] = {yf.x, yf.y}, ge[2] = {gf.x, gf.y};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * k + e;
        const bool pos = mb_base ? ((mb >> j) & 1u) != 0u : fmaf(ye[e], p_sc[j], p_sh[j]) > 0.f;
        const float dz = pos ? ge[e] : ge[e] * slope;
        const float xh = (ye[e] - p_mu[j]) * p_is[j];
        s0[j] += dz;
        s1[j] = fmaf(dz, xh, s1[j]);
      }
    }
  }
  // lanes that share a channel group (32 / cvec rows per warp) combine first, then one shared-memory atomic per channel and warp
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float a = s0[k], b = s1[k];
    for (int o = 16; o >= cvec; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) < cvec) {
      atomicAdd(red + c0 + k, a);
      atomicAdd(red + C + c0 + k, b);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, red[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) grid_barrier_arrive_and_wait(reinterpret_cast<unsigned*>(sums + 3 * C), gridDim.x);
  __syncthreads();
  const float invM = 1.f / (float)nvox;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const float tot = __ldcg(sums + i);
    prm[4 * C + i] = tot * invM;
    if (blockIdx.x == 0) {
      if (i < C) { if (dbeta) dbeta[i] = tot; }
      else if (dgamma) dgamma[i - C] = tot;
    }
  }
  __syncthreads();
  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float p_m0[8], p_m1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    p_m0[k] = prm[4 * C + c0 + k];
    p_m1[k] = prm[5 * C + c0 + k];
  }
  uint4* gyb = gy + row0 * cvec;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint4 yv = sy[v], gv = sg[v];
    const uint32_t mb = mb_base ? (uint32_t)sm[v] : 0u;
    const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 yf = unpack_bf16x2(yw[k]), gf = unpack_bf16x2(gw[k]);
      const float ye[2] = {yf.x, yf.y}, ge[2] = {gf.x, gf.y};
      float res[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * k + e;
        const bool pos = mb_base ? ((mb >> j) & 1u) != 0u : fmaf(ye[e], p_sc[j], p_sh[j]) > 0.f;
        const float dz = pos ? ge[e] : ge[e] * slope;
        const float xh = (ye[e] - p_mu[j]) * p_is[j];
        res[e] = p_sc[j] * (dz - p_m0[j] - xh * p_m1[j]);
      }
      ow[k] = pack_bf16x2(res[0], res[1]);
      if (want_chsum) {     // sum of the STORED (bf16) gradient: the bias gradient of the convolution in front of this BatchNorm
        const float2 st = unpack_bf16x2(ow[k]);
        csum[2 * k] += st.x;
        csum[2 * k + 1] += st.y;
      }
    }
    gyb[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
  if (want_chsum) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a = csum[k];
      for (int o = 16; o >= cvec; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if ((threadIdx.x & 31) < cvec) atomicAdd(red + 2 * C + c0 + k, a);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(sums + 2 * C + i, red[2 * C + i]);
  }
}

__global__ void __launch_bounds__(256) lrelu_bwd_kernel(const uint4* __restrict__ gout, const uint4* __restrict__ outv,
                                                        uint4* __restrict__ gz, long long nvec, float slope, int C,
                                                        float* __restrict__ chsum) {
  pdl_enter();
  extern __shared__ float red[];   // [C] when chsum
  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 g = __ldg(gout + i), o = __ldg(outv + i);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, ow[4] = {o.x, o.y, o.z, o.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 gf = unpack_bf16x2(gw[k]), of = unpack_bf16x2(ow[k]);
      r[k] = pack_bf16x2(of.x > 0.f ? gf.x : gf.x * slope, of.y > 0.f ? gf.y : gf.y * slope);
      if (chsum) {
        const float2 st = unpack_bf16x2(r[k]);
        csum[2 * k] += st.x;
        csum[2 * k + 1] += st.y;
      }
    }
    gz[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
  if (chsum) block_channel_sum(csum, (int)((threadIdx.x % (C >> 3)) << 3), C, red, chsum);
}

// BatchNorm (batch statistics) + LeakyReLU with statistics PER SAMPLE: stats is [N][2C] (hpvg_conv_forward_ex with
// stats_per_sample), blockIdx.y = sample.  A batched forward then computes what N separate batch-1 forwards compute — the
// reference generates every draw with batch size 1 (train_video.py:226-235), and G stays in train mode — without giving up
// the larger launches.  Inference only: nothing is saved for a backward pass and running statistics are not advanced.
__global__ void __launch_bounds__(256) bn_apply_lrelu_per_sample_kernel(const uint4* __restrict__ y, const float* __restrict__ stats,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                        float eps, long long count, uint4* __restrict__ out,
                                                                        long long nvec, int C, float slope) {
  pdl_enter();
  extern __shared__ float ss[];  // scale[C], shift[C]
  const int n = blockIdx.y;
  const float* st = stats + (size_t)n * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double inv = 1.0 / (double)count;
    const double mean = (double)st[c] * inv;
    double var = (double)st[C + c] * inv - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    ss[c] = sc;
    ss[C + c] = beta[c] - (float)mean * sc;
  }
  __syncthreads();
  const int cvec = C >> 3;
  const uint4* yn = y + (size_t)n * nvec;
  uint4* on = out + (size_t)n * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cvec) << 3;
    uint4 v = __ldg(yn + i);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = unpack_bf16x2(w[k]);
      float a = fmaf(f.x, ss[c0 + 2 * k], ss[C + c0 + 2 * k]);
      float b = fmaf(f.y, ss[c0 + 2 * k + 1], ss[C + c0 + 2 * k + 1]);
      a = a > 0.f ? a : a * slope;
      b = b > 0.f ? b : b * slope;
      w[k] = pack_bf16x2(a, b);
    }
    on[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ===============================================================================================================
// linear resize, align_corners=True      reference: utils/images.py:22-26 (F.interpolate trilinear), :9-19 (bilinear)
// ===============================================================================================================
__device__ __forceinline__ void lin_src(int o, float scale, int in, int& i0, int& i1, float& l0, float& l1) {
  const float s = scale * (float)o;   // align_corners=True: src = o * (in-1)/(out-1)
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                           const float* __restrict__ noise, float amp, int NC, int Di, int Hi, int Wi,
                                                           int Do, int Ho, int Wo, float sd, float sh, float sw) {
  pdl_enter();
  const long long total = (long long)NC * Do * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int ow = (int)(t % Wo); t /= Wo;
    const int oh = (int)(t % Ho); t /= Ho;
    const int od = (int)(t % Do);
    const long long nc = t / Do;
    int d0, d1, h0, h1, w0, w1;
    float ld0, ld1, lh0, lh1, lw0, lw1;
    lin_src(od, sd, Di, d0, d1, ld0, ld1);
    lin_src(oh, sh, Hi, h0, h1, lh0, lh1);
    lin_src(ow, sw, Wi, w0, w1, lw0, lw1);
    const float* p = x + nc * (long long)Di * Hi * Wi;
#define X_(d, h, w) __ldg(p + ((long long)(d) * Hi + (h)) * Wi + (w))
    float v = ld0 * (lh0 * (lw0 * X_(d0, h0, w0) + lw1 * X_(d0, h0, w1)) + lh1 * (lw0 * X_(d0, h1, w0) + lw1 * X_(d0, h1, w1))) +
              ld1 * (lh0 * (lw0 * X_(d1, h0, w0) + lw1 * X_(d1, h0, w1)) + lh1 * (lw0 * X_(d1, h1, w0) + lw1 * X_(d1, h1, w1)));
#undef X_
    if (noise) v = fmaf(amp, __ldg(noise + i), v);
    out[i] = v;
  }
}

// adjoint as a gather: every input voxel sums the outputs that interpolate from it (no atomics, deterministic)
__device__ __forceinline__ void adj_range(int i, float scale, int out, int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out - 1; return; }
  lo = (int)floorf(((float)i - 1.f) / scale) - 1;
  hi = (int)ceilf(((float)i + 1.f) / scale) + 1;
  if (lo < 0) lo = 0;
  if (hi > out - 1) hi = out - 1;
}
__device__ __forceinline__ float adj_weight(int o, int i, float scale, int in) {
  int i0, i1;
  float l0, l1;
  lin_src(o, scale, in, i0, i1, l0, l1);
  float w = 0.f;
  if (i0 == i) w += l0;
  if (i1 == i) w += l1;
  return w;
}

// exact range of the outputs with a non-zero weight on input index i (adj_range is conservative: up to 4 extra candidates per
// axis, i.e. 360 instead of ~45 candidate outputs per input voxel for the 6 x 54 x 54 -> 16 x 64 x 64 resize)
__device__ __forceinline__ void adj_trim(int i, float scale, int in, int out, int& lo, int& hi) {
  adj_range(i, scale, out, lo, hi);
  while (lo <= hi && adj_weight(lo, i, scale, in) == 0.f) ++lo;
  while (hi >= lo && adj_weight(hi, i, scale, in) == 0.f) --hi;
}

__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gx, int NC, int Di, int Hi,
                                                           int Wi, int Do, int Ho, int Wo, float sd, float sh, float sw) {
  pdl_enter();
  const long long total = (long long)NC * Di * Hi * Wi;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int iw = (int)(t % Wi); t /= Wi;
    const int ih = (int)(t % Hi); t /= Hi;
    const int id = (int)(t % Di);
    const long long nc = t / Di;
    int dlo, dhi, hlo, hhi, wlo, whi;
    adj_trim(id, sd, Di, Do, dlo, dhi);
    adj_trim(ih, sh, Hi, Ho, hlo, hhi);
    adj_trim(iw, sw, Wi, Wo, wlo, whi);
    const float* g = gout + nc * (long long)Do * Ho * Wo;
    float acc = 0.f;
    for (int od = dlo; od <= dhi; ++od) {
      const float wd = adj_weight(od, id, sd, Di);
      if (wd == 0.f) continue;
      for (int oh = hlo; oh <= hhi; ++oh) {
        const float wh = adj_weight(oh, ih, sh, Hi);
        if (wh == 0.f) continue;
        float row = 0.f;
        for (int ow = wlo; ow <= whi; ++ow) {
          const float ww = adj_weight(ow, iw, sw, Wi);
          if (ww != 0.f) row = fmaf(ww, __ldg(g + ((long long)od * Ho + oh) * Wo + ow), row);
        }
        acc = fmaf(wd * wh, row, acc);
      }
    }
    gx[i] = acc;
  }
}

// ===============================================================================================================
// tanh(+residual)                          reference: modules/networks_3d.py:377,404
// ===============================================================================================================
__global__ void tanh_add_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = a[i];
    if (b) v += b[i];
    out[i] = tanhf(v);
  }
}
__global__ void tanh_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out, float* __restrict__ g, long long n) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float o = out[i];
    g[i] = gout[i] * (1.f - o * o);
  }
}

// ===============================================================================================================
// VAE head                                 reference: modules/networks_3d.py:29-35, modules/losses.py:7-9
// ===============================================================================================================
// thread = (voxel, 8-channel group): NDHWC bf16 mu/logvar, NCDHW fp32 eps (torch's normal_() order), NDHWC bf16 z
__global__ void __launch_bounds__(256) reparam_fwd_kernel(const uint4* __restrict__ mu, const uint4* __restrict__ logvar,
                                                          const float* __restrict__ eps, uint4* __restrict__ z, int N, int C,
                                                          long long S) {
  pdl_enter();
  const int cvec = C >> 3;
  const long long total = (long long)N * S * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // adjacent threads walk voxels (coalesced eps reads), channel groups are the slow index
    const long long s = i % S;
    const int cg = (int)((i / S) % cvec);
    const long long n = i / (S * cvec);
    const long long vi = (n * S + s) * cvec + cg;
    const uint4 m = __ldg(mu + vi), l = __ldg(logvar + vi);
    const uint32_t mw[4] = {m.x, m.y, m.z, m.w}, lw[4] = {l.x, l.y, l.z, l.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 mf = unpack_bf16x2(mw[k]), lf = unpack_bf16x2(lw[k]);
      const float e0 = __ldg(eps + (n * C + (cg * 8 + 2 * k)) * S + s);
      const float e1 = __ldg(eps + (n * C + (cg * 8 + 2 * k + 1)) * S + s);
      r[k] = pack_bf16x2(fmaf(e0, expf(0.5f * lf.x), mf.x), fmaf(e1, expf(0.5f * lf.y), mf.y));
    }
    z[vi] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

__global__ void __launch_bounds__(256) reparam_bwd_kernel(const uint4* __restrict__ gz, const uint4* __restrict__ logvar,
                                                          const float* __restrict__ eps, uint4* __restrict__ gmu,
                                                          uint4* __restrict__ glogvar, int N, int C, long long S) {
  pdl_enter();
  const int cvec = C >> 3;
  const long long total = (long long)N * S * cvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i % S;
    const int cg = (int)((i / S) % cvec);
    const long long n = i / (S * cvec);
    const long long vi = (n * S + s) * cvec + cg;
    const uint4 g = __ldg(gz + vi), l = __ldg(logvar + vi);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, lw[4] = {l.x, l.y, l.z, l.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 gf = unpack_bf16x2(gw[k]), lf = unpack_bf16x2(lw[k]);
      const float e0 = __ldg(eps + (n * C + (cg * 8 + 2 * k)) * S + s);
      const float e1 = __ldg(eps + (n * C + (cg * 8 + 2 * k + 1)) * S + s);
      r[k] = pack_bf16x2(gf.x * e0 * 0.5f * expf(0.5f * lf.x), gf.y * e1 * 0.5f * expf(0.5f * lf.y));
    }
    gmu[vi] = g;
    glogvar[vi] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

__global__ void __launch_bounds__(256) kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, float* __restrict__ out,
                                                     long long n, float inv_n) {
  pdl_enter();
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = mu[i], l = logvar[i];
    acc += -0.5f * (1.f + l - m * m - expf(l));
  }
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, s * inv_n);
}
__global__ void kl_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ mu, const float* __restrict__ logvar,
                              float* __restrict__ gmu, float* __restrict__ glogvar, long long n, float inv_n) {
  pdl_enter();
  const float g = gout[0] * inv_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    gmu[i] = g * mu[i];
    glogvar[i] = g * 0.5f * (expf(logvar[i]) - 1.f);
  }
}

// ===============================================================================================================
// WGAN-GP penalty                          reference: modules/utils.py:18
// ===============================================================================================================
__global__ void __launch_bounds__(256) gp_fwd_kernel(const float* __restrict__ g, float* __restrict__ out, int N, int C, long long S,
                                                     float coef) {
  pdl_enter();
  __shared__ float red[8];
  float acc = 0.f;
  const long long total = (long long)N * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / S, s = i % S;
    float q = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = g[(n * C + c) * S + s];
      q = fmaf(v, v, q);
    }
    const float d = sqrtf(q) - 1.f;
    acc = fmaf(d, d, acc);
  }
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, s * coef);
}
__global__ void gp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ g, float* __restrict__ gg, int N, int C,
                              long long S, float coef) {
  pdl_enter();
  const float go = gout[0] * coef * 2.f;
  const long long total = (long long)N * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / S, s = i % S;
    float q = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = g[(n * C + c) * S + s];
      q = fmaf(v, v, q);
    }
    const float nrm = sqrtf(q);
    const float f = nrm > 0.f ? go * (nrm - 1.f) / nrm : 0.f;
    for (int c = 0; c < C; ++c) gg[(n * C + c) * S + s] = f * g[(n * C + c) * S + s];
  }
}

// ===============================================================================================================
// layout / dtype conversion, lerp, channel sums, weight packing
// ===============================================================================================================
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int C, long long S) {
  pdl_enter();
  const long long total = (long long)N * S * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long s = (i / C) % S, n = i / (C * S);
    dst[i] = f2bf(src[(n * C + c) * S + s]);
  }
}
__global__ void ndhwc_to_ncdhw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int N, int C, long long S) {
  pdl_enter();
  const long long total = (long long)N * S * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i % S;
    const int c = (int)((i / S) % C);
    const long long n = i / (S * C);
    dst[i] = bf2f(src[(n * S + s) * C + c]);
  }
}
__global__ void lerp_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                            const float* __restrict__ alpha_ptr, long long n) {
  pdl_enter();
  const float alpha = __ldg(alpha_ptr);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = alpha * a[i] + (1.f - alpha) * b[i];
}

// NDHWC bf16: block = 256 threads = (256/C rows) x C channels
__global__ void __launch_bounds__(256) channel_sum_ndhwc_kernel(const __nv_bfloat16* __restrict__ t, float* __restrict__ out, long long rows,
                                                                int C) {
  pdl_enter();
  __shared__ float part[256];
  const int rpb = blockDim.x / C;
  const int c = threadIdx.x % C, rl = threadIdx.x / C;
  float acc = 0.f;
  if (rl < rpb)
    for (long long r = (long long)blockIdx.x * rpb + rl; r < rows; r += (long long)gridDim.x * rpb) acc += bf2f(t[r * C + c]);
  part[threadIdx.x] = (rl < rpb) ? acc : 0.f;
  __syncthreads();
  if (threadIdx.x < C) {
    float s = 0.f;
    for (int r = 0; r < rpb; ++r) s += part[r * C + threadIdx.x];
    atomicAdd(out + threadIdx.x, s);
  }
}
// NCDHW fp32: blockIdx.y = n*C + c, blockIdx.x = chunk of the spatial run
__global__ void __launch_bounds__(256) channel_sum_ncdhw_kernel(const float* __restrict__ t, float* __restrict__ out, int C, long long S) {
  pdl_enter();
  __shared__ float red[8];
  const int c = blockIdx.y % C;
  const float* p = t + (long long)blockIdx.y * S;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) acc += p[i];
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) atomicAdd(out + c, s);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin, int taps, int transposed,
                                    const float* __restrict__ sigma, int rows_per_tap) {
  pdl_enter();
  const float inv = sigma ? 1.f / sigma[0] : 1.f;
  const long long total = (long long)taps * rows_per_tap * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % rows_per_tap);
    const int t = (int)(i / ((long long)Cin * rows_per_tap));
    float v = 0.f;
    if (co < Cout) v = transposed ? w[((size_t)ci * Cout + co) * taps + (taps - 1 - t)] : w[((size_t)co * Cin + ci) * taps + t];
    out[i] = f2bf(v * inv);
  }
}

// Both bf16 operand images (forward and data-gradient form) of up to HPVG_SN_MAX_LAYERS wide layers in ONE launch
// (blockIdx.y = layer): the critic repacks W / sigma in every pass, 2 launches per layer otherwise.  Reads are coalesced in
// the PyTorch layout [Cout][Cin][taps]; each element goes to fwd[t][co][ci] and tr[taps-1-t][ci][co], exactly the values
// pack_weights_kernel writes for transposed = 0 / 1.
struct PackBatch {
  const float* w[HPVG_SN_MAX_LAYERS];
  __nv_bfloat16* fwd[HPVG_SN_MAX_LAYERS];
  __nv_bfloat16* tr[HPVG_SN_MAX_LAYERS];
  int cout[HPVG_SN_MAX_LAYERS], cin[HPVG_SN_MAX_LAYERS], taps[HPVG_SN_MAX_LAYERS];
};
__global__ void __launch_bounds__(256) pack_pair_batched_kernel(const PackBatch b) {
  pdl_enter();
  const int l = blockIdx.y;
  const int Cout = b.cout[l], Cin = b.cin[l], taps = b.taps[l];
  const long long total = (long long)Cout * Cin * taps;
  const float* w = b.w[l];
  __nv_bfloat16* fwd = b.fwd[l];
  __nv_bfloat16* tr = b.tr[l];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((long long)taps * Cin));
    const __nv_bfloat16 v = f2bf(w[i]);
    if (fwd) fwd[((size_t)t * Cout + co) * Cin + ci] = v;
    if (tr) tr[((size_t)(taps - 1 - t) * Cin + ci) * Cout + co] = v;
  }
}

// ===============================================================================================================
// spectral normalisation                   reference: nn.utils.spectral_norm as used at modules/networks_3d.py:63
// ===============================================================================================================
// The two vector norms are summed by ONE block in a fixed order (sn_norms_256) rather than accumulated with atomics: sigma
// must be bit-reproducible, because a one-ulp change of W / sigma flips bf16 roundings of the packed weights and from there
// LeakyReLU signs several layers on (measured: 1.6e-2 run-to-run scatter of the critic's input gradient with atomics).
// v_raw[k] = sum_r W[r][k] u[r]
__global__ void __launch_bounds__(256) sn_wtu_kernel(const float* __restrict__ W, const float* __restrict__ u, float* __restrict__ v_raw,
                                                     int Cout, int K) {
  pdl_enter();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc = 0.f;
  for (int r = 0; r < Cout; ++r) acc = fmaf(W[(size_t)r * K + k], u[r], acc);
  v_raw[k] = acc;
}
// t_raw[r] = sum_k W[r][k] vec[k]
__global__ void __launch_bounds__(256) sn_wv_kernel(const float* __restrict__ W, const float* __restrict__ vec, float* __restrict__ t_raw,
                                                    int K) {
  pdl_enter();
  __shared__ float red[8];
  const int r = blockIdx.x;
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) acc = fmaf(W[(size_t)r * K + k], vec[k], acc);
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) t_raw[r] = s;
}
// sum of squares of p[0..n) by one 256-thread block, the same order on every run; result broadcast to all threads
__device__ __forceinline__ float sn_norm2_256(const float* __restrict__ p, int n, float* red, float* bcast) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fmaf(p[i], p[i], acc);
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) *bcast = s;
  __syncthreads();
  const float r = *bcast;
  __syncthreads();
  return r;
}
__global__ void __launch_bounds__(256) sn_finalize_kernel(float* __restrict__ u, float* __restrict__ v, float* __restrict__ sigma,
                                                          const float* __restrict__ v_raw, const float* __restrict__ t_raw,
                                                          int Cout, int K, int update_uv, float eps) {
  pdl_enter();
  __shared__ float red[8];
  __shared__ float bcast;
  if (update_uv) {
    const float nv = fmaxf(sqrtf(sn_norm2_256(v_raw, K, red, &bcast)), eps);
    const float tn = sqrtf(sn_norm2_256(t_raw, Cout, red, &bcast)) / nv;  // || W v ||
    const float nu = fmaxf(tn, eps);
    for (int k = threadIdx.x; k < K; k += blockDim.x) v[k] = v_raw[k] / nv;
    for (int r = threadIdx.x; r < Cout; r += blockDim.x) u[r] = (t_raw[r] / nv) / nu;
    if (threadIdx.x == 0) sigma[0] = tn * tn / nu;  // u^T (W v)
  } else {
    float acc = 0.f;
    for (int r = threadIdx.x; r < Cout; r += blockDim.x) acc = fmaf(u[r], t_raw[r], acc);
    const float s = block_sum_256(acc, red);
    if (threadIdx.x == 0) sigma[0] = s;
  }
}
__global__ void sn_scale_kernel(const float* __restrict__ w, const float* __restrict__ sigma, float* __restrict__ w_sn, long long n) {
  pdl_enter();
  const float inv = 1.f / sigma[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) w_sn[i] = w[i] * inv;
}
// out[blockIdx.x] = this block's share of sum a*b (gridDim.x <= HPVG_SN_DOT_PARTS); the consumer adds the shares in order
__global__ void __launch_bounds__(256) dot_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  pdl_enter();
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc = fmaf(a[i], b[i], acc);
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}
__device__ __forceinline__ float sn_ordered_sum(const float* __restrict__ parts, int n) {
  float d = 0.f;
  for (int i = 0; i < n; ++i) d += parts[i];
  return d;
}
__global__ void sn_bwd_kernel(const float* __restrict__ gw_sn, const float* __restrict__ u, const float* __restrict__ v,
                              const float* __restrict__ sigma, const float* __restrict__ dot, int parts, float* __restrict__ gw, int Cout,
                              int K) {
  pdl_enter();
  const float inv = 1.f / sigma[0], d = sn_ordered_sum(dot, parts);
  const long long total = (long long)Cout * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K), r = (int)(i / K);
    gw[i] = (gw_sn[i] - d * u[r] * v[k]) * inv;
  }
}


// ---- the same four steps for up to HPVG_SN_MAX_LAYERS layers at once (blockIdx.y = layer): one network's spectral-norm
// weights are all computed before its first convolution, in 4 launches instead of 4 per layer ----------------------------
struct SnBatch {
  int n;
  const float* w[HPVG_SN_MAX_LAYERS];
  float* u[HPVG_SN_MAX_LAYERS];
  float* v[HPVG_SN_MAX_LAYERS];
  float* sigma[HPVG_SN_MAX_LAYERS];
  float* w_sn[HPVG_SN_MAX_LAYERS];
  float* scratch[HPVG_SN_MAX_LAYERS];     // K + Cout + 4 floats: v_raw, t_raw, norms
  const float* gw_sn[HPVG_SN_MAX_LAYERS];  // backward only
  float* gw[HPVG_SN_MAX_LAYERS];
  float* u_saved[HPVG_SN_MAX_LAYERS];      // forward only, may be null: copies of the updated u / v for the backward pass
  float* v_saved[HPVG_SN_MAX_LAYERS];
  int cout[HPVG_SN_MAX_LAYERS], k[HPVG_SN_MAX_LAYERS];
};

// v_raw = W^T u: a block owns 64 columns, its four 64-thread groups each walk a quarter of the rows; the quarters are combined in
// a fixed order (bit-reproducible).  (One thread per column over all rows left 49 blocks on 148 SMs: 15 us per call.)
__global__ void __launch_bounds__(256) snb_wtu_kernel(const SnBatch b) {
  pdl_enter();
  __shared__ float part[4][64];
  const int l = blockIdx.y, K = b.k[l], Cout = b.cout[l];
  const int kc = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int k = blockIdx.x * 64 + kc;
  const float* W = b.w[l];
  const float* u = b.u[l];
  float acc = 0.f;
  if (k < K) {
    const int per = (Cout + 3) / 4;
    const int r1 = min(Cout, (rg + 1) * per);
    for (int r = rg * per; r < r1; ++r) acc = fmaf(W[(size_t)r * K + k], u[r], acc);
  }
  part[rg][kc] = acc;
  __syncthreads();
  if (rg == 0 && k < K) b.scratch[l][k] = ((part[0][kc] + part[1][kc]) + part[2][kc]) + part[3][kc];
}
__global__ void __launch_bounds__(256) snb_wv_kernel(const SnBatch b, int update_uv) {
  pdl_enter();
  __shared__ float red[8];
  const int l = blockIdx.y, K = b.k[l], Cout = b.cout[l];
  const int r = blockIdx.x;
  if (r >= Cout) return;
  const float* W = b.w[l];
  const float* vec = update_uv ? b.scratch[l] : b.v[l];
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) acc = fmaf(W[(size_t)r * K + k], vec[k], acc);
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) b.scratch[l][K + r] = s;
}
__global__ void __launch_bounds__(256) snb_finalize_kernel(const SnBatch b, int update_uv, float eps) {
  pdl_enter();
  __shared__ float red[8];
  const int l = blockIdx.x, K = b.k[l], Cout = b.cout[l];
  const float* v_raw = b.scratch[l];
  const float* t_raw = b.scratch[l] + K;
  __shared__ float bcast;
  if (update_uv) {
    const float nv = fmaxf(sqrtf(sn_norm2_256(v_raw, K, red, &bcast)), eps);
    const float tn = sqrtf(sn_norm2_256(t_raw, Cout, red, &bcast)) / nv;  // || W v ||
    const float nu = fmaxf(tn, eps);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const float x = v_raw[k] / nv;
      b.v[l][k] = x;
      if (b.v_saved[l]) b.v_saved[l][k] = x;
    }
    for (int r = threadIdx.x; r < Cout; r += blockDim.x) {
      const float x = (t_raw[r] / nv) / nu;
      b.u[l][r] = x;
      if (b.u_saved[l]) b.u_saved[l][r] = x;
    }
    if (threadIdx.x == 0) b.sigma[l][0] = tn * tn / nu;  // u^T (W v)
  } else {
    float acc = 0.f;
    for (int r = threadIdx.x; r < Cout; r += blockDim.x) acc = fmaf(b.u[l][r], t_raw[r], acc);
    const float s = block_sum_256(acc, red);
    if (threadIdx.x == 0) b.sigma[l][0] = s;
  }
}
__global__ void snb_scale_kernel(const SnBatch b) {
  pdl_enter();
  const int l = blockIdx.y;
  const long long n = (long long)b.cout[l] * b.k[l];
  const float inv = 1.f / b.sigma[l][0];
  const float* w = b.w[l];
  float* o = b.w_sn[l];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = w[i] * inv;
}
// backward: scratch[l][0] = sum gw_sn * w_sn ; gw = (gw_sn - dot * u v^T) / sigma      (b.w holds w_sn here)
__global__ void __launch_bounds__(256) snb_dot_kernel(const SnBatch b) {
  pdl_enter();
  __shared__ float red[8];
  const int l = blockIdx.y;
  const long long n = (long long)b.cout[l] * b.k[l];
  const float* a = b.gw_sn[l];
  const float* w = b.w[l];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc = fmaf(a[i], w[i], acc);
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) b.scratch[l][blockIdx.x] = s;
}
__global__ void snb_bwd_kernel(const SnBatch b, int parts) {
  pdl_enter();
  const int l = blockIdx.y, K = b.k[l];
  const long long total = (long long)b.cout[l] * K;
  const float inv = 1.f / b.sigma[l][0], d = sn_ordered_sum(b.scratch[l], parts);
  const float* g = b.gw_sn[l];
  const float* u = b.u[l];
  const float* v = b.v[l];
  float* o = b.gw[l];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K), r = (int)(i / K);
    o[i] = (g[i] - d * u[r] * v[k]) * inv;
  }
}

}  // namespace hpvg

// =================================================================================================================
// C-ABI
// =================================================================================================================
using namespace hpvg;

#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define MEMSET0(ptr, bytes, st, name)                                                   \
  do {                                                                                  \
    cudaError_t e_ = cudaMemsetAsync(ptr, 0, bytes, st);                                \
    if (e_ != cudaSuccess) {                                                            \
      set_error("%s: memset failed: %s", name, cudaGetErrorString(e_));                 \
      return -2;                                                                        \
    }                                                                                   \
  } while (0)

extern "C" {

int hpvg_bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     long long* nbt, float momentum, float eps, long long count, float* scale_shift, float* mean_invstd, int C,
                     void* stream) {
  HPVG_CHECK_ARG(C > 0 && count > 0, "bn_finalize: bad C=%d count=%lld", C, count);
  launch_k(bn_finalize_kernel, (unsigned)cdiv(C, 128), 128, 0, ST(stream), stats, gamma, beta, running_mean, running_var, nbt, momentum, eps,
                                                                      count, scale_shift, mean_invstd, C);
  HPVG_CHECK_LAUNCH("bn_finalize");
  return 0;
}

int hpvg_bn_apply_lrelu(const void* y, const float* scale_shift, void* out, long long nvox, int C, float slope, void* stream) {
  HPVG_CHECK_ARG(C % 8 == 0 && C <= 1024, "bn_apply_lrelu: C=%d must be a multiple of 8 (<= 1024)", C);
  const long long nvec = nvox * (C / 8);
  launch_k(bn_apply_lrelu_kernel, ew_blocks(nvec, 256), 256, 2 * C * sizeof(float), ST(stream), 
      reinterpret_cast<const uint4*>(y), scale_shift, reinterpret_cast<uint4*>(out), nvec, C, slope);
  HPVG_CHECK_LAUNCH("bn_apply_lrelu");
  return 0;
}

int hpvg_bn_finalize_apply_lrelu(const void* y, const float* stats, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, long long* nbt, float momentum, float eps, float* scale_shift,
                                 float* mean_invstd, void* out, long long nvox, int C, float slope, void* stream) {
  HPVG_CHECK_ARG(C % 8 == 0 && C <= 256 && nvox > 0, "bn_finalize_apply_lrelu: C=%d must be a multiple of 8 (<= 256)", C);
  const long long nvec = nvox * (C / 8);
  launch_k(bn_finalize_apply_lrelu_kernel, ew_blocks(nvec, 256), 256, 2 * C * sizeof(float), ST(stream), 
      reinterpret_cast<const uint4*>(y), stats, gamma, beta, running_mean, running_var, nbt, momentum, eps, nvox, scale_shift,
      mean_invstd, reinterpret_cast<uint4*>(out), nvec, C, slope);
  HPVG_CHECK_LAUNCH("bn_finalize_apply_lrelu");
  return 0;
}

int hpvg_bn_apply_lrelu_per_sample(const void* y, const float* stats, const float* gamma, const float* beta, float eps, void* out,
                                   int N, long long nvox_per_sample, int C, float slope, void* stream) {
  HPVG_CHECK_ARG(y && stats && gamma && beta && out, "bn_apply_lrelu_per_sample: null tensor");
  HPVG_CHECK_ARG(C % 8 == 0 && C <= 256 && nvox_per_sample > 0 && N > 0 && N <= 65535,
                 "bn_apply_lrelu_per_sample: C=%d must be a multiple of 8 (<= 256), 1 <= N <= 65535", C);
  const long long nvec = nvox_per_sample * (C / 8);
  const int bx = (int)max(1LL, min(cdiv(nvec, 256), cdiv((long long)num_sms() * 16, N)));
  launch_k(bn_apply_lrelu_per_sample_kernel, dim3(bx, N), 256, 2 * C * sizeof(float), ST(stream), reinterpret_cast<const uint4*>(y), stats,
           gamma, beta, eps, nvox_per_sample, reinterpret_cast<uint4*>(out), nvec, C, slope);
  HPVG_CHECK_LAUNCH("bn_apply_lrelu_per_sample");
  return 0;
}

int hpvg_bn_lrelu_bwd_reduce(const void* y, const void* gout, const float* scale_shift, const float* mean_invstd, float* sums,
                             long long nvox, int C, float slope, const void* mask_bits, void* stream) {
  HPVG_CHECK_ARG(C % 8 == 0 && C <= 256, "bn_lrelu_bwd_reduce: C=%d must be a multiple of 8 (<= 256)", C);
  MEMSET0(sums, 3 * C * sizeof(float), ST(stream), "bn_lrelu_bwd_reduce");   // [2C] sums + [C] bias-gradient accumulator
  const int rows = 256 / (C / 8);
  const size_t smem = (size_t)(4 * C + rows * 2 * C) * sizeof(float);
  const int blocks = (int)max(1LL, min(cdiv(nvox, rows), (long long)num_sms() * 4));
  launch_k(bn_lrelu_bwd_reduce_kernel, blocks, 256, smem, ST(stream), reinterpret_cast<const uint4*>(y), reinterpret_cast<const uint4*>(gout),
                                                                scale_shift, mean_invstd, sums, nvox, C, slope,
                                                                reinterpret_cast<const uint8_t*>(mask_bits));
  HPVG_CHECK_LAUNCH("bn_lrelu_bwd_reduce");
  return 0;
}

int hpvg_bn_lrelu_bwd_apply(const void* y, const void* gout, const float* scale_shift, const float* mean_invstd, float* sums,
                            void* gy, float* dgamma, float* dbeta, long long nvox, int C, float slope, int want_chsum,
                            const void* mask_bits, void* stream) {
  HPVG_CHECK_ARG(C % 8 == 0 && C <= 256 && (256 % (C / 8)) == 0, "bn_lrelu_bwd_apply: C=%d must be a multiple of 8 dividing 2048", C);
  const long long nvec = nvox * (C / 8);
  // with the fused channel sum every block ends with C global atomics on the same C addresses: keep the grid at 2 CTAs per SM
  const int blocks = want_chsum ? min(ew_blocks(nvec, 256), 2 * num_sms()) : ew_blocks(nvec, 256);
  launch_k(bn_lrelu_bwd_apply_kernel, blocks, 256, 7 * C * sizeof(float), ST(stream), 
      reinterpret_cast<const uint4*>(y), reinterpret_cast<const uint4*>(gout), scale_shift, mean_invstd, sums,
      reinterpret_cast<uint4*>(gy), dgamma, dbeta, nvox, C, slope, want_chsum ? sums + 2 * C : nullptr,
      reinterpret_cast<const uint8_t*>(mask_bits));
  HPVG_CHECK_LAUNCH("bn_lrelu_bwd_apply");
  return 0;
}


int hpvg_bn_running_update_batched(int n, float* const* running_mean, float* const* running_var, long long* const* num_batches_tracked,
                                   const float* const* mean_invstd, const long long* count, const int* C, const float* momentum,
                                   const float* eps, void* stream) {
  HPVG_CHECK_ARG(n >= 1 && n <= HPVG_BN_LOG_MAX, "bn_running_update_batched: 1..%d entries per call, got %d", HPVG_BN_LOG_MAX, n);
  BnRunBatch b;
  b.n = n;
  for (int i = 0; i < n; ++i) {
    HPVG_CHECK_ARG(mean_invstd[i] && C[i] > 0 && C[i] <= 256, "bn_running_update_batched: entry %d: C=%d must be in 1..256", i, C[i]);
    b.rm[i] = running_mean[i]; b.rv[i] = running_var[i]; b.nbt[i] = num_batches_tracked[i];
    b.mean_invstd[i] = mean_invstd[i]; b.count[i] = count[i]; b.C[i] = C[i]; b.momentum[i] = momentum[i]; b.eps[i] = eps[i];
  }
  launch_k(bn_running_update_kernel, n, 256, 0, ST(stream), b);
  HPVG_CHECK_LAUNCH("bn_running_update");
  return 0;
}

// rows of y / gout (+ mask bytes) a CTA can keep in shared memory: 9C floats of parameters + rows * (2 * 2C + C/8) bytes
static int bnb_rows_per_cta(long long nvox, int C, int& grid) {
  const int sms = num_sms();
  grid = (int)max(1LL, min((long long)sms, cdiv(nvox, 64)));
  const long long rows = cdiv(nvox, grid);
  const size_t need = 16 + (size_t)9 * C * sizeof(float) + (size_t)rows * (4 * C + C / 8);
  return need <= (size_t)220 * 1024 ? (int)rows : 0;
}

int hpvg_bn_lrelu_bwd_fused_supported(long long nvox, int C) {
  int grid;
  const int cvec = C / 8;      // lanes that share a channel group combine by shuffles: C / 8 must be a power of two <= 32
  return (C % 8 == 0 && C >= 8 && C <= 256 && (cvec & (cvec - 1)) == 0 && nvox > 0 && bnb_rows_per_cta(nvox, C, grid) > 0) ? 1 : 0;
}

int hpvg_bn_lrelu_bwd_fused(const void* y, const void* gout, const float* scale_shift, const float* mean_invstd, float* sums, void* gy,
                            float* dgamma, float* dbeta, long long nvox, int C, float slope, int want_chsum, const void* mask_bits,
                            void* stream) {
  HPVG_CHECK_ARG(y && gout && scale_shift && mean_invstd && sums && gy, "bn_lrelu_bwd_fused: null tensor");
  HPVG_CHECK_ARG(hpvg_bn_lrelu_bwd_fused_supported(nvox, C), "bn_lrelu_bwd_fused: %lld voxels x %d channels do not fit the SMs' shared memory", nvox, C);
  int grid;
  const int rows = bnb_rows_per_cta(nvox, C, grid);
  grid = (int)cdiv(nvox, rows);
  const size_t smem = 16 + (size_t)9 * C * sizeof(float) + (size_t)rows * (4 * C + C / 8) + 16;
  static std::atomic<unsigned long long> attr_mask{0};
  if (attr_pending(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(bn_lrelu_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) {
      set_error("bn_lrelu_bwd_fused: cannot opt in to shared memory: %s", cudaGetErrorString(e));
      return -2;
    }
    attr_set(attr_mask);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(BNB_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ST(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  static const bool coop = !(getenv("HPVG_FUSED_COOP") && atoi(getenv("HPVG_FUSED_COOP")) == 0);
  cfg.numAttrs = coop ? 1 : 0;
  cudaLaunchKernelEx(&cfg, bn_lrelu_bwd_fused_kernel, reinterpret_cast<const uint4*>(y), reinterpret_cast<const uint4*>(gout), scale_shift,
                     mean_invstd, sums, reinterpret_cast<uint4*>(gy), dgamma, dbeta, nvox, C, slope, want_chsum,
                     reinterpret_cast<const uint8_t*>(mask_bits), rows);
  HPVG_CHECK_LAUNCH("bn_lrelu_bwd_fused");
  return 0;
}

int hpvg_lrelu_bwd(const void* gout, const void* out_saved, void* gz, long long numel, float slope, int C, float* chsum,
                   void* stream) {
  HPVG_CHECK_ARG(numel % 8 == 0, "lrelu_bwd: numel=%lld must be a multiple of 8", numel);
  HPVG_CHECK_ARG(chsum == nullptr || (C % 8 == 0 && C <= 256 && (256 % (C / 8)) == 0 && numel % C == 0),
                 "lrelu_bwd: fused channel sum needs C (=%d) a multiple of 8 dividing 2048", C);
  if (chsum) MEMSET0(chsum, C * sizeof(float), ST(stream), "lrelu_bwd");
  const long long nvec = numel / 8;
  const int blocks = chsum ? min(ew_blocks(nvec, 256), 2 * num_sms()) : ew_blocks(nvec, 256);
  launch_k(lrelu_bwd_kernel, blocks, 256, chsum ? C * sizeof(float) : 0, ST(stream), 
      reinterpret_cast<const uint4*>(gout), reinterpret_cast<const uint4*>(out_saved), reinterpret_cast<uint4*>(gz), nvec, slope, C,
      chsum);
  HPVG_CHECK_LAUNCH("lrelu_bwd");
  return 0;
}

static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

int hpvg_upsample_linear_fwd(const float* x, float* out, const float* noise, float noise_amp, int NC, int Di, int Hi, int Wi, int Do,
                             int Ho, int Wo, void* stream) {
  HPVG_CHECK_ARG(NC > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "upsample_linear_fwd: bad extents");
  const long long total = (long long)NC * Do * Ho * Wo;
  launch_k(upsample_fwd_kernel, ew_blocks(total, 256), 256, 0, ST(stream), x, out, noise, noise_amp, NC, Di, Hi, Wi, Do, Ho, Wo,
                                                                     ac_scale(Di, Do), ac_scale(Hi, Ho), ac_scale(Wi, Wo));
  HPVG_CHECK_LAUNCH("upsample_linear_fwd");
  return 0;
}

int hpvg_upsample_linear_bwd(const float* gout, float* gx, int NC, int Di, int Hi, int Wi, int Do, int Ho, int Wo, void* stream) {
  HPVG_CHECK_ARG(NC > 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0, "upsample_linear_bwd: bad extents");
  const long long total = (long long)NC * Di * Hi * Wi;
  launch_k(upsample_bwd_kernel, ew_blocks(total, 256), 256, 0, ST(stream), gout, gx, NC, Di, Hi, Wi, Do, Ho, Wo, ac_scale(Di, Do),
                                                                     ac_scale(Hi, Ho), ac_scale(Wi, Wo));
  HPVG_CHECK_LAUNCH("upsample_linear_bwd");
  return 0;
}

int hpvg_tanh_add_fwd(const float* a, const float* b, float* out, long long numel, void* stream) {
  launch_k(tanh_add_fwd_kernel, ew_blocks(numel, 256), 256, 0, ST(stream), a, b, out, numel);
  HPVG_CHECK_LAUNCH("tanh_add_fwd");
  return 0;
}
int hpvg_tanh_bwd(const float* gout, const float* out, float* g, long long numel, void* stream) {
  launch_k(tanh_bwd_kernel, ew_blocks(numel, 256), 256, 0, ST(stream), gout, out, g, numel);
  HPVG_CHECK_LAUNCH("tanh_bwd");
  return 0;
}

int hpvg_reparam_fwd(const void* mu, const void* logvar, const float* eps, void* z, int N, int C, long long spatial, void* stream) {
  HPVG_CHECK_ARG(C % 8 == 0, "reparam_fwd: C=%d must be a multiple of 8", C);
  const long long total = (long long)N * spatial * (C / 8);
  launch_k(reparam_fwd_kernel, ew_blocks(total, 256), 256, 0, ST(stream), reinterpret_cast<const uint4*>(mu), reinterpret_cast<const uint4*>(logvar),
                                                                    eps, reinterpret_cast<uint4*>(z), N, C, spatial);
  HPVG_CHECK_LAUNCH("reparam_fwd");
  return 0;
}
int hpvg_reparam_bwd(const void* gz, const void* logvar, const float* eps, void* gmu, void* glogvar, int N, int C, long long spatial,
                     void* stream) {
  HPVG_CHECK_ARG(C % 8 == 0, "reparam_bwd: C=%d must be a multiple of 8", C);
  const long long total = (long long)N * spatial * (C / 8);
  launch_k(reparam_bwd_kernel, ew_blocks(total, 256), 256, 0, ST(stream), reinterpret_cast<const uint4*>(gz), reinterpret_cast<const uint4*>(logvar),
                                                                    eps, reinterpret_cast<uint4*>(gmu), reinterpret_cast<uint4*>(glogvar), N, C,
                                                                    spatial);
  HPVG_CHECK_LAUNCH("reparam_bwd");
  return 0;
}

int hpvg_kl_fwd(const float* mu, const float* logvar, float* out, long long numel, void* stream) {
  HPVG_CHECK_ARG(numel > 0, "kl_fwd: empty input");
  MEMSET0(out, sizeof(float), ST(stream), "kl_fwd");
  launch_k(kl_fwd_kernel, (int)min((long long)num_sms() * 2, cdiv(numel, 256)), 256, 0, ST(stream), mu, logvar, out, numel, 1.f / (float)numel);
  HPVG_CHECK_LAUNCH("kl_fwd");
  return 0;
}
int hpvg_kl_bwd(const float* gout, const float* mu, const float* logvar, float* gmu, float* glogvar, long long numel, void* stream) {
  launch_k(kl_bwd_kernel, ew_blocks(numel, 256), 256, 0, ST(stream), gout, mu, logvar, gmu, glogvar, numel, 1.f / (float)numel);
  HPVG_CHECK_LAUNCH("kl_bwd");
  return 0;
}

int hpvg_gp_penalty_fwd(const float* g, float* out, int N, int C, long long spatial, float lambda, void* stream) {
  HPVG_CHECK_ARG(N > 0 && C > 0 && spatial > 0, "gp_penalty_fwd: bad extents");
  MEMSET0(out, sizeof(float), ST(stream), "gp_penalty_fwd");
  const long long total = (long long)N * spatial;
  launch_k(gp_fwd_kernel, (int)min((long long)num_sms() * 2, cdiv(total, 256)), 256, 0, ST(stream), g, out, N, C, spatial, lambda / (float)total);
  HPVG_CHECK_LAUNCH("gp_penalty_fwd");
  return 0;
}
int hpvg_gp_penalty_bwd(const float* gout, const float* g, float* gg, int N, int C, long long spatial, float lambda, void* stream) {
  const long long total = (long long)N * spatial;
  launch_k(gp_bwd_kernel, ew_blocks(total, 256), 256, 0, ST(stream), gout, g, gg, N, C, spatial, lambda / (float)total);
  HPVG_CHECK_LAUNCH("gp_penalty_bwd");
  return 0;
}

int hpvg_convert_format(const void* src, int src_fmt, void* dst, int dst_fmt, int N, int C, long long spatial, void* stream) {
  const long long total = (long long)N * C * spatial;
  if (src_fmt == HPVG_FMT_NCDHW_F32 && dst_fmt == HPVG_FMT_NDHWC_BF16) {
    launch_k(ncdhw_to_ndhwc_kernel, ew_blocks(total, 256), 256, 0, ST(stream), reinterpret_cast<const float*>(src),
                                                                         reinterpret_cast<__nv_bfloat16*>(dst), N, C, spatial);
  } else if (src_fmt == HPVG_FMT_NDHWC_BF16 && dst_fmt == HPVG_FMT_NCDHW_F32) {
    launch_k(ndhwc_to_ncdhw_kernel, ew_blocks(total, 256), 256, 0, ST(stream), reinterpret_cast<const __nv_bfloat16*>(src),
                                                                         reinterpret_cast<float*>(dst), N, C, spatial);
  } else {
    set_error("convert_format: unsupported conversion %d -> %d", src_fmt, dst_fmt);
    return -1;
  }
  HPVG_CHECK_LAUNCH("convert_format");
  return 0;
}

int hpvg_lerp(const float* a, const float* b, float* out, const float* alpha, long long numel, void* stream) {
  HPVG_CHECK_ARG(alpha != nullptr, "lerp: alpha must point to a device float");
  launch_k(lerp_kernel, ew_blocks(numel, 256), 256, 0, ST(stream), a, b, out, alpha, numel);
  HPVG_CHECK_LAUNCH("lerp");
  return 0;
}

int hpvg_channel_sum(const void* t, int fmt, float* out, int N, int C, long long spatial, void* stream) {
  HPVG_CHECK_ARG(N > 0 && C > 0 && spatial > 0, "channel_sum: bad extents");
  MEMSET0(out, C * sizeof(float), ST(stream), "channel_sum");
  if (fmt == HPVG_FMT_NDHWC_BF16) {
    HPVG_CHECK_ARG(C <= 256, "channel_sum: C=%d too large for the NDHWC kernel", C);
    const long long rows = (long long)N * spatial;
    const int rpb = 256 / C;
    const int blocks = (int)max(1LL, min(cdiv(rows, (long long)rpb * 8), (long long)num_sms() * 4));
    launch_k(channel_sum_ndhwc_kernel, blocks, 256, 0, ST(stream), reinterpret_cast<const __nv_bfloat16*>(t), out, rows, C);
  } else {
    dim3 grid((unsigned)max(1LL, min(cdiv(spatial, 2048), 64LL)), (unsigned)(N * C));
    launch_k(channel_sum_ncdhw_kernel, grid, 256, 0, ST(stream), reinterpret_cast<const float*>(t), out, C, spatial);
  }
  HPVG_CHECK_LAUNCH("channel_sum");
  return 0;
}

int hpvg_pack_weights(const float* w_f32, void* w_packed, int Cout, int Cin, int taps, int transposed, const float* inv_scale_of,
                      int rows_per_tap, void* stream) {
  HPVG_CHECK_ARG(rows_per_tap >= Cout, "pack_weights: rows_per_tap (%d) < Cout (%d)", rows_per_tap, Cout);
  const long long total = (long long)taps * rows_per_tap * Cin;
  launch_k(pack_weights_kernel, ew_blocks(total, 256), 256, 0, ST(stream), w_f32, reinterpret_cast<__nv_bfloat16*>(w_packed), Cout, Cin, taps,
                                                                     transposed, inv_scale_of, rows_per_tap);
  HPVG_CHECK_LAUNCH("pack_weights");
  return 0;
}

int hpvg_pack_weights_pair_batched(int n, const float* const* w_f32, void* const* packed_fwd, void* const* packed_tr, const int* Cout,
                                   const int* Cin, const int* taps, void* stream) {
  HPVG_CHECK_ARG(n > 0 && n <= HPVG_SN_MAX_LAYERS, "pack_weights_pair_batched: %d layers (max %d)", n, HPVG_SN_MAX_LAYERS);
  PackBatch b;
  long long maxn = 0;
  for (int l = 0; l < n; ++l) {
    HPVG_CHECK_ARG(w_f32[l] && Cout[l] > 0 && Cin[l] > 0 && (taps[l] == 9 || taps[l] == 27), "pack_weights_pair_batched: bad layer %d", l);
    b.w[l] = w_f32[l];
    b.fwd[l] = reinterpret_cast<__nv_bfloat16*>(packed_fwd[l]);
    b.tr[l] = reinterpret_cast<__nv_bfloat16*>(packed_tr[l]);
    b.cout[l] = Cout[l]; b.cin[l] = Cin[l]; b.taps[l] = taps[l];
    maxn = max(maxn, (long long)Cout[l] * Cin[l] * taps[l]);
  }
  launch_k(pack_pair_batched_kernel, dim3((unsigned)min(cdiv(maxn, 256), 128LL), n), 256, 0, ST(stream), b);
  HPVG_CHECK_LAUNCH("pack_weights_pair_batched");
  return 0;
}

int hpvg_sn_power_iter(const float* w_orig, float* u, float* v, float* sigma, float* w_sn, float* scratch, int Cout, int K,
                       int update_uv, float eps, void* stream) {
  HPVG_CHECK_ARG(Cout > 0 && K > 0, "sn_power_iter: bad shape");
  float* v_raw = scratch;
  float* t_raw = scratch + K;
  if (update_uv) {
    launch_k(sn_wtu_kernel, (unsigned)cdiv(K, 256), 256, 0, ST(stream), w_orig, u, v_raw, Cout, K);
    HPVG_CHECK_LAUNCH("sn_wtu");
    launch_k(sn_wv_kernel, Cout, 256, 0, ST(stream), w_orig, v_raw, t_raw, K);
  } else {
    launch_k(sn_wv_kernel, Cout, 256, 0, ST(stream), w_orig, v, t_raw, K);
  }
  HPVG_CHECK_LAUNCH("sn_wv");
  launch_k(sn_finalize_kernel, 1, 256, 0, ST(stream), u, v, sigma, v_raw, t_raw, Cout, K, update_uv, eps);
  HPVG_CHECK_LAUNCH("sn_finalize");
  if (w_sn) {
    const long long n = (long long)Cout * K;
    launch_k(sn_scale_kernel, ew_blocks(n, 256), 256, 0, ST(stream), w_orig, sigma, w_sn, n);
    HPVG_CHECK_LAUNCH("sn_scale");
  }
  return 0;
}

int hpvg_sn_backward(const float* gw_sn, const float* w_sn, const float* u, const float* v, const float* sigma, float* gw_orig,
                     float* scratch, int Cout, int K, void* stream) {
  const long long n = (long long)Cout * K;
  const int parts = (int)min((long long)HPVG_SN_DOT_PARTS, cdiv(n, 256));
  launch_k(dot_kernel, parts, 256, 0, ST(stream), gw_sn, w_sn, scratch, n);
  HPVG_CHECK_LAUNCH("sn_dot");
  launch_k(sn_bwd_kernel, ew_blocks(n, 256), 256, 0, ST(stream), gw_sn, u, v, sigma, scratch, parts, gw_orig, Cout, K);
  HPVG_CHECK_LAUNCH("sn_bwd");
  return 0;
}

// Batched forms: `n` layers (<= HPVG_SN_MAX_LAYERS) in 4 (forward) / 2 (backward) launches.  Pointer arrays live in host
// memory and are copied into the kernel parameters.
static int sn_fill_batch(SnBatch& b, int n, const int* cout, const int* k, const char* who) {
  HPVG_CHECK_ARG(n > 0 && n <= HPVG_SN_MAX_LAYERS, "%s: %d layers (max %d)", who, n, HPVG_SN_MAX_LAYERS);
  b.n = n;
  for (int l = 0; l < n; ++l) {
    HPVG_CHECK_ARG(cout[l] > 0 && k[l] > 0, "%s: bad shape of layer %d", who, l);
    b.cout[l] = cout[l];
    b.k[l] = k[l];
    b.u_saved[l] = b.v_saved[l] = nullptr;
  }
  return 0;
}

int hpvg_sn_power_iter_batched(int n, const float* const* w_orig, float* const* u, float* const* v, float* const* sigma,
                               float* const* w_sn, float* const* scratch, const int* cout, const int* k, int update_uv, float eps,
                               void* stream) {
  return hpvg_sn_power_iter_batched_ex(n, w_orig, u, v, sigma, w_sn, scratch, cout, k, update_uv, eps, nullptr, nullptr, stream);
}

int hpvg_sn_power_iter_batched_ex(int n, const float* const* w_orig, float* const* u, float* const* v, float* const* sigma,
                                  float* const* w_sn, float* const* scratch, const int* cout, const int* k, int update_uv, float eps,
                                  float* const* u_saved, float* const* v_saved, void* stream) {
  SnBatch b;
  if (int rc = sn_fill_batch(b, n, cout, k, "sn_power_iter_batched")) return rc;
  for (int l = 0; l < n; ++l) {
    b.u_saved[l] = u_saved ? u_saved[l] : nullptr;
    b.v_saved[l] = v_saved ? v_saved[l] : nullptr;
  }
  int maxk = 0, maxc = 0;
  long long maxn = 0;
  for (int l = 0; l < n; ++l) {
    b.w[l] = w_orig[l]; b.u[l] = u[l]; b.v[l] = v[l]; b.sigma[l] = sigma[l]; b.w_sn[l] = w_sn[l]; b.scratch[l] = scratch[l];
    maxk = max(maxk, k[l]); maxc = max(maxc, cout[l]);
    maxn = max(maxn, (long long)cout[l] * k[l]);
  }
  if (update_uv) {
    launch_k(snb_wtu_kernel, dim3((unsigned)cdiv(maxk, 64), n), 256, 0, ST(stream), b);
    HPVG_CHECK_LAUNCH("snb_wtu");
  }
  launch_k(snb_wv_kernel, dim3(maxc, n), 256, 0, ST(stream), b, update_uv);
  HPVG_CHECK_LAUNCH("snb_wv");
  launch_k(snb_finalize_kernel, n, 256, 0, ST(stream), b, update_uv, eps);
  HPVG_CHECK_LAUNCH("snb_finalize");
  launch_k(snb_scale_kernel, dim3((unsigned)min(cdiv(maxn, 256), 64LL), n), 256, 0, ST(stream), b);
  HPVG_CHECK_LAUNCH("snb_scale");
  return 0;
}

int hpvg_sn_backward_batched(int n, const float* const* gw_sn, const float* const* w_sn, const float* const* u, const float* const* v,
                             const float* const* sigma, float* const* gw_orig, float* const* scratch, const int* cout, const int* k,
                             void* stream) {
  SnBatch b;
  if (int rc = sn_fill_batch(b, n, cout, k, "sn_backward_batched")) return rc;
  long long maxn = 0;
  for (int l = 0; l < n; ++l) {
    b.gw_sn[l] = gw_sn[l]; b.w[l] = w_sn[l]; b.u[l] = const_cast<float*>(u[l]); b.v[l] = const_cast<float*>(v[l]);
    b.sigma[l] = const_cast<float*>(sigma[l]); b.gw[l] = gw_orig[l]; b.scratch[l] = scratch[l];
    maxn = max(maxn, (long long)cout[l] * k[l]);
  }
  const unsigned bx = (unsigned)min(cdiv(maxn, 256), (long long)HPVG_SN_DOT_PARTS);
  launch_k(snb_dot_kernel, dim3(bx, n), 256, 0, ST(stream), b);
  HPVG_CHECK_LAUNCH("snb_dot");
  launch_k(snb_bwd_kernel, dim3((unsigned)min(cdiv(maxn, 256), 64LL), n), 256, 0, ST(stream), b, (int)bx);
  HPVG_CHECK_LAUNCH("snb_bwd");
  return 0;
}

}  // extern "C"

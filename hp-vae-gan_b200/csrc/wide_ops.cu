// Resize, zero-pad / crop and residual add on WIDE tensors (bf16 [N][D][H][W][C], C a multiple of 8).
//
// GeneratorCSG (reference modules/networks_3d.py:213-269) keeps nfc-channel feature maps between its pyramid stages: it
// zero-pads them (F.pad, :248,264), resizes them trilinearly with align_corners=True (utils.upscale / interpolate_3D,
// :252-259, utils/images.py:22-26,83-93), adds noise_amp * N(0,1) noise of the same shape (:260-261) and adds the stage
// output to the upscaled input (:265).  The hot path's other generators do all of that on 3-channel float32 tensors
// (elementwise.cu); these are the same operations for the bf16 channels-last storage the convolution kernels consume.
// All four are HBM-bound: one 16-byte vector (8 channels) per thread, every output element written exactly once, the
// adjoint of the resize in gather form (no atomics, deterministic).
#include "common.cuh"

namespace hpvg {

static inline int wide_blocks(long long work_items, int threads) {
  long long b = cdiv(work_items, threads);
  long long cap = (long long)num_sms() * 16;
  return (int)max(1LL, min(b, cap));
}
static inline float wide_ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

__device__ __forceinline__ void wlin_src(int o, float scale, int in, int& i0, int& i1, float& l0, float& l1) {
  const float s = scale * (float)o;   // align_corners=True: src = o * (in-1)/(out-1)
  i0 = (int)s;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}
__device__ __forceinline__ void wadj_range(int i, float scale, int out, int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out - 1; return; }
  lo = (int)floorf(((float)i - 1.f) / scale) - 1;
  hi = (int)ceilf(((float)i + 1.f) / scale) + 1;
  if (lo < 0) lo = 0;
  if (hi > out - 1) hi = out - 1;
}
__device__ __forceinline__ float wadj_weight(int o, int i, float scale, int in) {
  int i0, i1;
  float l0, l1;
  wlin_src(o, scale, in, i0, i1, l0, l1);
  float w = 0.f;
  if (i0 == i) w += l0;
  if (i1 == i) w += l1;
  return w;
}
__device__ __forceinline__ void fma8(float* acc, float w, const uint4 v) {
  float2 f;
  f = unpack_bf16x2(v.x); acc[0] = fmaf(w, f.x, acc[0]); acc[1] = fmaf(w, f.y, acc[1]);
  f = unpack_bf16x2(v.y); acc[2] = fmaf(w, f.x, acc[2]); acc[3] = fmaf(w, f.y, acc[3]);
  f = unpack_bf16x2(v.z); acc[4] = fmaf(w, f.x, acc[4]); acc[5] = fmaf(w, f.y, acc[5]);
  f = unpack_bf16x2(v.w); acc[6] = fmaf(w, f.x, acc[6]); acc[7] = fmaf(w, f.y, acc[7]);
}
__device__ __forceinline__ uint4 pack8(const float* a) {
  return make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
}

// out[n][od][oh][ow][c] = trilinear(x)[...] (+ amp * noise[n][c][od][oh][ow], noise float32 in the reference's NCDHW order).
// Thread = (output voxel, 8-channel chunk), voxel index fastest, so the eight noise reads of a warp are coalesced rows.
__global__ void __launch_bounds__(256) upsample_wide_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ out,
                                                                const float* __restrict__ noise, float amp, int N, int C8, int Di, int Hi,
                                                                int Wi, int Do, int Ho, int Wo, float sd, float sh, float sw) {
  pdl_enter();
  const long long Vo = (long long)Do * Ho * Wo;
  const long long total = (long long)N * C8 * Vo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pv = i % Vo;
    long long t = i / Vo;
    const int ch = (int)(t % C8);
    const int n = (int)(t / C8);
    const int ow = (int)(pv % Wo);
    const int oh = (int)((pv / Wo) % Ho);
    const int od = (int)(pv / ((long long)Wo * Ho));
    int d0, d1, h0, h1, w0, w1;
    float ld0, ld1, lh0, lh1, lw0, lw1;
    wlin_src(od, sd, Di, d0, d1, ld0, ld1);
    wlin_src(oh, sh, Hi, h0, h1, lh0, lh1);
    wlin_src(ow, sw, Wi, w0, w1, lw0, lw1);
    const uint4* p = x + (long long)n * Di * Hi * Wi * C8 + ch;
#define XW_(d, h, w) __ldg(p + (((long long)(d) * Hi + (h)) * Wi + (w)) * C8)
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    fma8(acc, ld0 * lh0 * lw0, XW_(d0, h0, w0));
    fma8(acc, ld0 * lh0 * lw1, XW_(d0, h0, w1));
    fma8(acc, ld0 * lh1 * lw0, XW_(d0, h1, w0));
    fma8(acc, ld0 * lh1 * lw1, XW_(d0, h1, w1));
    fma8(acc, ld1 * lh0 * lw0, XW_(d1, h0, w0));
    fma8(acc, ld1 * lh0 * lw1, XW_(d1, h0, w1));
    fma8(acc, ld1 * lh1 * lw0, XW_(d1, h1, w0));
    fma8(acc, ld1 * lh1 * lw1, XW_(d1, h1, w1));
#undef XW_
    if (noise) {
      const float* np = noise + ((long long)n * C8 * 8 + ch * 8) * Vo + pv;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(amp, __ldg(np + (long long)k * Vo), acc[k]);
    }
    out[((long long)n * Vo + pv) * C8 + ch] = pack8(acc);
  }
}

// gx[n][id][ih][iw][c] = sum over the output voxels that interpolate from this input voxel
__global__ void __launch_bounds__(256) upsample_wide_bwd_kernel(const uint4* __restrict__ gout, uint4* __restrict__ gx, int N, int C8, int Di,
                                                                int Hi, int Wi, int Do, int Ho, int Wo, float sd, float sh, float sw) {
  pdl_enter();
  const long long total = (long long)N * Di * Hi * Wi * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int ch = (int)(t % C8); t /= C8;
    const int iw = (int)(t % Wi); t /= Wi;
    const int ih = (int)(t % Hi); t /= Hi;
    const int id = (int)(t % Di);
    const int n = (int)(t / Di);
    int dlo, dhi, hlo, hhi, wlo, whi;
    wadj_range(id, sd, Do, dlo, dhi);
    wadj_range(ih, sh, Ho, hlo, hhi);
    wadj_range(iw, sw, Wo, wlo, whi);
    const uint4* g = gout + (long long)n * Do * Ho * Wo * C8 + ch;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int od = dlo; od <= dhi; ++od) {
      const float wd = wadj_weight(od, id, sd, Di);
      if (wd == 0.f) continue;
      for (int oh = hlo; oh <= hhi; ++oh) {
        const float wh = wadj_weight(oh, ih, sh, Hi);
        if (wh == 0.f) continue;
        for (int ow = wlo; ow <= whi; ++ow) {
          const float ww = wadj_weight(ow, iw, sw, Wi);
          if (ww != 0.f) fma8(acc, wd * wh * ww, __ldg(g + (((long long)od * Ho + oh) * Wo + ow) * C8));
        }
      }
    }
    gx[i] = pack8(acc);
  }
}

// out = x shifted by `pad` voxels on D, H and W into a volume of extents + 2 * pad, zeros outside (pad < 0: crop)
__global__ void __launch_bounds__(256) pad_wide_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int N, int C8, int Di, int Hi, int Wi,
                                                       int pad) {
  pdl_enter();
  const int Do = Di + 2 * pad, Ho = Hi + 2 * pad, Wo = Wi + 2 * pad;
  const long long total = (long long)N * Do * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int ch = (int)(t % C8); t /= C8;
    const int ow = (int)(t % Wo); t /= Wo;
    const int oh = (int)(t % Ho); t /= Ho;
    const int od = (int)(t % Do);
    const int n = (int)(t / Do);
    const int id = od - pad, ih = oh - pad, iw = ow - pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (id >= 0 && id < Di && ih >= 0 && ih < Hi && iw >= 0 && iw < Wi)
      v = __ldg(x + ((((long long)n * Di + id) * Hi + ih) * Wi + iw) * C8 + ch);
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256) add_wide_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                                                       long long nvec) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 va = __ldg(a + i), vb = __ldg(b + i);
    const uint32_t aw[4] = {va.x, va.y, va.z, va.w}, bw[4] = {vb.x, vb.y, vb.z, vb.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack_bf16x2(aw[k]), fb = unpack_bf16x2(bw[k]);
      r[k] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
    }
    out[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

}  // namespace hpvg

using namespace hpvg;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int hpvg_upsample_linear_wide_fwd(const void* x, void* out, const float* noise, float noise_amp, int N, int C, int Di, int Hi, int Wi, int Do,
                                  int Ho, int Wo, void* stream) {
  HPVG_CHECK_ARG(x && out, "upsample_linear_wide_fwd: null tensor");
  HPVG_CHECK_ARG(N > 0 && C > 0 && C % 8 == 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0,
                 "upsample_linear_wide_fwd: bad extents (C = %d must be a multiple of 8)", C);
  const long long total = (long long)N * (C / 8) * Do * Ho * Wo;
  launch_k(upsample_wide_fwd_kernel, wide_blocks(total, 256), 256, 0, ST(stream), reinterpret_cast<const uint4*>(x),
           reinterpret_cast<uint4*>(out), noise, noise_amp, N, C / 8, Di, Hi, Wi, Do, Ho, Wo, wide_ac_scale(Di, Do), wide_ac_scale(Hi, Ho),
           wide_ac_scale(Wi, Wo));
  HPVG_CHECK_LAUNCH("upsample_linear_wide_fwd");
  return 0;
}

int hpvg_upsample_linear_wide_bwd(const void* gout, void* gx, int N, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo, void* stream) {
  HPVG_CHECK_ARG(gout && gx, "upsample_linear_wide_bwd: null tensor");
  HPVG_CHECK_ARG(N > 0 && C > 0 && C % 8 == 0 && Di > 0 && Hi > 0 && Wi > 0 && Do > 0 && Ho > 0 && Wo > 0,
                 "upsample_linear_wide_bwd: bad extents (C = %d must be a multiple of 8)", C);
  const long long total = (long long)N * (C / 8) * Di * Hi * Wi;
  launch_k(upsample_wide_bwd_kernel, wide_blocks(total, 256), 256, 0, ST(stream), reinterpret_cast<const uint4*>(gout),
           reinterpret_cast<uint4*>(gx), N, C / 8, Di, Hi, Wi, Do, Ho, Wo, wide_ac_scale(Di, Do), wide_ac_scale(Hi, Ho), wide_ac_scale(Wi, Wo));
  HPVG_CHECK_LAUNCH("upsample_linear_wide_bwd");
  return 0;
}

int hpvg_pad_wide(const void* x, void* out, int N, int C, int D, int H, int W, int pad, void* stream) {
  HPVG_CHECK_ARG(x && out, "pad_wide: null tensor");
  HPVG_CHECK_ARG(N > 0 && C > 0 && C % 8 == 0 && D > 0 && H > 0 && W > 0, "pad_wide: bad extents (C = %d must be a multiple of 8)", C);
  HPVG_CHECK_ARG(D + 2 * pad > 0 && H + 2 * pad > 0 && W + 2 * pad > 0, "pad_wide: pad %d leaves nothing of %d x %d x %d", pad, D, H, W);
  const long long total = (long long)N * (C / 8) * (D + 2 * pad) * (H + 2 * pad) * (W + 2 * pad);
  launch_k(pad_wide_kernel, wide_blocks(total, 256), 256, 0, ST(stream), reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), N,
           C / 8, D, H, W, pad);
  HPVG_CHECK_LAUNCH("pad_wide");
  return 0;
}

int hpvg_add_wide(const void* a, const void* b, void* out, long long numel, void* stream) {
  HPVG_CHECK_ARG(a && b && out, "add_wide: null tensor");
  HPVG_CHECK_ARG(numel > 0 && numel % 8 == 0, "add_wide: numel = %lld must be a positive multiple of 8", numel);
  launch_k(add_wide_kernel, wide_blocks(numel / 8, 256), 256, 0, ST(stream), reinterpret_cast<const uint4*>(a),
           reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(out), numel / 8);
  HPVG_CHECK_LAUNCH("add_wide");
  return 0;
}

}  // extern "C"

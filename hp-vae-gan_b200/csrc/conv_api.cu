// C-ABI entry points of the convolution family: argument validation and dispatch between the tcgen05 kernels
// (conv_tc.cu, wgrad_tc.cu) and the CUDA-core kernels (conv_direct.cu).
#include "common.cuh"

namespace hpvg {
int conv_direct(const void* x, int x_fmt, const float* w, const float* bias, void* y, int y_fmt, const ConvGeom& g, int transposed,
                int act, float slope, float* stats, const void* mask_src, cudaStream_t st);
int wgrad_direct(const void* x, int x_fmt, const void* gy, int gy_fmt, float* dw, const ConvGeom& g, cudaStream_t st);
bool conv_tc_supported(int x_fmt, int y_fmt, const ConvGeom& g, const void* w_packed);
int conv_tc(const void* x, const void* w_packed, const float* bias, void* y, int y_fmt, const ConvGeom& g, int act, float slope,
            float* stats, const void* mask_src, cudaStream_t st);
bool expand_conv_supported(int x_fmt, int y_fmt, const ConvGeom& g, const void* mask_src);
int expand_conv(const void* x, const float* w, const float* w_tco, const float* bias, void* y, const ConvGeom& g, int transposed,
                int act, float slope, float* stats, cudaStream_t st);
int pack_expand(const float* w, float* out, int Cin, int taps, int transposed, cudaStream_t st);
bool narrow_wgrad_supported(int x_fmt, int gy_fmt, const ConvGeom& g);
size_t narrow_wgrad_workspace(int x_fmt, const ConvGeom& g);
int narrow_wgrad(const void* x, int x_fmt, const void* gy, float* dw, float* dbias_wide, const ConvGeom& g, void* workspace,
                 size_t ws_bytes, cudaStream_t st);
bool conv_tc_picks_column(int y_fmt, const ConvGeom& g, const void* w_packed);
bool wgrad_tc_supported(int x_fmt, int gy_fmt, const ConvGeom& g);
size_t wgrad_tc_workspace(const ConvGeom& g);
int wgrad_tc(const void* x, const void* gy, float* dw, const ConvGeom& g, void* workspace, size_t ws_bytes, cudaStream_t st);

bool thin_gs_supported(int x_fmt, int y_fmt, const ConvGeom& g);
bool expand_tc_supported(const ConvGeom& g);
bool narrow_wgrad_tc_supported(const ConvGeom& g, bool head);
int thin_conv_gs(const void* x, const void* w_packed, const float* bias, float* y, const ConvGeom& g, cudaStream_t st);
bool conv_bn_fused_supported(const ConvGeom& g);
int conv_bn_fused(const void* x, const void* w_packed, const float* bias, void* y, void* out, const ConvGeom& g, float slope,
                  const float* gamma, const float* beta, float* running_mean, float* running_var, long long* nbt, float momentum,
                  float eps, float* stats, unsigned* grid_counter, float* scale_shift, float* mean_invstd, uint32_t* mask_bits,
                  cudaStream_t st);

static int make_geom(ConvGeom& g, int N, int Cin, int Cout, int D, int H, int W, int KD, int pad, const char* who) {
  if (!(N > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0)) {
    set_error("%s: extents must be positive (N=%d Cin=%d Cout=%d D=%d H=%d W=%d)", who, N, Cin, Cout, D, H, W);
    return -1;
  }
  if (!(KD == 1 || KD == 3)) {
    set_error("%s: KD must be 1 (3x3) or 3 (3x3x3), got %d", who, KD);
    return -1;
  }
  if (pad < 0 || pad > 2) {
    set_error("%s: pad must be 0, 1 or 2, got %d", who, pad);
    return -1;
  }
  g.N = N; g.Cin = Cin; g.Cout = Cout;
  g.Di = D; g.Hi = H; g.Wi = W;
  g.KD = KD; g.taps = KD * 9; g.pad = pad; g.pad_d = (KD == 3) ? pad : 0;
  g.stats_stride = 0;
  g.Do = D + 2 * g.pad_d - (KD - 1);
  g.Ho = H + 2 * pad - 2;
  g.Wo = W + 2 * pad - 2;
  if (g.Do <= 0 || g.Ho <= 0 || g.Wo <= 0) {
    set_error("%s: empty output (%d x %d x %d) for input %d x %d x %d, pad %d", who, g.Do, g.Ho, g.Wo, D, H, W, pad);
    return -1;
  }
  return 0;
}
}  // namespace hpvg

using namespace hpvg;

extern "C" {

int hpvg_conv_forward(const void* x, int x_fmt, const float* w_f32, const void* w_packed, const float* bias, void* y, int y_fmt, int N,
                      int Cin, int Cout, int D, int H, int W, int KD, int pad, int transposed, int act, float lrelu_slope, float* stats,
                      const void* mask_src, void* stream) {
  return hpvg_conv_forward_ex(x, x_fmt, w_f32, w_packed, bias, y, y_fmt, N, Cin, Cout, D, H, W, KD, pad, transposed, act, lrelu_slope,
                              stats, 0, mask_src, stream);
}

int hpvg_conv_forward_ex(const void* x, int x_fmt, const float* w_f32, const void* w_packed, const float* bias, void* y, int y_fmt, int N,
                         int Cin, int Cout, int D, int H, int W, int KD, int pad, int transposed, int act, float lrelu_slope,
                         float* stats, int stats_per_sample, const void* mask_src, void* stream) {
  ConvGeom g;
  if (int rc = make_geom(g, N, Cin, Cout, D, H, W, KD, pad, "conv_forward")) return rc;
  g.stats_stride = (stats && stats_per_sample) ? 2 * Cout : 0;
  HPVG_CHECK_ARG(x && y, "conv_forward: null tensor");
  HPVG_CHECK_ARG((x_fmt == 0 || x_fmt == 1) && (y_fmt == 0 || y_fmt == 1), "conv_forward: unknown tensor format");
  HPVG_CHECK_ARG(act == HPVG_ACT_NONE || act == HPVG_ACT_LRELU, "conv_forward: unknown activation %d", act);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int backend = conv_backend();
  const bool tc_ok = conv_tc_supported(x_fmt, y_fmt, g, w_packed) &&
                     (y_fmt == HPVG_FMT_NDHWC_BF16 || (act == HPVG_ACT_NONE && !stats && !mask_src));
  // thin_gs.cu takes the packed image with exactly Cout rows per tap; the thin tcgen05 kernel of conv_tc.cu takes 16 rows per tap:
  // the caller says which one it packed through `transposed` bit 1 (hpvg.ops sets it when it packed for thin_gs)
  const bool gs_ok = w_packed != nullptr && (transposed & 2) && act == HPVG_ACT_NONE && !stats && !mask_src && thin_gs_supported(x_fmt, y_fmt, g);
  transposed &= 1;
  if (backend == HPVG_BACKEND_TCGEN05 && !tc_ok && !gs_ok) {
    set_error("conv_forward: tcgen05 backend required but shape/format unsupported (Cin=%d Cout=%d fmt %d->%d packed=%p)", Cin, Cout,
              x_fmt, y_fmt, w_packed);
    return -1;
  }
  const double flops = 2.0 * g.N * g.Do * g.Ho * g.Wo * (double)g.Cin * g.Cout * g.taps;
  if (backend != HPVG_BACKEND_DIRECT && gs_ok) {
    // wide -> thin (64 -> <= 4 channels): one GEMM per input slab + shift-add gather (thin_gs.cu); w_packed = [taps][Cout][64]
    void* ph = prof_begin(HPVG_PROF_CONV_THIN, flops, st);
    int rc = thin_conv_gs(x, w_packed, bias, reinterpret_cast<float*>(y), g, st);
    prof_end(ph, st);
    return rc;
  }
  if (tc_ok && backend != HPVG_BACKEND_DIRECT) {
    void* ph = prof_begin(HPVG_PROF_CONV_TC, flops, st);
    int rc = conv_tc(x, w_packed, bias, y, y_fmt, g, act, lrelu_slope, stats, mask_src, st);
    prof_end(ph, st);
    return rc;
  }
  HPVG_CHECK_ARG(w_f32 != nullptr, "conv_forward: the CUDA-core kernel needs the float32 weights");
  HPVG_CHECK_ARG(mask_src == nullptr || y_fmt == HPVG_FMT_NDHWC_BF16, "conv_forward: mask_src requires an NDHWC_BF16 output");
  if (backend != HPVG_BACKEND_DIRECT && expand_conv_supported(x_fmt, y_fmt, g, mask_src)) {
    void* ph = prof_begin(HPVG_PROF_CONV_EXPAND, flops, st);
    int rc = expand_conv(x, w_f32, reinterpret_cast<const float*>(w_packed), bias, y, g, transposed, act, lrelu_slope, stats, st);
    prof_end(ph, st);
    return rc;
  }
  HPVG_CHECK_ARG(g.stats_stride == 0, "conv_forward: per-sample statistics need the tcgen05 or the expand kernel (Cin=%d Cout=%d)", Cin,
                 Cout);
  void* ph = prof_begin(HPVG_PROF_CONV_DIRECT, flops, st);
  int rc = conv_direct(x, x_fmt, w_f32, bias, y, y_fmt, g, transposed, act, lrelu_slope, stats, mask_src, st);
  prof_end(ph, st);
  return rc;
}

int hpvg_conv_bn_lrelu_fused_supported(int N, int Cin, int Cout, int D, int H, int W, int KD, int pad) {
  ConvGeom g;
  if (make_geom(g, N, Cin, Cout, D, H, W, KD, pad, "conv_bn_lrelu_fused_supported")) return 0;
  return (conv_backend() != HPVG_BACKEND_DIRECT && conv_bn_fused_supported(g)) ? 1 : 0;
}

int hpvg_conv_bn_lrelu_fused(const void* x, const void* w_packed, const float* bias, void* y, void* out, int N, int Cin, int Cout,
                             int D, int H, int W, int KD, int pad, float slope, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                             float* stats, float* scale_shift, float* mean_invstd, void* mask_bits, void* stream) {
  ConvGeom g;
  if (int rc = make_geom(g, N, Cin, Cout, D, H, W, KD, pad, "conv_bn_lrelu_fused")) return rc;
  HPVG_CHECK_ARG(x && w_packed && out && gamma && beta && stats && scale_shift && mean_invstd, "conv_bn_lrelu_fused: null tensor");
  HPVG_CHECK_ARG(conv_bn_fused_supported(g), "conv_bn_lrelu_fused: layer not eligible (Cin=%d Cout=%d KD=%d, %d x %d x %d)", Cin, Cout,
                 KD, g.Do, g.Ho, g.Wo);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const double flops = 2.0 * g.N * g.Do * g.Ho * g.Wo * (double)g.Cin * g.Cout * g.taps;
  void* ph = prof_begin(HPVG_PROF_CONV_BN_FUSED, flops, st);
  int rc = conv_bn_fused(x, w_packed, bias, y, out, g, slope, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps,
                         stats, reinterpret_cast<unsigned*>(stats + 2 * Cout), scale_shift, mean_invstd,
                         reinterpret_cast<uint32_t*>(mask_bits), st);
  prof_end(ph, st);
  return rc;
}

int hpvg_narrow_kernel_choice(int c_thin, int KD) {
  ConvGeom g = {};
  g.N = 1; g.Cin = c_thin; g.Cout = 64; g.KD = KD; g.taps = KD * 9;
  return (expand_tc_supported(g) && narrow_wgrad_tc_supported(g, true)) ? 1 : 0;
}

int hpvg_conv_kernel_choice(int N, int Cin, int Cout, int D, int H, int W, int KD, int pad, int x_fmt, int y_fmt) {
  ConvGeom g;
  if (int rc = make_geom(g, N, Cin, Cout, D, H, W, KD, pad, "conv_kernel_choice")) return rc;
  const void* packed = reinterpret_cast<const void*>(uintptr_t(1));      // "a packed weight image exists": only compared with NULL
  if (conv_tc_supported(x_fmt, y_fmt, g, packed)) return conv_tc_picks_column(y_fmt, g, packed) ? HPVG_KERNEL_TC_COLUMN : HPVG_KERNEL_TC_BRICK;
  if (expand_conv_supported(x_fmt, y_fmt, g, nullptr)) return HPVG_KERNEL_EXPAND;
  return HPVG_KERNEL_DIRECT;
}

int hpvg_pack_weights_expand(const float* w_f32, float* w_tco, int Cin, int taps, int transposed, void* stream) {
  HPVG_CHECK_ARG(w_f32 && w_tco && Cin > 0 && Cin <= 4 && (taps == 9 || taps == 27), "pack_weights_expand: bad arguments");
  return pack_expand(w_f32, w_tco, Cin, taps, transposed, reinterpret_cast<cudaStream_t>(stream));
}

size_t hpvg_conv_wgrad_workspace(int N, int Cin, int Cout, int D, int H, int W, int KD, int pad, int x_fmt, int gy_fmt) {
  ConvGeom g;
  if (make_geom(g, N, Cin, Cout, D, H, W, KD, pad, "conv_wgrad_workspace")) return 0;
  if (wgrad_tc_supported(x_fmt, gy_fmt, g)) return wgrad_tc_workspace(g);
  if (narrow_wgrad_supported(x_fmt, gy_fmt, g)) return narrow_wgrad_workspace(x_fmt, g);
  return 0;
}

int hpvg_conv_wgrad(const void* x, int x_fmt, const void* gy, int gy_fmt, float* dw, float* dbias, int N, int Cin, int Cout, int D, int H,
                    int W, int KD, int pad, void* workspace, size_t workspace_bytes, void* stream) {
  ConvGeom g;
  if (int rc = make_geom(g, N, Cin, Cout, D, H, W, KD, pad, "conv_wgrad")) return rc;
  HPVG_CHECK_ARG(x && gy && dw, "conv_wgrad: null tensor");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int backend = conv_backend();
  const bool tc_ok = wgrad_tc_supported(x_fmt, gy_fmt, g);
  if (backend == HPVG_BACKEND_TCGEN05 && !tc_ok) {
    set_error("conv_wgrad: tcgen05 backend required but shape/format unsupported (Cin=%d Cout=%d)", Cin, Cout);
    return -1;
  }
  int rc;
  const double flops = 2.0 * g.N * g.Do * g.Ho * g.Wo * (double)g.Cin * g.Cout * g.taps;
  if (tc_ok && backend != HPVG_BACKEND_DIRECT) {
    void* ph = prof_begin(HPVG_PROF_WGRAD_TC, flops, st);
    rc = wgrad_tc(x, gy, dw, g, workspace, workspace_bytes, st);
    prof_end(ph, st);
  } else if (backend != HPVG_BACKEND_DIRECT && narrow_wgrad_supported(x_fmt, gy_fmt, g)) {
    void* ph = prof_begin(HPVG_PROF_WGRAD_NARROW, flops, st);
    const bool head = x_fmt == HPVG_FMT_NCDHW_F32;
    rc = narrow_wgrad(x, x_fmt, gy, dw, head ? dbias : nullptr, g, workspace, workspace_bytes, st);
    prof_end(ph, st);
    if (rc) return rc;
    if (head) return 0;      // the bias gradient (channel sum of the wide gy) was produced by the same kernel
  } else {
    void* ph = prof_begin(HPVG_PROF_WGRAD_DIRECT, flops, st);
    rc = wgrad_direct(x, x_fmt, gy, gy_fmt, dw, g, st);
    prof_end(ph, st);
  }
  if (rc) return rc;
  if (dbias) return hpvg_channel_sum(gy, gy_fmt, dbias, N, Cout, (long long)g.Do * g.Ho * g.Wo, stream);
  return 0;
}

}  // extern "C"

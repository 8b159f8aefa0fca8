// The end of a training iteration: torch.nn.utils.clip_grad_norm_ (train_video.py:201,
// train_image.py:216) and the Adam steps (train_video.py:55,88,183,202; train_video_baselines.py:51,70) as TWO launches
// over all parameter tensors of an optimizer: a fixed-order squared-norm reduction that ends in the clip coefficient,
// and one multi-tensor pass that scales the gradients, updates both moments and the parameters.  torch.optim.Adam with
// the reference's arguments: eps 1e-8, no weight decay, no amsgrad, bias correction on.
#include "common.cuh"

namespace hpvg {

struct OptBatch {
  int n;
  float* p[HPVG_OPT_MAX_TENSORS];
  float* g[HPVG_OPT_MAX_TENSORS];
  float* m[HPVG_OPT_MAX_TENSORS];     // nullptr: the tensor is not owned by the optimizer, only its gradient is scaled
  float* v[HPVG_OPT_MAX_TENSORS];
  long long numel[HPVG_OPT_MAX_TENSORS];
  float lr[HPVG_OPT_MAX_TENSORS];
};

// state[0] = step count, [1] = clip coefficient, [2] = total gradient norm, [4], [5] = block tickets (unsigned, self-resetting)

__device__ __forceinline__ float block_sum_256(float s, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
  }
  return t;     // valid in thread 0
}

// partials[slot_base + tensor * gridDim.x + block] = sum of squares of that block's share; when `finalize`, the block that
// arrives last adds all `total_slots` partials in index order (the same order on every run) and writes the coefficient
//   min(1, max_norm / (||g|| + 1e-6))                                   torch/nn/utils/clip_grad.py
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const OptBatch b, float* __restrict__ partials, int slot_base, int total_slots,
                                                          int finalize, float max_norm, float* state) {
  __shared__ float red[8];
  __shared__ int is_last;
  pdl_enter();
  const int l = blockIdx.y;
  const long long n = b.numel[l];
  const float* __restrict__ g = b.g[l];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float x = g[i];
    s = fmaf(x, x, s);
  }
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) {
    partials[slot_base + l * gridDim.x + blockIdx.x] = s;
    is_last = 0;
    if (finalize) {
      __threadfence();
      const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(state) + 4, 1u);
      is_last = (ticket == gridDim.x * gridDim.y - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < total_slots; i += 256) t += __ldcg(partials + i);
  t = block_sum_256(t, red);
  if (threadIdx.x == 0) {
    const float total = sqrtf(t);
    state[2] = total;
    state[1] = fminf(max_norm / (total + 1e-6f), 1.0f);
    reinterpret_cast<unsigned*>(state)[4] = 0u;
  }
}

// torch/optim/adam.py (_single_tensor_adam / the fused kernel's adam_math), float32 arithmetic in the same order:
//   g <- g * coef (written back: clip_grad_norm_ scales .grad in place);  m <- lerp(m, g, 1 - beta1);  v <- beta2 v + (1 - beta2) g g
//   p <- p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps),   t = step + 1
__global__ void __launch_bounds__(256) adam_step_kernel(const OptBatch b, double beta1, double beta2, double eps, int use_clip, int advance,
                                                        float* state) {
  __shared__ float s_bc1, s_bc2_sqrt, s_coef;
  pdl_enter();
  if (threadIdx.x == 0) {
    const double t = (double)state[0] + 1.0;
    s_bc1 = (float)(1.0 - pow(beta1, t));
    s_bc2_sqrt = (float)sqrt(1.0 - pow(beta2, t));
    s_coef = use_clip ? state[1] : 1.0f;
  }
  __syncthreads();
  const int l = blockIdx.y;
  const long long n = b.numel[l];
  float* __restrict__ gp = b.g[l];
  const float coef = s_coef;
  if (b.m[l] == nullptr) {
    if (use_clip)
      for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) gp[i] *= coef;
  } else {
    float* __restrict__ pp = b.p[l];
    float* __restrict__ mp = b.m[l];
    float* __restrict__ vp = b.v[l];
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), epsf = (float)eps;
    const float step_size = b.lr[l] / s_bc1, bc2s = s_bc2_sqrt;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
      float g = gp[i];
      if (use_clip) {
        g *= coef;
        gp[i] = g;
      }
      float m = mp[i], v = vp[i];
      const float d = g - m;
      m = (w1 < 0.5f) ? fmaf(w1, d, m) : g - d * (1.0f - w1);      // at::lerp's two branches
      v = b2 * v + w2 * g * g;
      const float denom = sqrtf(v) / bc2s + epsf;
      mp[i] = m;
      vp[i] = v;
      pp[i] -= step_size * m / denom;
    }
  }
  if (!advance) return;
  // every block read state[0] before taking its ticket; the last ticket holder advances the step count
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(state) + 5, 1u);
    if (ticket == gridDim.x * gridDim.y - 1) {
      state[0] += 1.0f;
      reinterpret_cast<unsigned*>(state)[5] = 0u;
    }
  }
}

static int fill(OptBatch& b, int n, const long long* numel, const char* who) {
  if (n <= 0 || n > HPVG_OPT_MAX_TENSORS) {
    set_error("%s: %d tensors per call (1..%d)", who, n, HPVG_OPT_MAX_TENSORS);
    return -1;
  }
  b.n = n;
  for (int l = 0; l < HPVG_OPT_MAX_TENSORS; ++l) {
    b.p[l] = b.g[l] = b.m[l] = b.v[l] = nullptr;
    b.numel[l] = 0;
    b.lr[l] = 0.f;
  }
  for (int l = 0; l < n; ++l) {
    if (numel[l] < 0) {
      set_error("%s: tensor %d has a negative size", who, l);
      return -1;
    }
    b.numel[l] = numel[l];
  }
  return 0;
}

}  // namespace hpvg

using namespace hpvg;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int hpvg_grad_clip_coef(int n, const float* const* grads, const long long* numel, float* partials, int slot_base, int total_slots,
                        int finalize, float max_norm, float* state, void* stream) {
  HPVG_CHECK_ARG(grads && numel && partials && state, "grad_clip_coef: null argument");
  HPVG_CHECK_ARG(slot_base >= 0 && slot_base + n * HPVG_OPT_BLOCKS <= total_slots, "grad_clip_coef: partial slots %d + %d x %d exceed %d",
                 slot_base, n, HPVG_OPT_BLOCKS, total_slots);
  HPVG_CHECK_ARG(max_norm > 0.f, "grad_clip_coef: max_norm must be positive");
  OptBatch b;
  if (int rc = fill(b, n, numel, "grad_clip_coef")) return rc;
  for (int l = 0; l < n; ++l) {
    HPVG_CHECK_ARG(grads[l] || numel[l] == 0, "grad_clip_coef: gradient %d is null", l);
    b.g[l] = const_cast<float*>(grads[l]);
  }
  launch_k(grad_sqnorm_kernel, dim3(HPVG_OPT_BLOCKS, n), 256, 0, ST(stream), b, partials, slot_base, total_slots, finalize, max_norm, state);
  HPVG_CHECK_LAUNCH("grad_sqnorm");
  return 0;
}

int hpvg_adam_step(int n, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const long long* numel, const float* lr, double beta1, double beta2, double eps, int use_clip, int advance_step,
                   float* state, void* stream) {
  HPVG_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel && lr && state, "adam_step: null argument");
  HPVG_CHECK_ARG(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "adam_step: bad hyper-parameters");
  OptBatch b;
  if (int rc = fill(b, n, numel, "adam_step")) return rc;
  for (int l = 0; l < n; ++l) {
    HPVG_CHECK_ARG(grads[l] || numel[l] == 0, "adam_step: gradient %d is null", l);
    HPVG_CHECK_ARG((exp_avg[l] == nullptr) == (exp_avg_sq[l] == nullptr), "adam_step: tensor %d has one moment but not the other", l);
    HPVG_CHECK_ARG(exp_avg[l] == nullptr || params[l], "adam_step: parameter %d is null", l);
    b.p[l] = params[l]; b.g[l] = grads[l]; b.m[l] = exp_avg[l]; b.v[l] = exp_avg_sq[l]; b.lr[l] = lr[l];
  }
  launch_k(adam_step_kernel, dim3(HPVG_OPT_BLOCKS, n), 256, 0, ST(stream), b, beta1, beta2, eps, use_clip, advance_step, state);
  HPVG_CHECK_LAUNCH("adam_step");
  return 0;
}

}  // extern "C"

// Data formats on either side of the path: byte <-> float conversions that the reference does on
// the host, per iteration, in DataLoader workers (datasets/video.py:44-92) and when writing results (utils/saver.py:8-19).
// Both are bit-exact restatements (IEEE round-to-nearest operations in the reference's order, no FMA contraction).
#include "common.cuh"

namespace hpvg {

// clip[c][t][h][w] = ((frames[f0 + t*every][h][w'][c] / 255) - 0.5) / 0.5 ,  w' = hflip ? W-1-w : w
//   datasets/video.py:52-54 (slice + /255), :75 K.hflip, :78 K.normalize(x, 0.5, 0.5), :81 permute to CTHW
__global__ void __launch_bounds__(256) clip_from_frames_kernel(const uint8_t* __restrict__ frames, float* __restrict__ clip, int f0,
                                                               int every, int T, int H, int W, int hflip) {
  pdl_enter();
  const long long total = (long long)3 * T * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int t = (int)(r % T);
    const int c = (int)(r / T);
    const int ws = hflip ? W - 1 - w : w;
    const uint8_t u = frames[(((size_t)(f0 + t * every) * H + h) * W + ws) * 3 + c];
    const float v = __fdiv_rn((float)u, 255.0f);
    clip[i] = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
  }
}

// out[t][h][w][c] = uint8((video[c][t][h][w] + 1) * 127.5)      utils/saver.py:16-18 (float32 arithmetic, C truncation)
__global__ void __launch_bounds__(256) frames_to_uint8_kernel(const float* __restrict__ video, uint8_t* __restrict__ out, int T, int H,
                                                              int W) {
  pdl_enter();
  const long long total = (long long)T * H * W * 3;
  const size_t plane = (size_t)T * H * W;
  video += (size_t)blockIdx.y * 3 * plane;      // blockIdx.y = video of a batch
  out += (size_t)blockIdx.y * 3 * plane;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    const long long vox = i / 3;
    const float v = __fmul_rn(__fadd_rn(video[(size_t)c * plane + vox], 1.0f), 127.5f);
    // numpy's float32 -> uint8 cast goes through a signed integer conversion and keeps the low byte
    out[i] = (uint8_t)(int)v;
  }
}

}  // namespace hpvg

using namespace hpvg;

extern "C" {

int hpvg_clip_from_frames(const uint8_t* frames, float* clip, int num_frames, int first, int every, int T, int H, int W, int hflip,
                          void* stream) {
  HPVG_CHECK_ARG(frames && clip && T > 0 && H > 0 && W > 0 && every > 0 && first >= 0, "clip_from_frames: bad arguments");
  HPVG_CHECK_ARG(first + (T - 1) * every < num_frames, "clip_from_frames: frames %d..%d step %d exceed the %d resident frames", first,
                 first + (T - 1) * every, every, num_frames);
  const long long total = (long long)3 * T * H * W;
  const int blocks = (int)max(1LL, min(cdiv(total, 256), (long long)num_sms() * 8));
  launch_k(clip_from_frames_kernel, blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream), frames, clip, first, every, T, H, W, hflip);
  HPVG_CHECK_LAUNCH("clip_from_frames");
  return 0;
}

int hpvg_frames_to_uint8(const float* video, uint8_t* out, int T, int H, int W, void* stream) {
  return hpvg_frames_to_uint8_batched(video, out, 1, T, H, W, stream);
}

int hpvg_frames_to_uint8_batched(const float* video, uint8_t* out, int N, int T, int H, int W, void* stream) {
  HPVG_CHECK_ARG(video && out && N > 0 && N <= 65535 && T > 0 && H > 0 && W > 0, "frames_to_uint8: bad arguments");
  const long long total = (long long)3 * T * H * W;
  const int blocks = (int)max(1LL, min(cdiv(total, 256), cdiv((long long)num_sms() * 8, N)));
  launch_k(frames_to_uint8_kernel, dim3(blocks, N), 256, 0, reinterpret_cast<cudaStream_t>(stream), video, out, T, H, W);
  HPVG_CHECK_LAUNCH("frames_to_uint8");
  return 0;
}

}  // extern "C"

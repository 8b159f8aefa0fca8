// Shared device/host helpers for libhpvg (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/hpvg.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL) && !defined(__CUDA_ARCH_FEAT_SM103_ALL)
#error "libhpvg is written for sm_100a; compile with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace hpvg {

// ---------------------------------------------------------------------------------------------------------------
// host side: error reporting + launch accounting
// ---------------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int conv_backend();
// per-launch device timing for bench.py (hpvg_profile_enable): returns an opaque handle or nullptr when off
void* prof_begin(int kind, double work, cudaStream_t st);
void prof_end(void* handle, cudaStream_t st);
long long* debug_clock_buffer();   // development aid: device buffer for in-kernel phase clocks (nullptr = off)

#define HPVG_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::hpvg::set_error(__VA_ARGS__);             \
      return -1;                                  \
    }                                             \
  } while (0)

#define HPVG_CHECK_LAUNCH(name)                                                            \
  do {                                                                                     \
    cudaError_t e_ = cudaGetLastError();                                                   \
    if (e_ != cudaSuccess) {                                                               \
      ::hpvg::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));            \
      return -2;                                                                           \
    }                                                                                      \
    ::hpvg::count_launch();                                                                \
  } while (0)

static inline int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

// cudaFuncSetAttribute applies to the CURRENT device: a process that drives several GPUs (nn.DataParallel's per-device
// threads, train_video.py:91-94) must opt in to large dynamic shared memory once per device, not once per process
static inline bool attr_pending(std::atomic<unsigned long long>& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  return (mask.load(std::memory_order_acquire) & (1ull << (dev & 63))) == 0;
}
static inline void attr_set(std::atomic<unsigned long long>& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
}

static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// Programmatic dependent launch (off by default, HPVG_PDL=1 / hpvg_set_pdl(1) turn it on): every libhpvg kernel is then launched with the programmatic
// stream-serialization attribute and starts with pdl_enter() (or, in the tcgen05 kernels, runs its barrier / TMEM /
// descriptor set-up first and then pdl_wait()).  The next kernel of the stream is scheduled while this one drains and
// blocks in griddepcontrol.wait until this grid has completed and flushed, so a chain of dependent launches — a pyramid
// pass is a few hundred of them — does not pay the launch latency between every pair.  Inside a stream capture the
// attribute becomes a programmatic edge of the CUDA graph.
bool pdl_enabled();
int set_pdl(int on);
// which tcgen05 kernel runs the 3-D 64-channel layers: -1 = per layer, by how well the brick kernel's units fill the SMs
// (default, rule in conv_tc.cu), 0 = brick kernel (conv_tc.cu) always, 1 = column-streaming kernel (conv_col.cu) whenever it
// supports the layer.  Environment HPVG_TC_COL sets the initial value.
int conv_col_mode();
int set_conv_col_mode(int mode);
// HPVG_CARVEOUT=1 (experimental, unmeasured): every libhpvg kernel asks for the maximum shared-memory carve-out, so that SMs
// do not have to drain and re-partition L1 / shared memory between the small-footprint kernels and the 220 KB tcgen05 kernels
// (kernels with different carve-outs cannot share an SM, which also limits the overlap of the iteration's streams)
bool carveout_enabled();
void prefer_max_smem(const void* kernel);
int wgrad_mode();               // 0 = one kd per CTA (default), 1 = kd-stacked N = 192 form where KD == 3, 2 = default kernel with the staged drain (wgrad_tc.cu)
int set_wgrad_mode(int mode);

#ifdef __CUDACC__
template <typename... P, typename... A>
static inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (carveout_enabled()) prefer_max_smem(reinterpret_cast<const void*>(kernel));
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
#endif

// geometry of one convolution call (k = 3 on H and W, KD in {1,3} on D, stride 1)
struct ConvGeom {
  int N, Cin, Cout;
  int Di, Hi, Wi;  // input extents
  int Do, Ho, Wo;  // output extents = input + 2*pad - 2 on every filtered axis
  int KD, taps, pad, pad_d;
  int stats_stride;  // floats between the BatchNorm-sum blocks of consecutive samples (0: one [2*Cout] block for the batch)
};

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// griddepcontrol.wait: returns once every grid this launch programmatically depends on has completed and its writes are
// visible (at once when there is no such dependency).  launch_dependents: lets the next kernel of the stream be scheduled
// as soon as every CTA of this grid has executed it (or exited).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_trigger();
  pdl_wait();
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(p);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a converged warp (elect.sync): ptxas treats the guarded region as single-threaded, so values feeding
// tcgen05 / TMA instructions move to uniform registers without a per-instruction broadcast loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must fail the launch (trap -> cudaErrorLaunchFailure), never hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) {
      printf("hpvg: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Grid-wide barrier for kernels whose whole grid is resident (at most one CTA per SM, cooperative launch): ONE thread per CTA
// arrives on a zero-initialised global counter and spins until all `expected` CTAs have; the caller brackets it with its own
// CTA-level barriers and makes its global writes visible with __threadfence() first.  Bounded like mbar_wait: a protocol or
// residency bug fails the launch instead of hanging the GPU.  (CTAs of a grid are dispatched before any CTA of a later grid, so
// a grid that fits the machine always becomes fully resident: measured with two such grids racing on two streams,
// experiments/coop_concurrency.py.)
__device__ __forceinline__ void grid_barrier_arrive_and_wait(unsigned* counter, unsigned expected) {
  __threadfence();
  atomicAdd(counter, 1u);
  unsigned seen;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    if (seen < expected && clock64() - t0 > 2000000000LL) {
      printf("hpvg: grid barrier timed out (block %d saw %u of %u)\n", blockIdx.x, seen, expected);
      __trap();
    }
  } while (seen < expected);
}

// ---- TMA ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(m), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(m), "r"(src), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
#ifndef HPVG_STORE_WAIT_READ
#define HPVG_STORE_WAIT_READ 1
#endif
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05 --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// shared-memory matrix descriptor (sm_100 format, see experiments/umma_desc_probe.cu for the verified semantics)
//   layout: 0 = no swizzle, 2 = 128-byte swizzle
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

// instruction descriptor: bf16 x bf16 -> f32, M x N, operand majors (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers (thread = lane, register j = column j)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------------------------
// host: tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// bf16 tensor map over a dense [.. outer dims ..][C] array: dims[0] = C (innermost), box[0] = 64 channels (128 B),
// 128-byte swizzle (swizzle_bytes = 64: box[0] = 32 channels, 64-byte swizzle).  rank <= 5.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint32_t* box, int swizzle_bytes = 128);

}  // namespace hpvg

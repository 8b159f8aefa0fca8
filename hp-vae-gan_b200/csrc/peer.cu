// Gradient averaging over NVLink peer memory: ONE kernel per backward instead of a library collective (include/hpvg.h,
// "Gradient averaging over NVLink peer memory").
//
// Replaces the reduction of the replicas' gradients in nn.DataParallel's backward (train_video.py:91-94, :182, :200) for the
// one-process-per-GPU mode.  The buckets are small (1.3 MB for the critic, 2.7 - 5.4 MB for the generator at BASELINE configs[1]),
// so what a collective costs here is latency, not bandwidth: the kernel is two flag exchanges around one pull of a 1/world slice
// from every peer and one push of the mean to every peer — every GPU reads and writes (world - 1) / world of a bucket over
// NVLink / NVSwitch, all pairs at once.
//
// Memory model: the flags are written with st.release.sys and polled with ld.acquire.sys; a rank's bucket was filled by earlier
// kernels of its stream (performed before this kernel starts), the pushed means are followed by __threadfence_system() in the
// writing thread and a CTA barrier before the flag thread releases.  Peer data is pulled with ld.volatile (never from L1).
#include "common.cuh"

namespace hpvg {

constexpr int PEER_MAX_BLOCKS = 64, PEER_THREADS = 256, PEER_UNROLL = 4;
constexpr int PEER_EPOCH_OFF = 2 * PEER_MAX_BLOCKS * HPVG_PEER_MAX_RANKS;      // unsigned index of the per-CTA call counts in the pad
static_assert((PEER_EPOCH_OFF + PEER_MAX_BLOCKS) * 4 <= HPVG_PEER_SIGNAL_BYTES, "signal pad layout");

struct PeerArgs {
  float4* buf[HPVG_PEER_MAX_RANKS];
  unsigned* sig[HPVG_PEER_MAX_RANKS];
  int rank, world;
  long long n4;      // float4 elements per rank slice
  float scale;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// CTA b of this rank meets CTA b of every other rank: thread q tells rank q "call e, phase ph reached" and waits for rank q's word
__device__ __forceinline__ void peer_barrier(const PeerArgs& a, int ph, unsigned e) {
  if ((int)threadIdx.x < a.world) {
    const int q = threadIdx.x;
    const int slot = (ph * PEER_MAX_BLOCKS + blockIdx.x) * HPVG_PEER_MAX_RANKS;
    st_release_sys(a.sig[q] + slot + a.rank, e);
    const unsigned* theirs = a.sig[a.rank] + slot + q;
    unsigned long long t0 = 0;
    while ((int)(ld_acquire_sys(theirs) - e) < 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      if (t - t0 > 60000000000ull) {
        printf("hpvg: peer all-reduce timed out (rank %d block %d phase %d waits for rank %d, call %u)\n", a.rank, blockIdx.x, ph, q, e);
        __trap();
      }
    }
  }
  __syncthreads();
}

// rank r's share of the work: pull slice r from every bucket, sum in rank order, push the mean into slice r of every bucket
__device__ __forceinline__ void peer_reduce_slice(const PeerArgs& a) {
  const int W = a.world;
  const long long base = (long long)a.rank * a.n4;
  const long long step = (long long)gridDim.x * PEER_THREADS;
  for (long long i0 = (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i0 < a.n4; i0 += step * PEER_UNROLL) {
    float4 v[PEER_UNROLL][HPVG_PEER_MAX_RANKS];
#pragma unroll
    for (int k = 0; k < PEER_UNROLL; ++k) {
      const long long i = i0 + k * step;
#pragma unroll
      for (int q = 0; q < HPVG_PEER_MAX_RANKS; ++q)
        if (q < W && i < a.n4) v[k][q] = ld_volatile_f4(a.buf[q] + base + i);
    }
#pragma unroll
    for (int k = 0; k < PEER_UNROLL; ++k) {
      const long long i = i0 + k * step;
      if (i >= a.n4) break;
      float4 s = v[k][0];      // rank order: the same sum on every rank, whoever arrives first
#pragma unroll
      for (int q = 1; q < HPVG_PEER_MAX_RANKS; ++q)
        if (q < W) {
          s.x += v[k][q].x; s.y += v[k][q].y; s.z += v[k][q].z; s.w += v[k][q].w;
        }
      s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
#pragma unroll
      for (int q = 0; q < HPVG_PEER_MAX_RANKS; ++q)
        if (q < W) a.buf[q][base + i] = s;
    }
  }
}

__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_kernel(const PeerArgs a) {
  __shared__ unsigned s_epoch;
  unsigned* epoch = a.sig[a.rank] + PEER_EPOCH_OFF + blockIdx.x;
  if (threadIdx.x == 0) s_epoch = *epoch + 1u;
  __syncthreads();
  const unsigned e = s_epoch;
  peer_barrier(a, 0, e);      // every rank's bucket is filled
  peer_reduce_slice(a);
  __threadfence_system();      // this thread's pushes are performed at system scope before the CTA's flags go out
  __syncthreads();
  peer_barrier(a, 1, e);       // all means have landed here; nobody reads this rank's bucket any more
  if (threadIdx.x == 0) *epoch = e;
}

// ---- the same exchange with the bucket filled from, and emptied into, the gradient tensors by the kernel itself -----------------
// Bucket layout: tensor t occupies float4 slots [start4[t], start4[t + 1]) (its last slot zero-padded), so every slot belongs to one
// tensor and is 16-byte aligned on both sides.  CTA b packs and unpacks exactly the slots CTA b of the peers pulls and pushes — slot
// i of every slice with (i / PEER_THREADS) % gridDim.x == b — so the per-CTA flag exchange still orders everything.
struct PeerTensors {
  float* ptr[HPVG_PEER_MAX_TENSORS];
  unsigned start4[HPVG_PEER_MAX_TENSORS + 1];
  unsigned numel[HPVG_PEER_MAX_TENSORS];
  int n;
};

template <bool PACK>
__device__ __forceinline__ void peer_copy_slots(const PeerArgs& a, float* const* s_ptr, const unsigned* s_start4, const unsigned* s_numel, int nt) {
  float4* mine = a.buf[a.rank];
  const long long total4 = s_start4[nt];
  const long long step = (long long)gridDim.x * PEER_THREADS;
  for (int q = 0; q < a.world; ++q) {
    const long long qbase = (long long)q * a.n4;
    if (qbase >= total4) break;      // this slice and the ones behind it are padding
    // PEER_UNROLL slots per pass: all look-ups, then all loads, then all stores (one slot at a time the loop ran at one memory round
    // trip per slot: 18 us for 2.7 MB of gradients on 64 CTAs)
    for (long long i0 = (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i0 < a.n4; i0 += step * PEER_UNROLL) {
      float* g[PEER_UNROLL];
      unsigned left[PEER_UNROLL];
      float4 v[PEER_UNROLL];
#pragma unroll
      for (int k = 0; k < PEER_UNROLL; ++k) {
        const long long i = i0 + k * step, idx = qbase + i;
        left[k] = 0u;
        g[k] = nullptr;
        if (i < a.n4 && idx < total4) {
          int lo = 0, hi = nt - 1;       // largest t with start4[t] <= idx
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_start4[mid] <= (unsigned)idx) lo = mid; else hi = mid - 1;
          }
          const unsigned off = ((unsigned)idx - s_start4[lo]) * 4u;
          left[k] = s_numel[lo] - off;      // > 0: floats of the tensor from this slot on
          g[k] = s_ptr[lo] + off;
        }
      }
#pragma unroll
      for (int k = 0; k < PEER_UNROLL; ++k) {
        if (left[k] == 0u) continue;
        if (PACK) {
          v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (left[k] >= 4u) v[k] = *reinterpret_cast<const float4*>(g[k]);
          else {
            v[k].x = g[k][0];
            if (left[k] > 1u) v[k].y = g[k][1];
            if (left[k] > 2u) v[k].z = g[k][2];
          }
        } else {
          v[k] = ld_volatile_f4(mine + qbase + i0 + k * step);
        }
      }
#pragma unroll
      for (int k = 0; k < PEER_UNROLL; ++k) {
        if (left[k] == 0u) continue;
        if (PACK) {
          mine[qbase + i0 + k * step] = v[k];
        } else if (left[k] >= 4u) {
          *reinterpret_cast<float4*>(g[k]) = v[k];
        } else {
          g[k][0] = v[k].x;
          if (left[k] > 1u) g[k][1] = v[k].y;
          if (left[k] > 2u) g[k][2] = v[k].z;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_tensors_kernel(const PeerArgs a, const PeerTensors t) {
  __shared__ unsigned s_epoch;
  __shared__ float* s_ptr[HPVG_PEER_MAX_TENSORS];
  __shared__ unsigned s_start4[HPVG_PEER_MAX_TENSORS + 1], s_numel[HPVG_PEER_MAX_TENSORS];
  unsigned* epoch = a.sig[a.rank] + PEER_EPOCH_OFF + blockIdx.x;
  if (threadIdx.x == 0) s_epoch = *epoch + 1u;
  if ((int)threadIdx.x < t.n) {
    s_ptr[threadIdx.x] = t.ptr[threadIdx.x];
    s_numel[threadIdx.x] = t.numel[threadIdx.x];
  }
  if ((int)threadIdx.x <= t.n) s_start4[threadIdx.x] = t.start4[threadIdx.x];
  __syncthreads();
  const unsigned e = s_epoch;
  peer_copy_slots<true>(a, s_ptr, s_start4, s_numel, t.n);      // gradients -> this rank's bucket
  __threadfence_system();
  __syncthreads();
  peer_barrier(a, 0, e);
  peer_reduce_slice(a);
  __threadfence_system();
  __syncthreads();
  peer_barrier(a, 1, e);
  peer_copy_slots<false>(a, s_ptr, s_start4, s_numel, t.n);     // averaged bucket -> gradients
  if (threadIdx.x == 0) *epoch = e;
}

}  // namespace hpvg

using namespace hpvg;

#define PEER_CUDA(call, what)                                                \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      (void)cudaGetLastError();                                              \
      set_error("%s: %s", what, cudaGetErrorString(e_));                     \
      return -2;                                                             \
    }                                                                        \
  } while (0)

extern "C" {

int hpvg_peer_alloc(size_t bytes, void** ptr) {
  HPVG_CHECK_ARG(ptr && bytes > 0, "peer_alloc: null argument or zero size");
  static_assert(sizeof(cudaIpcMemHandle_t) == HPVG_PEER_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  PEER_CUDA(cudaMalloc(&p, bytes), "peer_alloc: cudaMalloc");
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    (void)cudaGetLastError();
    set_error("peer_alloc: zero fill: %s", cudaGetErrorString(e));
    return -2;
  }
  *ptr = p;
  return 0;
}

int hpvg_peer_free(void* ptr) {
  if (ptr) PEER_CUDA(cudaFree(ptr), "peer_free");
  return 0;
}

int hpvg_peer_export(const void* ptr, void* handle) {
  HPVG_CHECK_ARG(ptr && handle, "peer_export: null argument");
  cudaIpcMemHandle_t h;
  PEER_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)), "peer_export: cudaIpcGetMemHandle");
  memcpy(handle, &h, sizeof(h));
  return 0;
}

int hpvg_peer_import(const void* handle, void** ptr) {
  HPVG_CHECK_ARG(ptr && handle, "peer_import: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  PEER_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "peer_import: cudaIpcOpenMemHandle");
  *ptr = p;
  return 0;
}

int hpvg_peer_close(void* ptr) {
  if (ptr) PEER_CUDA(cudaIpcCloseMemHandle(ptr), "peer_close");
  return 0;
}

int hpvg_peer_can_access(int device, int peer_device) {
  if (device == peer_device) return 1;
  int ok = 0;
  if (cudaDeviceCanAccessPeer(&ok, device, peer_device) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return ok;
}

static int peer_args(PeerArgs& a, void* const* bufs, void* const* signals, int rank, int world, long long numel, const char* who) {
  HPVG_CHECK_ARG(bufs && signals, "%s: null argument", who);
  HPVG_CHECK_ARG(world >= 1 && world <= HPVG_PEER_MAX_RANKS && rank >= 0 && rank < world, "%s: rank %d of %d (at most %d ranks)", who, rank, world,
                 HPVG_PEER_MAX_RANKS);
  HPVG_CHECK_ARG(numel > 0 && numel % (4LL * world) == 0, "%s: %lld floats are not a multiple of 4 x %d", who, numel, world);
  for (int q = 0; q < HPVG_PEER_MAX_RANKS; ++q) {
    a.buf[q] = nullptr;
    a.sig[q] = nullptr;
  }
  for (int q = 0; q < world; ++q) {
    HPVG_CHECK_ARG(bufs[q] && signals[q], "%s: rank %d's bucket or signal pad is not mapped", who, q);
    a.buf[q] = reinterpret_cast<float4*>(bufs[q]);
    a.sig[q] = reinterpret_cast<unsigned*>(signals[q]);
  }
  a.rank = rank;
  a.world = world;
  a.n4 = numel / (4LL * world);
  a.scale = 1.0f / (float)world;
  return 0;
}

// the CTAs of all ranks pair up by number: the grid is a function of the arguments every rank shares
static int peer_grid(const PeerArgs& a) { return (int)min((long long)PEER_MAX_BLOCKS, max(1LL, cdiv(a.n4, (long long)PEER_THREADS * PEER_UNROLL))); }

int hpvg_peer_allreduce_avg(void* const* bufs, void* const* signals, int rank, int world, long long numel, void* stream) {
  PeerArgs a;
  if (int rc = peer_args(a, bufs, signals, rank, world, numel, "peer_allreduce_avg")) return rc;
  // no programmatic launch here: the kernel's first act publishes what the preceding kernels of the stream wrote
  peer_allreduce_kernel<<<peer_grid(a), PEER_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  HPVG_CHECK_LAUNCH("peer_allreduce_kernel");
  return 0;
}

long long hpvg_peer_bucket_numel(int n, const long long* numel, int world) {
  if (n < 0 || !numel || world < 1) return -1;
  long long slots = 0;
  for (int i = 0; i < n; ++i) {
    if (numel[i] < 0) return -1;
    slots += (numel[i] + 3) / 4;
  }
  const long long per_rank = (max(slots, 1LL) + world - 1) / world;
  return per_rank * 4 * world;
}

int hpvg_peer_allreduce_avg_tensors(void* const* bufs, void* const* signals, int rank, int world, long long bucket_numel, int n,
                                    float* const* grads, const long long* numel, void* stream) {
  PeerArgs a;
  if (int rc = peer_args(a, bufs, signals, rank, world, bucket_numel, "peer_allreduce_avg_tensors")) return rc;
  HPVG_CHECK_ARG(grads && numel && n >= 1 && n <= HPVG_PEER_MAX_TENSORS, "peer_allreduce_avg_tensors: %d tensors (1 .. %d per call)", n, HPVG_PEER_MAX_TENSORS);
  HPVG_CHECK_ARG(bucket_numel >= hpvg_peer_bucket_numel(n, numel, world), "peer_allreduce_avg_tensors: the bucket (%lld floats) is smaller than the gradients need",
                 bucket_numel);
  PeerTensors t;
  unsigned long long slots = 0;
  for (int i = 0; i < n; ++i) {
    HPVG_CHECK_ARG(grads[i] && numel[i] > 0 && numel[i] < (1LL << 31), "peer_allreduce_avg_tensors: gradient %d is null, empty or too large", i);
    HPVG_CHECK_ARG((reinterpret_cast<uintptr_t>(grads[i]) & 15u) == 0, "peer_allreduce_avg_tensors: gradient %d is not 16-byte aligned", i);
    t.ptr[i] = grads[i];
    t.numel[i] = (unsigned)numel[i];
    t.start4[i] = (unsigned)slots;
    slots += (unsigned long long)((numel[i] + 3) / 4);
  }
  HPVG_CHECK_ARG(slots < (1ull << 31), "peer_allreduce_avg_tensors: bucket too large");
  t.start4[n] = (unsigned)slots;
  t.n = n;
  peer_allreduce_tensors_kernel<<<peer_grid(a), PEER_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, t);
  HPVG_CHECK_LAUNCH("peer_allreduce_tensors_kernel");
  return 0;
}

}  // extern "C"

// libhpvg: process-wide state of the C-ABI (error string, backend switch, launch counter, tensor-map encoder).
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace hpvg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_backend{HPVG_BACKEND_AUTO};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int conv_backend() { return g_backend.load(std::memory_order_relaxed); }

static int pdl_default() {
  // measured on B200 (bench.py, config 2): with the attribute on, dependent grids are made resident early and sit in
  // griddepcontrol.wait holding shared memory and registers that the kernels of the other streams of the recorded
  // iteration could use: 6.38 vs 6.20 ms per iteration, 37.7k vs 45.5k generated frames/s.  Off unless HPVG_PDL=1.
  const char* e = getenv("HPVG_PDL");
  return (e && atoi(e) != 0) ? 1 : 0;
}
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    v = pdl_default();
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
static std::atomic<int> g_col_mode{-2};
int conv_col_mode() {
  int v = g_col_mode.load(std::memory_order_relaxed);
  if (v == -2) {
    const char* e = getenv("HPVG_TC_COL");
    v = e ? atoi(e) : -1;     // default: chosen per layer by how well the brick kernel's units fill the SMs (conv_tc.cu)
    if (v < -1 || v > 1) v = -1;
    g_col_mode.store(v, std::memory_order_relaxed);
  }
  return v;
}
int set_conv_col_mode(int mode) {
  const int prev = conv_col_mode();
  g_col_mode.store(mode < 0 ? -1 : (mode > 0 ? 1 : 0), std::memory_order_relaxed);
  return prev;
}
static std::atomic<int> g_carveout{-1};
bool carveout_enabled() {
  int v = g_carveout.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("HPVG_CARVEOUT");
    v = (e && atoi(e) != 0) ? 1 : 0;
    g_carveout.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
void prefer_max_smem(const void* kernel) {
  static std::mutex mu;
  static std::map<const void*, bool> seen;
  std::lock_guard<std::mutex> lock(mu);
  if (seen.count(kernel)) return;
  seen[kernel] = true;
  // a hint: failure (or a driver that ignores it) changes nothing about correctness
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
    cudaGetLastError();
}
static std::atomic<int> g_wgrad_mode{-1};
int wgrad_mode() {
  int v = g_wgrad_mode.load(std::memory_order_relaxed);
  if (v < 0) {
    // 1 = kd-stacked kernel (default: measured on B200 29.7 vs 37.9 us per 64 -> 64 call at 16 x 64 x 64, kernel + reduction),
    // 0 = one kd per CTA, 2 = one kd per CTA with the staged drain
    const char* e = getenv("HPVG_WGRAD_STACK");
    v = e ? atoi(e) : 1;
    if (v < 0 || v > 2) v = 1;
    g_wgrad_mode.store(v, std::memory_order_relaxed);
  }
  return v;
}
int set_wgrad_mode(int mode) {
  const int prev = wgrad_mode();
  g_wgrad_mode.store((mode < 0 || mode > 2) ? 0 : mode, std::memory_order_relaxed);
  return prev;
}
int set_pdl(int on) {
  const int prev = pdl_enabled() ? 1 : 0;
  g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

// ---- optional per-launch device timing (bench.py's roofline leg): CUDA events around selected kernels, on the stream
// they are launched on.  Off by default; never used under stream capture.
struct ProfRecord {
  int kind;
  double work;
  cudaEvent_t start, stop;
};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRecord> g_prof;

bool profiling() { return g_prof_on.load(std::memory_order_relaxed) != 0; }
void* prof_begin(int kind, double work, cudaStream_t st) {
  if (!profiling()) return nullptr;
  ProfRecord r;
  r.kind = kind;
  r.work = work;
  if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return nullptr;
  cudaEventRecord(r.start, st);
  ProfRecord* out = new ProfRecord(r);
  return out;
}
void prof_end(void* h, cudaStream_t st) {
  if (!h) return;
  ProfRecord* r = reinterpret_cast<ProfRecord*>(h);
  cudaEventRecord(r->stop, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(*r);
  delete r;
}

static std::atomic<long long*> g_dbg{nullptr};
long long* debug_clock_buffer() { return g_dbg.load(std::memory_order_relaxed); }

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::atomic<int> state{0};
  if (state.load(std::memory_order_acquire) == 2) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  state.store(2, std::memory_order_release);
  return fn;
}

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return -3;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  uint64_t stride = 2;  // bytes of one bf16
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstrides[i] = stride;
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides, gbox, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu %llu)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
              (unsigned long long)(rank > 3 ? dims[3] : 0), (unsigned long long)(rank > 4 ? dims[4] : 0));
    return -3;
  }
  return 0;
}

}  // namespace hpvg

extern "C" {

const char* hpvg_last_error(void) { return hpvg::g_err; }
int hpvg_version(void) { return 100; }
int hpvg_set_conv_backend(int backend) {
  if (backend < 0 || backend > 2) {
    hpvg::set_error("hpvg_set_conv_backend: unknown backend %d", backend);
    return -1;
  }
  hpvg::g_backend.store(backend);
  return 0;
}
int hpvg_get_conv_backend(void) { return hpvg::g_backend.load(); }
long long hpvg_launch_count(void) { return hpvg::g_launches.load(); }

int hpvg_debug_set_clock_buffer(long long* device_buffer) {
  hpvg::g_dbg.store(device_buffer);
  return 0;
}

int hpvg_set_pdl(int on) { return hpvg::set_pdl(on); }
int hpvg_set_conv_col_mode(int mode) { return hpvg::set_conv_col_mode(mode); }
int hpvg_set_wgrad_mode(int mode) { return hpvg::set_wgrad_mode(mode); }

int hpvg_profile_enable(int on) {
  hpvg::g_prof_on.store(on ? 1 : 0);
  return 0;
}

int hpvg_profile_dump(double* rows, int max_rows) {
  using namespace hpvg;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::pair<int, double>, std::pair<long long, double>> agg;
  for (ProfRecord& r : g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
      auto& a = agg[std::make_pair(r.kind, r.work)];
      a.first += 1;
      a.second += ms;
    }
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  g_prof.clear();
  int n = 0;
  for (auto& kv : agg) {
    if (n >= max_rows) break;
    rows[4 * n + 0] = kv.first.first;
    rows[4 * n + 1] = kv.first.second;
    rows[4 * n + 2] = (double)kv.second.first;
    rows[4 * n + 3] = kv.second.second;
    ++n;
  }
  return n;
}

}  // extern "C"

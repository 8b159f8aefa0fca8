// Probe of tcgen05 shared-memory descriptor addressing on sm_100a.
//
// Design question for the conv kernels: can ONE halo slab in shared memory (rows = voxels, 128 B = 64 bf16
// channels per row, SWIZZLE_128B as written by TMA) be read by tcgen05.mma for EVERY filter tap just by moving
// the descriptor start address by a whole number of rows (not a multiple of the 1024 B swizzle atom)?
// That holds iff the hardware applies the 128B swizzle to the final absolute smem address.  This program decodes
// which (row, k) element the tensor core actually consumed for every (m, k) of A under several descriptor
// settings, and also dumps what TMA writes for a box that starts at a row offset / with out-of-bound coords.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_desc_probe umma_desc_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int SLAB_ROWS = 320;
constexpr int SLAB_BYTES = SLAB_ROWS * 128;       // 40960, multiple of 1024
constexpr int B_BYTES = 64 * 128;                 // 8192
constexpr int NTHREADS = 128;

struct TestCfg {
  int kind;         // 0: A K-major SW128 ; 1: A K-major no-swizzle (8 planes of rows x 16B) ; 2: A MN-major SW128 (B MN-major too)
  int row_off;      // start row inside the slab
  int sbo_bytes;    // SBO of A
  int lbo_bytes;    // LBO of A
  int base_off;     // descriptor base-offset field
  int kadv_bytes;   // start-address advance per K=16 step
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(cnt));
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_timeout(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) { if (clock64() - t0 > 400000000LL) return false; }
  return true;
}

__global__ void __launch_bounds__(NTHREADS, 1)
probe_kernel(const __nv_bfloat16* __restrict__ X,      // [SLAB_ROWS][64]
             const __nv_bfloat16* __restrict__ Bm,     // [64][64] identity
             const TestCfg* __restrict__ cfgs, int ncfg,
             float* __restrict__ Dout,                 // [ncfg][128][64]
             int* __restrict__ status)
{
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // manual 1024 alignment
  uint32_t base = smem_u32(smem_raw);
  uint32_t pad = (1024 - (base & 1023)) & 1023;
  uint8_t* smem = smem_raw + pad;
  uint8_t* slab = smem;
  uint8_t* bt = smem + SLAB_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SLAB_BYTES + B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SLAB_BYTES + B_BYTES + 16);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  uint32_t parity = 0;

  for (int t = 0; t < ncfg; ++t) {
    TestCfg c = cfgs[t];
    // ---- fill the slab ----
    for (int i = tid; i < SLAB_ROWS * 8; i += NTHREADS) {
      int row = i >> 3, ch = i & 7;
      uint4 v = *reinterpret_cast<const uint4*>(X + row * 64 + ch * 8);
      uint32_t off;
      if (c.kind == 1) off = ch * (SLAB_ROWS * 16) + row * 16;               // plane layout, no swizzle
      else { uint32_t L = row * 128 + ch * 16; off = L ^ (((L >> 7) & 7) << 4); }  // 128B swizzle on absolute bits
      *reinterpret_cast<uint4*>(slab + off) = v;
    }
    for (int i = tid; i < 64 * 8; i += NTHREADS) {
      int row = i >> 3, ch = i & 7;
      uint4 v = *reinterpret_cast<const uint4*>(Bm + row * 64 + ch * 8);
      uint32_t L = row * 128 + ch * 16; uint32_t off = L ^ (((L >> 7) & 7) << 4);
      *reinterpret_cast<uint4*>(bt + off) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      if (c.kind == 2) idesc |= (1u << 15) | (1u << 16);
      uint32_t a_layout = (c.kind == 1) ? 0u : 2u;
      for (int k = 0; k < 4; ++k) {
        uint32_t a_addr = smem_u32(slab) + (c.kind == 1 ? c.row_off * 16 : c.row_off * 128) + k * c.kadv_bytes;
        uint64_t ad = make_desc(a_addr, c.lbo_bytes, c.sbo_bytes, c.base_off, a_layout);
        uint64_t bd;
        if (c.kind == 2) bd = make_desc(smem_u32(bt) + k * 2048, 16, 1024, 0, 2);   // B MN-major: 16 k-rows per step
        else             bd = make_desc(smem_u32(bt) + k * 32, 16, 1024, 0, 2);     // B K-major: 32 B per K=16 step
        uint32_t acc = (k > 0);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
    }
    bool ok = mbar_wait_timeout(smem_u32(bar), parity);
    parity ^= 1;
    if (!ok) { if (tid == 0) status[0] = 100 + t; break; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- read the accumulator: warp w owns lanes 32w..32w+31 ; two loads of 32 columns ----
    uint32_t r[64];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#define LD32(off, R) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(R[0]),"=r"(R[1]),"=r"(R[2]),"=r"(R[3]),"=r"(R[4]),"=r"(R[5]),"=r"(R[6]),"=r"(R[7]),"=r"(R[8]),"=r"(R[9]),"=r"(R[10]),"=r"(R[11]),"=r"(R[12]),"=r"(R[13]),"=r"(R[14]),"=r"(R[15]), \
        "=r"(R[16]),"=r"(R[17]),"=r"(R[18]),"=r"(R[19]),"=r"(R[20]),"=r"(R[21]),"=r"(R[22]),"=r"(R[23]),"=r"(R[24]),"=r"(R[25]),"=r"(R[26]),"=r"(R[27]),"=r"(R[28]),"=r"(R[29]),"=r"(R[30]),"=r"(R[31]) \
      : "r"(taddr + off));
    LD32(0, r);
    uint32_t* r2 = r + 32;
    LD32(32, r2);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float* dst = Dout + ((size_t)t * 128 + tid) * 64;
    for (int j = 0; j < 64; ++j) dst[j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(64));
}

// ---------------- TMA layout probe ----------------
__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap map2d, const __grid_constant__ CUtensorMap map4d,
                 int smem_row_off, __nv_bfloat16* __restrict__ dump2d, __nv_bfloat16* __restrict__ dump4d, int* status)
{
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint32_t base = smem_u32(smem_raw);
  uint32_t pad = (1024 - (base & 1023)) & 1023;
  uint8_t* smem = smem_raw + pad;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * SLAB_BYTES);
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * SLAB_BYTES / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0xFFFFFFFFu;  // bf16 NaN pattern
  if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    uint32_t bytes2d = 20 * 128, bytes4d = 12 * 3 * 2 * 128;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes2d + bytes4d) : "memory");
    // 2-D box {64 ch, 20 rows} from row 7, destination = slab + smem_row_off rows (128 B aligned only)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem) + smem_row_off * 128), "l"(&map2d), "r"(smem_u32(bar)), "r"(0), "r"(7) : "memory");
    // 4-D box {64, 12, 3, 2} at (0,-1,-1,0): halo with zero fill, destination second slab (1024 aligned)
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(smem) + SLAB_BYTES), "l"(&map4d), "r"(smem_u32(bar)), "r"(0), "r"(-1), "r"(-1), "r"(0) : "memory");
  }
  bool ok = mbar_wait_timeout(smem_u32(bar), 0);
  if (!ok) { if (tid == 0) status[1] = 77; return; }
  for (int i = tid; i < 64 * 64; i += 128) dump2d[i] = reinterpret_cast<__nv_bfloat16*>(smem)[i];
  for (int i = tid; i < 96 * 64; i += 128) dump4d[i] = reinterpret_cast<__nv_bfloat16*>(smem + SLAB_BYTES)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  std::vector<TestCfg> cfgs;
  // kind, row_off, sbo, lbo, base_off, kadv
  cfgs.push_back({0, 0, 1024, 16, 0, 32});     // 0 canonical
  for (int r : {1, 2, 3, 5, 8, 13}) cfgs.push_back({0, r, 1024, 16, 0, 32});     // 1..6 row offsets, base_off 0
  for (int r : {1, 3, 5}) cfgs.push_back({0, r, 1024, 16, r & 7, 32});             // 7..9 with base_off = row&7
  cfgs.push_back({0, 0, 1280, 16, 0, 32});     // 10: SBO = 10 rows
  cfgs.push_back({0, 11, 1280, 16, 0, 32});    // 11: SBO = 10 rows + row offset 11
  cfgs.push_back({0, 3, 2176, 16, 0, 32});     // 12: SBO = 17 rows
  cfgs.push_back({1, 0, 128, SLAB_ROWS * 16, 0, 2 * SLAB_ROWS * 16});   // 13: no-swizzle canonical
  cfgs.push_back({1, 5, 128, SLAB_ROWS * 16, 0, 2 * SLAB_ROWS * 16});   // 14: no-swizzle + row offset 5
  cfgs.push_back({1, 5, 160, SLAB_ROWS * 16, 0, 2 * SLAB_ROWS * 16});   // 15: no-swizzle + SBO 10 rows
  cfgs.push_back({2, 0, 1024, 8192, 0, 2048});   // 16: MN-major, second 64-wide atom 64 rows further (canonical-ish)
  cfgs.push_back({2, 0, 1024, 128, 0, 2048});    // 17: MN-major, second atom shifted by ONE row (tap-pair trick)
  cfgs.push_back({2, 3, 1024, 256, 0, 2048});    // 18: MN-major, row off 3, second atom +2 rows
  cfgs.push_back({2, 3, 1280, 128, 0, 2560});    // 19: MN-major, K-groups 10 rows apart (brick rows inside a halo slab)
  cfgs.push_back({2, 3, 1280, 0, 0, 2560});      // 20: same, LBO = 0 (second atom duplicates the first)
  cfgs.push_back({2, 3, 1280, 1024, 0, 2560});   // 21: same, second atom 8 rows further
  int ncfg = (int)cfgs.size();

  std::vector<__nv_bfloat16> hB(64 * 64);
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dX, *dB; TestCfg* dC; float* dD; int* dS;
  CK(cudaMalloc(&dX, SLAB_ROWS * 64 * 2)); CK(cudaMalloc(&dB, 64 * 64 * 2)); CK(cudaMalloc(&dC, ncfg * sizeof(TestCfg)));
  CK(cudaMalloc(&dD, (size_t)ncfg * 128 * 64 * 4)); CK(cudaMalloc(&dS, 16)); CK(cudaMemset(dS, 0, 16));
  CK(cudaMemcpy(dB, hB.data(), 64 * 64 * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, cfgs.data(), ncfg * sizeof(TestCfg), cudaMemcpyHostToDevice));
  size_t smem_bytes = 2 * SLAB_BYTES + B_BYTES + 2048;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));

  std::vector<float> rowdec((size_t)ncfg * 128 * 64), kdec((size_t)ncfg * 128 * 64);
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<__nv_bfloat16> hX(SLAB_ROWS * 64);
    for (int r = 0; r < SLAB_ROWS; ++r) for (int k = 0; k < 64; ++k) hX[r * 64 + k] = __float2bfloat16(pass == 0 ? (float)(r % 256) : (float)k);
    CK(cudaMemcpy(dX, hX.data(), SLAB_ROWS * 64 * 2, cudaMemcpyHostToDevice));
    for (int t = 0; t < ncfg; ++t) {
      probe_kernel<<<1, NTHREADS, smem_bytes>>>(dX, dB, dC + t, 1, dD + (size_t)t * 128 * 64, dS);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("probe test %d pass %d failed: %s\n", t, pass, cudaGetErrorString(e)); return 1; }
    }
    CK(cudaMemcpy(pass == 0 ? rowdec.data() : kdec.data(), dD, (size_t)ncfg * 128 * 64 * 4, cudaMemcpyDeviceToHost));
  }
  int hS[4]; CK(cudaMemcpy(hS, dS, 16, cudaMemcpyDeviceToHost));
  printf("status %d %d\n", hS[0], hS[1]);
  for (int t = 0; t < ncfg; ++t) {
    TestCfg c = cfgs[t];
    int bad = 0; int first_m = -1, first_n = -1;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      int er, ek;
      if (c.kind == 0) { er = c.row_off + (m / 8) * (c.sbo_bytes / 128) + m % 8; ek = n; }
      else if (c.kind == 1) { er = c.row_off + (m / 8) * (c.sbo_bytes / 16) + m % 8; ek = n; }
      else { er = c.row_off + (m >= 64 ? c.lbo_bytes / 128 : 0) + (n / 16) * (c.kadv_bytes / 128) + ((n % 16) / 8) * (c.sbo_bytes / 128) + n % 8; ek = m % 64; }
      int gr = (int)rowdec[((size_t)t * 128 + m) * 64 + n], gk = (int)kdec[((size_t)t * 128 + m) * 64 + n];
      if (gr != er % 256 || gk != ek) { if (!bad) { first_m = m; first_n = n; } ++bad; }
    }
    printf("test %2d kind %d row_off %2d sbo %5d lbo %5d base_off %d : %s (%d mismatches)\n", t, c.kind, c.row_off, c.sbo_bytes,
           c.lbo_bytes, c.base_off, bad ? "MISMATCH" : "OK", bad);
    if (bad) {
      printf("   first mismatch at m=%d n=%d ; decoded (row,k) for m=0..17, n in {0,8,17}:\n   ", first_m, first_n);
      for (int m = 0; m < 18; ++m) {
        for (int n : {0, 8, 17}) printf("(%d,%d)", (int)rowdec[((size_t)t * 128 + m) * 64 + n], (int)kdec[((size_t)t * 128 + m) * 64 + n]);
        printf(" ");
      }
      printf("\n   m=64..69: ");
      for (int m = 64; m < 70; ++m) { for (int n : {0, 8, 17}) printf("(%d,%d)", (int)rowdec[((size_t)t * 128 + m) * 64 + n], (int)kdec[((size_t)t * 128 + m) * 64 + n]); printf(" "); }
      printf("\n");
    }
  }

  // ---------------- TMA probe ----------------
  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 0; }
  // X2: [SLAB_ROWS][64] value = row (pass 0) / k (pass 1).  X4: [D=2][H=6][W=10][C=64] value = flat voxel index+1 / k
  __nv_bfloat16 *dX4, *dDump2, *dDump4;
  CK(cudaMalloc(&dX4, 2 * 6 * 10 * 64 * 2)); CK(cudaMalloc(&dDump2, 64 * 64 * 2)); CK(cudaMalloc(&dDump4, 96 * 64 * 2));
  for (int smem_row_off : {0, 3}) {
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<__nv_bfloat16> hX(SLAB_ROWS * 64), hX4(120 * 64);
      for (int r = 0; r < SLAB_ROWS; ++r) for (int k = 0; k < 64; ++k) hX[r * 64 + k] = __float2bfloat16(pass == 0 ? (float)(r % 256) : (float)k);
      for (int v = 0; v < 120; ++v) for (int k = 0; k < 64; ++k) hX4[v * 64 + k] = __float2bfloat16(pass == 0 ? (float)(v + 1) : (float)k);
      CK(cudaMemcpy(dX, hX.data(), SLAB_ROWS * 64 * 2, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dX4, hX4.data(), 120 * 64 * 2, cudaMemcpyHostToDevice));
      CUtensorMap m2, m4;
      { cuuint64_t dims[2] = {64, (cuuint64_t)SLAB_ROWS}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, 20}; cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode 2d failed %d\n", (int)r); return 0; } }
      { cuuint64_t dims[4] = {64, 10, 6, 2}; cuuint64_t strides[3] = {128, 1280, 7680}; cuuint32_t box[4] = {64, 12, 3, 2}; cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&m4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dX4, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode 4d failed %d\n", (int)r); return 0; } }
      tma_probe_kernel<<<1, 128, smem_bytes>>>(m2, m4, smem_row_off, dDump2, dDump4, dS);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("tma probe (smem_row_off %d) failed: %s\n", smem_row_off, cudaGetErrorString(e)); return 0; }
      std::vector<__nv_bfloat16> h2(64 * 64), h4(96 * 64);
      CK(cudaMemcpy(h2.data(), dDump2, 64 * 64 * 2, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(h4.data(), dDump4, 96 * 64 * 2, cudaMemcpyDeviceToHost));
      // check 2d against absolute-swizzle hypothesis: logical (row j of box, chunk c) sits at physical row (off+j), chunk c ^ ((off+j)&7)
      int bad = 0;
      for (int j = 0; j < 20; ++j) for (int cch = 0; cch < 8; ++cch) {
        int prow = smem_row_off + j, pch = cch ^ (prow & 7);
        float v = __bfloat162float(h2[prow * 64 + pch * 8]);
        float ev = pass == 0 ? (float)(7 + j) : (float)(cch * 8);
        if (v != ev) ++bad;
      }
      printf("TMA 2d smem_row_off %d pass %d : absolute-swizzle hypothesis %s (%d bad)\n", smem_row_off, pass, bad ? "FAILS" : "holds", bad);
      if (bad) { printf("   phys rows %d..%d chunk-first values: ", smem_row_off, smem_row_off + 3);
        for (int pr = smem_row_off; pr < smem_row_off + 4; ++pr) { for (int pc = 0; pc < 8; ++pc) printf("%g ", __bfloat162float(h2[pr * 64 + pc * 8])); printf("| "); } printf("\n"); }
      // 4d halo: box rows order (w 12, h 3, d 2); element (wj,hj,dj) <- input (w=wj-1,h=hj-1,d=dj) or zero
      bad = 0;
      for (int dj = 0; dj < 2; ++dj) for (int hj = 0; hj < 3; ++hj) for (int wj = 0; wj < 12; ++wj) for (int cch = 0; cch < 8; ++cch) {
        int prow = (dj * 3 + hj) * 12 + wj, pch = cch ^ (prow & 7);
        float v = __bfloat162float(h4[prow * 64 + pch * 8]);
        int w = wj - 1, h = hj - 1, d = dj; bool in = (w >= 0 && w < 10 && h >= 0 && h < 6);
        float ev = in ? (pass == 0 ? (float)((d * 6 + h) * 10 + w + 1) : (float)(cch * 8)) : 0.f;
        if (v != ev) ++bad;
      }
      printf("TMA 4d halo pass %d : dense rows + zero fill hypothesis %s (%d bad)\n", pass, bad ? "FAILS" : "holds", bad);
    }
  }
  printf("done\n");
  return 0;
}

// Experiment: can a tcgen05 SWIZZLE_128B K-major smem descriptor start at a row offset that is a multiple of
// 128 B but not of 1024 B (shifted windows of one TMA-loaded halo tile)?  Tries base_offset = 0 and
// base_offset = (addr >> 7) & 7.  Also times an MMA-only loop per N to calibrate the tensor-pipe floor.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0; long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t base_off) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(base_off & 7) << 49) | ((uint64_t)2 << 61);
}
constexpr uint32_t idesc_for(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

constexpr int kRows = 256;   // halo tile rows
constexpr int kOffsets = 8;
__constant__ int c_offsets[kOffsets] = {0, 1, 3, 8, 9, 58, 59, 117};

// out[variant][offset][128][64]
__global__ void __launch_bounds__(128, 1) desc_test(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* g = smem_raw + (base - raw);
  const uint32_t a_s = base, b_s = base + kRows * 128, bar = b_s + 64 * 128, bar2 = bar + 8;
  volatile uint32_t* slot = (volatile uint32_t*)(g + kRows * 128 + 64 * 128 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, kRows * 128 + 64 * 128);
    tma_load_2d(a_s, &map_a, bar, 0, 0);
    tma_load_2d(b_s, &map_b, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  uint32_t ph = 0;
  for (int variant = 0; variant < 2; ++variant)
    for (int oi = 0; oi < kOffsets; ++oi) {
      const int t = c_offsets[oi];
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t sa = a_s + t * 128;
        const uint64_t ad = desc_sw128(sa, variant ? (sa >> 7) : 0), bd = desc_sw128(b_s, 0);
        for (int k = 0; k < 4; ++k) tc_mma(tmem, ad + 2 * k, bd + 2 * k, idesc_for(64), k ? 1u : 0u);
        tc_commit(bar2);
      }
      mbar_wait(bar2, ph); ph ^= 1; tc_fence_after();
      float* o = out + ((size_t)(variant * kOffsets + oi) * 128 + warp * 32 + lane) * 64;
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v); tmem_ld_wait();
        for (int j = 0; j < 32; ++j) o[c + j] = __uint_as_float(v[j]);
      }
      tc_fence_before(); __syncthreads();
    }
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory"); }
}

// MMA-only throughput: one CTA per SM, `iters` k-blocks of 4 MMAs (M=128, N, K=16 each) on fixed smem operands
template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate(long long* cycles, int iters, int shift_rows) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* g = smem_raw + (base - raw);
  const uint32_t a_s = base, b_s = base + 256 * 128, bar = b_s + 256 * 128;
  volatile uint32_t* slot = (volatile uint32_t*)(g + 512 * 128 + 16);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 512 * 128 / 4; i += 128) ((uint32_t*)g)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sa = a_s + ((it % 9) * shift_rows) * 128;
      const uint64_t ad = desc_sw128(sa, 0), bd = desc_sw128(b_s, 0);
      const uint32_t d = tmem + (uint32_t)((it & 1) * N);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma(d, ad + 2 * k, bd + 2 * k, idesc_for(N), 1u);
    }
    tc_commit(bar);
    mbar_wait(bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

// TMEM read throughput: 4 warps read `cols` columns repeatedly
__global__ void __launch_bounds__(128, 1) tmem_rate(long long* cycles, int iters, int cols, float* sink) {
  __shared__ uint32_t slot_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot_s;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    for (int c = 0; c < cols; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v); tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc += __uint_as_float(v[j]);
    }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)f;
  const int R = kRows;
  std::vector<__nv_bfloat16> hx((size_t)R * 64), hw(64 * 64);
  std::vector<float> fx((size_t)R * 64), fw(64 * 64);
  srand(1);
  for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 17 - 8); hx[i] = __float2bfloat16(fx[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { fw[i] = (float)(rand() % 9 - 4); hw[i] = __float2bfloat16(fw[i]); }
  __nv_bfloat16 *dx, *dw; float* dout;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  const size_t out_n = (size_t)2 * kOffsets * 128 * 64;
  CK(cudaMalloc(&dout, out_n * 4)); CK(cudaMemset(dout, 0, out_n * 4));
  CUtensorMap ma, mb;
  { cuuint64_t dims[2] = {64, (cuuint64_t)R}; cuuint64_t st[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)R}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dx, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return 1; } }
  { cuuint64_t dims[2] = {64, 64}; cuuint64_t st[1] = {128}; cuuint32_t box[2] = {64, 64}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return 1; } }
  const int smem = kRows * 128 + 64 * 128 + 1024 + 64;
  CK(cudaFuncSetAttribute(desc_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  desc_test<<<1, 128, smem>>>(ma, mb, dout);
  CK(cudaDeviceSynchronize());
  std::vector<float> ho(out_n);
  CK(cudaMemcpy(ho.data(), dout, out_n * 4, cudaMemcpyDeviceToHost));
  const int offs[kOffsets] = {0, 1, 3, 8, 9, 58, 59, 117};
  for (int v = 0; v < 2; ++v)
    for (int oi = 0; oi < kOffsets; ++oi) {
      int bad = 0; double maxd = 0;
      for (int j = 0; j < 128; ++j)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += fx[(size_t)(offs[oi] + j) * 64 + k] * fw[n * 64 + k];
          double d = fabs(ref - ho[((size_t)(v * kOffsets + oi) * 128 + j) * 64 + n]);
          if (d > 1e-3) ++bad;
          if (d > maxd) maxd = d;
        }
      printf("DESC variant=%s row_offset=%3d : mismatches %5d / 8192  max|d| %.3f\n", v ? "base_offset=(addr>>7)&7" : "base_offset=0", offs[oi], bad, maxd);
    }
  // ---- MMA floor
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  long long* dcy; CK(cudaMalloc(&dcy, sms * 8));
  std::vector<long long> hcy(sms);
  const int smem2 = 512 * 128 + 1024 + 64;
  auto report = [&](const char* name, int n, int iters, int shift) {
    CK(cudaMemcpy(hcy.data(), dcy, sms * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; ++i) avg += hcy[i]; avg /= sms;
    printf("MMA %s N=%3d shift=%d: %.1f cycles per MMA (M128 x N x K16), %.0f MAC/cyc/SM (floor N/2 = %d cyc)\n", name, n, shift, avg / (iters * 4.0), 128.0 * n * 16 * iters * 4 / avg, n / 2);
  };
  for (int shift = 0; shift <= 1; ++shift) {
    const int iters = 4000;
    CK(cudaFuncSetAttribute(mma_rate<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    CK(cudaFuncSetAttribute(mma_rate<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    CK(cudaFuncSetAttribute(mma_rate<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    for (int rep = 0; rep < 2; ++rep) {
      mma_rate<64><<<sms, 128, smem2>>>(dcy, iters, shift * 7); CK(cudaDeviceSynchronize()); if (rep) report("all-SM", 64, iters, shift * 7);
      mma_rate<128><<<sms, 128, smem2>>>(dcy, iters, shift * 7); CK(cudaDeviceSynchronize()); if (rep) report("all-SM", 128, iters, shift * 7);
      mma_rate<256><<<sms, 128, smem2>>>(dcy, iters, shift * 7); CK(cudaDeviceSynchronize()); if (rep) report("all-SM", 256, iters, shift * 7);
    }
  }
  float* sink; CK(cudaMalloc(&sink, 4));
  for (int cols : {64, 128, 256}) {
    tmem_rate<<<sms, 128>>>(dcy, 2000, cols, sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hcy.data(), dcy, sms * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; ++i) avg += hcy[i]; avg /= sms;
    printf("TMEM read 128 lanes x %d cols x4B: %.1f cycles per pass -> %.1f B/cyc/SM\n", cols, avg / 2000, 128.0 * cols * 4 * 2000 / avg);
  }
  printf("done\n");
  return 0;
}

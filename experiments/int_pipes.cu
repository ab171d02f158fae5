// Integer-pipe issue rates on B200 (per SM sub-partition), for the K1 (preprocess) instruction mix.
// Each kernel runs kIter iterations of 8 independent chains of one instruction per thread; 4 warps per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/int_pipes experiments/int_pipes.cu && ./experiments/int_pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIter = 4096;

#define BENCH(name, BODY)                                                              \
  __global__ void k_##name(unsigned* out, unsigned seed, long long* cyc) {             \
    unsigned r[8];                                                                     \
    for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;                     \
    unsigned a = seed | 1u, b = seed * 3u + 1u;                                        \
    long long t0 = clock64();                                                          \
    for (int it = 0; it < kIter; ++it) {                                               \
      _Pragma("unroll") for (int i = 0; i < 8; ++i) { BODY; }                          \
    }                                                                                  \
    long long t1 = clock64();                                                          \
    unsigned s = 0;                                                                    \
    for (int i = 0; i < 8; ++i) s += r[i];                                             \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + a + b;                            \
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;                           \
  }

BENCH(imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b)))
BENCH(imad_hi, asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b)))
BENCH(mul_hi, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a)))
BENCH(shr, asm volatile("shr.u32 %0, %0, 1; add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a)))
BENCH(lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b)))
BENCH(prmt, asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b & 0x7777)))
BENCH(dp2a, asm volatile("dp2a.lo.u32.u32 %0, %1, %0, %2;" : "+r"(r[i]) : "r"(a), "r"(b)))
BENCH(dp4a, asm volatile("dp4a.u32.u32 %0, %1, %0, %2;" : "+r"(r[i]) : "r"(a), "r"(b)))
BENCH(shf, asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b)))
BENCH(bfe, asm volatile("bfe.u32 %0, %0, 3, 9;" : "+r"(r[i])))
BENCH(iadd3, asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a)))

__global__ void k_atoms(unsigned* out, unsigned seed, long long* cyc, int mode) {
  __shared__ unsigned h[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // mode 0: 32 distinct banks; 1: all lanes one address; 2: 4 groups of 8 lanes, one address per group; 3: 8 groups of 4
  unsigned idx = mode == 0 ? lane : mode == 1 ? 0 : mode == 2 ? (lane & 3) * 33 : (lane & 7) * 33;
  idx += (threadIdx.x >> 5) * 256;
  long long t0 = clock64();
  for (int it = 0; it < kIter; ++it) atomicAdd(&h[idx], 1u);
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = h[threadIdx.x];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  unsigned* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMallocManaged(&cyc, 8);
  const int threads = 512;  // 16 warps per SM = 4 per SMSP
#define RUN(name)                                                                        \
  k_##name<<<148, threads>>>(out, 12345u, cyc);                                          \
  cudaDeviceSynchronize();                                                               \
  k_##name<<<148, threads>>>(out, 12345u, cyc);                                          \
  cudaDeviceSynchronize();                                                               \
  printf("%-8s %8.3f cycles per warp-instruction per SMSP (4 warps/SMSP, 8 chains)\n", #name, (double)*cyc / (kIter * 8.0 * 4.0));
  RUN(imad) RUN(imad_hi) RUN(mul_hi) RUN(shr) RUN(lop3) RUN(prmt) RUN(dp2a) RUN(dp4a) RUN(shf) RUN(bfe) RUN(iadd3)
  for (int mode = 0; mode < 4; ++mode) {
    k_atoms<<<148, 128>>>(out, 1u, cyc, mode);
    cudaDeviceSynchronize();
    printf("atoms mode %d: %8.2f cycles per warp-level ATOMS per SM (4 warps)\n", mode, (double)*cyc / (kIter * 4.0));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

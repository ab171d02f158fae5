// TMEM read bandwidth on B200: how many bytes per cycle per SM does tcgen05.ld deliver when 4 / 8 warps (one / two per
// TMEM lane quarter) read their quarters back to back?  Decides whether an epilogue may read 3x the accumulator columns
// (3 filter taps side by side in N, summed with lane shifts) without becoming the bottleneck of the 3x3 / small-N layers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I sykepic_b200/csrc -I include -o experiments/tmem_ld experiments/tmem_ld.cu && ./experiments/tmem_ld
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace spk::tc;

template <int X>
__global__ void __launch_bounds__(256, 1) k_ld(long long* cyc, unsigned* sink, int iters, int shfl) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  const uint32_t lane_base = base + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 512; c += 32 * X) {
      uint32_t v[X][32];
#pragma unroll
      for (int x = 0; x < X; ++x) tmem_ld32(lane_base + c + 32 * x, v[x]);
      tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < X; ++x)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          unsigned u = v[x][i];
          if (shfl) u = __shfl_down_sync(0xffffffffu, u, 1);
          acc += u;
        }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(base, 512);
  }
}

int main() {
  long long* d_cyc;
  unsigned* d_sink;
  cudaMalloc(&d_cyc, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 148 * 256 * sizeof(unsigned));
  const int iters = 200;
  for (int shfl = 0; shfl < 2; ++shfl)
    for (int warps : {4, 8}) {
      for (int x : {1, 2}) {
        for (int rep = 0; rep < 2; ++rep) {
          if (x == 1)
            k_ld<1><<<148, warps * 32>>>(d_cyc, d_sink, iters, shfl);
          else
            k_ld<2><<<148, warps * 32>>>(d_cyc, d_sink, iters, shfl);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(e));
            return 1;
          }
        }
        long long h[148];
        cudaMemcpy(h, d_cyc, sizeof h, cudaMemcpyDeviceToHost);
        // bytes read per CTA: warps * iters * 512 columns * 32 lanes * 4 B
        const double bytes = (double)warps * iters * 512.0 * 32.0 * 4.0;
        printf("warps %d  loads in flight %d  shfl %d: %lld cycles -> %.1f B/cycle/SM (%.2f cycles per 32x32b.x32 load per warp)\n", warps, x, shfl,
               h[0], bytes / (double)h[0], (double)h[0] / (iters * 16.0));
      }
    }
  return 0;
}

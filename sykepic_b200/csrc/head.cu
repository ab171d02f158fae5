// K3: global average pool + (folded) Linear head + softmax(logits * ln 1.3) + per-class
// threshold / label, one CTA per ROI, warp-shuffle reductions.
//
// Replaces
//   AdaptiveAvgPool2d(1) + view + head     sykepic/train/network.py:56-69 (Linear chain, no activations)
//   net_pass tail                          sykepic/compute/probability.py:189-195
//   row_prediction                         sykepic/compute/prediction.py:49-71
// The label rule of the reference works on the 5-decimal strings of the CSV; here it is
// evaluated on q = round_half_even(p * 1e5) (exact in double, see host.cpp) against
// thresholds quantised into the same integer domain (spk_threshold_quantize), which gives
// the identical decision.  Ties go to the lowest class index (the reference's idxmax; its
// descending sort is formally unstable).
#include <climits>

#include "spk_internal.h"
#include "tc_common.cuh"  // tc::pdl_* / tc::launch_pdl

namespace spk {
namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

template <typename T>
__device__ __forceinline__ float ld(const T* p);
template <>
__device__ __forceinline__ float ld<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float ld<SplitF>(const SplitF* p) { return split_load(p->v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// dynamic smem: float feat[F] | float logit[K] | int q[K]
template <typename T>
__global__ void __launch_bounds__(THREADS) head_kernel(const T* act, int hw, int F, const float* __restrict__ W,
                                                       const float* __restrict__ bias, int K, float scale,
                                                       const int32_t* __restrict__ thr_q, float* logits, float* probs,
                                                       int32_t* label, uint8_t* classified, unsigned long long* stamp) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* feat = reinterpret_cast<float*>(smem);
  float* logit = feat + F;
  int* q = reinterpret_cast<int*>(logit + K);
  __shared__ float red[WARPS];
  __shared__ float s_max, s_sum;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n = blockIdx.x;
  tc::pdl_trigger();
  tc::stamp_begin(stamp);
  tc::pdl_wait();  // the activations are the previous kernel's output

  // global average pool: sum over the hw positions in position order, then divide (fp32)
  const T* a = act + n * (long long)hw * F;
  const float inv = 1.0f / (float)hw;
  if constexpr (sizeof(T) == 2) {
    // bf16: two adjacent features per thread and load (same per-feature summation order), eight positions in flight
    if ((F & 1) == 0) {
      for (int f = 2 * tid; f < F; f += 2 * THREADS) {
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(a + f);
        const int ld2 = F >> 1;
        float s0 = 0.f, s1 = 0.f;
        int p = 0;
        for (; p + 8 <= hw; p += 8) {
          __nv_bfloat162 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = a2[(long long)(p + u) * ld2];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float2 t = __bfloat1622float2(v[u]);
            s0 += t.x;
            s1 += t.y;
          }
        }
        for (; p < hw; ++p) {
          const float2 t = __bfloat1622float2(a2[(long long)p * ld2]);
          s0 += t.x;
          s1 += t.y;
        }
        feat[f] = s0 * inv;
        feat[f + 1] = s1 * inv;
      }
    } else {
      for (int f = tid; f < F; f += THREADS) {
        float s = 0.f;
        for (int p = 0; p < hw; ++p) s += ld<T>(a + (long long)p * F + f);
        feat[f] = s * inv;
      }
    }
  } else {
    for (int f = tid; f < F; f += THREADS) {
      float s = 0.f;
      for (int p = 0; p < hw; ++p) s += ld<T>(a + (long long)p * F + f);
      feat[f] = s * inv;
    }
  }
  __syncthreads();

  // logits: one warp per class, lanes stride the features
  for (int k = warp; k < K; k += WARPS) {
    const float* w = W + (long long)k * F;
    float s = 0.f;
    for (int f = lane; f < F; f += 32) s = fmaf(feat[f], __ldg(w + f), s);
    s = warp_sum(s);
    if (lane == 0) {
      s += bias[k];
      logit[k] = s;
      if (logits) logits[n * K + k] = s;
    }
  }
  __syncthreads();

  // softmax over K of logit * scale (probability.py:192-194)
  float mx = -INFINITY;
  for (int k = tid; k < K; k += THREADS) mx = fmaxf(mx, scale != 0.f ? logit[k] * scale : logit[k]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    float m2 = red[0];
    for (int i = 1; i < WARPS; ++i) m2 = fmaxf(m2, red[i]);
    s_max = m2;
  }
  __syncthreads();
  float sum = 0.f;
  for (int k = tid; k < K; k += THREADS) {
    const float z = scale != 0.f ? logit[k] * scale : logit[k];
    const float e = expf(z - s_max);
    logit[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncthreads();  // red[] reads above are done
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float s2 = 0.f;
    for (int i = 0; i < WARPS; ++i) s2 += red[i];
    s_sum = s2;
  }
  __syncthreads();
  for (int k = tid; k < K; k += THREADS) {
    const float p = logit[k] / s_sum;
    probs[n * K + k] = p;
    q[k] = (int)rint((double)p * 100000.0);  // the CSV's "%.5f", as an integer
  }
  __syncthreads();

  if (label && tid == 0) {
    // idxmax over the decimals (first maximum), and the best class at or above its own threshold
    int best = 0, best_thr = -1;
    for (int k = 1; k < K; ++k)
      if (q[k] > q[best]) best = k;
    if (thr_q) {
      for (int k = 0; k < K; ++k)
        if (thr_q[k] != INT_MAX && q[k] >= thr_q[k] && (best_thr < 0 || q[k] > q[best_thr])) best_thr = k;
    }
    label[n] = best_thr >= 0 ? best_thr : best;
    if (classified) classified[n] = best_thr >= 0 ? 1 : 0;
  }
  tc::stamp_end(stamp);  // (thread 0 is the last one working)
}

}  // namespace

int launch_head(spk_ctx* ctx, const void* act, int act_dtype, int64_t n, int hw, int feat, const float* w_kf,
                const float* bias, int k, float softmax_scale, const int32_t* thr_q, float* logits, float* probs,
                int32_t* label, uint8_t* classified) {
  if (n == 0) return SPK_OK;
  const size_t smem = (size_t)feat * 4 + (size_t)k * 8;
  if (smem > 200 * 1024) return fail(ctx, SPK_ERR_UNSUPPORTED, "head: %d features x %d classes do not fit shared memory", feat, k);
  if (act_dtype == SPK_DTYPE_F32) {
    if (smem > 48 * 1024)
      SPK_CUDA_OK(ctx, cudaFuncSetAttribute(head_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SPK_CUDA_OK(ctx, tc::launch_pdl(head_kernel<float>, dim3((unsigned)n), dim3(THREADS), smem, ctx->stream, (const float*)act, hw, feat, w_kf,
                                    bias, k, softmax_scale, thr_q, logits, probs, label, classified, ctx->cur_stamp));
  } else if (act_dtype == SPK_DTYPE_BF16) {
    if (smem > 48 * 1024)
      SPK_CUDA_OK(ctx, cudaFuncSetAttribute(head_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SPK_CUDA_OK(ctx, tc::launch_pdl(head_kernel<__nv_bfloat16>, dim3((unsigned)n), dim3(THREADS), smem, ctx->stream,
                                    (const __nv_bfloat16*)act, hw, feat, w_kf, bias, k, softmax_scale, thr_q, logits, probs, label, classified, ctx->cur_stamp));
  } else if (act_dtype == SPK_DTYPE_SPLIT) {
    if (smem > 48 * 1024)
      SPK_CUDA_OK(ctx, cudaFuncSetAttribute(head_kernel<SplitF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SPK_CUDA_OK(ctx, tc::launch_pdl(head_kernel<SplitF>, dim3((unsigned)n), dim3(THREADS), smem, ctx->stream, (const SplitF*)act, hw, feat, w_kf,
                                    bias, k, softmax_scale, thr_q, logits, probs, label, classified, ctx->cur_stamp));
  } else {
    return fail(ctx, SPK_ERR_UNSUPPORTED, "head: dtype %d", act_dtype);
  }
  SPK_LAUNCH_CHECK(ctx);
  return SPK_OK;
}

}  // namespace spk

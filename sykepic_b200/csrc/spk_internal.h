// Internal declarations shared by the translation units of libsykepic_b200.so.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "sykepic_b200.h"

#include "spk_debug.h"

namespace spk {

constexpr int kMaxTarget = 512;  // largest supported target side (T_h, T_w)

struct Net;  // net.cu

}  // namespace spk

namespace spk {
// one timed launch (profiling mode only): CUDA events on the context's stream around the launch
struct ProfRec {
  int category;
  cudaEvent_t start, stop;  // event mode
  int slot;                 // stamp mode: index into spk_ctx::d_stamps (-1 in event mode)
  double flops, bytes;
  std::string what;
};
}  // namespace spk

struct spk_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string error;
  int64_t launches = 0;
  unsigned long long* d_faults = nullptr;  // device-side count of ROIs with invalid geometry
  float* d_default_lut = nullptr;          // 3*256: v/255 (true fp32 division)
  int* d_big_list = nullptr;               // K1: queue of heavy-tail ROIs for the cluster kernel
  unsigned* d_big_count = nullptr;
  long long big_cap = 0;
  cudaStream_t stream2 = nullptr;          // K1: the cluster kernel runs beside the warp kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  spk::Net* net = nullptr;
  int sm_count = 148;
  bool profiling = false;
  int prof_mode = 0;                          // SPK_PROFILE_EVENTS / SPK_PROFILE_STAMPS
  unsigned long long* d_stamps = nullptr;     // [2 * kStampCap]: start stamps, then end stamps
  unsigned long long* cur_stamp = nullptr;    // slot of the launch being issued (stamp mode, inside a ProfScope), else null
  std::vector<spk::ProfRec> prof;
};

namespace spk {

// thread-local message for failures that have no context
std::string& tls_error();

int fail(spk_ctx* ctx, int code, const char* fmt, ...);

// profiling scope: records events around the launches issued while it lives (no-op unless ctx->profiling)
struct ProfScope {
  spk_ctx* ctx;
  int idx = -1;
  ProfScope(spk_ctx* c, int category, double flops, double bytes, const char* fmt, ...);
  ~ProfScope();
};

#define SPK_CUDA_OK(ctx, expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return spk::fail((ctx), SPK_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,      \
                       cudaGetErrorString(_e));                                             \
  } while (0)

#define SPK_LAUNCH_CHECK(ctx)                                                               \
  do {                                                                                      \
    (ctx)->launches++;                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      return spk::fail((ctx), SPK_ERR_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__,  \
                       cudaGetErrorString(_e));                                             \
  } while (0)

// SplitF: an fp32 value kept as two fp16 in one 32-bit word, hi = fp16(x) in the low half (the even K index of the
// tensor-core view), lo = fp16(x - hi) in the high half: 22 significant bits (two bf16 gave 16, and a floor of ~5e-5 on
// the probabilities), exact sums in fp32.  Range: |x| <= 65504 (saturates; the convolution epilogue counts any value
// beyond it in the context's fault counter, which the host turns into an error); below 6e-5 the absolute error is
// at most 3e-8 (fp16 subnormals).
constexpr float kSplitMax = 65504.f;
struct SplitF {
  uint32_t v;
};
__device__ __forceinline__ float split_load(uint32_t v) {
  return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))) + __half2float(__ushort_as_half((unsigned short)(v >> 16)));
}
__device__ __forceinline__ uint32_t split_store(float f) {
  f = fminf(fmaxf(f, -kSplitMax), kSplitMax);
  const __half hi = __float2half_rn(f);
  const __half lo = __float2half_rn(f - __half2float(hi));
  return (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
}

// get_new_dims of sykepic/train/image.py:183-198, shared by host and device code
__host__ __device__ inline void new_dims(int h, int w, int th, int tw, int* nh, int* nw) {
  if (h > w) {
    double r = (double)th / (double)h;
    *nh = th;
    *nw = (int)((double)w * r);
  } else {
    double r = (double)tw / (double)w;
    *nh = (int)((double)h * r);
    *nw = tw;
  }
}

// ---- kernels' host launchers (each enqueues on ctx->stream) ---------------------------------
// preprocess.cu
int init_default_lut(spk_ctx* ctx);
// conv_simt.cu
struct ConvGeom {
  int n, h, w, cin;         // input NHWC
  int ho, wo, cout;         // output NHWC
  int kh, kw, stride, pad;
  int relu;
  // channel strides of the pixel rows (>= cin / cout when the tensor is a channel slice of a wider
  // concat buffer, DenseNet); x / y / res pointers already include the channel offset.
  int ldx, ldy, ldres;
};
int launch_conv_simt(spk_ctx* ctx, const ConvGeom& g, const void* x, int x_dtype, const float* w_kc /*[K][Cout]*/,
                     const float* bias, const void* res, void* y, int y_dtype);
int launch_maxpool(spk_ctx* ctx, int n, int h, int w, int c, int k, int stride, int pad, int ho, int wo, int ldy,
                   const void* x, void* y, int dtype);
int launch_avgpool(spk_ctx* ctx, int n, int h, int w, int c, int ldx, int k, int stride, int ho, int wo, int ldy,
                   const void* x, void* y, int dtype);
int launch_affine_relu(spk_ctx* ctx, long long pixels, int c, int ldx, const float* scale, const float* shift,
                       const void* x, void* y, int dtype, int relu);
// head.cu
int launch_head(spk_ctx* ctx, const void* act, int act_dtype, int64_t n, int hw, int feat, const float* w_kf /*[K][F]*/,
                const float* bias, int k, float softmax_scale, const int32_t* thr_q, float* logits, float* probs,
                int32_t* label, uint8_t* classified);
// conv_tc.cu (tcgen05 / TMEM / TMA implicit GEMM)
struct TcConvPlan;
bool tc_conv_supported(const ConvGeom& g);
// w_ds / bias_ds / ldy_ds: optional fused 1x1 stride-2 downsample branch (same input, same Cout) of a 3x3 stride-2 conv
bool tc_conv_ds_fusable(const ConvGeom& g3x3, const ConvGeom& g1x1);
// split = true: FP32-accurate mode on SplitF activations (bf16 hi | bf16 lo per 32-bit word); g in SplitF words
bool tc_conv_split_supported(const ConvGeom& g);
int tc_conv_plan_create(spk_ctx* ctx, const ConvGeom& g_max, const float* w_oihw_folded /*[Cout][kh][kw][cin] fp32*/,
                        const float* bias, TcConvPlan** out, const float* w_ds = nullptr, const float* bias_ds = nullptr,
                        int ldy_ds = 0, bool split = false);
// y = conv(relu(x * scale[c] + shift[c])) over the first `channels` input channels (device arrays of that length)
int tc_conv_plan_set_prologue(spk_ctx* ctx, TcConvPlan* p, const float* d_scale, const float* d_shift, int channels, int relu);
void tc_conv_plan_destroy(TcConvPlan* p);
int tc_conv_launch(spk_ctx* ctx, TcConvPlan* p, int n, const void* x, const void* res, void* y, void* y_ds = nullptr);
int64_t tc_conv_plan_bytes(const TcConvPlan* p);
// consecutive layers walk their tiles in alternating directions, so that a kernel starts on what the previous one left in L2
void tc_conv_plan_set_reverse(TcConvPlan* p, int reverse);
// conv_pair.cu (Cout >= 128: CTA pairs, tcgen05.mma.cta_group::2, each CTA stages half of the weight tile)
struct PairConvPlan;
bool pair_conv_supported(const ConvGeom& g);
// w_ds / bias_ds / ldy_ds: optional fused 1x1 stride-2 shortcut of a 3x3 stride-2 convolution (as tc_conv_plan_create)
int pair_conv_plan_create(spk_ctx* ctx, const ConvGeom& g_max, const float* w_oihw_folded /*[Cout][kh][kw][cin] fp32*/,
                          const float* bias, PairConvPlan** out, const float* w_ds = nullptr, const float* bias_ds = nullptr,
                          int ldy_ds = 0);
void pair_conv_plan_destroy(PairConvPlan* p);
int pair_conv_launch(spk_ctx* ctx, PairConvPlan* p, int n, const void* x, const void* res, void* y, void* y_ds = nullptr);
int64_t pair_conv_plan_bytes(const PairConvPlan* p);
void pair_conv_plan_set_reverse(PairConvPlan* p, int reverse);
// conv_hp.cu (3x3 / stride 1, Cin and Cout in {64, 128}: halo tiles on CTA pairs, filter bank resident, direct epilogue)
struct HpConvPlan;
bool hp_conv_supported(const ConvGeom& g);
int hp_conv_plan_create(spk_ctx* ctx, const ConvGeom& g_max, const float* w_oihw_folded /*[Cout][kh][kw][cin] fp32*/,
                        const float* bias, HpConvPlan** out);
void hp_conv_plan_destroy(HpConvPlan* p);
int hp_conv_launch(spk_ctx* ctx, HpConvPlan* p, int n, const void* x, const void* res, void* y);
int64_t hp_conv_plan_bytes(const HpConvPlan* p);
void hp_conv_plan_set_reverse(HpConvPlan* p, int reverse);
// stem.cu: conv 7x7/2 (1 gray plane -> 64) + bias + ReLU + maxpool 3x3/2 fused on tcgen05; u8 in -> bf16 NHWC out
bool stem_pool_supported(const ConvGeom& g, int pool_k, int pool_stride, int pool_pad);
int stem_pool_pack_weights(spk_ctx* ctx, const float* w_folded /*[64][7][7]*/, uint4** d_out, bool split);
int stem_pool_t_pack_weights(spk_ctx* ctx, const float* w_folded /*[64][7][7]*/, uint4** d_out);  // stem_t.cu
int launch_stem_pool_t(spk_ctx* ctx, int n, int th, int tw, const uint8_t* x, const uint4* w, const float* bias, __nv_bfloat16* y,
                       int hc, int wc, int hp, int wp, int ldy);
int launch_stem_pool(spk_ctx* ctx, int n, int th, int tw, const uint8_t* x, const uint4* w_sw, const float* bias,
                     __nv_bfloat16* y, int hc, int wc, int hp, int wp, int ldy, bool split = false);  // split: y holds SplitF words, ldy in words

}  // namespace spk

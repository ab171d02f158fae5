// Context + network graph of the C ABI (spk_create ... spk_forward).
//
// Replaces prepare_model's network construction + load_state_dict
// (sykepic/compute/probability.py:118-130, sykepic/train/config.py:63-77,
// sykepic/train/network.py:14-64) and TorchVisionNet.forward (network.py:66-72) + the
// net_pass tail (probability.py:189-195).  The Python host walks the state_dict and
// describes the graph op by op; this file folds eval-mode BatchNorm into the convolution
// weights (in double), packs them for the kernels, plans the NHWC activation workspace
// and replays the op list on the context's stream for every batch.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

#include <exception>

#include "spk_internal.h"

namespace spk {

enum OpKind { kOpConv = 0, kOpMaxPool = 1, kOpAvgPool = 2, kOpBnRelu = 3, kOpStemPool = 4, kOpNop = 5 };

struct Buffer {
  bool known = false;
  bool declared = false;  // spk_net_buffer: a concat buffer that ops fill slice by slice
  int h = 0, w = 0, c = 0;
  int dtype = SPK_DTYPE_F32;
  void* d = nullptr;
  size_t bytes = 0;
};

struct Op {
  int kind = kOpConv;
  int in = 0, out = 0, res = -1;
  int in_off = 0, out_off = 0;
  ConvGeom g{};            // n filled per forward
  int impl = SPK_CONV_SIMT;
  float* d_w = nullptr;    // SIMT: [K][Cout] fp32
  float* d_bias = nullptr; // conv: folded bias [Cout]; bn_relu: shift [C]
  float* d_scale = nullptr;  // bn_relu: scale [C]; conv with a fused pre-activation: its scale
  float* d_pre_shift = nullptr;  // conv with a fused pre-activation (DenseNet BN + ReLU applied to the A tiles): shift
  int pre_c = 0, pre_relu = 0;
  TcConvPlan* tc = nullptr;
  HpConvPlan* hpair = nullptr;
  PairConvPlan* pair = nullptr;
  // fused downsample branch (1x1 / stride 2 of the same input, computed by this 3x3 / stride 2 convolution's kernel)
  int ds_out = -1, ds_off = 0, ds_ld = 0;
  float* d_bias_ds = nullptr;
  std::vector<float> ds_w, ds_b;
  uint4* d_stem_w = nullptr;  // fused stem: swizzled bf16 weight tile
  int pool_out = -1, hp = 0, wp = 0, pool_ld = 0;  // fused stem: the max-pool's output
  std::vector<float> w_host;  // [Cout][kh][kw][cin] folded fp32, kept until net_end for the tcgen05 packer
  std::vector<float> b_host;
  int k = 0, stride = 0, pad = 0, channels = 0, relu = 0;
};

struct Net {
  int th = 0, tw = 0, in_c = 0, precision = SPK_PRECISION_FP32, max_batch = 0;
  int act_dtype = SPK_DTYPE_F32;
  std::vector<Buffer> bufs;
  std::vector<Op> ops;
  bool ended = false;
  bool has_head = false;
  int head_in = -1, feat = 0, classes = 0;
  float* d_head_w = nullptr;  // [K][F]
  float* d_head_b = nullptr;
  float* d_logits = nullptr;  // [max_batch][K]
  int64_t bytes = 0;
  int simt_layers = 0;  // convolutions of a tensor-core precision that had to take the CUDA-core kernel
};

static size_t dtype_size(int dt) { return (dt == SPK_DTYPE_F32 || dt == SPK_DTYPE_SPLIT) ? 4 : dt == SPK_DTYPE_BF16 ? 2 : 1; }

static void net_free(Net* net) {
  if (!net) return;
  for (auto& b : net->bufs)
    if (b.d) cudaFree(b.d);
  for (auto& op : net->ops) {
    if (op.d_w) cudaFree(op.d_w);
    if (op.d_bias) cudaFree(op.d_bias);
    if (op.d_scale) cudaFree(op.d_scale);
    if (op.d_pre_shift) cudaFree(op.d_pre_shift);
    if (op.d_bias_ds) cudaFree(op.d_bias_ds);
    if (op.d_stem_w) cudaFree(op.d_stem_w);
    if (op.tc) tc_conv_plan_destroy(op.tc);
    if (op.hpair) hp_conv_plan_destroy(op.hpair);
    if (op.pair) pair_conv_plan_destroy(op.pair);
  }
  if (net->d_head_w) cudaFree(net->d_head_w);
  if (net->d_head_b) cudaFree(net->d_head_b);
  if (net->d_logits) cudaFree(net->d_logits);
  delete net;
}

static int upload(spk_ctx* ctx, const float* host, size_t count, float** dev) {
  SPK_CUDA_OK(ctx, cudaMalloc(dev, count * sizeof(float)));
  SPK_CUDA_OK(ctx, cudaMemcpy(*dev, host, count * sizeof(float), cudaMemcpyHostToDevice));
  ctx->net->bytes += (int64_t)(count * sizeof(float));
  return SPK_OK;
}

static Buffer* get_buf(Net* net, int id, bool create) {
  if (id < 0 || id > 4096) return nullptr;
  if ((size_t)id >= net->bufs.size()) {
    if (!create) return nullptr;
    net->bufs.resize((size_t)id + 1);
  }
  return &net->bufs[(size_t)id];
}

// Every op that writes a whole buffer (re)defines its shape; a buffer id may be reused for
// tensors of different shapes, the allocation is the largest of them.
static void define_buf(Net* net, Buffer* b, int h, int w, int c) {
  b->known = true;
  b->declared = false;
  b->h = h;
  b->w = w;
  b->c = c;
  b->dtype = net->act_dtype;
  const size_t need = (size_t)net->max_batch * h * w * c * dtype_size(b->dtype);
  b->bytes = std::max(b->bytes, need);
}

static int need_net(spk_ctx* ctx, bool building, const char* who) {
  if (!ctx) return fail(nullptr, SPK_ERR_INVALID, "%s: null context", who);
  if (!ctx->net) return fail(ctx, SPK_ERR_STATE, "%s: call spk_net_begin first", who);
  if (building && ctx->net->ended) return fail(ctx, SPK_ERR_STATE, "%s: network already finalised", who);
  if (!building && !ctx->net->ended) return fail(ctx, SPK_ERR_STATE, "%s: call spk_net_end first", who);
  return SPK_OK;
}

static void fold_bn(int c, const float* g, const float* b, const float* m, const float* v, double eps,
                    std::vector<double>* scale, std::vector<double>* shift) {
  scale->assign((size_t)c, 1.0);
  shift->assign((size_t)c, 0.0);
  if (!v) return;
  for (int i = 0; i < c; ++i) {
    const double s = (g ? (double)g[i] : 1.0) / std::sqrt((double)v[i] + eps);
    (*scale)[i] = s;
    (*shift)[i] = (b ? (double)b[i] : 0.0) - (m ? (double)m[i] : 0.0) * s;
  }
}

}  // namespace spk

using namespace spk;

// No C++ exception may cross the C ABI (include/sykepic_b200.h): every entry point that allocates is a function-try-block.
#define SPK_ABI_CATCH(ctx_, name_)                                                                       \
  catch (const std::exception& e) {                                                                      \
    return fail(ctx_, SPK_ERR_STATE, name_ ": %s", e.what());                                            \
  }                                                                                                      \
  catch (...) {                                                                                          \
    return fail(ctx_, SPK_ERR_STATE, name_ ": unknown C++ exception");                                   \
  }

extern "C" {

int spk_create(int device, void* stream, spk_ctx** out) try {
  if (!out) return fail(nullptr, SPK_ERR_INVALID, "spk_create: null out");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, SPK_ERR_CUDA, "spk_create: no CUDA device (%s); there is no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(nullptr, SPK_ERR_INVALID, "spk_create: device %d of %d", device, count);
  std::unique_ptr<spk_ctx> ctx(new spk_ctx);
  ctx->device = device;
  ctx->stream = (cudaStream_t)stream;
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, SPK_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, SPK_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, SPK_ERR_UNSUPPORTED, "spk_create: device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  ctx->sm_count = prop.multiProcessorCount;
  e = cudaMalloc(&ctx->d_faults, sizeof(unsigned long long));
  if (e != cudaSuccess) return fail(nullptr, SPK_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  cudaMemsetAsync(ctx->d_faults, 0, sizeof(unsigned long long), ctx->stream);
  int rc = init_default_lut(ctx.get());
  if (rc != SPK_OK) {
    tls_error() = ctx->error;
    return rc;
  }
  *out = ctx.release();
  return SPK_OK;
} SPK_ABI_CATCH(nullptr, "spk_create")

int spk_destroy(spk_ctx* ctx) try {
  if (!ctx) return SPK_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  net_free(ctx->net);
  if (ctx->d_faults) cudaFree(ctx->d_faults);
  if (ctx->d_stamps) cudaFree(ctx->d_stamps);
  if (ctx->d_default_lut) cudaFree(ctx->d_default_lut);
  if (ctx->d_big_list) cudaFree(ctx->d_big_list);
  if (ctx->d_big_count) cudaFree(ctx->d_big_count);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  delete ctx;
  return SPK_OK;
} SPK_ABI_CATCH(nullptr, "spk_destroy")

int spk_set_stream(spk_ctx* ctx, void* stream) try {
  if (!ctx) return fail(nullptr, SPK_ERR_INVALID, "spk_set_stream: null context");
  ctx->stream = (cudaStream_t)stream;
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_set_stream")

int spk_synchronize(spk_ctx* ctx) try {
  if (!ctx) return fail(nullptr, SPK_ERR_INVALID, "spk_synchronize: null context");
  SPK_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_synchronize")

int64_t spk_launch_count(const spk_ctx* ctx) { return ctx ? ctx->launches : 0; }

int spk_fault_count(spk_ctx* ctx, int64_t* count) try {
  if (!ctx || !count) return fail(ctx, SPK_ERR_INVALID, "spk_fault_count: bad arguments");
  unsigned long long v = 0;
  SPK_CUDA_OK(ctx, cudaMemcpyAsync(&v, ctx->d_faults, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
  SPK_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  *count = (int64_t)v;
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_fault_count")

// ---------------------------------------------------------------------------------------- graph
int spk_net_begin(spk_ctx* ctx, int target_h, int target_w, int in_channels, int precision, int max_batch) try {
  if (!ctx) return fail(nullptr, SPK_ERR_INVALID, "spk_net_begin: null context");
  if (target_h < 1 || target_w < 1 || max_batch < 1) return fail(ctx, SPK_ERR_INVALID, "spk_net_begin: bad sizes");
  if (in_channels != 1 && in_channels != 3) return fail(ctx, SPK_ERR_UNSUPPORTED, "spk_net_begin: in_channels %d", in_channels);
  if (precision != SPK_PRECISION_FP32 && precision != SPK_PRECISION_BF16 && precision != SPK_PRECISION_FP32_TC)
    return fail(ctx, SPK_ERR_INVALID, "spk_net_begin: precision %d", precision);
  SPK_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  SPK_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  net_free(ctx->net);
  Net* net = new Net;
  ctx->net = net;
  net->th = target_h;
  net->tw = target_w;
  net->in_c = in_channels;
  net->precision = precision;
  // FP32_TC: activations as SplitF words (bf16 hi | lo) so that the convolutions can run on the tensor cores with
  // fp32-level accuracy (conv_tc.cu, kModeSplit).  FP32: plain fp32 activations, CUDA-core convolutions (exact fp32 FMA).
  net->act_dtype = precision == SPK_PRECISION_BF16 ? SPK_DTYPE_BF16 : (precision == SPK_PRECISION_FP32_TC ? SPK_DTYPE_SPLIT : SPK_DTYPE_F32);
  net->max_batch = max_batch;
  Buffer* b0 = get_buf(net, 0, true);
  b0->known = true;
  b0->h = target_h;
  b0->w = target_w;
  b0->c = in_channels;
  // 1 channel: the padded/resized bytes (ToTensor's v/255 is applied by the first convolution through
  // the 256-entry LUT); 3 channels: fp32 NHWC as the reference tensor holds it.
  b0->dtype = in_channels == 1 ? SPK_DTYPE_U8 : SPK_DTYPE_F32;
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_net_begin")

int spk_net_buffer(spk_ctx* ctx, int buf, int h, int w, int channels) try {
  int rc = need_net(ctx, true, "spk_net_buffer");
  if (rc) return rc;
  if (buf < 1 || h < 1 || w < 1 || channels < 1) return fail(ctx, SPK_ERR_INVALID, "spk_net_buffer: bad arguments");
  Buffer* b = get_buf(ctx->net, buf, true);
  if (!b) return fail(ctx, SPK_ERR_INVALID, "spk_net_buffer: buffer id %d", buf);
  define_buf(ctx->net, b, h, w, channels);
  b->declared = true;
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_net_buffer")

int spk_net_conv(spk_ctx* ctx, int in_buf, int in_c_off, int out_buf, int out_c_off, int res_buf, const float* weight,
                 int cout, int cin, int kh, int kw, int stride, int pad, const float* bn_gamma, const float* bn_beta,
                 const float* bn_mean, const float* bn_var, float bn_eps, const float* bias, int relu, int impl) try {
  int rc = need_net(ctx, true, "spk_net_conv");
  if (rc) return rc;
  Net* net = ctx->net;
  if (!weight || cout < 1 || cin < 1 || kh < 1 || kw < 1 || stride < 1 || pad < 0 || in_c_off < 0 || out_c_off < 0)
    return fail(ctx, SPK_ERR_INVALID, "spk_net_conv: bad arguments");
  Buffer* bi = get_buf(net, in_buf, false);
  if (!bi || !bi->known) return fail(ctx, SPK_ERR_STATE, "spk_net_conv: input buffer %d is undefined", in_buf);
  if (out_buf < 1 || out_buf == in_buf) return fail(ctx, SPK_ERR_INVALID, "spk_net_conv: output buffer %d", out_buf);
  // the 3 identical planes of the reference tensor (train/data.py:217-219) folded into 1
  const bool fold_rgb = (in_buf == 0 && net->in_c == 1 && cin == 3);
  const int cin_eff = fold_rgb ? 1 : cin;
  if (in_c_off + cin_eff > bi->c)
    return fail(ctx, SPK_ERR_INVALID, "spk_net_conv: channels [%d,%d) outside the %d of buffer %d", in_c_off,
                in_c_off + cin_eff, bi->c, in_buf);
  const int ho = (bi->h + 2 * pad - kh) / stride + 1, wo = (bi->w + 2 * pad - kw) / stride + 1;
  if (ho < 1 || wo < 1) return fail(ctx, SPK_ERR_INVALID, "spk_net_conv: empty output");
  const int in_h = bi->h, in_w = bi->w, in_ld = bi->c;  // bi may dangle after get_buf(create)
  Buffer* bo = get_buf(net, out_buf, true);
  if (!bo) return fail(ctx, SPK_ERR_INVALID, "spk_net_conv: output buffer id %d", out_buf);
  if (bo->declared) {
    if (bo->h != ho || bo->w != wo || out_c_off + cout > bo->c)
      return fail(ctx, SPK_ERR_STATE, "spk_net_conv: slice [%d,%d) of %dx%d does not fit declared buffer %d", out_c_off,
                  out_c_off + cout, ho, wo, out_buf);
  } else {
    if (out_c_off != 0) return fail(ctx, SPK_ERR_STATE, "spk_net_conv: buffer %d must be declared to be written by slices", out_buf);
    define_buf(net, bo, ho, wo, cout);
  }
  const int out_ld = bo->c;
  int res_ld = 0;
  if (res_buf >= 0) {
    Buffer* br = get_buf(net, res_buf, false);
    if (!br || !br->known || br->h != ho || br->w != wo || br->c != cout)
      return fail(ctx, SPK_ERR_INVALID, "spk_net_conv: residual buffer %d does not match %dx%dx%d", res_buf, ho, wo, cout);
    res_ld = br->c;
  }

  Op op;
  op.kind = kOpConv;
  op.in = in_buf;
  op.out = out_buf;
  op.res = res_buf >= 0 ? res_buf : -1;
  op.in_off = in_c_off;
  op.out_off = out_c_off;
  ConvGeom& g = op.g;
  g.n = 0;
  g.h = in_h;
  g.w = in_w;
  g.cin = cin_eff;
  g.ho = ho;
  g.wo = wo;
  g.cout = cout;
  g.kh = kh;
  g.kw = kw;
  g.stride = stride;
  g.pad = pad;
  g.relu = relu ? 1 : 0;
  g.ldx = in_ld;
  g.ldy = out_ld;
  g.ldres = res_ld;

  // ---- fold BatchNorm (double), then lay the weights out as [Cout][kh][kw][cin_eff]
  std::vector<double> scale, shift;
  fold_bn(cout, bn_gamma, bn_beta, bn_mean, bn_var, (double)bn_eps, &scale, &shift);
  const int taps = kh * kw;
  const int K = taps * cin_eff;
  op.w_host.assign((size_t)cout * K, 0.f);
  op.b_host.assign((size_t)cout, 0.f);
  for (int o = 0; o < cout; ++o) {
    for (int t = 0; t < taps; ++t) {
      if (fold_rgb) {
        double s = 0.0;
        for (int c = 0; c < cin; ++c) s += (double)weight[((size_t)o * cin + c) * taps + t];
        op.w_host[(size_t)o * K + t] = (float)(s * scale[o]);
      } else {
        for (int c = 0; c < cin; ++c)
          op.w_host[((size_t)o * taps + t) * cin + c] = (float)((double)weight[((size_t)o * cin + c) * taps + t] * scale[o]);
      }
    }
    op.b_host[o] = (float)((bias ? (double)bias[o] * scale[o] : 0.0) + shift[o]);
  }
  op.impl = impl;
  net->ops.push_back(std::move(op));
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_net_conv")

int spk_net_maxpool(spk_ctx* ctx, int in_buf, int out_buf, int k, int stride, int pad) try {
  int rc = need_net(ctx, true, "spk_net_maxpool");
  if (rc) return rc;
  Net* net = ctx->net;
  Buffer* bi = get_buf(net, in_buf, false);
  if (!bi || !bi->known || in_buf == 0) return fail(ctx, SPK_ERR_STATE, "spk_net_maxpool: input buffer %d", in_buf);
  if (k < 1 || stride < 1 || pad < 0 || out_buf < 1 || out_buf == in_buf) return fail(ctx, SPK_ERR_INVALID, "spk_net_maxpool: bad arguments");
  const int h = bi->h, w = bi->w, c = bi->c;
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  Buffer* bo = get_buf(net, out_buf, true);
  if (!bo) return fail(ctx, SPK_ERR_INVALID, "spk_net_maxpool: buffer id %d", out_buf);
  if (bo->declared) {
    if (bo->h != ho || bo->w != wo || c > bo->c) return fail(ctx, SPK_ERR_STATE, "spk_net_maxpool: declared buffer %d does not fit", out_buf);
  } else {
    define_buf(net, bo, ho, wo, c);
  }
  Op op;
  op.g.ldy = bo->c;
  op.kind = kOpMaxPool;
  op.in = in_buf;
  op.out = out_buf;
  op.k = k;
  op.stride = stride;
  op.pad = pad;
  op.g.h = h;
  op.g.w = w;
  op.g.cin = c;
  op.g.ho = ho;
  op.g.wo = wo;
  net->ops.push_back(std::move(op));
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_net_maxpool")

int spk_net_avgpool(spk_ctx* ctx, int in_buf, int out_buf, int k, int stride) try {
  int rc = need_net(ctx, true, "spk_net_avgpool");
  if (rc) return rc;
  Net* net = ctx->net;
  Buffer* bi = get_buf(net, in_buf, false);
  if (!bi || !bi->known || in_buf == 0) return fail(ctx, SPK_ERR_STATE, "spk_net_avgpool: input buffer %d", in_buf);
  if (k < 1 || stride < 1 || out_buf < 1 || out_buf == in_buf) return fail(ctx, SPK_ERR_INVALID, "spk_net_avgpool: bad arguments");
  const int h = bi->h, w = bi->w, c = bi->c;
  const int ho = (h - k) / stride + 1, wo = (w - k) / stride + 1;
  if (ho < 1 || wo < 1) return fail(ctx, SPK_ERR_INVALID, "spk_net_avgpool: empty output");
  Buffer* bo = get_buf(net, out_buf, true);
  if (!bo) return fail(ctx, SPK_ERR_INVALID, "spk_net_avgpool: buffer id %d", out_buf);
  // a declared (wider) concat buffer: the pooled tensor becomes its first c channels
  if (bo->declared) {
    if (bo->h != ho || bo->w != wo || c > bo->c) return fail(ctx, SPK_ERR_STATE, "spk_net_avgpool: declared buffer %d does not fit", out_buf);
  } else {
    define_buf(net, bo, ho, wo, c);
  }
  Op op;
  op.kind = kOpAvgPool;
  op.in = in_buf;
  op.out = out_buf;
  op.k = k;
  op.stride = stride;
  op.g.h = h;
  op.g.w = w;
  op.g.cin = c;
  op.g.ldx = c;
  op.g.ho = ho;
  op.g.wo = wo;
  op.g.ldy = bo->c;
  net->ops.push_back(std::move(op));
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_net_avgpool")

int spk_net_bn_relu(spk_ctx* ctx, int in_buf, int out_buf, int channels, const float* bn_gamma, const float* bn_beta,
                    const float* bn_mean, const float* bn_var, float bn_eps, int relu) try {
  int rc = need_net(ctx, true, "spk_net_bn_relu");
  if (rc) return rc;
  Net* net = ctx->net;
  Buffer* bi = get_buf(net, in_buf, false);
  if (!bi || !bi->known || in_buf == 0) return fail(ctx, SPK_ERR_STATE, "spk_net_bn_relu: input buffer %d", in_buf);
  if (channels < 1 || channels > bi->c || out_buf < 1 || out_buf == in_buf || !bn_var)
    return fail(ctx, SPK_ERR_INVALID, "spk_net_bn_relu: bad arguments");
  const int h = bi->h, w = bi->w, ld = bi->c;
  Buffer* bo = get_buf(net, out_buf, true);
  if (!bo) return fail(ctx, SPK_ERR_INVALID, "spk_net_bn_relu: buffer id %d", out_buf);
  define_buf(net, bo, h, w, channels);
  Op op;
  op.kind = kOpBnRelu;
  op.in = in_buf;
  op.out = out_buf;
  op.channels = channels;
  op.relu = relu ? 1 : 0;
  op.g.h = h;
  op.g.w = w;
  op.g.ldx = ld;
  std::vector<double> scale, shift;
  fold_bn(channels, bn_gamma, bn_beta, bn_mean, bn_var, (double)bn_eps, &scale, &shift);
  std::vector<float> s((size_t)channels), t((size_t)channels);
  for (int i = 0; i < channels; ++i) {
    s[i] = (float)scale[i];
    t[i] = (float)shift[i];
  }
  rc = upload(ctx, s.data(), s.size(), &op.d_scale);
  if (rc) return rc;
  rc = upload(ctx, t.data(), t.size(), &op.d_bias);
  if (rc) return rc;
  net->ops.push_back(std::move(op));
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_net_bn_relu")

int spk_net_head(spk_ctx* ctx, int in_buf, int n_layers, const float* const* weights, const float* const* biases,
                 const int* dims) try {
  int rc = need_net(ctx, true, "spk_net_head");
  if (rc) return rc;
  Net* net = ctx->net;
  Buffer* bi = get_buf(net, in_buf, false);
  if (!bi || !bi->known || in_buf == 0) return fail(ctx, SPK_ERR_STATE, "spk_net_head: input buffer %d", in_buf);
  if (n_layers < 1 || !weights || !biases || !dims) return fail(ctx, SPK_ERR_INVALID, "spk_net_head: bad arguments");
  if (dims[0] != bi->c) return fail(ctx, SPK_ERR_INVALID, "spk_net_head: head expects %d features, base gives %d", dims[0], bi->c);
  if (net->has_head) return fail(ctx, SPK_ERR_STATE, "spk_net_head: head already set");
  // Fold the activation-free Linear chain into one affine map, in double:
  // W = W_L ... W_1, b = W_L(... (W_2 b_1 + b_2) ...) + b_L.
  const int F = dims[0];
  std::vector<double> W((size_t)dims[1] * F), b((size_t)dims[1]);
  for (size_t i = 0; i < W.size(); ++i) W[i] = weights[0][i];
  for (int i = 0; i < dims[1]; ++i) b[i] = biases[0] ? biases[0][i] : 0.0;
  for (int l = 1; l < n_layers; ++l) {
    const int din = dims[l], dout = dims[l + 1];
    std::vector<double> W2((size_t)dout * F, 0.0), b2((size_t)dout, 0.0);
    for (int o = 0; o < dout; ++o) {
      const float* wrow = weights[l] + (size_t)o * din;
      double acc_b = biases[l] ? biases[l][o] : 0.0;
      double* dst = &W2[(size_t)o * F];
      for (int i = 0; i < din; ++i) {
        const double wv = wrow[i];
        acc_b += wv * b[i];
        const double* src = &W[(size_t)i * F];
        for (int f = 0; f < F; ++f) dst[f] += wv * src[f];
      }
      b2[o] = acc_b;
    }
    W.swap(W2);
    b.swap(b2);
  }
  const int K = dims[n_layers];
  std::vector<float> Wf(W.size()), bf(b.size());
  for (size_t i = 0; i < W.size(); ++i) Wf[i] = (float)W[i];
  for (size_t i = 0; i < b.size(); ++i) bf[i] = (float)b[i];
  rc = upload(ctx, Wf.data(), Wf.size(), &net->d_head_w);
  if (rc) return rc;
  rc = upload(ctx, bf.data(), bf.size(), &net->d_head_b);
  if (rc) return rc;
  net->has_head = true;
  net->head_in = in_buf;
  net->feat = F;
  net->classes = K;
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_net_head")

int spk_net_end(spk_ctx* ctx) try {
  int rc = need_net(ctx, true, "spk_net_end");
  if (rc) return rc;
  Net* net = ctx->net;
  if (!net->has_head) return fail(ctx, SPK_ERR_STATE, "spk_net_end: no head");
  SPK_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  // activation workspace: one allocation per buffer id, sized for max_batch
  for (size_t i = 1; i < net->bufs.size(); ++i) {
    Buffer& b = net->bufs[i];
    if (!b.known || b.bytes == 0) continue;
    SPK_CUDA_OK(ctx, cudaMalloc(&b.d, b.bytes));
    net->bytes += (int64_t)b.bytes;
  }
  SPK_CUDA_OK(ctx, cudaMalloc(&net->d_logits, (size_t)net->max_batch * net->classes * sizeof(float)));
  net->bytes += (int64_t)net->max_batch * net->classes * 4;
  // ---- stem fusion: conv 7x7/2 on the u8 input followed by max-pool 3x3/2 -> one tcgen05 kernel
  if (net->precision != SPK_PRECISION_FP32 && net->ops.size() >= 2) {  // BF16, and FP32_TC (SplitF output)
    Op& c0 = net->ops[0];
    Op& m1 = net->ops[1];
    // is the conv output read by anything but the pool before its buffer id is written again?
    bool other_reader = false;
    for (size_t i = 2; i < net->ops.size(); ++i) {
      if (net->ops[i].in == c0.out || net->ops[i].res == c0.out) {
        other_reader = true;
        break;
      }
      if (net->ops[i].out == c0.out) break;  // recycled id: overwritten first
    }
    if (c0.kind == kOpConv && c0.in == 0 && net->bufs[0].dtype == SPK_DTYPE_U8 && c0.impl != SPK_CONV_SIMT &&
        m1.kind == kOpMaxPool && m1.in == c0.out && c0.res < 0 && c0.out_off == 0 && net->head_in != c0.out &&
        stem_pool_supported(c0.g, m1.k, m1.stride, m1.pad) && (m1.g.ldy % 8) == 0 && !other_reader) {
      rc = stem_pool_pack_weights(ctx, c0.w_host.data(), &c0.d_stem_w, net->act_dtype == SPK_DTYPE_SPLIT);
      if (rc) return rc;
      net->bytes += 16384;
      rc = upload(ctx, c0.b_host.data(), c0.b_host.size(), &c0.d_bias);
      if (rc) return rc;
      c0.kind = kOpStemPool;
      c0.pool_out = m1.out;
      c0.hp = m1.g.ho;
      c0.wp = m1.g.wo;
      c0.pool_ld = m1.g.ldy;
      m1.kind = kOpNop;
      std::vector<float>().swap(c0.w_host);
      std::vector<float>().swap(c0.b_host);
    }
  }
  // ---- downsample fusion: a ResNet block's 1x1 / stride-2 shortcut reads exactly the pixels the centre tap of the
  // block's 3x3 / stride-2 convolution reads; the pair becomes one launch with two accumulators
  if (net->precision == SPK_PRECISION_BF16 && !debug_env("SPK_NO_DS_FUSION")) {
    for (size_t i = 0; i + 1 < net->ops.size(); ++i) {
      Op& a = net->ops[i];      // the shortcut (the host declares it first)
      Op& b = net->ops[i + 1];  // the 3x3
      if (a.kind != kOpConv || b.kind != kOpConv || a.in != b.in || a.in_off != b.in_off || a.res >= 0 || b.res >= 0) continue;
      if (a.impl == SPK_CONV_SIMT || b.impl == SPK_CONV_SIMT || a.out == b.out || a.in == 0) continue;
      if (net->bufs[(size_t)a.in].dtype != SPK_DTYPE_BF16) continue;
      ConvGeom ga = a.g, gb = b.g;
      ga.n = gb.n = net->max_batch;
      if (!tc_conv_ds_fusable(gb, ga)) continue;
      b.ds_out = a.out;
      b.ds_off = a.out_off;
      b.ds_ld = a.g.ldy;
      b.ds_w = std::move(a.w_host);
      b.ds_b = std::move(a.b_host);
      a.kind = kOpNop;
    }
  }
  // ---- pre-activation fusion (DenseNet): bn_relu(concat[:, :c]) -> 1x1 convolution becomes one launch; the BatchNorm +
  // ReLU is applied to the convolution's A tiles in shared memory (conv_tc.cu, kModePre).  Only unpadded convolutions:
  // a padding pixel must stay 0, not relu(shift).
  if (net->precision == SPK_PRECISION_BF16 && !debug_env("SPK_NO_PRE_FUSION")) {
    for (size_t i = 0; i + 1 < net->ops.size(); ++i) {
      Op& a = net->ops[i];      // bn_relu
      Op& b = net->ops[i + 1];  // its only consumer
      if (a.kind != kOpBnRelu || b.kind != kOpConv || b.in != a.out || b.in_off != 0 || b.pad != 0 || b.g.pad != 0) continue;
      if (b.impl == SPK_CONV_SIMT || b.impl == SPK_CONV_TCGEN05_TAPS || b.g.cin != a.channels || a.channels > 1024) continue;
      if (net->bufs[(size_t)a.in].dtype != SPK_DTYPE_BF16) continue;
      bool other_reader = false, overwritten = false;
      for (size_t j = i + 2; j < net->ops.size(); ++j) {
        const Op& o = net->ops[j];
        if (o.kind == kOpNop) continue;
        if (o.in == a.out || o.res == a.out) {
          other_reader = true;
          break;
        }
        if (o.out == a.out || o.pool_out == a.out || o.ds_out == a.out) {  // recycled id: overwritten first
          overwritten = true;
          break;
        }
      }
      if (!overwritten && net->head_in == a.out) other_reader = true;  // still live when the head reads it
      if (other_reader) continue;
      ConvGeom g = b.g;
      g.ldx = a.g.ldx;
      g.n = net->max_batch;
      if (!tc_conv_supported(g)) continue;
      b.in = a.in;
      b.g.ldx = a.g.ldx;
      b.d_scale = a.d_scale;
      b.d_pre_shift = a.d_bias;
      b.pre_c = a.channels;
      b.pre_relu = a.relu;
      a.d_scale = nullptr;
      a.d_bias = nullptr;
      a.kind = kOpNop;
    }
  }
  for (auto& op : net->ops) {
    if (op.kind != kOpConv) continue;
    const ConvGeom& g = op.g;
    const int K = g.kh * g.kw * g.cin;
    int impl = op.impl;
    const bool split_in = net->bufs[(size_t)op.in].dtype == SPK_DTYPE_SPLIT && net->bufs[(size_t)op.out].dtype == SPK_DTYPE_SPLIT &&
                          (op.res < 0 || net->bufs[(size_t)op.res].dtype == SPK_DTYPE_SPLIT);
    const bool tc_ok = net->precision == SPK_PRECISION_BF16
                           ? (net->bufs[(size_t)op.in].dtype == SPK_DTYPE_BF16 && tc_conv_supported(g))
                           : (split_in && tc_conv_split_supported(g));
    const bool taps_only = impl == SPK_CONV_TCGEN05_TAPS;
    if (taps_only) impl = SPK_CONV_TCGEN05;
    if (impl == SPK_CONV_AUTO) {
      impl = tc_ok ? SPK_CONV_TCGEN05 : SPK_CONV_SIMT;
      if (!tc_ok && net->precision != SPK_PRECISION_FP32 && op.in != 0) {
        // no silent 27x cliff: say which layer left the tensor cores (spk_net_simt_layers reports the count)
        ++net->simt_layers;
        fprintf(stderr, "spk: WARNING: convolution %dx%d/%d %d->%d on %dx%d has no tcgen05 kernel in this precision; it runs on CUDA cores\n",
                g.kh, g.kw, g.stride, g.cin, g.cout, g.h, g.w);
      }
    }
    if (impl == SPK_CONV_TCGEN05 && !tc_ok)
      return fail(ctx, SPK_ERR_UNSUPPORTED, "spk_net_end: tcgen05 convolution does not support %dx%d s%d cin %d cout %d here",
                  g.kh, g.kw, g.stride, g.cin, g.cout);
    op.impl = impl;
    rc = upload(ctx, op.b_host.data(), op.b_host.size(), &op.d_bias);
    if (rc) return rc;
    if (impl == SPK_CONV_TCGEN05) {
      ConvGeom gm = g;
      gm.n = net->max_batch;
      if (op.pre_c > 0) {
        rc = tc_conv_plan_create(ctx, gm, op.w_host.data(), op.d_bias, &op.tc);
        if (rc) return rc;
        rc = tc_conv_plan_set_prologue(ctx, op.tc, op.d_scale, op.d_pre_shift, op.pre_c, op.pre_relu);
        if (rc) return rc;
        net->bytes += tc_conv_plan_bytes(op.tc);
      } else if (net->precision != SPK_PRECISION_BF16) {
        rc = tc_conv_plan_create(ctx, gm, op.w_host.data(), op.d_bias, &op.tc, nullptr, nullptr, 0, true);
        if (rc) return rc;
        net->bytes += tc_conv_plan_bytes(op.tc);
      } else if (!taps_only && hp_conv_supported(gm)) {
        rc = hp_conv_plan_create(ctx, gm, op.w_host.data(), op.d_bias, &op.hpair);
        if (rc) return rc;
        net->bytes += hp_conv_plan_bytes(op.hpair);
      } else if (!taps_only && op.ds_out < 0 && pair_conv_supported(gm)) {
        rc = pair_conv_plan_create(ctx, gm, op.w_host.data(), op.d_bias, &op.pair);
        if (rc) return rc;
        net->bytes += pair_conv_plan_bytes(op.pair);
      } else if (op.ds_out >= 0) {
        rc = upload(ctx, op.ds_b.data(), op.ds_b.size(), &op.d_bias_ds);
        if (rc) return rc;
        if (!taps_only && pair_conv_supported(gm) && !debug_env("SPK_NO_PAIR_DS")) {
          rc = pair_conv_plan_create(ctx, gm, op.w_host.data(), op.d_bias, &op.pair, op.ds_w.data(), op.d_bias_ds, op.ds_ld);
          if (rc) return rc;
          net->bytes += pair_conv_plan_bytes(op.pair);
        } else {
          rc = tc_conv_plan_create(ctx, gm, op.w_host.data(), op.d_bias, &op.tc, op.ds_w.data(), op.d_bias_ds, op.ds_ld);
          if (rc) return rc;
          net->bytes += tc_conv_plan_bytes(op.tc);
        }
        std::vector<float>().swap(op.ds_w);
        std::vector<float>().swap(op.ds_b);
      } else {
        rc = tc_conv_plan_create(ctx, gm, op.w_host.data(), op.d_bias, &op.tc);
        if (rc) return rc;
        net->bytes += tc_conv_plan_bytes(op.tc);
      }
    } else {
      // SIMT layout: [K][Cout], K ordered (r, s, c)
      std::vector<float> wk((size_t)K * g.cout);
      for (int o = 0; o < g.cout; ++o)
        for (int k = 0; k < K; ++k) wk[(size_t)k * g.cout + o] = op.w_host[(size_t)o * K + k];
      rc = upload(ctx, wk.data(), wk.size(), &op.d_w);
      if (rc) return rc;
    }
    std::vector<float>().swap(op.w_host);
    std::vector<float>().swap(op.b_host);
  }
  // ---- L2-friendly traversal: launch k walks its tiles in the direction opposite to launch k-1 (the stem goes forward)
  if (!debug_env("SPK_NO_ZIGZAG")) {
    int dir = 0;
    for (auto& op : net->ops) {
      if (op.kind == kOpNop) continue;
      if (op.kind == kOpConv && op.impl == SPK_CONV_TCGEN05) {
        dir ^= 1;
        if (op.hpair) hp_conv_plan_set_reverse(op.hpair, dir);
        else if (op.pair) pair_conv_plan_set_reverse(op.pair, dir);
        else if (op.tc) tc_conv_plan_set_reverse(op.tc, dir);
        else dir = 0;
      } else {
        dir = 0;  // every other kernel walks forward
      }
    }
  }
  net->ended = true;
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_net_end")

int64_t spk_net_bytes(const spk_ctx* ctx) { return (ctx && ctx->net) ? ctx->net->bytes : 0; }

int spk_net_simt_layers(const spk_ctx* ctx) { return (ctx && ctx->net) ? ctx->net->simt_layers : 0; }

// ---------------------------------------------------------------------------------------- forward
int spk_forward(spk_ctx* ctx, const void* x, int64_t n, float softmax_scale, const int32_t* thr_q, float* probs,
                int32_t* label, uint8_t* classified) try {
  int rc = need_net(ctx, false, "spk_forward");
  if (rc) return rc;
  Net* net = ctx->net;
  if (n == 0) return SPK_OK;
  if (!x || !probs || n < 0) return fail(ctx, SPK_ERR_INVALID, "spk_forward: null buffer");
  if (n > net->max_batch) return fail(ctx, SPK_ERR_CAPACITY, "spk_forward: batch %lld > max_batch %d", (long long)n, net->max_batch);
  // channel offset `c_off` into buffer `id` (DenseNet concat slices); every op takes the whole batch
  auto ptr = [&](int id, int c_off, size_t = 0) -> char* {
    Buffer& b = net->bufs[(size_t)id];
    char* base = id == 0 ? (char*)const_cast<void*>(x) : (char*)b.d;
    return base + (size_t)c_off * dtype_size(b.dtype);
  };
  for (size_t oi = 0; oi < net->ops.size(); ++oi) {
    Op& op = net->ops[oi];
    const Buffer& bi = net->bufs[(size_t)op.in];
    const Buffer& bo = net->bufs[(size_t)op.out];
    switch (op.kind) {
      case kOpConv: {
        ConvGeom g = op.g;
        g.n = (int)n;
        const double px = (double)n * g.ho * g.wo;
        const double ds_taps = op.ds_out >= 0 ? 1.0 : 0.0;  // the fused 1x1 shortcut: one more tap, one more output
        ProfScope prof(ctx, op.in == 0 ? SPK_PROF_STEM : (op.impl == SPK_CONV_TCGEN05 ? SPK_PROF_CONV_TC : SPK_PROF_CONV_SIMT),
                       2.0 * px * g.cout * (g.kh * g.kw + ds_taps) * g.cin,
                       (double)n * g.h * g.w * g.cin * dtype_size(bi.dtype) + px * g.cout * dtype_size(bo.dtype) * ((op.res >= 0 ? 2 : 1) + ds_taps) +
                           (double)g.cout * (g.kh * g.kw + ds_taps) * g.cin * (op.impl == SPK_CONV_TCGEN05 ? 2 : 4),
                       "conv%dx%d/%d %d->%d in %dx%d out %dx%d n=%d%s%s%s", g.kh, g.kw, g.stride, g.cin, g.cout, g.h, g.w, g.ho,
                       g.wo, (int)n, op.res >= 0 ? " +res" : "", g.relu ? " relu" : "", op.hpair ? " [halo pair]" : (op.pair ? (op.ds_out >= 0 ? " [pair +1x1/2 shortcut]" : " [pair]") : (op.ds_out >= 0 ? " [+1x1/2 shortcut]" : (op.pre_c > 0 ? " [bn+relu on A]" : ""))));
        const size_t img_in = (size_t)g.h * g.w * g.ldx, img_out = (size_t)g.ho * g.wo * g.ldy;
        const void* res = op.res >= 0 ? ptr(op.res, 0, (size_t)g.ho * g.wo * g.ldres) : nullptr;
        if (op.hpair)
          rc = hp_conv_launch(ctx, op.hpair, (int)n, ptr(op.in, op.in_off, img_in), res, ptr(op.out, op.out_off, img_out));
        else if (op.pair)
          rc = pair_conv_launch(ctx, op.pair, (int)n, ptr(op.in, op.in_off, img_in), res, ptr(op.out, op.out_off, img_out),
                                op.ds_out >= 0 ? ptr(op.ds_out, op.ds_off, (size_t)g.ho * g.wo * op.ds_ld) : nullptr);
        else if (op.impl == SPK_CONV_TCGEN05)
          rc = tc_conv_launch(ctx, op.tc, (int)n, ptr(op.in, op.in_off, img_in), res, ptr(op.out, op.out_off, img_out),
                              op.ds_out >= 0 ? ptr(op.ds_out, op.ds_off, (size_t)g.ho * g.wo * op.ds_ld) : nullptr);
        else
          rc = launch_conv_simt(ctx, g, ptr(op.in, op.in_off), bi.dtype, op.d_w, op.d_bias, res, ptr(op.out, op.out_off),
                                bo.dtype);
        break;
      }
      case kOpNop:
        break;
      case kOpStemPool: {
        const ConvGeom& g = op.g;
        const Buffer& bp = net->bufs[(size_t)op.pool_out];
        ProfScope prof(ctx, SPK_PROF_STEM, 2.0 * n * g.ho * g.wo * g.cout * 49.0,
                       (double)n * g.h * g.w + (double)n * op.hp * op.wp * g.cout * 2.0,
                       "stem conv7x7/2 1->64 + relu + maxpool3/2 fused, in %dx%d out %dx%d n=%d", g.h, g.w, op.hp, op.wp, (int)n);
        (void)bp;
        rc = launch_stem_pool(ctx, (int)n, g.h, g.w, (const uint8_t*)ptr(0, 0, (size_t)g.h * g.w), op.d_stem_w, op.d_bias,
                              (__nv_bfloat16*)ptr(op.pool_out, 0, (size_t)op.hp * op.wp * op.pool_ld), g.ho,
                              g.wo, op.hp, op.wp, op.pool_ld, net->act_dtype == SPK_DTYPE_SPLIT);
        break;
      }
      case kOpMaxPool: {
        ProfScope prof(ctx, SPK_PROF_POOL, 0.0, (double)n * ((double)op.g.h * op.g.w + (double)op.g.ho * op.g.wo) * op.g.cin * dtype_size(bo.dtype),
                       "maxpool%d/%d c=%d in %dx%d n=%d", op.k, op.stride, op.g.cin, op.g.h, op.g.w, (int)n);
        rc = launch_maxpool(ctx, (int)n, op.g.h, op.g.w, op.g.cin, op.k, op.stride, op.pad, op.g.ho, op.g.wo, op.g.ldy,
                            ptr(op.in, 0), ptr(op.out, 0), bo.dtype);
        break;
      }
      case kOpAvgPool: {
        ProfScope prof(ctx, SPK_PROF_POOL, 0.0, (double)n * ((double)op.g.h * op.g.w + (double)op.g.ho * op.g.wo) * op.g.cin * dtype_size(bo.dtype),
                       "avgpool%d/%d c=%d in %dx%d n=%d", op.k, op.stride, op.g.cin, op.g.h, op.g.w, (int)n);
        rc = launch_avgpool(ctx, (int)n, op.g.h, op.g.w, op.g.cin, op.g.ldx, op.k, op.stride, op.g.ho, op.g.wo, op.g.ldy,
                            ptr(op.in, 0), ptr(op.out, 0), bo.dtype);
        break;
      }
      case kOpBnRelu: {
        ProfScope prof(ctx, SPK_PROF_BN_RELU, 0.0, 2.0 * n * op.g.h * op.g.w * op.channels * dtype_size(bo.dtype),
                       "bn_relu c=%d %dx%d n=%d", op.channels, op.g.h, op.g.w, (int)n);
        rc = launch_affine_relu(ctx, (long long)n * op.g.h * op.g.w, op.channels, op.g.ldx, op.d_scale, op.d_bias,
                                ptr(op.in, 0), ptr(op.out, 0), bo.dtype, op.relu);
        break;
      }
      default:
        rc = fail(ctx, SPK_ERR_STATE, "spk_forward: unknown op");
    }
    if (rc) return rc;
  }
  const Buffer& bh = net->bufs[(size_t)net->head_in];
  ProfScope prof(ctx, SPK_PROF_HEAD, 2.0 * n * net->feat * net->classes,
                 (double)n * bh.h * bh.w * net->feat * dtype_size(bh.dtype) + (double)n * net->classes * 4,
                 "head pool %dx%d f=%d k=%d n=%d", bh.h, bh.w, net->feat, net->classes, (int)n);
  return launch_head(ctx, bh.d, bh.dtype, n, bh.h * bh.w, net->feat, net->d_head_w, net->d_head_b, net->classes,
                     softmax_scale, thr_q, net->d_logits, probs, label, classified);
} SPK_ABI_CATCH(ctx, "spk_forward")

const float* spk_last_logits(const spk_ctx* ctx) { return (ctx && ctx->net) ? ctx->net->d_logits : nullptr; }

int spk_net_read_buffer(spk_ctx* ctx, int buf, int64_t n, float* host_out, int64_t cap_elems, int* h, int* w, int* c) try {
  int rc = need_net(ctx, false, "spk_net_read_buffer");
  if (rc) return rc;
  Net* net = ctx->net;
  Buffer* b = get_buf(net, buf, false);
  if (!b || !b->known || buf == 0 || !b->d) return fail(ctx, SPK_ERR_INVALID, "spk_net_read_buffer: buffer %d", buf);
  if (h) *h = b->h;
  if (w) *w = b->w;
  if (c) *c = b->c;
  const int64_t elems = n * b->h * b->w * b->c;
  if (!host_out) return SPK_OK;
  if (elems > cap_elems || n > net->max_batch) return fail(ctx, SPK_ERR_CAPACITY, "spk_net_read_buffer: need %lld elements", (long long)elems);
  SPK_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (b->dtype == SPK_DTYPE_F32) {
    SPK_CUDA_OK(ctx, cudaMemcpy(host_out, b->d, (size_t)elems * 4, cudaMemcpyDeviceToHost));
  } else if (b->dtype == SPK_DTYPE_SPLIT) {
    std::vector<uint32_t> tmp((size_t)elems);
    SPK_CUDA_OK(ctx, cudaMemcpy(tmp.data(), b->d, (size_t)elems * 4, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < elems; ++i) {
      const float fh = __half2float(__ushort_as_half((unsigned short)(tmp[(size_t)i] & 0xffffu)));
      const float fl = __half2float(__ushort_as_half((unsigned short)(tmp[(size_t)i] >> 16)));
      host_out[i] = fh + fl;
    }
  } else {
    std::vector<uint16_t> tmp((size_t)elems);
    SPK_CUDA_OK(ctx, cudaMemcpy(tmp.data(), b->d, (size_t)elems * 2, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < elems; ++i) {
      const uint32_t bits = (uint32_t)tmp[(size_t)i] << 16;
      memcpy(&host_out[i], &bits, 4);
    }
  }
  return SPK_OK;
} SPK_ABI_CATCH(ctx, "spk_net_read_buffer")

}  // extern "C"

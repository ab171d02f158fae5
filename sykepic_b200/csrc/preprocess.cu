// K1: ROI decode + mode border + cv2-exact bilinear resize + centred pad + ToTensor (LUT).
//
// Replaces, for a whole batch of ROIs in one launch, the reference's
//   ifcb.raw_to_png            sykepic/utils/ifcb.py:76-118        (slice the .roi byte stream)
//   ImageDataset.__getitem__   sykepic/train/data.py:210-231       (PNG read, gray -> 3 equal planes)
//   Compose.__call__           sykepic/train/image.py:25-56        (mode border, get_new_dims)
//   resize_with_border         sykepic/train/image.py:201-226      (cv2.resize INTER_LINEAR + copyMakeBorder)
//   ToTensor (+ Normalize)     sykepic/train/config.py:52-56       (a 3x256 fp32 LUT is exact)
// Integer semantics are OpenCV's 11-bit fixed-point path (resize.cpp, INTER_RESIZE_COEF_BITS = 11),
// restated in oracle/preprocess.py and pinned bit-for-bit there.
//
// Mapping: one CTA per (ROI, slab of output rows).  HBM-bound: the ROI bytes are read once
// (the histogram pass uses 16-byte loads; the horizontal pass re-reads them through L1/L2),
// horizontally interpolated source rows are staged in shared memory as u16, and every output
// pixel is written exactly once with 16-byte stores.
#include "spk_internal.h"

namespace spk {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kHBytes = 28 * 1024;  // shared-memory budget of the horizontally resized rows

enum ResizeKind { kBilinear = 0, kCopy = 1, kArea2x = 2 };

struct Params {
  const uint8_t* roi;
  long long roi_len;
  const long long* start;
  const int* w;
  const int* h;
  int th, tw;
  int border_mode;
  int channels;
  int out_dtype;
  int out_layout;
  const float* lut;
  void* out;
  unsigned long long* faults;
  int slabs;
};

// OpenCV: f = float((d + 0.5) * scale - 0.5); s = floor(f); f -= s.  No FMA contraction allowed.
__device__ __forceinline__ void src_coord(int d, double scale, int* s, float* f) {
  double t = __dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  float ff = __double2float_rn(t);
  float fl = floorf(ff);
  *s = (int)fl;
  *f = __fsub_rn(ff, fl);
}

// cv::saturate_cast<short>(float): round half to even
__device__ __forceinline__ int coef(float v) { return __float2int_rn(__fmul_rn(v, 2048.0f)); }

template <typename T>
__device__ __forceinline__ T cvt(float v);
template <>
__device__ __forceinline__ float cvt<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct Tables {
  unsigned short xofs[kMaxTarget];
  short xa0[kMaxTarget];
  short xa1[kMaxTarget];
  int ysy[kMaxTarget];  // un-clamped source row of the first vertical tap
  short yb0[kMaxTarget];
  short yb1[kMaxTarget];
};

// Store 4 consecutive output pixels (x .. x+3, same row) of every channel.
template <typename T>
__device__ __forceinline__ void store4(const Params& p, const float* lut_s, long long n, int r, int x, const int v[4], int count) {
  T* out = (T*)p.out;
  const int C = p.channels;
  if (p.out_layout == SPK_LAYOUT_NCHW || C == 1) {
    for (int c = 0; c < C; ++c) {
      T* dst = out + ((n * C + c) * p.th + r) * (long long)p.tw + x;
      const float* l = lut_s + c * 256;
      if (count == 4 && (p.tw & 3) == 0) {
        if constexpr (sizeof(T) == 4) {
          float4 q = make_float4(l[v[0]], l[v[1]], l[v[2]], l[v[3]]);
          *reinterpret_cast<float4*>(dst) = q;
        } else {
          __nv_bfloat162 a = __floats2bfloat162_rn(l[v[0]], l[v[1]]);
          __nv_bfloat162 b = __floats2bfloat162_rn(l[v[2]], l[v[3]]);
          uint2 q;
          q.x = *reinterpret_cast<unsigned*>(&a);
          q.y = *reinterpret_cast<unsigned*>(&b);
          *reinterpret_cast<uint2*>(dst) = q;
        }
      } else {
        for (int i = 0; i < count; ++i) dst[i] = cvt<T>(l[v[i]]);
      }
    }
  } else {  // NHWC, C == 3: 12 contiguous elements
    T* dst = out + ((n * p.th + r) * (long long)p.tw + x) * C;
    if (count == 4 && (p.tw & 3) == 0) {
      float e[12];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) e[i * 3 + c] = lut_s[c * 256 + v[i]];
      if constexpr (sizeof(T) == 4) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        d4[0] = make_float4(e[0], e[1], e[2], e[3]);
        d4[1] = make_float4(e[4], e[5], e[6], e[7]);
        d4[2] = make_float4(e[8], e[9], e[10], e[11]);
      } else {
        uint2* d2 = reinterpret_cast<uint2*>(dst);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          __nv_bfloat162 a = __floats2bfloat162_rn(e[j * 4 + 0], e[j * 4 + 1]);
          __nv_bfloat162 b = __floats2bfloat162_rn(e[j * 4 + 2], e[j * 4 + 3]);
          uint2 q;
          q.x = *reinterpret_cast<unsigned*>(&a);
          q.y = *reinterpret_cast<unsigned*>(&b);
          d2[j] = q;
        }
      }
    } else {
      for (int i = 0; i < count; ++i)
        for (int c = 0; c < C; ++c) dst[i * C + c] = cvt<T>(lut_s[c * 256 + v[i]]);
    }
  }
}

__device__ __forceinline__ void store4_u8(const Params& p, long long n, int r, int x, const int v[4], int count) {
  uint8_t* dst = (uint8_t*)p.out + (n * p.th + r) * (long long)p.tw + x;
  if (count == 4 && (p.tw & 3) == 0) {
    *reinterpret_cast<unsigned*>(dst) = (unsigned)v[0] | ((unsigned)v[1] << 8) | ((unsigned)v[2] << 16) | ((unsigned)v[3] << 24);
  } else {
    for (int i = 0; i < count; ++i) dst[i] = (uint8_t)v[i];
  }
}

__global__ void __launch_bounds__(kThreads) preprocess_kernel(Params p) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  unsigned short* hbuf = reinterpret_cast<unsigned short*>(dyn_smem);  // [hrows][new_w]
  __shared__ unsigned hist[kWarps][256];
  __shared__ Tables tb;
  __shared__ float lut_s[3 * 256];
  __shared__ int s_mode;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n = blockIdx.x / p.slabs;
  const int slab = blockIdx.x % p.slabs;
  const int th = p.th, tw = p.tw;

  for (int i = tid; i < 3 * 256; i += kThreads) lut_s[i] = p.lut[i];

  const int w = p.w[n], h = p.h[n];
  const long long start = p.start[n];
  bool valid = (w >= 1) && (h >= 1) && (start >= 0) && (start + (long long)w * h <= p.roi_len);
  int nh = 0, nw = 0;
  if (valid) {
    new_dims(h, w, th, tw, &nh, &nw);
    valid = (nh >= 1) && (nw >= 1) && (nh <= th) && (nw <= tw);
  }
  const int rows_per_slab = (th + p.slabs - 1) / p.slabs;
  const int r0 = slab * rows_per_slab;
  const int r1 = min(th, r0 + rows_per_slab);
  if (!valid) {
    if (slab == 0 && tid == 0) atomicAdd(p.faults, 1ULL);
    nh = 0;
    nw = 0;
  }
  const uint8_t* src = p.roi + (valid ? start : 0);
  const int top = (th - nh) / 2, left = (tw - nw) / 2;
  const int img_r0 = max(r0, top), img_r1 = min(r1, top + nh);  // image rows inside this slab
  const bool has_border = (r0 < top) || (r1 > top + nh) || (left > 0) || (left + nw < tw);

  int kind = kBilinear;
  if (nw == w && nh == h) kind = kCopy;
  else if (w == 2 * nw && h == 2 * nh) kind = kArea2x;

  // ---- border value: 256-bin histogram of the ORIGINAL ROI, lowest value wins ties ----------
  int border = p.border_mode == SPK_BORDER_WHITE ? 255 : 0;
  if (p.border_mode == SPK_BORDER_MODE && valid && has_border) {
    for (int i = tid; i < kWarps * 256; i += kThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    const long long total = (long long)w * h;
    unsigned* my = hist[warp];
    // unaligned head, 16-byte body, tail
    const unsigned long long addr = (unsigned long long)src;
    long long head = (long long)((16 - (addr & 15)) & 15);
    if (head > total) head = total;
    for (long long i = tid; i < head; i += kThreads) atomicAdd(&my[src[i]], 1u);
    const long long nvec = (total - head) / 16;
    const uint4* v4 = reinterpret_cast<const uint4*>(src + head);
    for (long long i = tid; i < nvec; i += kThreads) {
      uint4 q = __ldg(v4 + i);
      unsigned wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&my[wds[j] & 255u], 1u);
        atomicAdd(&my[(wds[j] >> 8) & 255u], 1u);
        atomicAdd(&my[(wds[j] >> 16) & 255u], 1u);
        atomicAdd(&my[wds[j] >> 24], 1u);
      }
    }
    for (long long i = head + nvec * 16 + tid; i < total; i += kThreads) atomicAdd(&my[src[i]], 1u);
    __syncthreads();
    // 256 threads: one bin each, then argmax with the lowest index winning ties
    unsigned cnt = 0;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) cnt += hist[k][tid];
    // pack (count, 255 - bin) so that max picks the highest count, then the lowest bin
    unsigned long long key = ((unsigned long long)cnt << 8) | (unsigned)(255 - tid);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other > key ? other : key;
    }
    __syncthreads();  // everyone has read hist[][tid]; reuse hist[0] for the warp maxima
    if (lane == 0) reinterpret_cast<unsigned long long*>(&hist[0][0])[warp] = key;
    __syncthreads();
    if (tid == 0) {
      unsigned long long best = 0;
      for (int k = 0; k < kWarps; ++k) {
        unsigned long long v = reinterpret_cast<unsigned long long*>(&hist[0][0])[k];
        best = v > best ? v : best;
      }
      s_mode = 255 - (int)(best & 255u);
    }
    __syncthreads();
    border = s_mode;
  }

  // ---- coefficient tables ------------------------------------------------------------------
  if (kind == kBilinear && img_r1 > img_r0) {
    const double scale_x = __ddiv_rn(1.0, __ddiv_rn((double)nw, (double)w));
    const double scale_y = __ddiv_rn(1.0, __ddiv_rn((double)nh, (double)h));
    for (int dx = tid; dx < nw; dx += kThreads) {
      int s;
      float f;
      src_coord(dx, scale_x, &s, &f);
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= w - 1) { s = w - 1; f = 0.f; }
      tb.xofs[dx] = (unsigned short)s;
      tb.xa0[dx] = (short)coef(__fsub_rn(1.0f, f));
      tb.xa1[dx] = (short)coef(f);
    }
    for (int dy = img_r0 - top + tid; dy < img_r1 - top; dy += kThreads) {
      int s;
      float f;
      src_coord(dy, scale_y, &s, &f);
      tb.ysy[dy] = s;
      tb.yb0[dy] = (short)coef(__fsub_rn(1.0f, f));
      tb.yb1[dy] = (short)coef(f);
    }
  }
  __syncthreads();

  const int groups = (tw + 3) >> 2;  // 4-pixel groups per output row
  const float inv_groups = 1.0f / (float)groups;

  // Writes output rows [ra, rb); rows inside [img_r0, img_r1) take pixels from hbuf (source row ylo == hbuf row 0).
  auto write_rows = [&](int ra, int rb, int ylo) {
    const int total = (rb - ra) * groups;
    for (int it = tid; it < total; it += kThreads) {
      int q = __float2int_rz(((float)it + 0.5f) * inv_groups);  // it / groups without an integer divide
      q -= (q * groups > it);
      q += ((q + 1) * groups <= it);
      const int r = ra + q;
      const int x = (it - q * groups) << 2;
      const int count = min(4, tw - x);
      int v[4];
      const bool img_row = (r >= top) && (r < top + nh);
      if (!img_row) {
        v[0] = v[1] = v[2] = v[3] = border;
      } else {
        const int dy = r - top;
        if (kind == kBilinear) {
          const int sy = tb.ysy[dy];
          const int y0 = min(max(sy, 0), h - 1) - ylo, y1 = min(max(sy + 1, 0), h - 1) - ylo;
          const int b0 = tb.yb0[dy], b1 = tb.yb1[dy];
          const unsigned short* h0 = hbuf + y0 * nw;
          const unsigned short* h1 = hbuf + y1 * nw;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int dx = x + i - left;
            int val = border;
            if (dx >= 0 && dx < nw) val = (((b0 * (int)h0[dx]) >> 16) + ((b1 * (int)h1[dx]) >> 16) + 2) >> 2;
            v[i] = val;
          }
        } else if (kind == kCopy) {
          const uint8_t* row = src + (long long)dy * w;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int dx = x + i - left;
            v[i] = (dx >= 0 && dx < nw) ? (int)__ldg(row + dx) : border;
          }
        } else {  // exact 2x decimation: INTER_LINEAR silently becomes INTER_AREA
          const uint8_t* row = src + (long long)(2 * dy) * w;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int dx = x + i - left;
            int val = border;
            if (dx >= 0 && dx < nw)
              val = ((int)__ldg(row + 2 * dx) + (int)__ldg(row + 2 * dx + 1) + (int)__ldg(row + w + 2 * dx) +
                     (int)__ldg(row + w + 2 * dx + 1) + 2) >> 2;
            v[i] = val;
          }
        }
      }
      if (p.out_dtype == SPK_DTYPE_F32) store4<float>(p, lut_s, n, r, x, v, count);
      else if (p.out_dtype == SPK_DTYPE_BF16) store4<__nv_bfloat16>(p, lut_s, n, r, x, v, count);
      else store4_u8(p, n, r, x, v, count);
    }
  };

  // rows above the image
  if (r0 < img_r0 || img_r1 <= img_r0) write_rows(r0, img_r1 > img_r0 ? img_r0 : r1, 0);
  if (img_r1 > img_r0) {
    if (kind != kBilinear) {
      write_rows(img_r0, img_r1, 0);
    } else {
      const int hrows_cap = max(2, (kHBytes / 2) / nw);
      int c0 = img_r0;
      while (c0 < img_r1) {
        // greedy chunk of output rows whose source rows fit the staging buffer
        const int ylo = min(max(tb.ysy[c0 - top], 0), h - 1);
        int c1 = c0 + 1;
        while (c1 < img_r1 && min(max(tb.ysy[c1 - top] + 1, 0), h - 1) - ylo + 1 <= hrows_cap) ++c1;
        const int yhi = min(max(tb.ysy[c1 - 1 - top] + 1, 0), h - 1);
        const int nrows = yhi - ylo + 1;
        // horizontal pass: H = (S[sx]*a0 + S[sx+1]*a1) >> 4, one warp per source row
        for (int yy = warp; yy < nrows; yy += kWarps) {
          const uint8_t* row = src + (long long)(ylo + yy) * w;
          unsigned short* hrow = hbuf + yy * nw;
          for (int dx = lane; dx < nw; dx += 32) {
            const int sx = tb.xofs[dx];
            const int sx1 = min(sx + 1, w - 1);
            const int acc = (int)__ldg(row + sx) * (int)tb.xa0[dx] + (int)__ldg(row + sx1) * (int)tb.xa1[dx];
            hrow[dx] = (unsigned short)(acc >> 4);
          }
        }
        __syncthreads();
        write_rows(c0, c1, ylo);
        __syncthreads();
        c0 = c1;
      }
    }
    // rows below the image
    if (r1 > img_r1) write_rows(img_r1, r1, 0);
  }
}

__global__ void fill_default_lut(float* lut) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * 256) lut[i] = __fdiv_rn((float)(i & 255), 255.0f);  // ToTensor: true division
}

}  // namespace

int init_default_lut(spk_ctx* ctx) {
  SPK_CUDA_OK(ctx, cudaMalloc(&ctx->d_default_lut, 3 * 256 * sizeof(float)));
  fill_default_lut<<<3, 256, 0, ctx->stream>>>(ctx->d_default_lut);
  SPK_LAUNCH_CHECK(ctx);
  return SPK_OK;
}

}  // namespace spk

using namespace spk;

extern "C" int spk_preprocess(spk_ctx* ctx, const uint8_t* roi_bytes, int64_t roi_len, const int64_t* start,
                              const int32_t* width, const int32_t* height, int64_t n, int target_h, int target_w,
                              int border_mode, int channels, int out_dtype, int out_layout, const float* lut, void* out) {
  if (!ctx) return fail(nullptr, SPK_ERR_INVALID, "spk_preprocess: null context");
  if (n == 0) return SPK_OK;
  if (!roi_bytes || !start || !width || !height || !out || n < 0)
    return fail(ctx, SPK_ERR_INVALID, "spk_preprocess: null buffer");
  if (target_h < 1 || target_w < 1 || target_h > kMaxTarget || target_w > kMaxTarget)
    return fail(ctx, SPK_ERR_UNSUPPORTED, "spk_preprocess: target %dx%d outside [1,%d]", target_h, target_w, kMaxTarget);
  if (border_mode < 0 || border_mode > 2) return fail(ctx, SPK_ERR_INVALID, "spk_preprocess: border_mode %d", border_mode);
  if (channels != 1 && channels != 3) return fail(ctx, SPK_ERR_UNSUPPORTED, "spk_preprocess: channels %d", channels);
  if (out_dtype == SPK_DTYPE_U8 && channels != 1)
    return fail(ctx, SPK_ERR_INVALID, "spk_preprocess: u8 output has one channel");
  if (out_dtype < 0 || out_dtype > 2 || out_layout < 0 || out_layout > 1)
    return fail(ctx, SPK_ERR_INVALID, "spk_preprocess: bad dtype/layout");
  Params p;
  p.roi = roi_bytes;
  p.roi_len = roi_len;
  p.start = reinterpret_cast<const long long*>(start);
  p.w = width;
  p.h = height;
  p.th = target_h;
  p.tw = target_w;
  p.border_mode = border_mode;
  p.channels = channels;
  p.out_dtype = out_dtype;
  p.out_layout = out_layout;
  p.lut = lut ? lut : ctx->d_default_lut;
  p.out = out;
  p.faults = ctx->d_faults;
  p.slabs = 4;
  const double out_bytes = (double)n * target_h * target_w * channels * (out_dtype == SPK_DTYPE_F32 ? 4 : out_dtype == SPK_DTYPE_BF16 ? 2 : 1);
  // input bytes are data dependent (sum of w*h); the caller adds them -- recorded here: descriptors + output
  ProfScope prof(ctx, SPK_PROF_PREPROCESS, 0.0, out_bytes + 16.0 * n, "preprocess T=%dx%d c=%d dtype=%d n=%lld", target_h, target_w,
                 channels, out_dtype, (long long)n);
  const long long blocks = n * p.slabs;
  if (blocks > 0x7fffffffLL) return fail(ctx, SPK_ERR_UNSUPPORTED, "spk_preprocess: batch too large");
  preprocess_kernel<<<(unsigned)blocks, kThreads, kHBytes, ctx->stream>>>(p);
  SPK_LAUNCH_CHECK(ctx);
  return SPK_OK;
}

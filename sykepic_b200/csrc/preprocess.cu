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
#include <cooperative_groups.h>

#include <exception>

#include "spk_internal.h"

namespace cg = cooperative_groups;

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
  int* big_list;        // u8 path: ROIs over `big_bytes` are queued here by the warp kernel ...
  unsigned* big_count;  // ... and resized by clusters of kBigSlabs CTAs (preprocess_big_kernel)
  long long big_bytes;
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

// One (ROI, slab of output rows) per CTA of kThreads threads.  kCluster: the CTAs of a thread-block cluster hold the
// slabs of one ROI; each histograms 1/slabs of the ROI bytes and the partial histograms are summed through
// distributed shared memory (otherwise every slab CTA histograms the whole ROI).
template <bool kCluster>
__device__ __forceinline__ void roi_slab(const Params& p, const long long n, const int slab) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  unsigned short* hbuf = reinterpret_cast<unsigned short*>(dyn_smem);  // [hrows][new_w]
  __shared__ unsigned hist[kWarps][256];
  __shared__ unsigned hist_cta[256];
  __shared__ Tables tb;
  __shared__ float lut_s[3 * 256];
  __shared__ int s_mode;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int th = p.th, tw = p.tw;

  if (p.out_dtype != SPK_DTYPE_U8)
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
  if (p.border_mode == SPK_BORDER_MODE && valid && (has_border || kCluster)) {
    for (int i = tid; i < kWarps * 256; i += kThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    const long long total = (long long)w * h;
    // this CTA's share of the ROI bytes: everything, or the slab-th part in 16-byte units (cluster)
    long long lo = 0, hi = total;
    if (kCluster) {
      const long long per = (((total + p.slabs - 1) / p.slabs) + 15) & ~15LL;
      lo = min(total, per * slab);
      hi = min(total, lo + per);
    }
    unsigned* my = hist[warp];
    const uint8_t* part = src + lo;
    const long long part_total = hi - lo;
    // unaligned head, 16-byte body, tail
    const unsigned long long addr = (unsigned long long)part;
    long long head = (long long)((16 - (addr & 15)) & 15);
    if (head > part_total) head = part_total;
    for (long long i = tid; i < head; i += kThreads) atomicAdd(&my[part[i]], 1u);
    const long long nvec = (part_total - head) / 16;
    const uint4* v4 = reinterpret_cast<const uint4*>(part + head);
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
    for (long long i = head + nvec * 16 + tid; i < part_total; i += kThreads) atomicAdd(&my[part[i]], 1u);
    __syncthreads();
    // 256 threads: one bin each, then argmax with the lowest index winning ties
    unsigned cnt = 0;
#pragma unroll
    for (int k = 0; k < kWarps; ++k) cnt += hist[k][tid];
    if (kCluster) {
      cg::cluster_group cluster = cg::this_cluster();
      hist_cta[tid] = cnt;
      cluster.sync();
      cnt = 0;
      for (int r = 0; r < p.slabs; ++r) cnt += cluster.map_shared_rank(hist_cta, r)[tid];
      cluster.sync();  // nobody leaves (or overwrites hist_cta) while a peer still reads it
    }
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



__global__ void __launch_bounds__(kThreads) preprocess_kernel(Params p) {
  roi_slab<false>(p, blockIdx.x / p.slabs, blockIdx.x % p.slabs);
}

// u8 path, ROIs over p.big_bytes (the ~1 % heavy tail of IFCB ROIs, up to 1380x1034): one cluster of kBigSlabs CTAs per
// queued ROI, looping over the queue the warp kernel filled.
constexpr int kBigSlabs = 8;
constexpr int kBigClusters = 74;  // 592 CTAs = 4 per SM
constexpr long long kBigBytes = 24 * 1024;  // ROIs over this many bytes go to the cluster kernel

__global__ void __cluster_dims__(kBigSlabs, 1, 1) __launch_bounds__(kThreads) preprocess_big_kernel(Params p) {
  const unsigned count = *p.big_count;
  const int slab = blockIdx.x % kBigSlabs;
  for (unsigned i = blockIdx.x / kBigSlabs; i < count; i += gridDim.x / kBigSlabs) {
    roi_slab<true>(p, p.big_list[i], slab);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// u8 fast path (the engine's path): ONE WARP PER ROI, every lane owns 8 adjacent output columns.
//   * The ROI bytes are read from HBM once, with 16-byte loads, into a per-warp shared-memory stage
//     (ROIs over kStage bytes -- ~5 % of IFCB ROIs -- are re-read through L1/L2 instead) while the
//     256-bin histogram of the mode border is built (4 lane-interleaved sub-histograms of packed
//     16-bit counters, flushed into per-lane 32-bit totals every kHistSeg bytes).
//   * A lane keeps the two horizontally interpolated source rows of its 8 columns in registers and
//     walks down the output rows (the vertical taps of a column only need that column), so there is
//     no intermediate buffer and no block-wide barrier; rows are written with 8-byte stores, the
//     border rows above/below the image with 16-byte stores.
// Algorithmic HBM bytes per ROI: w*h read + 16 B descriptor + T*T written.
constexpr int kWarpsPerCta = 4;
constexpr int kStage = 8192;           // staged ROI bytes per warp
constexpr int kCols = 8;               // output columns per lane
constexpr int kHistSeg = 128 * 1024;   // bytes per histogram segment: <= 4 sub-histograms * 32767 per packed counter

struct RowCoef {
  int sy;   // un-clamped source row of the first vertical tap
  int b01;  // b0 | b1 << 16
};

__device__ __forceinline__ RowCoef row_coef(int dy, double scale_y) {
  int s;
  float f;
  src_coord(dy, scale_y, &s, &f);
  RowCoef r;
  r.sy = s;
  r.b01 = coef(__fsub_rn(1.0f, f)) | (coef(f) << 16);
  return r;
}

__device__ __forceinline__ void warp_fill(uint8_t* dst, long long bytes, int value, int lane) {
  if (bytes <= 0) return;
  const unsigned long long a = (unsigned long long)dst;
  long long head = (long long)((16 - (a & 15)) & 15);
  if (head > bytes) head = bytes;
  if (lane < head) dst[lane] = (uint8_t)value;
  const unsigned v4 = (unsigned)value * 0x01010101u;
  const uint4 q = make_uint4(v4, v4, v4, v4);
  uint4* body = reinterpret_cast<uint4*>(dst + head);
  const long long nvec = (bytes - head) >> 4;
  for (long long i = lane; i < nvec; i += 32) body[i] = q;
  const long long done = head + (nvec << 4);
  if (lane < bytes - done) dst[done + lane] = (uint8_t)value;
}

// The ROI bytes as seen by one warp: its shared-memory stage (32-bit addresses) or global memory.
template <bool kShared>
struct SrcView {
  unsigned sbase;    // shared address of ROI byte 0
  const uint8_t* g;  // global address of ROI byte 0
  // 8 bytes starting at ROI offset `off` (any alignment): three aligned words, funnel-shifted
  __device__ __forceinline__ void window(int off, unsigned* lo, unsigned* hi) const {
    unsigned w0, w1, w2, sh;
    if constexpr (kShared) {
      const unsigned a = sbase + (unsigned)off, al = a & ~3u;
      sh = a * 8u;  // shf.r.wrap uses the low 5 bits: (a & 3) * 8
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(al));
      asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(al));
      asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(al));
    } else {
      const unsigned long long a = (unsigned long long)(g + off);
      const unsigned* al = reinterpret_cast<const unsigned*>(a & ~3ull);
      sh = (unsigned)a * 8u;
      w0 = __ldg(al);
      w1 = __ldg(al + 1);
      w2 = __ldg(al + 2);
    }
    *lo = __funnelshift_r(w0, w1, sh);
    *hi = __funnelshift_r(w1, w2, sh);
  }
};

// Image rows [dy_lo, dy_hi) of one up-scaled (or mildly down-scaled) ROI for the 8
// columns starting at output column x0.  Per source row a lane loads two 8-byte windows (columns 0-3 and 4-7);
// PRMT picks each tap pair out of its window and DP2A does S0*a0 + S1*a1.  The vertical pass is done on pixel
// pairs: PRMT gathers the high halves of two 32-bit products, so the fma and alu pipes stay balanced.
template <bool kShared>
__device__ __forceinline__ void image_rows_fast(const SrcView<kShared> S, int w, int h, int nw, int nh, int dy_lo, int dy_hi, int tw,
                                                int left, int border,
                                                int x0, uint8_t* __restrict__ out_rows /* row `top`, column 0 */, bool aligned8,
                                                uint4* rowrec /* 32 records of this warp, shared */, int lane) {
  const int count = min(kCols, tw - x0);
  const double scale_x = __ddiv_rn(1.0, __ddiv_rn((double)nw, (double)w));
  const double scale_y = __ddiv_rn(1.0, __ddiv_rn((double)nh, (double)h));
  unsigned a01[kCols];  // a0 | a1 << 16 (both in [0, 2048])
  unsigned sel[kCols];  // PRMT selector of the tap pair within the window of the column's group
  int smin[2];
  unsigned keep[2] = {0u, 0u};  // byte mask of the image columns of each 4-column group (the others are left / right border)
#pragma unroll
  for (int i = 0; i < kCols; ++i) {
    int s;
    float f;
    const int dx = x0 + i - left;
    if (dx >= 0 && dx < nw) keep[i >> 2] |= 0xffu << (8 * (i & 3));
    src_coord(min(max(dx, 0), nw - 1), scale_x, &s, &f);  // border columns repeat the nearest image column (then masked)
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= w - 1) { s = w - 1; f = 0.f; }
    const unsigned a1 = (unsigned)coef(f);
    a01[i] = (unsigned)coef(__fsub_rn(1.0f, f)) | (a1 << 16);
    if ((i & 3) == 0) smin[i >> 2] = s;
    const unsigned k0 = (unsigned)(s - smin[i >> 2]);
    sel[i] = k0 | ((k0 + (a1 ? 1u : 0u)) << 4);  // a zero weight never reads its tap (cv2 clamps it instead)
  }
  int H0[kCols], H1[kCols];
  auto hrow = [&](int roff, int* H) {
    unsigned lo, hi;
    S.window(roff + smin[0], &lo, &hi);
#pragma unroll
    for (int i = 0; i < 4; ++i) H[i] = (int)(__dp2a_lo(a01[i], __byte_perm(lo, hi, sel[i]), 0u) >> 4);
    S.window(roff + smin[1], &lo, &hi);
#pragma unroll
    for (int i = 4; i < 8; ++i) H[i] = (int)(__dp2a_lo(a01[i], __byte_perm(lo, hi, sel[i]), 0u) >> 4);
  };
  uint8_t* dst = out_rows + x0;
  const unsigned bword = (unsigned)border * 0x01010101u;
  const bool full = aligned8 && count == kCols;
  int py0 = -1, py1 = -1;  // source rows of the previous output row (warp-uniform)
  for (int base = dy_lo; base < dy_hi; base += 32) {
    // lane L prepares output row base+L: vertical weights, source rows, and what has to be (re)loaded; the warp then
    // reads one row record per output row from its shared scratch (a broadcast load, cheaper than shuffles)
    const RowCoef mine = row_coef(base + lane, scale_y);
    const int y0 = min(max(mine.sy, 0), h - 1), y1 = min(max(mine.sy + 1, 0), h - 1);
    int q0 = __shfl_up_sync(0xffffffffu, y0, 1), q1 = __shfl_up_sync(0xffffffffu, y1, 1);
    if (lane == 0) { q0 = py0; q1 = py1; }
    const unsigned code = (y0 == q0 && y1 == q1) ? 0u : (y0 == q1 ? 1u : 2u);  // 0 keep, 1 shift + load y1, 2 load both
    py0 = __shfl_sync(0xffffffffu, y0, 31);
    py1 = __shfl_sync(0xffffffffu, y1, 31);
    __syncwarp();
    rowrec[lane] = make_uint4((unsigned)(y1 * w), (unsigned)mine.b01, code, (unsigned)(y0 * w));
    __syncwarp();
    const int rows = min(32, dy_hi - base);
    for (int j = 0; j < rows; ++j) {
      const uint4 rec = rowrec[j];
      if (rec.z) {
        if (rec.z == 1u) {
#pragma unroll
          for (int i = 0; i < kCols; ++i) H0[i] = H1[i];
        } else {
          hrow((int)rec.w, H0);
        }
        hrow((int)rec.x, H1);
      }
      const unsigned b0 = rec.y & 0xffffu, b1 = rec.y >> 16;
      unsigned wd[2];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        unsigned r6[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int i = g * 4 + k * 2;
          // high halves of the four products of a pixel pair: (b*H) >> 16 for two pixels per PRMT
          const unsigned P = __byte_perm(b0 * (unsigned)H0[i], b0 * (unsigned)H0[i + 1], 0x7632);
          const unsigned Q = __byte_perm(b1 * (unsigned)H1[i], b1 * (unsigned)H1[i + 1], 0x7632);
          // ((P + Q + 2) >> 2) of both halves, left in bytes 1 and 3: ((P + Q) << 6) + (2 << 6)
          r6[k] = (P + Q) * 64u + 0x00800080u;
        }
        wd[g] = (__byte_perm(r6[0], r6[1], 0x7531) & keep[g]) | (bword & ~keep[g]);
      }
      uint8_t* d = dst + (long long)(base + j) * tw;
      if (full) {
        *reinterpret_cast<uint2*>(d) = make_uint2(wd[0], wd[1]);
      } else if (count > 0) {  // ragged row end or unaligned rows (T not a multiple of 8)
#pragma unroll
        for (int i = 0; i < kCols; ++i)
          if (i < count) d[i] = (uint8_t)(wd[i >> 2] >> (8 * (i & 3)));
      }
    }
  }
}

// Geometry of ROI n and whether the warp kernel takes it.  The warp kernel resizes what IFCB ROIs almost always are:
// bilinear (not the identity / exact-2x special cases), horizontal scale <= 1.6 so that the taps of 4 columns fit an
// 8-byte window, and small.  Everything else -- ~1 % of real ROIs -- goes to the cluster kernel.
struct RoiGeom {
  int w, h, nw, nh;
  long long start, total;
  bool valid, fast, staged;
  int mis;
};
__device__ __forceinline__ RoiGeom roi_geom(const Params& p, long long n) {
  RoiGeom g;
  g.w = p.w[n];
  g.h = p.h[n];
  g.start = p.start[n];
  g.valid = (g.w >= 1) && (g.h >= 1) && (g.w < 65536) && (g.start >= 0) && (g.start + (long long)g.w * g.h <= p.roi_len);
  g.nh = g.nw = 0;
  if (g.valid) {
    new_dims(g.h, g.w, p.th, p.tw, &g.nh, &g.nw);
    g.valid = (g.nh >= 1) && (g.nw >= 1) && (g.nh <= p.th) && (g.nw <= p.tw);
  }
  if (!g.valid) g.nh = g.nw = 0;
  g.total = g.valid ? (long long)g.w * g.h : 0;
  const uint8_t* src = p.roi + (g.valid ? g.start : 0);
  g.mis = (int)((unsigned long long)src & 15);
  g.staged = g.total > 0 && (g.mis + g.total <= kStage + 16);
  g.fast = !g.valid || ((g.total <= p.big_bytes) && !(g.nw == g.w && g.nh == g.h) && !(g.w == 2 * g.nw && g.h == 2 * g.nh) &&
                        ((long long)g.w * 10 <= (long long)g.nw * 16) && (g.staged || (src + g.total + 12 <= p.roi + p.roi_len)));
  return g;
}

// one thread per ROI: queue what the warp kernel will not take, so that the cluster kernel can run BESIDE it (second stream)
__global__ void preprocess_classify_kernel(Params p, long long n_rois) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_rois) return;
  const RoiGeom g = roi_geom(p, n);
  if (!g.fast) p.big_list[atomicAdd(p.big_count, 1u)] = (int)n;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, 5) preprocess_u8_kernel(Params p, long long n_rois) {
  __shared__ __align__(16) uint8_t stage_s[kWarpsPerCta][kStage + 32];  // + misalignment (<= 15) + window over-read (<= 11)
  __shared__ __align__(16) unsigned hist_s[kWarpsPerCta][4 * 128];  // [sub-histogram][bin pair]: two 16-bit counters per word;
                                                                   // afterwards the 32 row records of the warp

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // `parts` warps share a ROI (each takes a band of its output rows, and repeats the histogram): small launches would
  // otherwise leave most of the GPU idle behind the latency of one warp per ROI
  const int parts = p.slabs;
  const long long item = (long long)blockIdx.x * kWarpsPerCta + warp;
  const long long n = item / parts;
  const int part = (int)(item - n * parts);
  if (n >= n_rois) return;
  const int th = p.th, tw = p.tw;
  const RoiGeom g = roi_geom(p, n);
  if (!g.fast) return;  // queued by preprocess_classify_kernel for the cluster kernel
  const int w = g.w, h = g.h, nw = g.nw, nh = g.nh;
  const bool valid = g.valid, staged = g.staged;
  const long long total = g.total;
  const int mis = g.mis;
  const uint8_t* src = p.roi + (valid ? g.start : 0);
  if (!valid && lane == 0 && part == 0) atomicAdd(p.faults, 1ULL);
  uint8_t* out = (uint8_t*)p.out + n * (long long)th * tw;
  const int top = (th - nh) / 2, left = (tw - nw) / 2;
  const bool has_border = (nh < th) || (nw < tw);
  const bool want_mode = (p.border_mode == SPK_BORDER_MODE) && valid && has_border;
  uint8_t* stage = stage_s[warp];
  unsigned* hist = hist_s[warp];

  // ---- one pass over the ROI bytes: stage + histogram -------------------------------------------
  int border = p.border_mode == SPK_BORDER_WHITE ? 255 : 0;
  if (staged || want_mode) {
    unsigned tot[8];  // this lane's bins lane*8 .. lane*8+7
#pragma unroll
    for (int k = 0; k < 8; ++k) tot[k] = 0;
    unsigned* my = hist + (lane & 3) * 128;
    auto count_byte = [&](unsigned b) { atomicAdd(&my[b >> 1], 1u << ((b & 1) * 16)); };
    auto flush = [&]() {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // word lane*4+k holds bins lane*8+2k, +2k+1
        unsigned c = 0, d = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const unsigned v = hist[s * 128 + lane * 4 + k];
          c += v & 0xffffu;
          d += v >> 16;
          hist[s * 128 + lane * 4 + k] = 0;
        }
        tot[2 * k] += c;
        tot[2 * k + 1] += d;
      }
      __syncwarp();
    };
    if (want_mode) {
      for (int i = lane; i < 4 * 128; i += 32) hist[i] = 0;
      __syncwarp();
    }
    // the aligned 16-byte chunks that cover [src, src + total): chunk c = bytes [c*16 - mis, c*16 - mis + 16) of the ROI
    const uint4* base = reinterpret_cast<const uint4*>(src - mis);
    const long long nchunk = (mis + total + 15) >> 4;
    const uint8_t* buf_lo = p.roi;
    const uint8_t* buf_hi = p.roi + p.roi_len;
    long long seg_left = kHistSeg;
    for (long long c0 = 0; c0 < nchunk; c0 += 32) {
      const long long c = c0 + lane;
      if (c < nchunk) {
        const uint8_t* cp = reinterpret_cast<const uint8_t*>(base + c);
        const long long lo = c * 16 - mis;  // ROI offset of the chunk's first byte
        unsigned wds[4];
        const bool inside = (cp >= buf_lo) && (cp + 16 <= buf_hi);
        if (inside) {
          const uint4 q = __ldg(base + c);
          wds[0] = q.x; wds[1] = q.y; wds[2] = q.z; wds[3] = q.w;
        } else {  // first / last chunk of the whole .roi buffer: only the bytes that exist
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            unsigned v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const uint8_t* bp = cp + j * 4 + b;
              if (bp >= buf_lo && bp < buf_hi) v |= (unsigned)(*bp) << (8 * b);
            }
            wds[j] = v;
          }
        }
        if (staged) *reinterpret_cast<uint4*>(stage + c * 16) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        if (want_mode) {
          if (lo >= 0 && lo + 16 <= total) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              count_byte(wds[j] & 255u);
              count_byte((wds[j] >> 8) & 255u);
              count_byte((wds[j] >> 16) & 255u);
              count_byte(wds[j] >> 24);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const long long o = lo + j;
              if (o >= 0 && o < total) count_byte((wds[j >> 2] >> (8 * (j & 3))) & 255u);
            }
          }
        }
      }
      if (want_mode) {
        seg_left -= 32 * 16;
        if (seg_left <= 0) {
          flush();
          seg_left = kHistSeg;
        }
      }
    }
    if (want_mode) {
      flush();
      // argmax, lowest bin wins ties: key = count << 8 | (255 - bin)
      unsigned long long key = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const unsigned long long cand = ((unsigned long long)tot[k] << 8) | (unsigned)(255 - (lane * 8 + k));
        key = cand > key ? cand : key;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
      }
      border = 255 - (int)(key & 255u);
    }
    __syncwarp();
  }

  // ---- border rows above and below the image (this warp's band of the th output rows) -----------------
  const int band_lo = (int)((long long)th * part / parts), band_hi = (int)((long long)th * (part + 1) / parts);
  {
    const int a0 = band_lo, a1 = min(band_hi, top);            // rows above the image
    const int b0 = max(band_lo, top + nh), b1 = band_hi;       // rows below
    if (a1 > a0) warp_fill(out + (long long)a0 * tw, (long long)(a1 - a0) * tw, border, lane);
    if (b1 > b0) warp_fill(out + (long long)b0 * tw, (long long)(b1 - b0) * tw, border, lane);
  }
  const int dy_lo = max(band_lo - top, 0), dy_hi = min(band_hi - top, nh);  // image rows of the band
  if (dy_hi <= dy_lo) return;

  uint8_t* rows = out + (long long)top * tw;
  const bool aligned8 = ((tw & 7) == 0) && (((unsigned long long)p.out & 7) == 0);
  for (int x0 = lane * kCols; x0 < ((tw + 255) & ~255); x0 += 32 * kCols) {
    if (staged) {
      SrcView<true> S{(unsigned)__cvta_generic_to_shared(stage + mis), nullptr};
      image_rows_fast<true>(S, w, h, nw, nh, dy_lo, dy_hi, tw, left, border, x0, rows, aligned8, reinterpret_cast<uint4*>(hist), lane);
    } else {
      SrcView<false> S{0u, src};
      image_rows_fast<false>(S, w, h, nw, nh, dy_lo, dy_hi, tw, left, border, x0, rows, aligned8, reinterpret_cast<uint4*>(hist), lane);
    }
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
                              int border_mode, int channels, int out_dtype, int out_layout, const float* lut, void* out) try {
  if (!ctx) return fail(nullptr, SPK_ERR_INVALID, "spk_preprocess: null context");
  if (n == 0) return SPK_OK;
  SPK_CUDA_OK(ctx, cudaSetDevice(ctx->device));
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
  p.big_list = nullptr;
  p.big_count = nullptr;
  p.big_bytes = 0;
  const double out_bytes = (double)n * target_h * target_w * channels * (out_dtype == SPK_DTYPE_F32 ? 4 : out_dtype == SPK_DTYPE_BF16 ? 2 : 1);
  // input bytes are data dependent (sum of w*h); the caller adds them -- recorded here: descriptors + output
  ProfScope prof(ctx, SPK_PROF_PREPROCESS, 0.0, out_bytes + 16.0 * n, "preprocess T=%dx%d c=%d dtype=%d n=%lld", target_h, target_w,
                 channels, out_dtype, (long long)n);
  const long long blocks = n * p.slabs;
  if (blocks > 0x7fffffffLL) return fail(ctx, SPK_ERR_UNSUPPORTED, "spk_preprocess: batch too large");
  if (out_dtype == SPK_DTYPE_U8) {
    if (n > 0x7fffffffLL) return fail(ctx, SPK_ERR_UNSUPPORTED, "spk_preprocess: batch too large");
    if (ctx->big_cap < n) {
      if (ctx->d_big_list) cudaFree(ctx->d_big_list);
      ctx->d_big_list = nullptr;
      ctx->big_cap = 0;
      const long long cap = n < 4096 ? 4096 : n + n / 2;
      SPK_CUDA_OK(ctx, cudaMalloc(&ctx->d_big_list, cap * sizeof(int)));
      ctx->big_cap = cap;
    }
    if (!ctx->d_big_count) SPK_CUDA_OK(ctx, cudaMalloc(&ctx->d_big_count, sizeof(unsigned)));
    SPK_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_big_count, 0, sizeof(unsigned), ctx->stream));
    if (!ctx->stream2) {
      SPK_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
      SPK_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
      SPK_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    p.big_list = ctx->d_big_list;
    p.big_count = ctx->d_big_count;
    p.big_bytes = kBigBytes;
    preprocess_classify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, (long long)n);
    SPK_LAUNCH_CHECK(ctx);
    // the cluster kernel (heavy tail) runs on a second stream beside the warp kernel
    SPK_CUDA_OK(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    SPK_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    Params pb = p;
    pb.slabs = kBigSlabs;
    preprocess_big_kernel<<<kBigClusters * kBigSlabs, kThreads, kHBytes, ctx->stream2>>>(pb);
    const cudaError_t e_big = cudaGetLastError();
    ctx->launches++;
    // whatever happened on the side stream, the main stream joins it again before this call returns
    const cudaError_t e_rec = cudaEventRecord(ctx->ev_join, ctx->stream2);
    p.slabs = n <= 2048 ? 4 : n <= 8192 ? 2 : 1;  // warps per ROI (fills the partial last wave of a bin-sized launch)
    const long long items = n * p.slabs;
    preprocess_u8_kernel<<<(unsigned)((items + kWarpsPerCta - 1) / kWarpsPerCta), kWarpsPerCta * 32, 0, ctx->stream>>>(p, (long long)n);
    const cudaError_t e_join = e_rec == cudaSuccess ? cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0) : e_rec;
    if (e_big != cudaSuccess) return fail(ctx, SPK_ERR_CUDA, "spk_preprocess: cluster kernel launch: %s", cudaGetErrorString(e_big));
    if (e_join != cudaSuccess) return fail(ctx, SPK_ERR_CUDA, "spk_preprocess: stream join: %s", cudaGetErrorString(e_join));
  } else {
    preprocess_kernel<<<(unsigned)blocks, kThreads, kHBytes, ctx->stream>>>(p);
  }
  SPK_LAUNCH_CHECK(ctx);
  return SPK_OK;
} catch (const std::exception& e) {
  return fail(ctx, SPK_ERR_STATE, "spk_preprocess: %s", e.what());
} catch (...) {
  return fail(ctx, SPK_ERR_STATE, "spk_preprocess: unknown C++ exception");
}

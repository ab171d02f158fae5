// K2 stem: conv 7x7 / stride 2 / pad 3 (1 gray plane -> 64 channels, folded BatchNorm) + ReLU +
// max-pool 3x3 / stride 2 / pad 1, fused, on the tcgen05 tensor cores.  u8 image in, bf16 NHWC out.
//
// Replaces conv1 / bn1 / relu / maxpool of torchvision's ResNet (and conv0 / norm0 / relu0 / pool0
// of DenseNet) as run by TorchVisionNet.forward (sykepic/train/network.py:66-68) on the three
// identical planes cv2.imread produces (sykepic/train/data.py:217-219); the planes are folded into
// one (w' = sum_c w[:, c]) and ToTensor's 1/255 is folded into the weights, so the A operand is the
// exact integer pixel value in bf16.
//
// One CTA = one image x one strip of 4 pooled rows (9 conv rows), two CTAs per SM.  The kernel is bound by
// instruction issue (ncu: 65 % issue-active), so the epilogue does one FMNMX per accumulator element.
// GEMM per conv output row: M = 128 (the row's <= 128 output columns), N = 64, K = 2 x 64 (k = 8 r + s;
// s = 7 and r = 7 carry zero weights; the operand is multiplied by the bf16 hi part and then by the bf16
// lo part of the weight, w = hi + lo to 2^-17, into one accumulator -- the stem keeps fp32-weight
// accuracy, which matters: bf16 stem weights alone raise the final probability error 1.5-3x).
//
// The A operand is never materialised per conv row.  For every INPUT row y the builder warps write
//   E_y[j][0..8) = bf16(x[y][2j-3 .. 2j+5))        j = 0..127, 16 bytes each, 2 KB per input row,
// once; conv row i then reads E_{2i-3} .. E_{2i+4} as the eight 16-byte K chunks of its operand through
// a NO-SWIZZLE ("interleaved") K-major descriptor: rows 16 bytes apart, 8-row groups 128 bytes apart
// (SBO), the two K chunks of an MMA one input row = 2048 bytes apart (LBO).  Each E row serves 3.5 conv
// rows, so the u8 -> bf16 expansion costs 2 input rows per conv row instead of 7.
//   warps 0-3  stage the u8 strip, build E rows (2 per step), signal one mbarrier per conv row;
//   warp 4     allocates TMEM, issues tcgen05.mma (one thread), commits to mbarriers;
//   warps 5-12 epilogue (two per TMEM lane quarter, 32 channels each): tcgen05.ld, running vertical max
//              of the 3 conv rows of a pooled row on the raw fp32 sums (max commutes with the monotone
//              bias / ReLU / rounding, applied once per pooled row), horizontal max through shared
//              memory, coalesced 16-byte stores.
// The 112x112x64 conv output (411 MB per 256 images in bf16) never exists in HBM.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include <cstdlib>
#include <cstring>

#include "spk_internal.h"
#include "tc_common.cuh"

namespace spk {
namespace {
using namespace tc;

constexpr int kBuilders = 128;
constexpr int kEpiWarps = 8;            // two per TMEM lane quarter, 32 of the 64 channels each
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 160 + kEpiThreads;  // 4 builder warps + 1 MMA warp + 8 epilogue warps
constexpr int kSlots = 4;               // TMEM accumulator ring: 4 x 64 columns (two CTAs share the SM's 512)
constexpr int kPoolRowsPerStrip = 4;
constexpr int kMaxConvRows = 2 * kPoolRowsPerStrip + 1;
constexpr int kSteps = kMaxConvRows + 3;  // builder steps (2 input rows each)
constexpr int kERows = 2 * kSteps;        // E rows of a strip (the last one only meets zero weights)
constexpr int kERowBytes = 128 * 16;      // one E row: 128 output columns x 8 bf16
constexpr int kEBytes = kERows * kERowBytes;
constexpr int kEGroups = kERows / 4;      // the builders signal every 4 E rows
constexpr int kBTile = 8 * 64 * 16;       // one weight tile: 8 K chunks x 64 rows x 8 bf16
constexpr int kBBytes = 2 * kBTile;       // bf16 hi tile + bf16 lo tile
constexpr int kPoolBytes = 128 * 128;
constexpr int kMaxT = 256;
constexpr int kXOff = 16;       // column of pixel x = 0 in a strip row
constexpr int kPitchPad = 32;   // kXOff + the builders' over-read past the last pixel

struct StemParams {
  const uint8_t* x;      // [n, th, tw] u8
  const uint4* w_il;     // 16 KB: the weight tile in the interleaved (no-swizzle) K-major layout
  const float* bias;     // [64]
  __nv_bfloat16* y;      // [n, hp, wp, ldy]
  int n, th, tw, hc, wc, hp, wp, ldy;
  int strips, pitch;
  int groups, rb_pitch;  // 16-pixel groups per strip row that are converted to bf16; byte pitch of the bf16 row buffer
  long long* trace;      // debug: clock64 timestamps of a few CTAs (SPK_STEM_TRACE=1), else nullptr
  int use_tma;           // strip staged by one TMA box (tw % 16 == 0, tw <= 224); else by the builder threads
  unsigned long long* faults;  // kSplit: counts CTAs with an output beyond the fp16 range of the SplitF format
};

// no-swizzle K-major descriptor: rows 16 B apart inside an 8-row core matrix, SBO between core matrices
// along M / N, LBO between the two 16-byte K chunks of one MMA
__device__ __forceinline__ uint64_t smem_desc_interleaved(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
constexpr uint32_t kIdesc = idesc_bf16(128, 64);
constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // fp16 A and B, fp32 D

// four bytes -> two packed bf16x2 of their integer values, exact.
// 0x4B000000 | v is the float 2^23 + v; subtracting 2^23 gives float(v) without the I2F pipe.
__device__ __forceinline__ uint2 bytes4_to_bf16x4(uint32_t x) {
  const float f0 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7440)) - 8388608.0f;
  const float f1 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7441)) - 8388608.0f;
  const float f2 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7442)) - 8388608.0f;
  const float f3 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7443)) - 8388608.0f;
  __nv_bfloat162 a = __floats2bfloat162_rn(f0, f1), b = __floats2bfloat162_rn(f2, f3);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  return r;
}

// four bytes -> two packed fp16x2 of their integer values, exact: 0x6400 | v is the half 1024 + v.
__device__ __forceinline__ uint2 bytes4_to_f16x4(uint32_t x) {
  const uint32_t k1024 = 0x64006400u;
  uint32_t a = __byte_perm(x, k1024, 0x7170), b = __byte_perm(x, k1024, 0x7372);
  __half2 ha = __hsub2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<const __half2*>(&k1024));
  __half2 hb = __hsub2(*reinterpret_cast<__half2*>(&b), *reinterpret_cast<const __half2*>(&k1024));
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&ha);
  r.y = *reinterpret_cast<uint32_t*>(&hb);
  return r;
}

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// per-CTA clock trace, compiled in only with -DSPK_STEM_TRACE_BUILD (the modulo and the clock reads cost ~40 of the ~240
// instructions an epilogue warp spends per conv row)
#ifdef SPK_STEM_TRACE_BUILD
#define STEM_TRACE(slot)                                                                         \
  do {                                                                                           \
    if (p.trace && (blockIdx.x % 1000) == 500) p.trace[(blockIdx.x / 1000) * 64 + (slot)] = clock64(); \
  } while (0)
#else
#define STEM_TRACE(slot) do { } while (0)
#endif

// kSplit (with kHalf): the output is SplitF words (fp16 hi | lo, FP32_TC precision) instead of bf16: the accumulators carry
// fp32 accuracy (exact fp16 pixels x per-channel scaled fp16 hi + lo weights, TWO passes); the pool buffers double (one
// CTA per SM then).
// kHalf: pixels and weights in fp16; for bf16 output ONE pass over K = 64.  The weights of channel c are scaled by a
// power of two into the top of the fp16 range (exact) and the sums scaled back in the epilogue's FFMA; an fp16 weight
// is good to 2^-12 relative, eight times finer than the bf16 rounding of the output, so the second (lo) pass buys
// nothing there -- and it was half of the kernel's MMA shared-memory traffic (48 KB per conv row).
template <bool kSplit, bool kHalf>
__global__ void __launch_bounds__(kThreads, kSplit ? 1 : 2) stem_pool_kernel(const __grid_constant__ CUtensorMap map_x, const StemParams p) {
  constexpr int kPoolBuf = kSplit ? 2 * kPoolBytes : kPoolBytes;  // one pooled-row buffer: 128 columns x 64 channels
  constexpr int kPoolRow = kSplit ? 256 : 128;                    // bytes per column
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - raw);
  // layout: E | B | pool[2] | image strip | bias | barriers | tmem slot
  const uint32_t e_off = 0, b_off = kEBytes, pool_off = b_off + kBBytes, img_off = pool_off + 2 * kPoolBuf;
  const int img_bytes = kERows * p.pitch;
  const uint32_t bias_off = (img_off + img_bytes + 15u) & ~15u;
  const uint32_t bar_off = bias_off + 128 * 4;  // bias[64], scale[64]
  auto e_ready = [&](int g) { return base + bar_off + 8u * g; };  // E rows [4g, 4g + 4) are built
  auto t_full = [&](int s) { return base + bar_off + 8u * (kEGroups + s); };
  auto t_empty = [&](int s) { return base + bar_off + 8u * (kEGroups + kSlots + s); };
  const uint32_t load_bar = base + bar_off + 8u * (kEGroups + 2 * kSlots);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + bar_off + 8 * (kEGroups + 1 + 2 * kSlots));
  // every E row is built: the pool buffers (which alias the row buffer the builders read) may be written
  const uint32_t built_bar = base + bar_off + 8u * (kEGroups + 2 + 2 * kSlots);
  unsigned char* rowbuf = gbase + pool_off;  // bf16 copy of the strip; aliases the pool buffers, which are idle until the first MMA is done
  unsigned char* img = gbase + img_off;
  float* bias_sm = reinterpret_cast<float*>(gbase + bias_off);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  pdl_trigger();
  pdl_wait();  // the strip is the previous kernel's output
  const int image = blockIdx.x / p.strips;
  const int strip = blockIdx.x - image * p.strips;
  const int p0 = strip * kPoolRowsPerStrip;
  const int p1 = min(p0 + kPoolRowsPerStrip, p.hp) - 1;  // last pooled row of the strip
  const int c_lo = max(0, 2 * p0 - 1), c_hi = min(p.hc - 1, 2 * p1 + 1);  // conv rows needed
  const int y_base = 2 * c_lo - 3;                                         // input row of E row 0 (may be negative)
  const int n_rows = c_hi - c_lo + 1;

  if (tid == 0) {
    STEM_TRACE(0);
    if (p.trace && (blockIdx.x % 1000) == 500) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[(blockIdx.x / 1000) * 64 + 6] = smid;
    }
  }
  if (tid < 64) {
    bias_sm[tid] = __ldg(p.bias + tid);
    bias_sm[64 + tid] = __ldg(reinterpret_cast<const float*>(p.w_il) + kBBytes / 4 + tid);  // 1 / weight scale of the channel
  }
  if (warp == 4) {
    if (lane == 0) {
      for (int g = 0; g < kEGroups; ++g) mbar_init(e_ready(g), 4);  // one arrival per builder warp
      for (int s = 0; s < kSlots; ++s) {
        mbar_init(t_full(s), 1);
        mbar_init(t_empty(s), kEpiWarps);
      }
      mbar_init(load_bar, 1);
      mbar_init(built_bar, 4);  // one arrival per builder warp
      mbar_init_fence();
      // weight tile (16 KB, bulk copy) and, when the geometry allows, the u8 strip as ONE TMA box: pixel x of
      // input row y lands at img[(y - y_base) * pitch + x + kXOff]; out-of-image rows / columns are zero-filled
      constexpr uint32_t kWBytes = (kHalf && !kSplit) ? kBTile : kBBytes;  // (split: fp16 hi AND lo tiles)
      mbar_expect_tx(load_bar, kWBytes + (p.use_tma ? (uint32_t)(kERows * p.pitch) : 0u));
      bulk_load(base + b_off, p.w_il, kWBytes, load_bar);
      // (the box must start on a 16-byte boundary of the row: x = -16, not -3)
      if (p.use_tma) tma_load_3d(base + img_off, &map_x, load_bar, -kXOff, y_base, image);
    }
    __syncwarp();
    tmem_alloc(smem_u32((const void*)tmem_slot), kSlots * 64);
  }
  if (!p.use_tma) {
    // manual staging (row pitch not a multiple of 16 bytes, or T > 240): zero the strip, then copy the rows
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4* i4 = reinterpret_cast<uint4*>(img);  // img_off is 16-byte aligned
    for (int i = tid; i < (img_bytes + 15) / 16; i += kThreads) i4[i] = z;
    __syncthreads();
    const uint8_t* src = p.x + (size_t)image * p.th * p.tw;
    const int rows = 2 * (n_rows + 3);
    for (int e = tid; e < rows * p.tw; e += kThreads) {
      const int rr = e / p.tw, c = e - rr * p.tw;
      const int gr = y_base + rr;
      if (gr < 0 || gr >= p.th) continue;
      img[rr * p.pitch + kXOff + c] = __ldg(src + (size_t)gr * p.tw + c);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) STEM_TRACE(1);

  // ---- u8 -> bf16, ONCE per pixel (all threads): element e of a row-buffer row is pixel x = e - 3, so that the
  // operand chunk of output column j (pixels 2j-3 .. 2j+4) starts at element 2j = byte 4j
  mbar_wait(load_bar, 0);
  if (tid == 0) STEM_TRACE(2);
  for (int task = tid; task < kERows * p.groups; task += kThreads) {
    const int rr = task / p.groups, g = task - rr * p.groups;
    // pixels 16g-3 .. 16g+12 = strip bytes [16g + 13, 16g + 29): two aligned 16-byte loads, shifted by one byte
    const uint4* src = reinterpret_cast<const uint4*>(img + rr * p.pitch + 16 * g);
    const uint4 lo4 = src[0], hi4 = src[1];
    const uint32_t u0 = __funnelshift_r(lo4.w, hi4.x, 8), u1 = __funnelshift_r(hi4.x, hi4.y, 8);
    const uint32_t u2 = __funnelshift_r(hi4.y, hi4.z, 8), u3 = __funnelshift_r(hi4.z, hi4.w, 8);
    uint2 c0, c1, c2, c3;
    if constexpr (kHalf) {
      c0 = bytes4_to_f16x4(u0), c1 = bytes4_to_f16x4(u1), c2 = bytes4_to_f16x4(u2), c3 = bytes4_to_f16x4(u3);
    } else {
      c0 = bytes4_to_bf16x4(u0), c1 = bytes4_to_bf16x4(u1), c2 = bytes4_to_bf16x4(u2), c3 = bytes4_to_bf16x4(u3);
    }
    uint4* dst = reinterpret_cast<uint4*>(rowbuf + rr * p.rb_pitch + 32 * g);
    dst[0] = make_uint4(c0.x, c0.y, c1.x, c1.y);
    dst[1] = make_uint4(c2.x, c2.y, c3.x, c3.y);
  }
  __syncthreads();
  if (tid == 0) STEM_TRACE(3);

  if (warp < 4) {
    // ===== E builders: pure copies.  E[j] = bytes [4j, 4j + 16) of the row buffer; consecutive lanes take
    // consecutive columns, so both the 4-byte loads and the 16-byte stores of a warp are contiguous
    // (a thread-owns-4-columns mapping was measured: 16-way bank conflicts on the stores, 3200 cycles) =====
    for (int yy = warp; yy < kERows; yy += 4) {
      const unsigned char* r = rowbuf + yy * p.rb_pitch;
      unsigned char* e = gbase + e_off + yy * kERowBytes;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = lane + 32 * c;
        if (j < p.wc) {
          const uint32_t* w = reinterpret_cast<const uint32_t*>(r + 4 * j);
          *reinterpret_cast<uint4*>(e + j * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(e_ready(yy >> 2));
    }
    if (lane == 0) mbar_arrive(built_bar);
    if (tid == 0) STEM_TRACE(4);
  } else if (warp == 4) {
    // ===== MMA issuer: the whole warp runs the loop with warp-uniform values, one elected lane issues =====
    {
      const uint32_t b_s = base + b_off;
      int groups_seen = 0;
      for (int idx = 0; idx < n_rows; ++idx) {
        const int slot = idx % kSlots;
        for (; groups_seen <= (2 * idx + 7) >> 2; ++groups_seen) mbar_wait(e_ready(groups_seen), 0);  // E rows 2idx .. 2idx+7
        mbar_wait(t_empty(slot), (((uint32_t)(idx / kSlots)) & 1u) ^ 1u);
        tc_fence_after();
        if (lane == 0) STEM_TRACE(8 + idx);
        const uint32_t a_s = base + e_off + (uint32_t)(2 * idx) * kERowBytes;  // E row of filter row 0
        const uint32_t d = tmem_base + (uint32_t)(slot * 64);
        // K = 128: the eight 16-byte K chunks (filter rows) against the hi weights, then again against the lo weights
        // (kHalf: one pass, fp16 operands)
#pragma unroll
        for (int pass = 0; pass < ((kHalf && !kSplit) ? 1 : 2); ++pass)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_w(d, smem_desc_interleaved(a_s + 2u * k * kERowBytes, kERowBytes, 128),
                   smem_desc_interleaved(b_s + (uint32_t)pass * kBTile + 2u * k * 1024u, 1024, 128), kHalf ? kIdescF16 : kIdesc,
                   (pass | k) != 0 ? 1u : 0u);
        tc_commit_w(t_full(slot));
      }
    }
  } else {
    // ===== epilogue: warp w reads TMEM lanes [32 * (w % 4), +32) and channels [32 * half, +32) =====
    const int q = warp & 3;
    const int half = (warp - 5) >> 2;
    const int wo = q * 32 + lane;    // conv output column == TMEM lane
    const int et = tid - 160;        // 0..255 among the epilogue threads
    float acc[32];                   // running vertical max of the raw conv sums (bias / ReLU / rounding are monotone:
                                     // they are applied once per pooled row, after the max)
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = -INFINITY;
    int emitted = 0;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
    // A pooled row p is the maximum over conv rows 2p-1, 2p, 2p+1.  Even conv row 2p: acc = max(row 2p-1, row 2p), where row
    // 2p-1 (the row that closed the previous window) is READ AGAIN from its accumulator slot, which is only released now;
    // odd row 2p+1: acc = max(acc, row 2p+1), emit.  (Restarting the running maximum from registers instead cost a
    // predicated move and a PLOP3 per element and row: 4 k of the 27 k instructions of a strip.)
    bool ovf = false;
    for (int idx = 0; idx < n_rows; ++idx) {
      const int i = c_lo + idx, slot = idx % kSlots;
      const bool odd = (i & 1) != 0;
      if (odd && idx == 0) continue;  // the odd row a strip starts on only seeds the first window: read with the next row
      __syncwarp();  // tcgen05.ld below is warp-collective
      mbar_wait(t_full(slot), ((uint32_t)(idx / kSlots)) & 1u);
      const bool reread = !odd && idx > 0;  // the previous (odd) row's slot is still ours
      if (reread && idx == 1) mbar_wait(t_full((idx - 1) % kSlots), 0);  // (row 0 of the strip was never waited for)
      tc_fence_after();
      if (tid == 160) STEM_TRACE(24 + idx);
      const bool last_of_window = odd || (i == p.hc - 1);
      const int prow = i >> 1;
      const bool emit = last_of_window && prow >= p0 && prow <= p1;
      unsigned char* pool = gbase + pool_off + (emitted & 1) * kPoolBuf;
      // The pool buffers alias the bf16 row buffer: wait until the builders have read all of it.  (With the MMA warp
      // issuing from uniform registers the first pooled row is ready before the builders finish; without this wait
      // the last conv rows of a strip were built from overwritten pixels.)
      if (emit && emitted == 0) mbar_wait(built_bar, 0);
      const int pslot = (idx + kSlots - 1) % kSlots;
#pragma unroll
      for (int qc = 0; qc < 2; ++qc) {  // 16 channels at a time: acc[32] + v[32] would not fit the 72-register budget
        uint32_t v[16];
        tmem_ld16(taddr0 + (uint32_t)(slot * 64 + qc * 16), v);
        if (odd) {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[qc * 16 + j] = fmaxf(acc[qc * 16 + j], __uint_as_float(v[j]));
        } else if (reread) {
          uint32_t u[16];
          tmem_ld16(taddr0 + (uint32_t)(pslot * 64 + qc * 16), u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[qc * 16 + j] = fmaxf(__uint_as_float(u[j]), __uint_as_float(v[j]));
        } else {  // the first conv row of the image: the row above is padding
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[qc * 16 + j] = __uint_as_float(v[j]);
        }
        if (qc == 1) {  // accumulator slots are free as soon as they have been read: an even row's at once (together with
                        // the odd row before it), an odd row's when the next row has re-read it -- or now, if it is the last
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (!odd || idx == n_rows - 1) mbar_arrive(t_empty(slot));
            if (reread) mbar_arrive(t_empty(pslot));
          }
        }
        if (emit) {
          // + bias, ReLU, bf16; row wo of the pool buffer in 16-byte chunks swizzled by (wo & 7)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if constexpr (kSplit) {  // 8 channels = two 16-byte chunks of SplitF words
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int e = qc * 16 + c * 8 + h2 * 4 + j;
                  const float v = fmaxf(fmaf(acc[e], bias_sm[64 + half * 32 + e], bias_sm[half * 32 + e]), 0.f);  // x 2^-k(channel): exact
                  ovf |= (wo < p.wc) && !(v <= kSplitMax);  // (columns beyond the image hold junk)
                  o[j] = split_store(v);
                }
                const int chunk = 8 * half + 4 * qc + 2 * c + h2;  // 16 chunks of 4 channels per column
                *reinterpret_cast<uint4*>(pool + wo * kPoolRow + ((chunk ^ (wo & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
              }
            } else {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = qc * 16 + c * 8 + 2 * j;
              const float2 b2 = *reinterpret_cast<const float2*>(bias_sm + half * 32 + e);
              __nv_bfloat162 t;
              if constexpr (kHalf) {
                const float2 s2 = *reinterpret_cast<const float2*>(bias_sm + 64 + half * 32 + e);
                t = __floats2bfloat162_rn(fmaxf(fmaf(acc[e], s2.x, b2.x), 0.f), fmaxf(fmaf(acc[e + 1], s2.y, b2.y), 0.f));
              } else {
                t = __floats2bfloat162_rn(fmaxf(acc[e] + b2.x, 0.f), fmaxf(acc[e + 1] + b2.y, 0.f));
              }
              o[j] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(pool + wo * 128 + (((4 * half + 2 * qc + c) ^ (wo & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
      if (tid == 160 && idx == 2) STEM_TRACE(56);
      if (tid == 415 && idx == 2) STEM_TRACE(59);
      if (emit) {
        named_bar_sync(1, kEpiThreads);  // the epilogue warps
        if (tid == 160 && idx == 2) STEM_TRACE(57);
        // horizontal max over conv columns 2pw-1, 2pw, 2pw+1 and a coalesced store of the pooled row
        if constexpr (kSplit) {
          // SplitF words: the maximum of three is one of them -- pick by the decoded value
          uint32_t* yrow = reinterpret_cast<uint32_t*>(p.y) + ((size_t)image * p.hp + prow) * p.wp * p.ldy;
          for (int o = et; o < p.wp * 16; o += kEpiThreads) {
            const int pw = o >> 4, j = o & 15;
            const int w1 = 2 * pw;
            uint4 m4 = *reinterpret_cast<const uint4*>(pool + w1 * kPoolRow + ((j ^ (w1 & 7)) << 4));
            uint32_t* mm = reinterpret_cast<uint32_t*>(&m4);
            auto take = [&](int wn) {
              const uint4 t4 = *reinterpret_cast<const uint4*>(pool + wn * kPoolRow + ((j ^ (wn & 7)) << 4));
              const uint32_t* tt = reinterpret_cast<const uint32_t*>(&t4);
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (split_load(tt[e]) > split_load(mm[e])) mm[e] = tt[e];
            };
            if (w1 - 1 >= 0) take(w1 - 1);
            if (w1 + 1 < p.wc) take(w1 + 1);
            *reinterpret_cast<uint4*>(yrow + (size_t)pw * p.ldy + j * 4) = m4;
          }
          ++emitted;
          continue;
        }
        __nv_bfloat16* yrow = p.y + ((size_t)image * p.hp + prow) * p.wp * p.ldy;
        for (int o = et; o < p.wp * 8; o += kEpiThreads) {
          const int pw = o >> 3, j = o & 7;
          const int w1 = 2 * pw;
          uint4 m4 = *reinterpret_cast<const uint4*>(pool + w1 * 128 + ((j ^ (w1 & 7)) << 4));
          __nv_bfloat162* mm = reinterpret_cast<__nv_bfloat162*>(&m4);
          if (w1 - 1 >= 0) {
            const int w0 = w1 - 1;
            const uint4 t4 = *reinterpret_cast<const uint4*>(pool + w0 * 128 + ((j ^ (w0 & 7)) << 4));
            const __nv_bfloat162* tt = reinterpret_cast<const __nv_bfloat162*>(&t4);
#pragma unroll
            for (int e = 0; e < 4; ++e) mm[e] = __hmax2(mm[e], tt[e]);
          }
          if (w1 + 1 < p.wc) {
            const int w2 = w1 + 1;
            const uint4 t4 = *reinterpret_cast<const uint4*>(pool + w2 * 128 + ((j ^ (w2 & 7)) << 4));
            const __nv_bfloat162* tt = reinterpret_cast<const __nv_bfloat162*>(&t4);
#pragma unroll
            for (int e = 0; e < 4; ++e) mm[e] = __hmax2(mm[e], tt[e]);
          }
          *reinterpret_cast<uint4*>(yrow + (size_t)pw * p.ldy + j * 8) = m4;
        }
        ++emitted;  // the other pool buffer is used next: one barrier per pooled row is enough
      }
      if (tid == 160) STEM_TRACE(40 + idx);
    }
    if constexpr (kSplit) {
      if (__any_sync(0xffffffffu, ovf) && lane == 0) atomicAdd(p.faults, 1ULL);
    }
    (void)ovf;
  }

  tc_fence_before();
  __syncthreads();
  if (tid == 0) STEM_TRACE(5);
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kSlots * 64);
  }
}

}  // namespace

bool stem_pool_supported(const ConvGeom& g, int pool_k, int pool_stride, int pool_pad) {
  if (g.kh != 7 || g.kw != 7 || g.stride != 2 || g.pad != 3 || g.cin != 1 || g.cout != 64 || !g.relu) return false;
  if (pool_k != 3 || pool_stride != 2 || pool_pad != 1) return false;
  if (g.w > kMaxT || g.h < 7 || g.w < 7 || g.wo > 128 || g.wo < 2) return false;
  return true;
}

// w: folded [64][7][7] fp32 (already includes BatchNorm); the 1/255 of ToTensor is folded here.
// Two tiles (bf16 hi part, bf16 lo part of the weight) in the interleaved K-major layout: K chunk kc = filter
// row r (8 of them, the last all zero), row n = output channel, element e = filter column s (e = 7 zero):
// byte offset kc * 1024 + n * 16 + e * 2.
// split = false (bf16 activations): ONE fp16 tile of w * 2^k(o), 2^k(o) the power of two that brings the channel's largest
// weight into [2^13, 2^14); the 64 factors 2^-k(o) follow the tiles (floats at byte kBBytes).
// Which kernel serves bf16 output -- A/B switch SPK_STEM: "t" (default) = the transposed kernel of stem_t.cu, "half" = this
// file's kernel with fp16 operands in one pass, "hilo" = this file's kernel with bf16 hi + lo weights in two passes.
// SplitF output (FP32_TC) always takes the hi + lo kernel.
enum { kStemHilo = 0, kStemHalf = 1, kStemT = 2 };
static int stem_variant(bool split) {
  static const int v = [] {
    const char* e = debug_env("SPK_STEM");
    if (e && !strcmp(e, "hilo")) return (int)kStemHilo;
    if (e && !strcmp(e, "half")) return (int)kStemHalf;
    return (int)kStemT;
  }();
  return split ? (int)kStemHilo : v;
}
static bool stem_hilo_forced() { return stem_variant(false) == kStemHilo; }
int stem_pool_pack_weights(spk_ctx* ctx, const float* w, uint4** d_out, bool split) {
  if (stem_variant(split) == kStemT) return stem_pool_t_pack_weights(ctx, w, d_out);
  std::vector<uint16_t> tile(kBBytes / 2 + 128, 0);
  float* inv_scale = reinterpret_cast<float*>(tile.data() + kBBytes / 2);
  for (int o = 0; o < 64; ++o) inv_scale[o] = 1.f;
  const bool half = !split && !stem_hilo_forced();
  // fp16 weights scaled per channel by a power of two (largest weight into [2^13, 2^14)); `half`: one tile, 2^-12 relative;
  // `split` (FP32_TC): hi + lo tiles, 22 significant bits, both normal fp16 numbers
  for (int o = 0; (half || split) && o < 64; ++o) {
    double amax = 0.0;
    for (int t = 0; t < 49; ++t) amax = std::max(amax, std::fabs((double)w[o * 49 + t] / 255.0));
    int k = 0;
    if (amax > 0.0 && std::isfinite(amax)) {
      int e;
      std::frexp(amax, &e);  // amax = m * 2^e, m in [0.5, 1)
      k = 14 - e;            // amax * 2^k in [2^13, 2^14)
      k = std::max(-60, std::min(60, k));
    }
    inv_scale[o] = (float)std::ldexp(1.0, -k);
    for (int r = 0; r < 7; ++r)
      for (int s = 0; s < 7; ++s) {
        const float v = (float)std::ldexp((double)w[(o * 7 + r) * 7 + s] / 255.0, k);
        const __half hv = __float2half_rn(v);
        uint16_t b;
        memcpy(&b, &hv, 2);
        tile[(size_t)r * 512 + (size_t)o * 8 + s] = b;
        if (split) {
          const __half lv = __float2half_rn(v - __half2float(hv));
          memcpy(&b, &lv, 2);
          tile[(size_t)kBTile / 2 + (size_t)r * 512 + (size_t)o * 8 + s] = b;
        }
      }
  }
  for (int o = 0; !half && !split && o < 64; ++o)
    for (int r = 0; r < 7; ++r)
      for (int s = 0; s < 7; ++s) {
        const float v = (float)((double)w[(o * 7 + r) * 7 + s] / 255.0);
        const __nv_bfloat16 hi = __float2bfloat16(v);
        const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
        uint16_t bh, bl;
        memcpy(&bh, &hi, 2);
        memcpy(&bl, &lo, 2);
        tile[(size_t)r * 512 + (size_t)o * 8 + s] = bh;
        tile[(size_t)kBTile / 2 + (size_t)r * 512 + (size_t)o * 8 + s] = bl;
      }
  SPK_CUDA_OK(ctx, cudaMalloc(d_out, kBBytes + 256));
  SPK_CUDA_OK(ctx, cudaMemcpy(*d_out, tile.data(), kBBytes + 256, cudaMemcpyHostToDevice));
  return SPK_OK;
}

int launch_stem_pool(spk_ctx* ctx, int n, int th, int tw, const uint8_t* x, const uint4* w_il, const float* bias,
                     __nv_bfloat16* y, int hc, int wc, int hp, int wp, int ldy, bool split) {
  if (n <= 0) return SPK_OK;
  if (stem_variant(split) == kStemT) return launch_stem_pool_t(ctx, n, th, tw, x, w_il, bias, y, hc, wc, hp, wp, ldy);
  StemParams p;
  p.x = x;
  p.w_il = w_il;
  p.bias = bias;
  p.y = y;
  p.n = n;
  p.th = th;
  p.tw = tw;
  p.hc = hc;
  p.wc = wc;
  p.hp = hp;
  p.wp = wp;
  p.ldy = ldy;
  p.strips = (hp + kPoolRowsPerStrip - 1) / kPoolRowsPerStrip;
  static const bool no_tma = debug_env("SPK_STEM_NO_TMA") != nullptr;  // A/B switch: stage the strip with the builder threads
  p.use_tma = (!no_tma && tw % 16 == 0 && tw <= 224 && ((uintptr_t)x & 15) == 0 && encode_fn() != nullptr) ? 1 : 0;
  p.groups = (tw + 16 + 15) / 16;                       // row-buffer elements [0, 16 * groups) cover pixels up to tw + 12
  p.rb_pitch = 32 * p.groups + 32;                      // + slack for the dead columns' over-read
  p.pitch = p.use_tma ? 256 : (16 * p.groups + 32);     // strip bytes read: up to 16 * groups + 15
  // the tensor map of the u8 batch {tw, th, n}, box {256, kERows, 1}: cached per (pointer, geometry)
  static thread_local struct { const void* x; int n, th, tw; CUtensorMap map; } cache = {nullptr, 0, 0, 0, {}};
  if (p.use_tma && (cache.x != x || cache.n < n || cache.th != th || cache.tw != tw)) {
    cuuint64_t dims[3] = {(cuuint64_t)tw, (cuuint64_t)th, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)tw, (cuuint64_t)tw * th};
    cuuint32_t box[3] = {256u, (cuuint32_t)kERows, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn()(&cache.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(x), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "stem: cuTensorMapEncodeTiled(u8 batch) failed: %d", (int)r);
    cache.x = x;
    cache.n = n;
    cache.th = th;
    cache.tw = tw;
  }
  const size_t smem = 1024 + kEBytes + kBBytes + 2 * kPoolBytes * (split ? 2 : 1) + (size_t)kERows * p.pitch + 32 + 128 * 4 +
                      8 * (kEGroups + 3 + 2 * kSlots) + 16;
  const bool half = !split && !stem_hilo_forced();  // (must agree with stem_pool_pack_weights)
  if (split)
    SPK_CUDA_OK(ctx, cudaFuncSetAttribute(stem_pool_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (half)
    SPK_CUDA_OK(ctx, cudaFuncSetAttribute(stem_pool_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    SPK_CUDA_OK(ctx, cudaFuncSetAttribute(stem_pool_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static const bool want_trace = debug_env("SPK_STEM_TRACE") != nullptr;
  static long long* d_trace = nullptr;
  static int trace_left = 3;
  p.trace = nullptr;
  p.faults = ctx->d_faults;
  if (want_trace && trace_left > 0) {
    if (!d_trace) cudaMalloc(&d_trace, 16 * 64 * sizeof(long long));
    cudaMemsetAsync(d_trace, 0, 16 * 64 * sizeof(long long), ctx->stream);
    p.trace = d_trace;
  }
  if (split)
    SPK_CUDA_OK(ctx, launch_pdl(stem_pool_kernel<true, true>, dim3((unsigned)(n * p.strips)), dim3(kThreads), smem, ctx->stream, cache.map, p));
  else if (half)
    SPK_CUDA_OK(ctx, launch_pdl(stem_pool_kernel<false, true>, dim3((unsigned)(n * p.strips)), dim3(kThreads), smem, ctx->stream, cache.map, p));
  else
    SPK_CUDA_OK(ctx, launch_pdl(stem_pool_kernel<false, false>, dim3((unsigned)(n * p.strips)), dim3(kThreads), smem, ctx->stream, cache.map, p));
  SPK_LAUNCH_CHECK(ctx);
  if (p.trace) {
    --trace_left;
    std::vector<long long> h(16 * 64);
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    for (int c = 0; c < 4; ++c) {
      const long long* t = &h[(size_t)c * 64];
      if (!t[0]) continue;
      fprintf(stderr, "stem trace cta %d sm %lld: sync1 %lld load %lld convert %lld e_done %lld end %lld\n  mma:", c * 1000 + 500, t[6], t[1] - t[0],
              t[2] - t[0], t[3] - t[0], t[4] - t[0], t[5] - t[0]);
      for (int i = 0; i < 9; ++i) fprintf(stderr, " %lld", t[8 + i] ? t[8 + i] - t[0] : -1);
      fprintf(stderr, "\n  epi wake:");
      for (int i = 0; i < 9; ++i) fprintf(stderr, " %lld", t[24 + i] ? t[24 + i] - t[0] : -1);
      fprintf(stderr, "\n  epi done:");
      for (int i = 0; i < 9; ++i) fprintf(stderr, " %lld", t[40 + i] ? t[40 + i] - t[0] : -1);
      fprintf(stderr, "\n  row2: before barrier %lld (last warp %lld) after barrier %lld\n", t[56] - t[0], t[59] - t[0], t[57] - t[0]);
    }
  }
  return SPK_OK;
}

}  // namespace spk

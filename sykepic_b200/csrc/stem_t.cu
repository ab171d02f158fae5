// K2 stem, transposed form (bf16 output): conv 7x7 / stride 2 / pad 3 (1 gray plane -> 64 channels, folded BatchNorm) + ReLU +
// max-pool 3x3 / stride 2 / pad 1, fused, on the tcgen05 tensor cores.  u8 image in, bf16 NHWC out.
//
// Same operation, inputs and staging as stem.cu (which it replaces for bf16 activations; stem.cu remains the SplitF /
// FP32_TC kernel and the A/B baseline, SPK_STEM=half|hilo).  What changes is which operand is which:
//
//   stem.cu    D[conv column (TMEM lane)][channel (TMEM column)]   = E (pixels) x W^T
//   here       D[channel, conv row parity (lane)][conv column (TMEM column)] = W2 x E^T
//
// stem.cu reads every conv row's accumulator 1.5 times out of TMEM and spends 31 k warp-instructions per strip of 4 pooled
// rows: with conv columns on the lanes, the horizontal half of the 3x3 max needs a trip through shared memory at full conv
// resolution, and bias / ReLU / rounding run on 2 x 64 lanes' worth of elements per pooled one.  With the conv columns of
// one channel in ONE thread's registers the horizontal max is one FMNMX3 per pooled element, bias and scale are per-thread
// scalars, and only pooled-width bf16 rows go through shared memory for the vertical max and the NHWC transpose
// (17 k warp-instructions per strip).
//
// GEMM per accumulator (two conv rows i, i + 1): M = 128 = 64 channels x 2 conv rows, N = the row's conv columns
// rounded up to 16 (112 at T = 224), K = 10 x 8: K chunk q is INPUT row 2i - 3 + q (E row, 8 taps wide, as in
// stem.cu); M rows 0-63 carry filter row q in chunk q (q <= 6), M rows 64-127 carry filter row q - 2 (2 <= q <= 8):
// the second conv row is the same filter two input rows further down.  Pixels and weights are fp16 (exact integers x
// per-channel power-of-two scaled weights, 2^-12 relative; see stem.cu kHalf).
//
//   all warps   convert the strip to fp16 once, then build the E rows (two each)
//   warps 0-3   then: STORE warps -- per pooled row, wait for its three conv rows in the ring (mbarriers per ring row),
//               maximum of the three (16-byte chunks, packed bf16 max), coalesced stores, ring rows handed back
//   warp 4      TMEM allocation, 5 MMAs (K = 16 each) per accumulator, commits
//   warps 5-12  epilogue: TMEM lane quarter q = warp % 4 is (conv row q / 2 of the pair, channels 32 (q % 2) + lane);
//               the two warps of a quarter split the columns.  16 columns at a time (the next tcgen05.ld in flight):
//               max over columns 2pw-1, 2pw, 2pw+1, x scale + bias, ReLU + bf16 (cvt.rn.relu.bf16x2), 2-byte stores
//               into the conv row's slot of a 6-row ring [pooled column][channel]
// Measured (finding 6): 0.127 ms per 256 images against 0.161 ms for stem.cu; latency- and
// occupancy-bound (two CTAs per SM, one instruction per ~12 cycles and warp), the step around it power-capped.
// Replaces conv1 / bn1 / relu / maxpool of torchvision's ResNet as run by TorchVisionNet.forward
// (sykepic/train/network.py:66-68).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "spk_internal.h"
#include "tc_common.cuh"

namespace spk {
namespace {
using namespace tc;

constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 160 + kEpiThreads;  // 4 store warps + 1 MMA warp + 8 epilogue warps
constexpr int kSlots = 2;                    // TMEM accumulator ring: 2 x 128 columns (two CTAs share the SM's 512)
constexpr int kPoolRowsPerStrip = 4;
constexpr int kERows = 26;                   // accumulator t reads E rows 4t .. 4t + 9, t <= 4
constexpr int kEGroups = kERows / 2;         // the builders signal every 2 E rows
constexpr int kWChunks = 10;
constexpr int kWBytes = kWChunks * 128 * 16;  // weight operand: 10 K chunks x 128 rows x 8 fp16
constexpr int kRing = 6;                      // conv rows (pooled width, bf16) in flight between the epilogue and the store warps
constexpr int kXOff = 16;  // column of pixel x = 0 in a strip row

struct StemTParams {
  const uint8_t* x;   // [n, th, tw] u8
  const uint4* w;     // kWBytes of fp16 weights (interleaved K-major), then 64 floats: 1 / scale of the channel
  const float* bias;  // [64]
  __nv_bfloat16* y;   // [n, hp, wp, ldy]
  int n, th, tw, hc, wc, hp, wp, ldy;
  int strips, pitch;
  int groups, rb_pitch;  // 16-pixel groups per strip row converted to fp16; byte pitch of the fp16 row buffer
  int ncols, e_pitch;    // N of the MMA (wc rounded up to 16); bytes per E row (16 per column)
  int ring_pitch, region;  // bytes per ring row (128 per pooled column); bytes of the ring / (row buffer + strip) region
  int rb_bytes;            // bytes of the row buffer (rounded to 128): the strip follows it
  int use_tma;
  unsigned long long* stamp;  // profiling (stamp mode): global-timer slot of this launch, else nullptr
};

__device__ __forceinline__ uint64_t smem_desc_interleaved(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {  // fp16 A and B (K-major), fp32 D
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return __uint_as_float(v);
}
// bf16x2 of (max(lo, 0), max(hi, 0)), round to nearest even
__device__ __forceinline__ uint32_t relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// shared-window addresses (32 bits) rather than generic pointers: with the generic forms the compiler rebuilt the window
// base (S2UR SR_CgaCtaId + uniform-datapath arithmetic) in every 16-column piece
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
  return r;
}

__global__ void __launch_bounds__(kThreads, 2) stem_pool_t_kernel(const __grid_constant__ CUtensorMap map_x, const StemTParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  unsigned char* gbase = smem_raw + (base - raw);
  // layout: E | W | region: the ring, which aliases (fp16 row buffer | u8 image strip) | barriers | tmem slot
  const uint32_t e_off = 0, w_off = (uint32_t)(kERows * p.e_pitch), ring_off = w_off + kWBytes, img_off = ring_off + (uint32_t)p.rb_bytes;
  const int img_bytes = kERows * p.pitch;
  const uint32_t bar_off = ring_off + (uint32_t)p.region;
  auto e_ready = [&](int g) { return base + bar_off + 8u * g; };  // E rows 2g, 2g + 1 are built
  auto t_full = [&](int s) { return base + bar_off + 8u * (kEGroups + s); };
  auto t_empty = [&](int s) { return base + bar_off + 8u * (kEGroups + kSlots + s); };
  const uint32_t load_bar = base + bar_off + 8u * (kEGroups + 2 * kSlots);
  const uint32_t built_bar = base + bar_off + 8u * (kEGroups + 2 * kSlots + 1);  // every E row is built: the ring may be written
  // ring row slot s: written by the 4 epilogue warps of its conv row (ring_full), read by the 4 store warps (ring_empty)
  auto ring_full = [&](int s) { return base + bar_off + 8u * (kEGroups + 2 * kSlots + 2 + s); };
  auto ring_empty = [&](int s) { return base + bar_off + 8u * (kEGroups + 2 * kSlots + 2 + kRing + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + bar_off + 8 * (kEGroups + 2 * kSlots + 2 + 2 * kRing));
  const uint32_t ring_s = base + ring_off;
  auto ring_slot = [](int idx) { return idx >= kRing ? idx - kRing : idx; };  // idx < 2 kRing
  unsigned char* ring = gbase + ring_off;
  unsigned char* rowbuf = ring;  // fp16 copy of the strip; idle once the E rows are built
  unsigned char* img = gbase + img_off;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  pdl_trigger();
  stamp_begin(p.stamp);
  pdl_wait();  // the strip is the previous kernel's output
  const int image = blockIdx.x / p.strips;
  const int strip = blockIdx.x - image * p.strips;
  const int p0 = strip * kPoolRowsPerStrip;
  const int p1 = min(p0 + kPoolRowsPerStrip, p.hp) - 1;  // last pooled row of the strip
  const int c_lo = max(0, 2 * p0 - 1), c_hi = min(p.hc - 1, 2 * p1 + 1);  // conv rows needed
  const int y_base = 2 * c_lo - 3;                                         // input row of E row 0 (may be negative)
  const int n_rows = c_hi - c_lo + 1;
  const int n_acc = (n_rows + 1) >> 1;  // accumulator t = conv rows c_lo + 2t, c_lo + 2t + 1

  if (warp == 4) {
    if (lane == 0) {
      for (int g = 0; g < kEGroups; ++g) mbar_init(e_ready(g), 2);  // one arrival per E row
      for (int s = 0; s < kSlots; ++s) {
        mbar_init(t_full(s), 1);
        mbar_init(t_empty(s), kEpiWarps);
      }
      mbar_init(load_bar, 1);
      mbar_init(built_bar, kThreads / 32);  // one arrival per warp
      for (int s = 0; s < kRing; ++s) {
        mbar_init(ring_full(s), 4);
        mbar_init(ring_empty(s), 4);
      }
      mbar_init_fence();
      // weights (20 KB, bulk copy) and, when the geometry allows, the u8 strip as ONE TMA box: pixel x of input row y
      // lands at img[(y - y_base) * pitch + x + kXOff]; out-of-image rows / columns are zero-filled
      mbar_expect_tx(load_bar, (uint32_t)kWBytes + (p.use_tma ? (uint32_t)(kERows * p.pitch) : 0u));
      bulk_load(base + w_off, p.w, kWBytes, load_bar);
      if (p.use_tma) tma_load_3d(base + img_off, &map_x, load_bar, -kXOff, y_base, image);
    }
    __syncwarp();
    tmem_alloc(smem_u32((const void*)tmem_slot), kSlots * 128);
  }
  if (!p.use_tma) {
    // manual staging (row pitch not a multiple of 16 bytes, or T > 224): zero the strip, then copy the rows
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4* i4 = reinterpret_cast<uint4*>(img);
    for (int i = tid; i < (img_bytes + 15) / 16; i += kThreads) i4[i] = z;
    __syncthreads();
    const uint8_t* src = p.x + (size_t)image * p.th * p.tw;
    for (int e = tid; e < kERows * p.tw; e += kThreads) {
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

  // ---- u8 -> fp16, ONCE per pixel (all threads): element e of a row-buffer row is pixel x = e - 3, so that the operand
  // chunk of conv column j (pixels 2j-3 .. 2j+4) starts at element 2j = byte 4j
  mbar_wait(load_bar, 0);
  for (int task = tid; task < kERows * p.groups; task += kThreads) {
    const int rr = task / p.groups, g = task - rr * p.groups;
    // pixels 16g-3 .. 16g+12 = strip bytes [16g + 13, 16g + 29): two aligned 16-byte loads, shifted by one byte
    const uint4* src = reinterpret_cast<const uint4*>(img + rr * p.pitch + 16 * g);
    const uint4 lo4 = src[0], hi4 = src[1];
    const uint32_t u0 = __funnelshift_r(lo4.w, hi4.x, 8), u1 = __funnelshift_r(hi4.x, hi4.y, 8);
    const uint32_t u2 = __funnelshift_r(hi4.y, hi4.z, 8), u3 = __funnelshift_r(hi4.z, hi4.w, 8);
    const uint2 c0 = bytes4_to_f16x4(u0), c1 = bytes4_to_f16x4(u1), c2 = bytes4_to_f16x4(u2), c3 = bytes4_to_f16x4(u3);
    uint4* dst = reinterpret_cast<uint4*>(rowbuf + rr * p.rb_pitch + 32 * g);
    dst[0] = make_uint4(c0.x, c0.y, c1.x, c1.y);
    dst[1] = make_uint4(c2.x, c2.y, c3.x, c3.y);
  }
  __syncthreads();

  {
    // ===== E rows: pure copies, by ALL 13 warps (two rows each) before they take up their roles: the MMA and epilogue warps
    // have nothing to do until the first ten rows exist, and a single warp gets through a row in ~385 cycles (ncu: one
    // instruction per 12.5 cycles and warp; four builder warps took 2500 cycles).  E[j] = bytes [4j, 4j + 16) of the row
    // buffer; consecutive lanes take consecutive columns, so both the 4-byte loads and the 16-byte stores of a warp are
    // contiguous =====
    for (int yy = warp; yy < kERows; yy += kThreads / 32) {
      const unsigned char* r = rowbuf + yy * p.rb_pitch;
      unsigned char* e = gbase + e_off + yy * p.e_pitch;
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
      if (lane == 0) mbar_arrive(e_ready(yy >> 1));
    }
    if (lane == 0) mbar_arrive(built_bar);
  }
  if (warp < 4) {
    // ===== store warps (the builders' second job): per pooled row, wait for its three conv rows in the ring, take the
    // vertical max (16-byte chunks of packed bf16), store coalesced, hand the ring rows that are no longer needed back.
    // Runs beside the epilogue warps' next accumulator: with both jobs on the epilogue warps, separated by named barriers,
    // an accumulator took 1100 + 1100 cycles =====
    mbar_wait(built_bar, 0);  // (the ring aliases the row buffer; also orders this warp's own ring reads after every builder's reads)
    int released = c_lo;      // conv rows below this one have been handed back
    for (int prow = p0; prow <= p1; ++prow) {
      const int r1 = 2 * prow;
      const bool has0 = r1 - 1 >= 0, has2 = r1 + 1 <= p.hc - 1;
      const int i1 = r1 - c_lo;
      if (has0) mbar_wait(ring_full(ring_slot(i1 - 1)), (uint32_t)((i1 - 1) >= kRing));
      mbar_wait(ring_full(ring_slot(i1)), (uint32_t)(i1 >= kRing));
      if (has2) mbar_wait(ring_full(ring_slot(i1 + 1)), (uint32_t)((i1 + 1) >= kRing));
      const uint32_t b1 = ring_s + (uint32_t)(ring_slot(i1) * p.ring_pitch);
      const uint32_t b0 = has0 ? ring_s + (uint32_t)(ring_slot(i1 - 1) * p.ring_pitch) : b1;  // (a missing row reads the middle one again)
      const uint32_t b2 = has2 ? ring_s + (uint32_t)(ring_slot(i1 + 1) * p.ring_pitch) : b1;
      __nv_bfloat16* yrow = p.y + ((size_t)image * p.hp + prow) * p.wp * p.ldy;
      for (int o = tid; o < p.wp * 8; o += 128) {
        const int pw = o >> 3, j = o & 7;
        const uint32_t off = (uint32_t)o * 16u;  // = pw * 128 + j * 16
        uint4 m4 = lds_v4(b1 + off);
        const uint4 t0 = lds_v4(b0 + off), t2 = lds_v4(b2 + off);
        __nv_bfloat162* mm = reinterpret_cast<__nv_bfloat162*>(&m4);
        const __nv_bfloat162* q0 = reinterpret_cast<const __nv_bfloat162*>(&t0);
        const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&t2);
#pragma unroll
        for (int e = 0; e < 4; ++e) mm[e] = __hmax2(__hmax2(mm[e], q0[e]), q2[e]);
        *reinterpret_cast<uint4*>(yrow + (size_t)pw * p.ldy + j * 8) = m4;
      }
      // conv rows up to 2 prow are done with (2 prow + 1 also feeds the next pooled row)
      __syncwarp();
      for (; released <= r1; ++released)
        if (lane == 0) mbar_arrive(ring_empty(ring_slot(released - c_lo)));
    }
  } else if (warp == 4) {
    // ===== MMA issuer: the whole warp runs the loop with warp-uniform values, one elected lane issues =====
    const uint32_t w_s = base + w_off;
    const uint32_t idesc = idesc_f16(128, p.ncols);
    int groups_seen = 0;
    for (int t = 0; t < n_acc; ++t) {
      const int slot = t & 1;
      for (; groups_seen <= min(2 * t + 4, kEGroups - 1); ++groups_seen) mbar_wait(e_ready(groups_seen), 0);  // E rows 4t .. 4t+9
      mbar_wait(t_empty(slot), (((uint32_t)(t >> 1)) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t e_s = base + e_off + (uint32_t)(4 * t) * p.e_pitch;
      const uint32_t d = tmem_base + (uint32_t)(slot * 128);
#pragma unroll
      for (int m = 0; m < kWChunks / 2; ++m)  // K chunks 2m, 2m + 1 = input rows 2 (c_lo + 2t) - 3 + 2m, + 1
        tc_mma_w(d, smem_desc_interleaved(w_s + (uint32_t)m * 4096u, 2048, 128),
                 smem_desc_interleaved(e_s + (uint32_t)(2 * m) * p.e_pitch, (uint32_t)p.e_pitch, 128), idesc, m != 0 ? 1u : 0u);
      tc_commit_w(t_full(slot));
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 5) >> 2;  // which part of the columns
    const int rowsel = q >> 1;         // conv row of the accumulator's pair
    const int ch = (q & 1) * 32 + lane;
    const float bias = __ldg(p.bias + ch);
    const float scale = __ldg(reinterpret_cast<const float*>(p.w) + kWBytes / 4 + ch);
    const int np = p.ncols >> 4;  // 16-column pieces
    const int kh = (np + 1) >> 1;
    const int k0 = half ? kh : 0, k1 = half ? np : kh;
    for (int t = 0; t < n_acc; ++t) {
      const int slot = t & 1;
      const int idx = 2 * t + rowsel;
      const bool row_ok = idx < n_rows;
      __syncwarp();  // tcgen05.ld below is warp-collective
      mbar_wait(t_full(slot), ((uint32_t)(t >> 1)) & 1u);
      tc_fence_after();
      // the ring aliases the fp16 row buffer: wait until the builders have read all of it
      if (t == 0) mbar_wait(built_bar, 0);
      // the ring row's previous occupant (conv row idx - kRing) has been consumed by the store warps
      if (row_ok && idx >= kRing) mbar_wait(ring_empty(ring_slot(idx)), 0);
      const uint32_t hrow = ring_s + (uint32_t)(ring_slot(idx) * p.ring_pitch + ch * 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 128);
      float carry = -INFINITY;  // conv column 16k - 1
      // 16 columns at a time, the next piece's tcgen05.ld in flight while this one is processed (a load + wait per piece
      // left a ~100-cycle bubble in each)
      auto piece = [&](const uint32_t (&v)[16], float cr, int k) {
        const uint32_t hp_ = hrow + (uint32_t)((8 * k) * 128);
        if (16 * k + 16 <= p.wc) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float l0 = j == 0 ? cr : __uint_as_float(v[2 * j - 1]);
            const float m0 = fmaxf(fmaxf(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), l0);
            const float m1 = fmaxf(fmaxf(__uint_as_float(v[2 * j + 2]), __uint_as_float(v[2 * j + 3])), __uint_as_float(v[2 * j + 1]));
            const uint32_t o = relu_bf16x2(fmaf(m0, scale, bias), fmaf(m1, scale, bias));
            sts_u16(hp_ + j * 128, o);
            sts_u16(hp_ + (j + 1) * 128, o >> 16);
          }
        } else {  // the piece that holds the last conv column
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int pw = 8 * k + j;
            if (pw < p.wp) {
              float m = __uint_as_float(v[2 * j]);
              if (2 * pw + 1 < p.wc) m = fmaxf(m, __uint_as_float(v[2 * j + 1]));
              m = fmaxf(m, j == 0 ? cr : __uint_as_float(v[2 * j - 1]));
              sts_u16(hp_ + j * 128, relu_bf16x2(fmaf(m, scale, bias), 0.f));
            }
          }
        }
      };
      auto release = [&]() {  // this warp has read its part of the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty(slot));
      };
      uint32_t va[16], vb[16];
      if (k0 < k1) {
        if (k0 > 0) carry = tmem_ld1(taddr + (uint32_t)(16 * k0 - 1));
        tmem_ld16(taddr + (uint32_t)(16 * k0), va);
      }
      for (int k = k0; k < k1; k += 2) {
        tmem_ld_wait();  // piece k (and, the first time, the carry)
        const bool more1 = k + 1 < k1;
        if (more1)
          tmem_ld16(taddr + (uint32_t)(16 * (k + 1)), vb);
        else
          release();
        if (row_ok) piece(va, carry, k);
        carry = __uint_as_float(va[15]);
        if (more1) {
          tmem_ld_wait();  // piece k + 1
          if (k + 2 < k1)
            tmem_ld16(taddr + (uint32_t)(16 * (k + 2)), va);
          else
            release();
          if (row_ok) piece(vb, carry, k + 1);
          carry = __uint_as_float(vb[15]);
        }
      }
      if (k0 >= k1) release();  // (a row of at most 16 conv columns: the second warp of the quarter only keeps the barrier counts)
      if (row_ok) {  // this warp's part of conv row idx is in the ring
        __syncwarp();
        if (lane == 0) mbar_arrive(ring_full(ring_slot(idx)));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  stamp_end(p.stamp);
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kSlots * 128);
  }
}

// ======================================================================================================================
// Persistent, pipelined form of the same kernel (the default when the strip can be staged by TMA).
//
// stem_pool_t_kernel runs one CTA per strip of 4 pooled rows with its phases one after the other (load, u8 -> fp16, E rows,
// MMAs, epilogue, stores), two CTAs per SM; 26 input rows are staged and expanded for 16 new ones, 10 conv rows computed
// for 8 new ones, and the 20 KB weight operand is fetched by each of the 3584 CTAs.  ncu (profiles/r2_ncu_stem_pool_t.txt):
// 121 us per 256 images, tensor pipe 16.5 %, DRAM 5.9 %, issue slots half empty: latency-bound.
// Here ONE CTA per SM stays resident and walks over units of kBand = 14 pooled rows (a quarter of a 224 image: 29 conv rows
// = 15 accumulators for 14 pooled rows), with every role running concurrently on rings of mbarriers:
//   warp 5       producer: weights once, then the u8 strip of unit i + 1 / i + 2 (one TMA box of 68 rows, double-buffered)
//   warps 6-9    builders: E rows straight from the u8 strip (8 taps -> fp16, 16 bytes per conv column), one row of every
//                QUAD (4 E rows = what one accumulator adds to the previous one's window) each, into a ring of 8 quads
//   warp 4       MMA issuer: accumulator t reads quads t, t + 1 and half of t + 2 of its unit; 5 MMAs; commits publish the
//                accumulator (ring of 4 TMEM slots) and hand quad t back to the builders
//   warps 10-25  epilogue, two groups of 8 warps that take alternate accumulators (an accumulator's epilogue is a ~1100-cycle
//                dependent chain per warp: two in flight), same register-side horizontal max as above, into a ring of 8
//                conv rows at pooled width
//   warps 0-3    store warps: vertical max of three ring rows, NHWC stores, ring rows handed back
// All counters (quads, accumulators, conv rows) run on across units, so the next unit's strip, E rows and MMAs start while
// the current unit's epilogue and stores drain.
constexpr int kBand = 14;                      // pooled rows per unit
constexpr int kMaxAcc = kBand + 1;             // 2 * kBand + 1 conv rows
constexpr int kStripRowsP = 4 * (kMaxAcc + 2); // input rows staged per unit (E rows 0 .. 4 nacc + 5, rounded to quads)
constexpr int kQuads = 8;                      // E ring, in quads
constexpr int kSlotsP = 4;                     // TMEM accumulator ring (4 x 128 columns)
constexpr int kRingP = 8;                      // conv rows between the epilogue and the store warps
constexpr int kBuilders = 4;
constexpr int kEpiWarpsP = 16;
constexpr int kStoreWarps = 4;  // (8 measured slower: 0.115 vs 0.096 ms -- more warps contending for issue slots and the TMEM read pipe)
constexpr int kMmaWarp = kStoreWarps, kProdWarp = kStoreWarps + 1, kBuild0 = kStoreWarps + 2, kEpi0 = kBuild0 + kBuilders;
constexpr int kThreadsP = 32 * (kStoreWarps + 1 + 1 + kBuilders + kEpiWarpsP);
constexpr int kBarsP = 1 + 2 + 2 + kQuads + kQuads + kSlotsP + kSlotsP + kRingP + kRingP;

// debug build: cycles a role of CTA 0 spends in each of its waits (SPK_STEM_TRACE=1), printed when the kernel ends
#ifdef SPK_DEBUG_SWITCHES
#define STEMP_TIMED(acc, stmt) do { const long long _t0 = clock64(); stmt; acc += clock64() - _t0; } while (0)
#else
#define STEMP_TIMED(acc, stmt) do { stmt; } while (0)
#endif

struct StemPParams {
  int trace;
  const uint4* w;     // as StemTParams
  const float* bias;  // [64]
  __nv_bfloat16* y;   // [n, hp, wp, ldy]
  int n, th, tw, hc, wc, hp, wp, ldy;
  int bands, units;   // units = n * bands
  int ncols, e_pitch, ring_pitch;
  unsigned long long* stamp;
};

__global__ void __launch_bounds__(kThreadsP, 1) stem_pool_p_kernel(const __grid_constant__ CUtensorMap map_x, const StemPParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  unsigned char* gbase = smem_raw + (base - raw);
  // layout: E ring | W | two u8 strips | conv-row ring | barriers | tmem slot
  const uint32_t e_off = 0, w_off = (uint32_t)(4 * kQuads * p.e_pitch), strip_off = w_off + kWBytes;
  const uint32_t ring_off = strip_off + 2u * kStripRowsP * 256u, bar_off = ring_off + (uint32_t)(kRingP * p.ring_pitch);
  int bi = 0;
  const uint32_t bar0 = base + bar_off;
  const uint32_t w_bar = bar0 + 8u * bi; bi += 1;
  const uint32_t strip_full0 = bar0 + 8u * bi; bi += 2;
  const uint32_t strip_empty0 = bar0 + 8u * bi; bi += 2;
  const uint32_t quad_full0 = bar0 + 8u * bi; bi += kQuads;
  const uint32_t quad_empty0 = bar0 + 8u * bi; bi += kQuads;
  const uint32_t t_full0 = bar0 + 8u * bi; bi += kSlotsP;
  const uint32_t t_empty0 = bar0 + 8u * bi; bi += kSlotsP;
  const uint32_t ring_full0 = bar0 + 8u * bi; bi += kRingP;
  const uint32_t ring_empty0 = bar0 + 8u * bi; bi += kRingP;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + bar_off + 8 * kBarsP);
  const uint32_t ring_s = base + ring_off;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler
  pdl_trigger();
  stamp_begin(p.stamp);

  if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_init(w_bar, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(strip_full0 + 8u * s, 1);
        mbar_init(strip_empty0 + 8u * s, kBuilders);
      }
      for (int s = 0; s < kQuads; ++s) {
        mbar_init(quad_full0 + 8u * s, kBuilders);
        mbar_init(quad_empty0 + 8u * s, 1);
      }
      for (int s = 0; s < kSlotsP; ++s) {
        mbar_init(t_full0 + 8u * s, 1);
        mbar_init(t_empty0 + 8u * s, kEpiWarpsP / 2);
      }
      for (int s = 0; s < kRingP; ++s) {
        mbar_init(ring_full0 + 8u * s, 4);
        mbar_init(ring_empty0 + 8u * s, kStoreWarps);
      }
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32((const void*)tmem_slot), kSlotsP * 128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> image, pooled rows [p0, p1], first conv row, conv rows, accumulators, input row of E row 0
  struct Unit { int image, p0, p1, c_lo, n_rows, nacc, y_base; };
  auto unit_geom = [&](int unit) {
    Unit u;
    u.image = unit / p.bands;
    u.p0 = (unit - u.image * p.bands) * kBand;
    u.p1 = min(u.p0 + kBand, p.hp) - 1;
    u.c_lo = max(0, 2 * u.p0 - 1);
    const int c_hi = min(p.hc - 1, 2 * u.p1 + 1);
    u.n_rows = c_hi - u.c_lo + 1;
    u.nacc = (u.n_rows + 1) >> 1;
    u.y_base = 2 * u.c_lo - 3;
    return u;
  };

  if (warp == kProdWarp) {
    // ===== producer =====
    if (lane == 0) {
      mbar_expect_tx(w_bar, (uint32_t)kWBytes);
      bulk_load(base + w_off, p.w, kWBytes, w_bar);
      pdl_wait();  // the image planes are the previous kernel's output
      int i = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++i) {
        const Unit u = unit_geom(unit);
        const int b = i & 1;
        if (i >= 2) mbar_wait(strip_empty0 + 8u * b, (uint32_t)(((i >> 1) - 1) & 1));
        mbar_expect_tx(strip_full0 + 8u * b, (uint32_t)(kStripRowsP * 256));
        // pixel x of input row y lands at strip[(y - y_base) * 256 + x + kXOff]; out-of-image rows / columns are zero-filled
        tma_load_3d(base + strip_off + (uint32_t)b * (kStripRowsP * 256), &map_x, strip_full0 + 8u * b, -kXOff, u.y_base, u.image);
      }
    }
  } else if (warp >= kBuild0 && warp < kBuild0 + kBuilders) {
    // ===== builders: E[j] = fp16 of pixels 2j-3 .. 2j+4 = strip bytes [2j + 13, 2j + 21) of the row =====
    const int b = warp - kBuild0;
    int i = 0, qg = 0;
    long long w_strip = 0, w_quad = 0;
    const long long t_begin = clock64();
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++i) {
      const Unit u = unit_geom(unit);
      const int nq = u.nacc + 2;
      STEMP_TIMED(w_strip, mbar_wait(strip_full0 + 8u * (i & 1), (uint32_t)((i >> 1) & 1)));
      const unsigned char* strip = gbase + strip_off + (size_t)(i & 1) * (kStripRowsP * 256);
      for (int q = 0; q < nq; ++q) {
        const int g = qg + q, slot = g & (kQuads - 1);
        if (g >= kQuads) STEMP_TIMED(w_quad, mbar_wait(quad_empty0 + 8u * slot, (uint32_t)(((g / kQuads) - 1) & 1)));
        const uint32_t* srow = reinterpret_cast<const uint32_t*>(strip + (4 * q + b) * 256);
        unsigned char* erow = gbase + e_off + (size_t)(slot * 4 + b) * p.e_pitch;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = lane + 32 * c;
          if (j < p.wc) {
            const int a = 2 * j + 13, wi = a >> 2;
            const uint32_t sh = (uint32_t)(a & 3) * 8u;
            const uint32_t w0 = srow[wi], w1 = srow[wi + 1], w2 = srow[wi + 2];
            const uint2 c0 = bytes4_to_f16x4(__funnelshift_r(w0, w1, sh)), c1 = bytes4_to_f16x4(__funnelshift_r(w1, w2, sh));
            *reinterpret_cast<uint4*>(erow + j * 16) = make_uint4(c0.x, c0.y, c1.x, c1.y);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(quad_full0 + 8u * slot);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(strip_empty0 + 8u * (i & 1));
      qg += nq;
    }
#ifdef SPK_DEBUG_SWITCHES
    if (p.trace && blockIdx.x == 0 && b == 0 && lane == 0)
      printf("stem_p builder: total %lld cycles, %d quads; waiting strip_full %lld, quad_empty %lld\n", clock64() - t_begin, qg, w_strip, w_quad);
#endif
    (void)t_begin; (void)w_strip; (void)w_quad;
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer: the whole warp runs the loop with warp-uniform values, one elected lane issues =====
    mbar_wait(w_bar, 0);
    const uint32_t w_s = base + w_off, e_s = base + e_off;
    const uint32_t idesc = idesc_f16(128, p.ncols);
    int qg = 0, tg = 0, qseen = 0;
    long long w_qfull = 0, w_tempty = 0;
    const long long t_begin = clock64();
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const Unit u = unit_geom(unit);
      for (int t = 0; t < u.nacc; ++t, ++tg) {
        for (; qseen < qg + t + 3; ++qseen)
          STEMP_TIMED(w_qfull, mbar_wait(quad_full0 + 8u * (qseen & (kQuads - 1)), (uint32_t)((qseen / kQuads) & 1)));
        const int slot = tg & (kSlotsP - 1);
        if (tg >= kSlotsP) STEMP_TIMED(w_tempty, mbar_wait(t_empty0 + 8u * slot, (uint32_t)(((tg / kSlotsP) - 1) & 1)));
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(slot * 128);
#pragma unroll
        for (int m = 0; m < kWChunks / 2; ++m) {  // K chunks 2m, 2m + 1 = E rows 4t + 2m, + 1 of the unit (never split by the ring's wrap)
          const uint32_t rr = (uint32_t)((((qg + t) << 2) + 2 * m) & (4 * kQuads - 1));
          tc_mma_w(d, smem_desc_interleaved(w_s + (uint32_t)m * 4096u, 2048, 128),
                   smem_desc_interleaved(e_s + rr * (uint32_t)p.e_pitch, (uint32_t)p.e_pitch, 128), idesc, m != 0 ? 1u : 0u);
        }
        tc_commit_w(t_full0 + 8u * slot);
        tc_commit_w(quad_empty0 + 8u * ((qg + t) & (kQuads - 1)));  // quad t is not read by later accumulators
      }
      // the two trailing quads of the unit
      tc_commit_w(quad_empty0 + 8u * ((qg + u.nacc) & (kQuads - 1)));
      tc_commit_w(quad_empty0 + 8u * ((qg + u.nacc + 1) & (kQuads - 1)));
      qg += u.nacc + 2;
    }
#ifdef SPK_DEBUG_SWITCHES
    if (p.trace && blockIdx.x == 0 && lane == 0)
      printf("stem_p mma: total %lld cycles, %d accumulators; waiting quad_full %lld, t_empty %lld\n", clock64() - t_begin, tg, w_qfull, w_tempty);
#endif
    (void)t_begin; (void)w_qfull; (void)w_tempty;
  } else if (warp >= kEpi0) {
    // ===== epilogue: group g takes the accumulators with (running index & 1) == g =====
    const int e = warp - kEpi0;
    const int grp = e >> 3;
    const int half = (e >> 2) & 1;     // which part of the columns
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int rowsel = q >> 1;         // conv row of the accumulator's pair
    const int ch = (q & 1) * 32 + lane;
    const float bias = __ldg(p.bias + ch);
    const float scale = __ldg(reinterpret_cast<const float*>(p.w) + kWBytes / 4 + ch);
    const int np = p.ncols >> 4;  // 16-column pieces
    const int kh = (np + 1) >> 1;
    const int k0 = half ? kh : 0, k1 = half ? np : kh;
    int tg0 = 0, rg0 = 0;
    long long w_tfull = 0, w_rempty = 0;
    const long long t_begin = clock64();
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const Unit u = unit_geom(unit);
      for (int t = 0; t < u.nacc; ++t) {
        const int tg = tg0 + t;
        if ((tg & 1) != grp) continue;
        const int slot = tg & (kSlotsP - 1);
        const int idx = 2 * t + rowsel;
        const bool row_ok = idx < u.n_rows;
        const int rg = rg0 + idx, rs = rg & (kRingP - 1);
        __syncwarp();  // tcgen05.ld below is warp-collective
        STEMP_TIMED(w_tfull, mbar_wait(t_full0 + 8u * slot, (uint32_t)((tg / kSlotsP) & 1)));
        tc_fence_after();
        // the ring row's previous occupant (conv row rg - kRingP) has been consumed by the store warps
        if (row_ok && rg >= kRingP) STEMP_TIMED(w_rempty, mbar_wait(ring_empty0 + 8u * rs, (uint32_t)(((rg / kRingP) - 1) & 1)));
        const uint32_t hrow = ring_s + (uint32_t)(rs * p.ring_pitch + ch * 2);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 128);
        float carry = -INFINITY;  // conv column 16k - 1
        auto piece = [&](const uint32_t (&v)[16], float cr, int k) {
          const uint32_t hp_ = hrow + (uint32_t)((8 * k) * 128);
          if (16 * k + 16 <= p.wc) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const float l0 = j == 0 ? cr : __uint_as_float(v[2 * j - 1]);
              const float m0 = fmaxf(fmaxf(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), l0);
              const float m1 = fmaxf(fmaxf(__uint_as_float(v[2 * j + 2]), __uint_as_float(v[2 * j + 3])), __uint_as_float(v[2 * j + 1]));
              const uint32_t o = relu_bf16x2(fmaf(m0, scale, bias), fmaf(m1, scale, bias));
              sts_u16(hp_ + j * 128, o);
              sts_u16(hp_ + (j + 1) * 128, o >> 16);
            }
          } else {  // the piece that holds the last conv column
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int pw = 8 * k + j;
              if (pw < p.wp) {
                float m = __uint_as_float(v[2 * j]);
                if (2 * pw + 1 < p.wc) m = fmaxf(m, __uint_as_float(v[2 * j + 1]));
                m = fmaxf(m, j == 0 ? cr : __uint_as_float(v[2 * j - 1]));
                sts_u16(hp_ + j * 128, relu_bf16x2(fmaf(m, scale, bias), 0.f));
              }
            }
          }
        };
        auto release = [&]() {  // this warp has read its part of the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty0 + 8u * slot);
        };
        uint32_t va[16], vb[16];
        if (k0 < k1) {
          if (k0 > 0) carry = tmem_ld1(taddr + (uint32_t)(16 * k0 - 1));
          tmem_ld16(taddr + (uint32_t)(16 * k0), va);
        }
        for (int k = k0; k < k1; k += 2) {
          tmem_ld_wait();  // piece k (and, the first time, the carry)
          const bool more1 = k + 1 < k1;
          if (more1)
            tmem_ld16(taddr + (uint32_t)(16 * (k + 1)), vb);
          else
            release();
          if (row_ok) piece(va, carry, k);
          carry = __uint_as_float(va[15]);
          if (more1) {
            tmem_ld_wait();  // piece k + 1
            if (k + 2 < k1)
              tmem_ld16(taddr + (uint32_t)(16 * (k + 2)), va);
            else
              release();
            if (row_ok) piece(vb, carry, k + 1);
            carry = __uint_as_float(vb[15]);
          }
        }
        if (k0 >= k1) release();  // (a row of at most 16 conv columns: the second warp of the quarter only keeps the barrier counts)
        if (row_ok) {  // this warp's part of conv row idx is in the ring
          __syncwarp();
          if (lane == 0) mbar_arrive(ring_full0 + 8u * rs);
        }
      }
      tg0 += u.nacc;
      rg0 += u.n_rows;
    }
#ifdef SPK_DEBUG_SWITCHES
    if (p.trace && blockIdx.x == 0 && e == 0 && lane == 0)
      printf("stem_p epilogue: total %lld cycles; waiting t_full %lld, ring_empty %lld\n", clock64() - t_begin, w_tfull, w_rempty);
#endif
    (void)t_begin; (void)w_tfull; (void)w_rempty;
  } else {
    // ===== store warps (0 .. kStoreWarps - 1): per pooled row the vertical max of its three conv rows, coalesced NHWC stores =====
    pdl_wait();  // (the output buffer may still be read by the previous step's kernels)
    int rg0 = 0;
    long long w_rfull = 0;
    const long long t_begin = clock64();
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const Unit u = unit_geom(unit);
      auto rfull = [&](int idx) {
        const int rg = rg0 + idx;
        STEMP_TIMED(w_rfull, mbar_wait(ring_full0 + 8u * (rg & (kRingP - 1)), (uint32_t)((rg / kRingP) & 1)));
      };
      auto raddr = [&](int idx) { return ring_s + (uint32_t)(((rg0 + idx) & (kRingP - 1)) * p.ring_pitch); };
      int released = 0;  // unit-relative conv rows below this one have been handed back
      for (int prow = u.p0; prow <= u.p1; ++prow) {
        const int r1 = 2 * prow;
        const bool has0 = r1 - 1 >= 0, has2 = r1 + 1 <= p.hc - 1;
        const int i1 = r1 - u.c_lo;
        if (has0) rfull(i1 - 1);
        rfull(i1);
        if (has2) rfull(i1 + 1);
        const uint32_t b1 = raddr(i1);
        const uint32_t b0 = has0 ? raddr(i1 - 1) : b1;  // (a missing row reads the middle one again)
        const uint32_t b2 = has2 ? raddr(i1 + 1) : b1;
        __nv_bfloat16* yrow = p.y + ((size_t)u.image * p.hp + prow) * p.wp * p.ldy;
        for (int o = tid; o < p.wp * 8; o += 32 * kStoreWarps) {
          const int pw = o >> 3, j = o & 7;
          const uint32_t off = (uint32_t)o * 16u;  // = pw * 128 + j * 16
          uint4 m4 = lds_v4(b1 + off);
          const uint4 t0 = lds_v4(b0 + off), t2 = lds_v4(b2 + off);
          __nv_bfloat162* mm = reinterpret_cast<__nv_bfloat162*>(&m4);
          const __nv_bfloat162* q0 = reinterpret_cast<const __nv_bfloat162*>(&t0);
          const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&t2);
#pragma unroll
          for (int e = 0; e < 4; ++e) mm[e] = __hmax2(__hmax2(mm[e], q0[e]), q2[e]);
          *reinterpret_cast<uint4*>(yrow + (size_t)pw * p.ldy + j * 8) = m4;
        }
        // conv rows up to 2 prow are done with (2 prow + 1 also feeds the next pooled row)
        __syncwarp();
        for (; released <= i1; ++released)
          if (lane == 0) mbar_arrive(ring_empty0 + 8u * ((rg0 + released) & (kRingP - 1)));
      }
      __syncwarp();
      for (; released < u.n_rows; ++released)  // the unit's last odd conv row
        if (lane == 0) mbar_arrive(ring_empty0 + 8u * ((rg0 + released) & (kRingP - 1)));
      rg0 += u.n_rows;
    }
#ifdef SPK_DEBUG_SWITCHES
    if (p.trace && blockIdx.x == 0 && tid == 0) printf("stem_p store: total %lld cycles; waiting ring_full %lld\n", clock64() - t_begin, w_rfull);
#endif
    (void)t_begin; (void)w_rfull;
  }

  tc_fence_before();
  __syncthreads();
  stamp_end(p.stamp);
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kSlotsP * 128);
  }
}

}  // namespace

bool stem_pool_t_supported(int wc, int wp) { return wc >= 2 && wc <= 128 && wp >= 1 && wp <= 64; }

// w: folded [64][7][7] fp32 (already includes BatchNorm); the 1/255 of ToTensor is folded here.  fp16 of w * 2^k(o),
// 2^k(o) the power of two that brings the channel's largest weight into [2^13, 2^14); byte offset of (K chunk q, M row m,
// tap s) = q * 2048 + m * 16 + s * 2.  The 64 factors 2^-k(o) follow (floats at byte kWBytes).
int stem_pool_t_pack_weights(spk_ctx* ctx, const float* w, uint4** d_out) {
  std::vector<uint16_t> tile(kWBytes / 2 + 128, 0);
  float* inv_scale = reinterpret_cast<float*>(tile.data() + kWBytes / 2);
  for (int o = 0; o < 64; ++o) {
    double amax = 0.0;
    for (int t = 0; t < 49; ++t) amax = std::max(amax, std::fabs((double)w[o * 49 + t] / 255.0));
    int k = 0;
    if (amax > 0.0 && std::isfinite(amax)) {
      int e;
      std::frexp(amax, &e);  // amax = m * 2^e, m in [0.5, 1)
      k = std::max(-60, std::min(60, 14 - e));
    }
    inv_scale[o] = (float)std::ldexp(1.0, -k);
    for (int r = 0; r < 7; ++r)
      for (int s = 0; s < 7; ++s) {
        const __half hv = __float2half_rn((float)std::ldexp((double)w[(o * 7 + r) * 7 + s] / 255.0, k));
        uint16_t b;
        memcpy(&b, &hv, 2);
        tile[(size_t)r * 1024 + (size_t)o * 8 + s] = b;               // first conv row of the pair: chunk q = r
        tile[(size_t)(r + 2) * 1024 + (size_t)(64 + o) * 8 + s] = b;  // second: the same filter two input rows down
      }
  }
  SPK_CUDA_OK(ctx, cudaMalloc(d_out, kWBytes + 256));
  SPK_CUDA_OK(ctx, cudaMemcpy(*d_out, tile.data(), kWBytes + 256, cudaMemcpyHostToDevice));
  return SPK_OK;
}

int launch_stem_pool_t(spk_ctx* ctx, int n, int th, int tw, const uint8_t* x, const uint4* w, const float* bias, __nv_bfloat16* y,
                       int hc, int wc, int hp, int wp, int ldy) {
  if (n <= 0) return SPK_OK;
  StemTParams p;
  p.x = x;
  p.w = w;
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
  static const bool no_tma = debug_env("SPK_STEM_NO_TMA") != nullptr;  // A/B switch: stage the strip with all threads
  p.use_tma = (!no_tma && tw % 16 == 0 && tw <= 224 && ((uintptr_t)x & 15) == 0 && encode_fn() != nullptr) ? 1 : 0;
  p.groups = (tw + 16 + 15) / 16;                    // row-buffer elements [0, 16 * groups) cover pixels up to tw + 12
  p.rb_pitch = 32 * p.groups + 32;                   // + slack for the dead columns' over-read
  p.pitch = p.use_tma ? 256 : (16 * p.groups + 32);  // strip bytes read: up to 16 * groups + 15
  p.ncols = (wc + 15) / 16 * 16;
  p.e_pitch = p.ncols * 16;
  p.ring_pitch = wp * 128;
  p.rb_bytes = (kERows * p.rb_pitch + 127) / 128 * 128;
  p.region = (std::max(kRing * p.ring_pitch, p.rb_bytes + kERows * p.pitch) + 127) / 128 * 128;
  // the tensor map of the u8 batch {tw, th, n}, box {256, rows, 1}: cached per (pointer, geometry)
  static const bool no_persistent = debug_env("SPK_STEM_STRIPS") != nullptr;  // A/B switch: the CTA-per-strip kernel
  const bool persistent = p.use_tma && !no_persistent;
  const int box_rows = persistent ? kStripRowsP : kERows;
  static thread_local struct { const void* x; int n, th, tw, rows; CUtensorMap map; } cache = {nullptr, 0, 0, 0, 0, {}};
  if (p.use_tma && (cache.x != x || cache.n < n || cache.th != th || cache.tw != tw || cache.rows != box_rows)) {
    cuuint64_t dims[3] = {(cuuint64_t)tw, (cuuint64_t)th, (cuuint64_t)n};
    cuuint64_t strides[2] = {(cuuint64_t)tw, (cuuint64_t)tw * th};
    cuuint32_t box[3] = {256u, (cuuint32_t)box_rows, 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn()(&cache.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(x), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "stem: cuTensorMapEncodeTiled(u8 batch) failed: %d", (int)r);
    cache.x = x;
    cache.n = n;
    cache.th = th;
    cache.tw = tw;
    cache.rows = box_rows;
  }
  if (persistent) {
    StemPParams q;
    static int trace_left = debug_env("SPK_STEM_TRACE") ? 3 : 0;
    q.trace = trace_left > 0 ? 1 : 0;
    if (trace_left > 0) --trace_left;
    q.w = w;
    q.bias = bias;
    q.y = y;
    q.n = n;
    q.th = th;
    q.tw = tw;
    q.hc = hc;
    q.wc = wc;
    q.hp = hp;
    q.wp = wp;
    q.ldy = ldy;
    q.bands = (hp + kBand - 1) / kBand;
    q.units = n * q.bands;
    q.ncols = p.ncols;
    q.e_pitch = p.e_pitch;
    q.ring_pitch = p.ring_pitch;
    q.stamp = ctx->cur_stamp;
    const size_t smem_p = 128 + (size_t)4 * kQuads * q.e_pitch + kWBytes + 2 * (size_t)kStripRowsP * 256 + (size_t)kRingP * q.ring_pitch +
                          8 * kBarsP + 16;
    SPK_CUDA_OK(ctx, cudaFuncSetAttribute(stem_pool_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    const int grid = std::min(q.units, ctx->sm_count);
    SPK_CUDA_OK(ctx, launch_pdl(stem_pool_p_kernel, dim3((unsigned)grid), dim3(kThreadsP), smem_p, ctx->stream, cache.map, q));
    SPK_LAUNCH_CHECK(ctx);
    return SPK_OK;
  }
  const size_t smem = 128 + (size_t)kERows * p.e_pitch + kWBytes + (size_t)p.region + 8 * (kEGroups + 2 * kSlots + 3 + 2 * kRing) + 16;
  SPK_CUDA_OK(ctx, cudaFuncSetAttribute(stem_pool_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.stamp = ctx->cur_stamp;
  SPK_CUDA_OK(ctx, launch_pdl(stem_pool_t_kernel, dim3((unsigned)(n * p.strips)), dim3(kThreads), smem, ctx->stream, cache.map, p));
  SPK_LAUNCH_CHECK(ctx);
  return SPK_OK;
}

}  // namespace spk

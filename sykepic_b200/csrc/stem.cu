// K2 stem: conv 7x7 / stride 2 / pad 3 (1 gray plane -> 64 channels, folded BatchNorm) + ReLU +
// max-pool 3x3 / stride 2 / pad 1, fused, on the tcgen05 tensor cores.  u8 image in, bf16 NHWC out.
//
// Replaces conv1 / bn1 / relu / maxpool of torchvision's ResNet (and conv0 / norm0 / relu0 / pool0
// of DenseNet) as run by TorchVisionNet.forward (sykepic/train/network.py:66-68) on the three
// identical planes cv2.imread produces (sykepic/train/data.py:217-219); the planes are folded into
// one (w' = sum_c w[:, c]) and ToTensor's 1/255 is folded into the weights, so the A operand is the
// exact integer pixel value in bf16.
//
// One CTA = one image x one strip of 7 pooled rows.  GEMM per conv output row: M = 128 (the row's
// <= 128 output columns), N = 64, K = 64 (k = 8 r + s; s = 7 and r = 7 carry zero weights).
//   warps 0-3  build the A operand: gather the 7x8 input patch of every output column from the
//              image strip staged in shared memory, convert u8 -> bf16 and store it in the K-major
//              SWIZZLE_128B layout (3-stage ring);
//   warp 4     allocates TMEM, issues tcgen05.mma (one thread), commits to mbarriers;
//   warps 5-8  epilogue: tcgen05.ld the row's accumulator (4-slot TMEM ring), running vertical max
//              of the 3 conv rows of a pooled row in registers, + bias, ReLU, bf16, horizontal max
//              through shared memory, coalesced 16-byte stores of the pooled row.
// The 112x112x64 conv output (411 MB per 256 images in bf16) never exists in HBM.
#include <cuda.h>

#include <cstring>

#include "spk_internal.h"

namespace spk {
namespace {

constexpr int kBuilders = 128;
constexpr int kThreads = 288;  // 4 builder warps + 1 MMA warp + 4 epilogue warps
constexpr int kAStages = 3;
constexpr int kSlots = 4;      // TMEM accumulator ring: 4 x 64 columns
constexpr int kABytes = 128 * 128;
constexpr int kBTile = 64 * 128;
constexpr int kBBytes = 2 * kBTile;  // weights as bf16 hi + bf16 lo tiles (w = hi + lo to 2^-17): two MMA passes over one A tile
constexpr int kPoolBytes = 128 * 128;
constexpr int kPoolRowsPerStrip = 7;
constexpr int kMaxImgRows = 2 * (2 * kPoolRowsPerStrip + 1 - 1) + 7 + 1;  // input rows of a strip + slack
constexpr int kMaxT = 256;
constexpr int kPitchPad = 16;

struct StemParams {
  const uint8_t* x;      // [n, th, tw] u8
  const uint4* w_sw;     // 16 KB: the 64 x 64 bf16 weight tiles (hi, lo), already in the swizzled smem layout
  const float* bias;     // [64]
  __nv_bfloat16* y;      // [n, hp, wp, ldy]
  int n, th, tw, hc, wc, hp, wp, ldy;
  int strips, m_count, pitch;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin == 64) t0 = clock64();
    if (spin > 64 && (spin & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();  // never hang the GPU
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// D = f32, A = B = bf16, K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// two bytes (x = b0 | b1 << 8) -> packed bf16x2 of their integer values, exact.
// 0x4B000000 | v is the float 2^23 + v; subtracting 2^23 gives float(v) without the I2F pipe.
__device__ __forceinline__ uint32_t bytes2_to_bf16x2(uint32_t x) {
  const float f0 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7440)) - 8388608.0f;
  const float f1 = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7441)) - 8388608.0f;
  __nv_bfloat162 p = __floats2bfloat162_rn(f0, f1);
  return *reinterpret_cast<uint32_t*>(&p);
}

__global__ void __launch_bounds__(kThreads, 1) stem_pool_kernel(const StemParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - raw);
  // layout: A[3] | B | pool[2] | image | barriers | tmem slot
  const uint32_t a_off = 0, b_off = kAStages * kABytes, pool_off = b_off + kBBytes, img_off = pool_off + 2 * kPoolBytes;
  const int img_bytes = (kMaxImgRows + 1) * p.pitch;
  const uint32_t bar_off = (img_off + img_bytes + 15u) & ~15u;
  auto a_full = [&](int s) { return base + bar_off + 8u * s; };
  auto a_empty = [&](int s) { return base + bar_off + 8u * (kAStages + s); };
  auto t_full = [&](int s) { return base + bar_off + 8u * (2 * kAStages + s); };
  auto t_empty = [&](int s) { return base + bar_off + 8u * (2 * kAStages + kSlots + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + bar_off + 8 * (2 * kAStages + 2 * kSlots));
  unsigned char* img = gbase + img_off;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int image = blockIdx.x / p.strips;
  const int strip = blockIdx.x - image * p.strips;
  const int p0 = strip * kPoolRowsPerStrip;
  const int p1 = min(p0 + kPoolRowsPerStrip, p.hp) - 1;  // last pooled row of the strip
  const int c_lo = max(0, 2 * p0 - 1), c_hi = min(p.hc - 1, 2 * p1 + 1);  // conv rows needed
  const int i_lo = 2 * c_lo - 3;                                          // first input row (may be negative)
  const int n_img_rows = 2 * (c_hi - c_lo) + 7;

  // ---- zero the A ring (chunk 7 and rows >= m_count stay zero) and the image strip (padding)
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* a4 = reinterpret_cast<uint4*>(gbase + a_off);
    for (int i = tid; i < kAStages * kABytes / 16; i += kThreads) a4[i] = z;
    uint4* i4 = reinterpret_cast<uint4*>(img);  // img_off is 16-byte aligned
    for (int i = tid; i < (img_bytes + 15) / 16; i += kThreads) i4[i] = z;
  }
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < kAStages; ++s) {
        mbar_init(a_full(s), 4);
        mbar_init(a_empty(s), 1);
      }
      for (int s = 0; s < kSlots; ++s) {
        mbar_init(t_full(s), 1);
        mbar_init(t_empty(s), 4);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)(kSlots * 64))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  // ---- stage the weights tile and the image strip
  {
    uint4* b4 = reinterpret_cast<uint4*>(gbase + b_off);
    for (int i = tid; i < kBBytes / 16; i += kThreads) b4[i] = __ldg(p.w_sw + i);
    const uint8_t* src = p.x + (size_t)image * p.th * p.tw;
    if ((p.tw & 3) == 0) {
      const int q_per_row = p.tw >> 2;
      for (int e = tid; e < n_img_rows * q_per_row; e += kThreads) {
        const int rr = e / q_per_row, c4 = e - rr * q_per_row;
        const int gr = i_lo + rr;
        if (gr < 0 || gr >= p.th) continue;
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)gr * p.tw) + c4);
        unsigned char* d = img + rr * p.pitch + 3 + 4 * c4;  // pixel x lives at column x + 3
        d[0] = (unsigned char)(v & 255u);
        d[1] = (unsigned char)((v >> 8) & 255u);
        d[2] = (unsigned char)((v >> 16) & 255u);
        d[3] = (unsigned char)(v >> 24);
      }
    } else {
      for (int e = tid; e < n_img_rows * p.tw; e += kThreads) {
        const int rr = e / p.tw, c = e - rr * p.tw;
        const int gr = i_lo + rr;
        if (gr < 0 || gr >= p.th) continue;
        img[rr * p.pitch + 3 + c] = __ldg(src + (size_t)gr * p.tw + c);
      }
    }
  }
  // generic-proxy writes of B (and the zeroed A ring) must be visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_rows = c_hi - c_lo + 1;

  if (warp < 4) {
    // ===== A builders =====
    const int chunks = 7 * p.m_count;
    for (int idx = 0; idx < n_rows; ++idx) {
      const int s = idx % kAStages;
      mbar_wait(a_empty(s), (((uint32_t)(idx / kAStages)) & 1u) ^ 1u);
      unsigned char* a = gbase + a_off + s * kABytes;
      const unsigned char* irow = img + (2 * idx) * p.pitch;  // input row 2*(c_lo+idx) - 3 == strip row 2*idx
      int m = tid, r = 0;
      while (m >= p.m_count) {
        m -= p.m_count;
        ++r;
      }
      for (int c = tid; c < chunks; c += kBuilders) {
        const unsigned short* src = reinterpret_cast<const unsigned short*>(irow + r * p.pitch + 2 * m);
        uint4 o;
        o.x = bytes2_to_bf16x2(src[0]);
        o.y = bytes2_to_bf16x2(src[1]);
        o.z = bytes2_to_bf16x2(src[2]);
        o.w = bytes2_to_bf16x2(src[3]);
        *reinterpret_cast<uint4*>(a + m * 128 + ((r ^ (m & 7)) << 4)) = o;
        m += kBuilders;
        while (m >= p.m_count) {
          m -= p.m_count;
          ++r;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(s));
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint64_t b_hi = smem_desc_sw128(base + b_off);
      const uint64_t b_lo = smem_desc_sw128(base + b_off + kBTile);
      for (int idx = 0; idx < n_rows; ++idx) {
        const int s = idx % kAStages, slot = idx % kSlots;
        mbar_wait(t_empty(slot), (((uint32_t)(idx / kSlots)) & 1u) ^ 1u);
        mbar_wait(a_full(s), ((uint32_t)(idx / kAStages)) & 1u);
        tc_fence_after();
        const uint64_t a_desc = smem_desc_sw128(base + a_off + s * kABytes);
        const uint32_t d = tmem_base + (uint32_t)(slot * 64);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma(d, a_desc + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), kIdesc, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma(d, a_desc + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), kIdesc, 1u);
        tc_commit(a_empty(s));
        tc_commit(t_full(slot));
      }
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;          // TMEM lane quarter this warp may read
    const int wo = q * 32 + lane;    // conv output column == TMEM lane
    const int et = (warp - 5) * 32 + lane;  // 0..127 among the epilogue threads
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = -INFINITY;
    int emitted = 0;
    for (int idx = 0; idx < n_rows; ++idx) {
      const int i = c_lo + idx, slot = idx % kSlots;
      __syncwarp();  // tcgen05.ld below is warp-collective
      mbar_wait(t_full(slot), ((uint32_t)(idx / kSlots)) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 64);
      const bool last_of_window = (i & 1) || (i == p.hc - 1);
      const int prow = i >> 1;
      const bool emit = last_of_window && prow >= p0 && prow <= p1;
      {
        uint32_t v[32];
        tmem_ld32(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], __uint_as_float(v[j]));
        tmem_ld32(taddr + 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 + j] = fmaxf(acc[32 + j], __uint_as_float(v[j]));
      }
      if (emit) {
        unsigned char* pool = gbase + pool_off + (emitted & 1) * kPoolBytes;
        // + bias, ReLU, bf16; row wo of the pool buffer, 16-byte chunks swizzled by (wo & 7)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias) + 2 * j);
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias) + 2 * j + 1);
          __nv_bfloat162 t0 = __floats2bfloat162_rn(fmaxf(acc[8 * j + 0] + b0.x, 0.f), fmaxf(acc[8 * j + 1] + b0.y, 0.f));
          __nv_bfloat162 t1 = __floats2bfloat162_rn(fmaxf(acc[8 * j + 2] + b0.z, 0.f), fmaxf(acc[8 * j + 3] + b0.w, 0.f));
          __nv_bfloat162 t2 = __floats2bfloat162_rn(fmaxf(acc[8 * j + 4] + b1.x, 0.f), fmaxf(acc[8 * j + 5] + b1.y, 0.f));
          __nv_bfloat162 t3 = __floats2bfloat162_rn(fmaxf(acc[8 * j + 6] + b1.z, 0.f), fmaxf(acc[8 * j + 7] + b1.w, 0.f));
          uint4 o;
          o.x = *reinterpret_cast<uint32_t*>(&t0);
          o.y = *reinterpret_cast<uint32_t*>(&t1);
          o.z = *reinterpret_cast<uint32_t*>(&t2);
          o.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(pool + wo * 128 + ((j ^ (wo & 7)) << 4)) = o;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps
        // horizontal max over conv columns 2pw-1, 2pw, 2pw+1 and a coalesced store of the pooled row
        __nv_bfloat16* yrow = p.y + ((size_t)image * p.hp + prow) * p.wp * p.ldy;
        for (int o = et; o < p.wp * 8; o += 128) {
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
      if (last_of_window) {
        // conv row 2p+1 is also the first row of pooled row p+1: restart the running max from it
        // (re-read from TMEM rather than keeping 64 more registers live across the emit)
        if (i & 1) {
          uint32_t v[32];
          tmem_ld32(taddr, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
          tmem_ld32(taddr + 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[32 + j] = __uint_as_float(v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j) acc[j] = -INFINITY;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty(slot));  // done with this accumulator slot
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(kSlots * 64)) : "memory");
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
int stem_pool_pack_weights(spk_ctx* ctx, const float* w, uint4** d_out) {
  std::vector<uint16_t> tile(kBBytes / 2, 0);
  for (int o = 0; o < 64; ++o)
    for (int r = 0; r < 7; ++r)
      for (int s = 0; s < 7; ++s) {
        const float v = (float)((double)w[(o * 7 + r) * 7 + s] / 255.0);
        const __nv_bfloat16 hi = __float2bfloat16(v);
        const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
        uint16_t bh, bl;
        memcpy(&bh, &hi, 2);
        memcpy(&bl, &lo, 2);
        // row o, 16-byte chunk r at position (r ^ (o & 7)), element s
        const size_t at = (size_t)o * 64 + ((r ^ (o & 7)) * 8) + s;
        tile[at] = bh;
        tile[kBTile / 2 + at] = bl;
      }
  SPK_CUDA_OK(ctx, cudaMalloc(d_out, kBBytes));
  SPK_CUDA_OK(ctx, cudaMemcpy(*d_out, tile.data(), kBBytes, cudaMemcpyHostToDevice));
  return SPK_OK;
}

int launch_stem_pool(spk_ctx* ctx, int n, int th, int tw, const uint8_t* x, const uint4* w_sw, const float* bias,
                     __nv_bfloat16* y, int hc, int wc, int hp, int wp, int ldy) {
  if (n <= 0) return SPK_OK;
  StemParams p;
  p.x = x;
  p.w_sw = w_sw;
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
  p.m_count = (wc + 7) & ~7;
  p.pitch = (tw + kPitchPad + 15) & ~15;
  const size_t smem = 1024 + kAStages * kABytes + kBBytes + 2 * kPoolBytes + (size_t)(kMaxImgRows + 1) * p.pitch + 16 +
                      8 * (2 * kAStages + 2 * kSlots) + 16;
  SPK_CUDA_OK(ctx, cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stem_pool_kernel<<<(unsigned)(n * p.strips), kThreads, smem, ctx->stream>>>(p);
  SPK_LAUNCH_CHECK(ctx);
  return SPK_OK;
}

}  // namespace spk

// K2 (3x3 / stride 1 / pad 1, Cin and Cout in {64, 128}): halo-resident implicit GEMM on CTA PAIRS with the whole
// filter bank resident in shared memory.
//
// Replaces the 3x3 conv / BatchNorm / ReLU / residual-add sequence of torchvision's ResNet blocks as run by
// TorchVisionNet.forward (sykepic/train/network.py:66-68); library calls in the reference.
//
// Why: the MMA-issuer trace of conv_halo.cu on the ResNet-18 layer1 / layer2 shapes (SPK_HALO_TRACE) shows the
// tensor pipe waiting on SHARED-MEMORY BANDWIDTH (128 B / cycle / SM), not on loads or the epilogue:
//   64->64 @56x56:   66 cycles per M128 N64 K16 MMA instead of 32 -- every MMA reads 4 KB of A and 2 KB of B, and TMA
//                    fills, output staging, residual tiles and TMA-store reads add 86 KB per tile;
//   128->128 @28x28: ~100 cycles per M128 N128 K16 MMA instead of 64 -- 8 KB per MMA plus a 16 KB weight tile streamed
//                    per tap.
// Here two CTAs issue ONE M = 256 MMA (tcgen05 cta_group::2): each CTA reads its own 128 A rows but only HALF of the
// weight tile (5 KB instead of 6 KB per MMA at N = 64, 6 KB instead of 8 KB at N = 128), the filter bank is loaded
// once per kernel (each CTA keeps the [Cout/2][9][Cin] half it feeds to both tensor cores: 36 KB .. 144 KB), and the
// epilogue goes TMEM -> registers -> global memory directly (64 contiguous bytes per thread; the residual is
// prefetched from global memory into registers before the accumulator is waited for), so the only shared-memory
// traffic left besides the MMA operand reads is the TMA fill of the input halo tile.
//
//   * M tile = hb whole image rows at halo pitch W + 2, as in conv_halo.cu: one 4-D TMA box per 64-channel chunk
//     (zero fill = padding); tap (r, s) = the same tile read (r * pitch + s) * 128 bytes further in;
//   * work unit = two consecutive M tiles, one per CTA of the pair; persistent clusters (74 x 2 CTAs);
//   * warp 0 = TMA producer (both CTAs; all loads complete on the LEADER's barriers), warp 1 = TMEM allocation + MMA
//     issuer (leader only; multicast tcgen05.commit frees stages / publishes accumulators in both CTAs), warps 3-10 =
//     epilogue (two warps per TMEM lane quarter, 32 of a slab's 64 channels each); accumulators double-buffered.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "spk_internal.h"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace spk {
namespace {
using namespace tc;

constexpr int kEpiWarps = 8;
constexpr int kThreads = 96 + kEpiWarps * 32;
constexpr int kMaxRing = 8;
constexpr size_t kSmemMax = 232448;  // 227 KB opt-in maximum per CTA on sm_100

struct alignas(64) HpParams {
  CUtensorMap map_x, map_w;
  const float* bias;
  __nv_bfloat16* y;
  const __nv_bfloat16* res;
  int n, h, w, cin, cout, relu, has_res;
  int ldy, ldres;
  int pitch, hb, tiles_h, kchunks;
  int m_tiles;      // n * tiles_h
  int units;        // ceil(m_tiles / 2)
  int a_stage;      // bytes per A stage (multiple of 1024)
  int a_box_bytes;  // (hb + 2) * (w + 2) * 128
  int na;           // A ring depth
  int reverse;      // walk the units from the last image to the first: the tensors the previous kernel touched last
                    // are still in L2 (consecutive layers alternate direction)
  unsigned long long* stamp;  // profiling (stamp mode): global-timer slot of this launch, else nullptr
  int dbg;          // debug build only (SPK_HP_DBG): 1 = skip the global stores, 2 = skip the residual loads
  long long* trace; // debug (SPK_HP_TRACE=1): clock64 stamps of the leader MMA issuer of cluster 3, else nullptr
};

// 256-bit global accesses (sm_100): a thread's 32 channels are two full 32-byte sectors, so every request moves
// whole sectors (with 16-byte accesses each sector is fetched twice: measured 0.089 ms vs 0.060 ms without residual)
struct alignas(32) U8 {
  uint32_t v[8];
};
__device__ __forceinline__ U8 ldg_nc_v8(const void* p) {
  U8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_v8(void* p, const U8& r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]),
               "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
               : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) conv3x3_hp_kernel(const __grid_constant__ HpParams p) {
  constexpr int kBHalf = (BN / 2) * 128;  // bytes of this CTA's half of one (chunk, tap) weight tile
  constexpr int kTmemCols = 2 * BN;       // double-buffered accumulator
  constexpr int kSlabs = BN >= 64 ? BN / 64 : 1;  // 64-channel slabs (BN = 32: one half slab, taken by the `half == 0` warps)
  constexpr int kTmemAlloc = kTmemCols < 32 ? 32 : kTmemCols;
  constexpr uint32_t kIdesc = idesc_bf16(256, BN);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gen = smem_raw + (base - raw);
  const uint32_t size_b = 9u * p.kchunks * kBHalf;
  const uint32_t b_s = base;
  const uint32_t a_s = b_s + size_b;
  const uint32_t bias_off = size_b + (uint32_t)p.na * p.a_stage;
  const uint32_t bar0 = base + ((bias_off + (uint32_t)p.cout * 4u + 15u) & ~15u);
  float* bias_sm = reinterpret_cast<float*>(gen + bias_off);
  auto a_full = [&](int s) { return bar0 + 8u * s; };                 // leader only: both CTAs' halo boxes
  auto a_empty = [&](int s) { return bar0 + 8u * (kMaxRing + s); };   // multicast tcgen05.commit
  auto t_full = [&](int s) { return bar0 + 8u * (2 * kMaxRing + s); };       // multicast tcgen05.commit
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * kMaxRing + 2 + s); };  // leader only: epilogue warps of both CTAs
  const uint32_t b_full = bar0 + 8u * (2 * kMaxRing + 4);                    // leader only: both halves of the filter bank
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 8 * (2 * kMaxRing + 5));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  pdl_trigger();
  stamp_begin(p.stamp);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), 2 * kEpiWarps);
    }
    mbar_init(b_full, 1);
    mbar_init_fence();
    tma_prefetch_desc(&p.map_x);
    tma_prefetch_desc(&p.map_w);
  }
  if (warp == 1) tmem2_alloc(smem_u32((const void*)tmem_slot), kTmemAlloc);
  for (int i = threadIdx.x; i < p.cout; i += kThreads) bias_sm[i] = __ldg(p.bias + i);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> this CTA's M tile: image and first output row (img == p.n: the odd tile out, TMA zero-fills, nothing stored)
  auto tile_coords = [&](int u, int& img, int& h0) {
    const int m = 2 * (p.reverse ? p.units - 1 - u : u) + (int)rank;
    if (m >= p.m_tiles) {
      img = p.n;
      h0 = 0;
      return false;
    }
    img = m / p.tiles_h;
    h0 = (m - img * p.tiles_h) * p.hb;
    return true;
  };

  if (warp == 0) {
    // ===== TMA producer (one thread per CTA) =====
    if (lane == 0) {
      // the filter bank, once: rows [rank * BN/2, +BN/2) of every (chunk, tap) tile; both halves land on the leader's barrier
      const uint32_t bfull_leader = mapa_rank(b_full, 0);
      if (leader) mbar_expect_tx(b_full, 2u * size_b);
      for (int ch = 0; ch < p.kchunks; ++ch)
        for (int tap = 0; tap < 9; ++tap)
          tma2_load_2d(b_s + (uint32_t)(ch * 9 + tap) * kBHalf, &p.map_w, bfull_leader, tap * p.cin + ch * 64, (int)rank * (BN / 2));
      pdl_wait();  // the filter bank does not depend on the previous kernel; the activations do
      int as = 0;
      uint32_t aph = 0;
      for (int u = cluster_id; u < p.units; u += n_clusters) {
        int img, h0;
        tile_coords(u, img, h0);
        for (int ch = 0; ch < p.kchunks; ++ch) {
          mbar_wait(a_empty(as), aph ^ 1u);
          if (leader) mbar_expect_tx(a_full(as), 2u * (uint32_t)p.a_box_bytes);
          tma2_load_4d(a_s + (uint32_t)as * p.a_stage, &p.map_x, mapa_rank(a_full(as), 0), ch * 64, -1, h0 - 1, img);
          if (++as == p.na) {
            as = 0;
            aph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA): the whole warp runs the loop with warp-uniform values, one elected lane issues.
    // A satisfied mbarrier wait + tcgen05 fence still costs ~150-200 cycles of this warp, during which the MMA queue
    // would run dry.  The waits for the NEXT chunk (its halo stage, and the accumulator buffer when it starts a unit)
    // are therefore done after the FIFTH tap of the current chunk, while four taps' worth of MMAs are still queued.
    // Trace (SPK_HP_TRACE, 64 -> 64 at 56x56): 20 MMAs issue in ~540 cycles, t_empty ~220, a_full ~160, fence ~90,
    // 16 MMAs ~450: ~1500 cycles per unit -- which is this shape's shared-memory bound, not the waits: each K=16 MMA
    // reads 128 x 32 B of A and 32 x 32 B of B per CTA (5 KB, 40 cycles at 128 B/clk), 36 of them = 1440 cycles.
    // With a residual the t_empty wait grows to 900-1400 cycles (the epilogue, i.e. HBM, is the limiter there; an L2
    // prefetch of the residual two units ahead changed nothing). =====
    if (leader) {
      mbar_wait(b_full, 0);
      int as = 0, acc = 0, tr = 0;
      uint32_t aph = 0, accph = 0;
      const bool tracing = p.trace != nullptr && cluster_id == 3 && lane == 0;
      // flat sequence of chunks: chunk c of unit u; `ready` = the barriers of the chunk about to be issued have been waited for
      auto wait_chunk = [&](int ch, int as_, uint32_t aph_, int acc_, uint32_t accph_) {
        if (tracing && tr < 250) p.trace[tr++] = clock64();
        if (ch == 0) mbar_wait_cluster(t_empty(acc_), accph_ ^ 1u);  // both CTAs have drained this accumulator buffer
        if (tracing && tr < 250) p.trace[tr++] = clock64();
        mbar_wait(a_full(as_), aph_);
        if (tracing && tr < 250) p.trace[tr++] = clock64();
        tc_fence_after();
        if (tracing && tr < 250) p.trace[tr++] = clock64();
      };
      if (cluster_id < p.units) wait_chunk(0, as, aph, acc, accph);
      for (int u = cluster_id; u < p.units; u += n_clusters) {
        if (tracing && tr < 250) p.trace[tr++] = clock64();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int ch = 0; ch < p.kchunks; ++ch) {
          const uint32_t sa = a_s + (uint32_t)as * p.a_stage;
          // what comes after this chunk
          int as_n = as + 1;
          uint32_t aph_n = aph;
          if (as_n == p.na) {
            as_n = 0;
            aph_n ^= 1u;
          }
          const bool last_ch = ch + 1 == p.kchunks;
          const bool has_next = !last_ch || (u + n_clusters < p.units);
          int acc_n = acc;
          uint32_t accph_n = accph;
          if (last_ch && ++acc_n == 2) {
            acc_n = 0;
            accph_n ^= 1u;
          }
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap - 3 * r;
            // tap (r, s) = the halo tile read (r * pitch + s) rows further in (same offset in both CTAs)
            const uint64_t a_desc = smem_desc_sw128(sa + (uint32_t)(r * p.pitch + s) * 128u);
            const uint64_t b_desc = smem_desc_sw128(b_s + (uint32_t)(ch * 9 + tap) * kBHalf);
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 16 bf16 = 32 bytes along K: +2 in the (addr >> 4) field
              tc2_mma_w(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), kIdesc, (ch | tap | k) != 0 ? 1u : 0u);
            if (tap == 4 && has_next) wait_chunk(last_ch ? 0 : ch + 1, as_n, aph_n, acc_n, accph_n);
          }
          tc2_commit_mc_w(a_empty(as), 3);  // frees the halo stage in both CTAs once these MMAs have read it
          as = as_n;
          aph = aph_n;
        }
        tc2_commit_mc_w(t_full(acc), 3);  // accumulator complete, both CTAs
        if (++acc == 2) {
          acc = 0;
          accph ^= 1u;
        }
      }
    }
  } else if (warp >= 3) {
    // ===== epilogue: warps 3-10; warp w may touch TMEM lanes [32 * (w % 4), +32) and takes channels [32 * half, +32)
    // of every 64-channel slab =====
    const int q = warp & 3;
    const int half = (warp - 3) >> 2;
    const int j = q * 32 + lane;  // TMEM lane == linear halo-pitch pixel of this CTA's tile
    const int jr = j / p.pitch, jc = j - jr * p.pitch;
    const uint32_t t_empty_leader = mapa_rank(t_empty(0), 0);
    const bool works = BN >= 64 || half == 0;  // BN = 32: the second warp of a lane quarter only keeps the barrier counts
    pdl_wait();  // residual reads and output writes wait for the previous kernel
    int acc = 0;
    uint32_t accph = 0;
    // residual: this thread's 32 channels of every slab, straight from global memory into registers ONE UNIT AHEAD
    // (an L2 / HBM round trip is about as long as a unit's MMAs)
    auto res_load = [&](int u, U8 (&r)[kSlabs][2]) {
      int img, h0;
      const bool tile_ok = tile_coords(u, img, h0);
      if (works && tile_ok && (jc < p.w) && (jr < p.hb) && (h0 + jr < p.h) && !(p.dbg & 2)) {
        const __nv_bfloat16* rp = p.res + (((long long)img * p.h + (h0 + jr)) * p.w + jc) * p.ldres + half * 32;
#pragma unroll
        for (int slab = 0; slab < kSlabs; ++slab)
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) r[slab][c8] = ldg_nc_v8(rp + slab * 64 + c8 * 16);
      }
    };
    U8 rnext[kSlabs][2];
#pragma unroll
    for (int slab = 0; slab < kSlabs; ++slab)
#pragma unroll
      for (int c8 = 0; c8 < 2; ++c8)
#pragma unroll
        for (int e = 0; e < 8; ++e) rnext[slab][c8].v[e] = 0u;
    if (p.has_res && cluster_id < p.units) res_load(cluster_id, rnext);
    for (int u = cluster_id; u < p.units; u += n_clusters) {
      int img, h0;
      const bool tile_ok = tile_coords(u, img, h0);
      const bool inside = works && tile_ok && (jc < p.w) && (jr < p.hb) && (h0 + jr < p.h);
      const long long pix = ((long long)img * p.h + (h0 + jr)) * p.w + jc;
      U8 rv[kSlabs][2];
#pragma unroll
      for (int slab = 0; slab < kSlabs; ++slab)
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8) rv[slab][c8] = rnext[slab][c8];
      if (p.has_res && u + n_clusters < p.units) res_load(u + n_clusters, rnext);
      mbar_wait(t_full(acc), accph);
      tc_fence_after();
#pragma unroll
      for (int slab = 0; slab < kSlabs; ++slab) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + slab * 64 + (BN >= 64 ? half * 32 : 0));
        uint32_t v[32];
        tmem_ld32(taddr, v);  // (BN = 32: both warps of a quarter read the same 32 columns; only one of them stores)
        tmem_ld_wait();
        if (slab == kSlabs - 1) {  // accumulator buffer drained: tell the leader's MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(t_empty_leader + 8u * acc);
        }
        if (inside) {
          const float* bsm = bias_sm + slab * 64 + half * 32;
          __nv_bfloat16* yp = p.y + pix * p.ldy + slab * 64 + half * 32;
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) {
            U8 o;
#pragma unroll
            for (int c4 = 0; c4 < 2; ++c4) {
              float f[8];
              const float* bq = bsm + c8 * 16 + c4 * 8;
              const float4 ba = *reinterpret_cast<const float4*>(bq), bb = *reinterpret_cast<const float4*>(bq + 4);
              const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[c8 * 16 + c4 * 8 + e]) + bv[e];
              if (p.has_res) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 rf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rv[slab][c8].v[c4 * 4 + t]));
                  f[2 * t] += rf.x;
                  f[2 * t + 1] += rf.y;
                }
              }
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
              }
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * t], f[2 * t + 1]);
                o.v[c4 * 4 + t] = *reinterpret_cast<const uint32_t*>(&b2);
              }
            }
            if (!(p.dbg & 1)) stg_v8(yp + c8 * 16, o);
          }
        }
      }
      if (++acc == 2) {
        acc = 0;
        accph ^= 1u;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still read this CTA's shared memory
  stamp_end(p.stamp);
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, kTmemAlloc);
  }
}

int halo_rows(const ConvGeom& g, int* hb_out) {
  const int pitch = g.w + 2;
  int hb = 128 / pitch;
  if (hb > g.ho) hb = g.ho;
  *hb_out = hb;
  return pitch;
}

// shared-memory plan: -> A ring depth (0 = does not fit)
int plan_smem(const ConvGeom& g, int* a_stage_out, size_t* smem_out) {
  int hb;
  const int pitch = halo_rows(g, &hb);
  const int kchunks = g.cin / 64;
  const int box = (hb + 2) * pitch * 128;
  int a_stage = ((2 * pitch + 2 + 128) * 128 + 1023) & ~1023;  // the last tap reads 128 rows from row 2 * pitch + 2
  if (a_stage < box) a_stage = (box + 1023) & ~1023;
  const size_t size_b = (size_t)9 * kchunks * (g.cout / 2) * 128;
  const size_t fixed = 1024 /*alignment*/ + (size_t)g.cout * 4 + 16 + 8 * (2 * kMaxRing + 6) + 16;
  if (size_b + fixed + (size_t)a_stage > kSmemMax) return 0;
  int na = (int)std::min<size_t>(kMaxRing, (kSmemMax - fixed - size_b) / (size_t)a_stage);
  if (na > 3 * kchunks) na = 3 * kchunks;  // (6 stages measured the same as 3: the a_full wait in the trace is the ~160-cycle
                                           // cost of a satisfied wait, not TMA latency)
  if (na < kchunks + 1 && na < 3) return 0;  // the next tile's first chunk must be loadable while this tile computes
  *a_stage_out = a_stage;
  *smem_out = fixed + size_b + (size_t)na * a_stage;
  return na;
}

}  // namespace

struct HpConvPlan {
  ConvGeom g;
  HpParams prm;
  __nv_bfloat16* d_w = nullptr;
  int64_t bytes = 0;
  size_t smem = 0;
  const void* x_ptr = nullptr;
};

bool hp_conv_supported(const ConvGeom& g) {
  static const bool off = debug_env("SPK_NO_HP") != nullptr;  // A/B switch
  if (off) return false;
  if (g.kh != 3 || g.kw != 3 || g.stride != 1 || g.pad != 1) return false;
  if ((g.cout != 32 && g.cout != 64 && g.cout != 128) || (g.cin != 64 && g.cin != 128)) return false;
  if (g.ldx % 8 != 0 || g.ldy % 16 != 0 || g.ldres % 16 != 0) return false;  // TMA rows; 32-byte epilogue accesses
  if (g.w + 2 > 128 || g.ho != g.h || g.wo != g.w) return false;
  int hb;
  halo_rows(g, &hb);
  if (hb < 1) return false;
  const int tiles_h = (g.h + hb - 1) / hb;
  const double eff = (double)(g.w * hb) / 128.0 * (double)g.h / (double)(tiles_h * hb);
  if (eff < 0.75) return false;  // share of the 128 MMA rows that are real output pixels (14x14 maps: 0.77)
  if (encode_fn() == nullptr) return false;
  int a_stage;
  size_t smem;
  return plan_smem(g, &a_stage, &smem) > 0;
}

int hp_conv_plan_create(spk_ctx* ctx, const ConvGeom& g_max, const float* w, const float* d_bias, HpConvPlan** out) {
  if (!hp_conv_supported(g_max)) return fail(ctx, SPK_ERR_UNSUPPORTED, "halo-pair convolution: unsupported geometry");
  HpConvPlan* p = new HpConvPlan;
  p->g = g_max;
  const ConvGeom& g = p->g;
  memset(&p->prm, 0, sizeof p->prm);
  HpParams& prm = p->prm;
  prm.bias = d_bias;
  prm.h = g.h;
  prm.w = g.w;
  prm.cin = g.cin;
  prm.cout = g.cout;
  prm.relu = g.relu;
  prm.ldy = g.ldy;
  prm.ldres = g.ldres;
  prm.pitch = halo_rows(g, &prm.hb);
  prm.tiles_h = (g.h + prm.hb - 1) / prm.hb;
  prm.kchunks = g.cin / 64;
  prm.a_box_bytes = (prm.hb + 2) * prm.pitch * 128;
  prm.na = plan_smem(g, &prm.a_stage, &p->smem);

  // ---- weights: bf16 [Cout][tap][Cin], round to nearest
  const size_t kk = (size_t)9 * g.cin;
  std::vector<__nv_bfloat16> wb16((size_t)g.cout * kk);
  for (size_t i = 0; i < wb16.size(); ++i) wb16[i] = __float2bfloat16(w[i]);
  cudaError_t e = cudaMalloc(&p->d_w, wb16.size() * 2);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_w, wb16.data(), wb16.size() * 2, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    hp_conv_plan_destroy(p);
    return fail(ctx, SPK_ERR_CUDA, "halo-pair convolution: weight upload: %s", cudaGetErrorString(e));
  }
  p->bytes = (int64_t)wb16.size() * 2;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kk, (cuuint64_t)g.cout};
    cuuint64_t strides[1] = {(cuuint64_t)kk * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)(g.cout / 2)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&prm.map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->d_w, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      hp_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "halo-pair convolution: cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    }
  }
  cudaError_t ea = g.cout == 128  ? cudaFuncSetAttribute(conv3x3_hp_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax)
                   : g.cout == 64 ? cudaFuncSetAttribute(conv3x3_hp_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax)
                                  : cudaFuncSetAttribute(conv3x3_hp_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
  if (ea != cudaSuccess) {
    hp_conv_plan_destroy(p);
    return fail(ctx, SPK_ERR_CUDA, "halo-pair convolution: cudaFuncSetAttribute: %s", cudaGetErrorString(ea));
  }
  *out = p;
  return SPK_OK;
}

void hp_conv_plan_destroy(HpConvPlan* p) {
  if (!p) return;
  if (p->d_w) cudaFree(p->d_w);
  delete p;
}

int64_t hp_conv_plan_bytes(const HpConvPlan* p) { return p ? p->bytes : 0; }

void hp_conv_plan_set_reverse(HpConvPlan* p, int reverse) {
  if (p) p->prm.reverse = reverse ? 1 : 0;
}

int hp_conv_launch(spk_ctx* ctx, HpConvPlan* p, int n, const void* x, const void* res, void* y) {
  if (n <= 0) return SPK_OK;
  const ConvGeom& g = p->g;
  if (n > g.n) return fail(ctx, SPK_ERR_CAPACITY, "halo-pair convolution: batch %d > planned %d", n, g.n);
  HpParams& prm = p->prm;
  if (x != p->x_ptr) {
    CUresult r = encode_nhwc_bf16(&prm.map_x, x, g.cin, g.w, g.h, g.n, g.ldx, 64, g.w + 2, prm.hb + 2, 1);
    if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "halo-pair convolution: tensor map (x) failed: %d", (int)r);
    p->x_ptr = x;
  }
  if (((uintptr_t)y & 31) || (res && ((uintptr_t)res & 31)))
    return fail(ctx, SPK_ERR_INVALID, "halo-pair convolution: output / residual not 32-byte aligned");
  prm.y = reinterpret_cast<__nv_bfloat16*>(y);
  prm.res = reinterpret_cast<const __nv_bfloat16*>(res);
  prm.has_res = res ? 1 : 0;
  prm.n = n;
  prm.m_tiles = n * prm.tiles_h;
  prm.units = (prm.m_tiles + 1) / 2;
  const int clusters = std::min(prm.units, ctx->sm_count / 2);
  static const int dbg = debug_env("SPK_HP_DBG") ? atoi(debug_env("SPK_HP_DBG")) : 0;
  prm.dbg = dbg;
  prm.stamp = ctx->cur_stamp;
  static const bool want_trace = debug_env("SPK_HP_TRACE") != nullptr;
  static long long* d_trace = nullptr;
  static int trace_left = 6;
  prm.trace = nullptr;
  if (want_trace && trace_left > 0) {
    if (!d_trace) cudaMalloc(&d_trace, 256 * sizeof(long long));
    cudaMemsetAsync(d_trace, 0, 256 * sizeof(long long), ctx->stream);
    prm.trace = d_trace;
  }
  if (g.cout == 128)
    SPK_CUDA_OK(ctx, launch_pdl(conv3x3_hp_kernel<128>, dim3(2 * clusters), dim3(kThreads), p->smem, ctx->stream, prm));
  else if (g.cout == 32)
    SPK_CUDA_OK(ctx, launch_pdl(conv3x3_hp_kernel<32>, dim3(2 * clusters), dim3(kThreads), p->smem, ctx->stream, prm));
  else
    SPK_CUDA_OK(ctx, launch_pdl(conv3x3_hp_kernel<64>, dim3(2 * clusters), dim3(kThreads), p->smem, ctx->stream, prm));
  SPK_LAUNCH_CHECK(ctx);
  if (prm.trace) {
    --trace_left;
    long long h[256];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(h, d_trace, sizeof h, cudaMemcpyDeviceToHost);
    fprintf(stderr, "hp trace cin=%d cout=%d w=%d units=%d na=%d res=%d (cycles since the first stamp; per unit: start, then mid-chunk: before t_empty, after, after a_full, after fence):\n ",
            g.cin, g.cout, g.w, prm.units, prm.na, prm.has_res);
    for (int i = 0; i < 250 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
    fprintf(stderr, "\n");
  }
  return SPK_OK;
}

}  // namespace spk

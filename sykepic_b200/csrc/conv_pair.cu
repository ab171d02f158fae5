// K2 (Cout >= 128): implicit-GEMM convolution on CTA PAIRS -- tcgen05.mma.cta_group::2.
//
// Replaces the conv / BatchNorm / ReLU / add sequence of torchvision's ResNet blocks as run by
// TorchVisionNet.forward (sykepic/train/network.py:66-68); library calls in the reference.
//
// Why pairs: with both operands in shared memory a single-CTA tcgen05.mma reads its whole A (128 x 16) and
// B (N x 16) slices for every instruction, and TMA writes the same bytes once more: at N = 256 that is
// 768 shared-memory wavefronts per 512 tensor-pipe cycles (measured: the tap-per-TMA kernel saturates at
// ~65 % of the tensor peak, 1 wavefront = 128 B per cycle per SM); at N = 128 the MMA reads alone need 100 %.
// Two CTAs of a cluster issue ONE M = 256 MMA: each CTA stages its own 128 pixels of A and only HALF of
// the weight tile, the tensor cores of both SMs read both halves.  Per CTA the traffic drops to
// 512 wavefronts per 512 cycles at N = 256.
//
//   * work unit = two M tiles (boxes of 128 output pixels, one per CTA) x one N tile (256 or 128 channels);
//   * A: one 4-D TMA box per filter tap and 64-channel chunk (as conv_tc.cu; zero fill = padding, one
//     tensor map per stride-2 parity); B: rows [rank * BN/2, +BN/2) of the tile; all loads of both CTAs
//     complete on the LEADER's mbarrier (cp.async.bulk.tensor ... .cta_group::2);
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 and frees the stage / publishes the
//     accumulator in BOTH CTAs with multicast tcgen05.commit; each CTA's TMEM holds its 128 rows;
//   * epilogue per CTA as in conv_halo.cu: TMEM -> registers -> (+bias, +residual, ReLU, bf16) -> swizzled
//     shared staging -> one TMA store per 64-channel slab; residual tiles arrive by TMA from their own
//     producer warp; the "accumulator drained" barrier lives in the leader and counts the epilogue warps of
//     both CTAs (remote mbarrier.arrive).
// Persistent clusters (74 x 2 CTAs), warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocation +
// MMA issuer (leader only), warp 2 = residual producer, warps 3-10 = epilogue.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "spk_internal.h"
#include "tc_common.cuh"
#include "tc_pair.cuh"

namespace spk {
namespace {
using namespace tc;

constexpr int kEpiWarps = 8;  // two per TMEM lane quarter, 32 of a slab's 64 channels each
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 96 + kEpiThreads;
constexpr int kBK = 64;
constexpr int kMaxTaps = 9;
constexpr int kABytes = 128 * kBK * 2;
constexpr int kIoSlot = 128 * 128;  // one 128-pixel x 64-channel slab
constexpr int kMaxStages = 8;
constexpr int kMaxResSlots = 4;  // residual slabs in flight per CTA (64-channel slabs of 16 KB)
constexpr size_t kSmemMax = 232448;

struct alignas(64) PairParams {
  CUtensorMap map_a[4];
  CUtensorMap map_b, map_y, map_res;
  CUtensorMap map_b2, map_y2;  // fused 1x1 / stride-2 shortcut: weights [Cout][cin_pad] on the centre tap's A tiles, its own output
  const float* bias;
  const float* bias2;
  int ds_tap;
  int n, ho, wo, cout, relu, has_res;
  int wb, hb, nb;
  int tiles_w, tiles_h, tiles_img, tiles_n;
  int m_tiles, units;
  int taps, kchunks, cin_pad;
  int stages, res_slots;
  int reverse;  // walk the units last-to-first (consecutive layers alternate direction: the previous kernel's last tiles are in L2)
  unsigned long long* stamp;  // profiling (stamp mode): global-timer slot of this launch, else nullptr
  long long* trace;  // debug (SPK_PAIR_TRACE=1): clock64 stamps of one CTA's epilogue thread 0
  signed char tap_map[kMaxTaps + 3], tap_dh[kMaxTaps + 3], tap_dw[kMaxTaps + 3];
};

// DS: the block's 1x1 / stride-2 shortcut rides on the centre tap of this 3x3 / stride-2 convolution (same pixels): a second
// weight half-tile per centre-tap stage, a second accumulator beside the main one (BN = 128: 2 x 2 x 128 TMEM columns).
template <int BN, bool DS = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) conv_pair_kernel(const __grid_constant__ PairParams p) {
  static_assert(!DS || BN == 128, "the fused shortcut needs two accumulators per buffer");
  constexpr int kBHalfBytes = (BN / 2) * kBK * 2;
  constexpr int kStageBytes = kABytes + kBHalfBytes * (DS ? 2 : 1);
  constexpr int kAccCols = DS ? 2 * BN : BN;
  constexpr int kTmemCols = 2 * kAccCols;  // double-buffered accumulator(s): 512 or 256 columns
  constexpr int kSlabs = BN / 64;
  constexpr uint32_t kIdesc = idesc_bf16(256, BN);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gen = smem_raw + (base - raw);
  const uint32_t res_off = (uint32_t)p.stages * kStageBytes;
  const uint32_t out_off = res_off + (uint32_t)p.res_slots * kIoSlot;  // no residual ring when the layer has none
  const uint32_t bias_off = out_off + 2u * kIoSlot;
  const uint32_t bar0 = base + ((bias_off + (uint32_t)p.cout * 4u + 15u) & ~15u);
  float* bias_sm = reinterpret_cast<float*>(gen + bias_off);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto r_full = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
  auto r_empty = [&](int s) { return bar0 + 8u * (2 * kMaxStages + kMaxResSlots + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 2 * kMaxResSlots + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 2 * kMaxResSlots + 2 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 8 * (2 * kMaxStages + 2 * kMaxResSlots + 4));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  pdl_trigger();
  stamp_begin(p.stamp);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(full_bar(s), 1);   // used in the leader only: its producer's arrive.expect_tx
      mbar_init(empty_bar(s), 1);  // multicast tcgen05.commit
    }
    for (int s = 0; s < kMaxResSlots; ++s) {
      mbar_init(r_full(s), 1);
      mbar_init(r_empty(s), kEpiWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full(s), 1);               // multicast tcgen05.commit
      mbar_init(t_empty(s), 2 * kEpiWarps);  // used in the leader only: the epilogue warps of both CTAs
    }
    mbar_init_fence();
    tma_prefetch_desc(&p.map_a[0]);
    tma_prefetch_desc(&p.map_b);
    tma_prefetch_desc(&p.map_y);
    if (p.has_res) tma_prefetch_desc(&p.map_res);
    if (DS) {
      tma_prefetch_desc(&p.map_b2);
      tma_prefetch_desc(&p.map_y2);
    }
  }
  if (warp == 1) tmem2_alloc(smem_u32((const void*)tmem_slot), kTmemCols);
  for (int i = threadIdx.x; i < p.cout; i += kThreads) bias_sm[i] = __ldg(p.bias + i);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything below reads / writes tensors the previous kernel may still be using

  // unit -> this CTA's M tile (box origin) and the N tile
  auto decode = [&](int u, int& nt, int& w0, int& h0, int& n0) {
    if (p.reverse) u = p.units - 1 - u;
    nt = u % p.tiles_n;
    int m = (u / p.tiles_n) * 2 + (int)rank;
    if (m >= p.m_tiles) {  // the odd tile out: a box past the batch -- TMA zero-fills the loads and clips the store
      w0 = 0;
      h0 = 0;
      n0 = p.tiles_img * p.nb;
      return;
    }
    const int twi = m % p.tiles_w;
    m /= p.tiles_w;
    const int thi = m % p.tiles_h;
    const int ng = m / p.tiles_h;
    w0 = twi * p.wb;
    h0 = thi * p.hb;
    n0 = ng * p.nb;
  };
  const int kblocks = p.taps * p.kchunks;

  if (warp == 0) {
    // ===== TMA producer (per CTA): whole warp, one elected lane issues =====
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < p.units; u += n_clusters) {
        int nt, w0, h0, n0;
        decode(u, nt, w0, h0, n0);
        for (int kb = 0; kb < kblocks; ++kb) {
          const int tap = kb / p.kchunks;
          const int c0 = (kb - tap * p.kchunks) * kBK;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * kStageBytes;
          const uint32_t full_leader = mapa_rank(full_bar(stage), 0);
          const bool ds = DS && tap == p.ds_tap;
          if (leader) mbar_expect_tx_w(full_bar(stage), 2u * (uint32_t)(kABytes + kBHalfBytes * (ds ? 2 : 1)));  // both CTAs' bytes land on this barrier
          tma2_load_4d_w(sa, &p.map_a[p.tap_map[tap]], full_leader, c0, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], n0);
          tma2_load_2d_w(sa + kABytes, &p.map_b, full_leader, tap * p.cin_pad + c0, nt * BN + (int)rank * (BN / 2));
          if (ds) tma2_load_2d_w(sa + kABytes + kBHalfBytes, &p.map_b2, full_leader, c0, nt * BN + (int)rank * (BN / 2));
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA): the whole warp runs the loop with warp-uniform values, one elected lane issues =====
    if (leader) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, accph = 0;
      for (int u = cluster_id; u < p.units; u += n_clusters) {
        mbar_wait_cluster(t_empty(acc), accph ^ 1u);  // both CTAs have drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * kStageBytes;
          const uint64_t a_desc = smem_desc_sw128(sa);
          const uint64_t b_desc = smem_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            tc2_mma_w(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), kIdesc, (kb | k) != 0 ? 1u : 0u);
          if (DS) {
            const int tap = kb / p.kchunks;
            if (tap == p.ds_tap) {
              const int c = kb - tap * p.kchunks;
              const uint64_t b2_desc = smem_desc_sw128(sa + kABytes + kBHalfBytes);
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                tc2_mma_w(d_tmem + (uint32_t)BN, a_desc + (uint64_t)(2 * k), b2_desc + (uint64_t)(2 * k), kIdesc, (c | k) != 0 ? 1u : 0u);
            }
          }
          tc2_commit_mc_w(empty_bar(stage), 3);  // frees the stage in both CTAs
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc2_commit_mc_w(t_full(acc), 3);  // accumulator complete, both CTAs
        if (++acc == 2) {
          acc = 0;
          accph ^= 1u;
        }
      }
    }
  } else if (warp == 2) {
    // ===== residual producer (one thread per CTA, local barriers) =====
    if (lane == 0 && p.has_res) {
      int rs = 0;
      uint32_t rph = 0;
      for (int u = cluster_id; u < p.units; u += n_clusters) {
        int nt, w0, h0, n0;
        decode(u, nt, w0, h0, n0);
        for (int slab = 0; slab < kSlabs; ++slab) {
          mbar_wait(r_empty(rs), rph ^ 1u);
          mbar_expect_tx(r_full(rs), (uint32_t)kIoSlot);
          tma_load_4d(base + res_off + (uint32_t)rs * kIoSlot, &p.map_res, r_full(rs), nt * BN + slab * 64, w0, h0, n0);
          if (++rs == p.res_slots) {
            rs = 0;
            rph ^= 1u;
          }
        }
      }
    }
  } else {
    // ===== epilogue: warps 3-10; warp w may touch TMEM lanes [32 * (w % 4), +32) and takes channels
    // [32 * half, +32) of every 64-channel slab =====
    const int q = warp & 3;
    const int half = (warp - 3) >> 2;
    const int et = threadIdx.x - 96;
    const int row = q * 32 + lane;  // TMEM lane == pixel of the box == row of the staging tile
    const uint32_t sw = (uint32_t)(row & 7);
    int acc = 0, rs = 0, os = 0;
    uint32_t accph = 0, rph = 0;
    int tr = 0;
    const bool tracing = p.trace != nullptr && blockIdx.x == 10 && et == 0;
#ifdef SPK_PAIR_TRACE_BUILD  // epilogue clock trace: compiled in on request only
#define PAIR_TRACE() do { if (tracing && tr < 250) p.trace[tr++] = clock64(); } while (0)
#else
#define PAIR_TRACE() do { (void)tracing; (void)tr; } while (0)
#endif
    for (int u = cluster_id; u < p.units; u += n_clusters) {
      int nt, w0, h0, n0;
      decode(u, nt, w0, h0, n0);
      PAIR_TRACE();
      mbar_wait(t_full(acc), accph);
      tc_fence_after();
      PAIR_TRACE();
#pragma unroll 1
      for (int s2 = 0; s2 < kSlabs * (DS ? 2 : 1); ++s2) {
        const bool dsp = DS && s2 >= kSlabs;  // the shortcut's accumulator: its own bias and output, no residual, no ReLU
        const int slab = dsp ? s2 - kSlabs : s2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccCols + (dsp ? BN : 0) + slab * 64 + half * 32);
        uint32_t v[32];
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        if (s2 == kSlabs * (DS ? 2 : 1) - 1) {  // accumulator buffer drained: tell the leader's MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_rank(t_empty(acc), 0));
        }
        PAIR_TRACE();
        if (et == 0) tma_store_wait_read<1>();  // staging slot `os` was last read by the store issued two slabs ago
        PAIR_TRACE();
        const bool with_res = p.has_res && !dsp;
        if (with_res) mbar_wait(r_full(rs), rph);
        PAIR_TRACE();
        named_bar_sync(1, kEpiThreads);
        PAIR_TRACE();
        {
          const float* bsm = dsp ? p.bias2 + nt * BN + slab * 64 + half * 32 : bias_sm + nt * BN + slab * 64 + half * 32;
          unsigned char* orow_p = gen + out_off + (uint32_t)os * kIoSlot + (uint32_t)row * 128u;
          const unsigned char* rrow_p = gen + res_off + (uint32_t)rs * kIoSlot + (uint32_t)row * 128u;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            float f[8];
            const float4 ba = *reinterpret_cast<const float4*>(bsm + c4 * 8), bb = *reinterpret_cast<const float4*>(bsm + c4 * 8 + 4);
            const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[c4 * 8 + e]) + bv[e];
            const uint32_t chunk = ((uint32_t)(half * 4 + c4) ^ sw) << 4;
            if (with_res) {
              const uint4 r4 = *reinterpret_cast<const uint4*>(rrow_p + chunk);
              const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 rf = __bfloat1622float2(r2[t]);
                f[2 * t] += rf.x;
                f[2 * t + 1] += rf.y;
              }
            }
            if (p.relu && !dsp) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            uint4 o;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(f[2], f[3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(f[6], f[7]);
            o.x = *reinterpret_cast<uint32_t*>(&t0);
            o.y = *reinterpret_cast<uint32_t*>(&t1);
            o.z = *reinterpret_cast<uint32_t*>(&t2);
            o.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(orow_p + chunk) = o;
          }
        }
        if (with_res) {
          __syncwarp();
          if (lane == 0) mbar_arrive(r_empty(rs));
          if (++rs == p.res_slots) {
            rs = 0;
            rph ^= 1u;
          }
        }
        PAIR_TRACE();
        fence_async_smem();
        named_bar_sync(2, kEpiThreads);
        PAIR_TRACE();
        if (et == 0) {
          tma_store_4d(dsp ? &p.map_y2 : &p.map_y, base + out_off + (uint32_t)os * kIoSlot, nt * BN + slab * 64, w0, h0, n0);
          tma_store_commit();
        }
        os ^= 1;
      }
      if (++acc == 2) {
        acc = 0;
        accph ^= 1u;
      }
    }
    if (et == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still read this CTA's shared memory
  stamp_end(p.stamp);
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

struct PairConvPlan {
  ConvGeom g;
  int bn = 0;
  bool ds = false;
  __nv_bfloat16* d_w2 = nullptr;
  const void* y2_ptr = nullptr;
  int ldy2 = 0;
  PairParams prm;
  __nv_bfloat16* d_w = nullptr;
  int64_t bytes = 0;
  size_t smem = 0;
  const void *x_ptr = nullptr, *y_ptr = nullptr, *res_ptr = nullptr;
  int n_maps = 0;
  int map_hp[4], map_wp[4];
};

bool pair_conv_supported(const ConvGeom& g) {
  static const bool off = debug_env("SPK_NO_PAIR") != nullptr;  // A/B switch
  if (off) return false;
  if (!tc_conv_supported(g)) return false;
  if (g.cout % 128 != 0 || g.cout > 4096) return false;
  return true;
}

int pair_conv_plan_create(spk_ctx* ctx, const ConvGeom& g_max, const float* w, const float* d_bias, PairConvPlan** out,
                          const float* w_ds, const float* d_bias_ds, int ldy_ds) {
  if (!pair_conv_supported(g_max)) return fail(ctx, SPK_ERR_UNSUPPORTED, "pair convolution: unsupported geometry");
  PairConvPlan* p = new PairConvPlan;
  p->g = g_max;
  const ConvGeom& g = p->g;
  memset(&p->prm, 0, sizeof p->prm);
  PairParams& prm = p->prm;
  p->ds = w_ds != nullptr;
  p->ldy2 = ldy_ds;
  prm.ds_tap = p->ds ? 4 : -1;  // (r, s) = (1, 1) of a 3x3 / pad 1 filter
  prm.bias2 = d_bias_ds;
  p->bn = (g.cout % 256 == 0 && !p->ds) ? 256 : 128;  // (the fused shortcut takes a second accumulator: N = 128)
  prm.bias = d_bias;
  prm.ho = g.ho;
  prm.wo = g.wo;
  prm.cout = g.cout;
  prm.relu = g.relu;
  prm.taps = g.kh * g.kw;
  prm.kchunks = (g.cin + kBK - 1) / kBK;
  prm.cin_pad = prm.kchunks * kBK;
  prm.tiles_n = g.cout / p->bn;

  // ---- tile box: wb * hb * nb = 128, least padded MMA rows; ties -> larger spatial footprint, wider rows
  double best = -1;
  for (int wb = 1; wb <= 128; wb *= 2)
    for (int hb = 1; wb * hb <= 128; hb *= 2) {
      const int nb = 128 / (wb * hb);
      const long long cover = (long long)((g.wo + wb - 1) / wb) * wb * ((g.ho + hb - 1) / hb) * hb * ((g.n + nb - 1) / nb) * nb;
      const double eff = (double)g.wo * g.ho * g.n / (double)cover;
      const double score = eff + 1e-4 * (wb * hb) + 1e-6 * wb;
      if (score > best) {
        best = score;
        prm.wb = wb;
        prm.hb = hb;
        prm.nb = nb;
      }
    }
  prm.tiles_w = (g.wo + prm.wb - 1) / prm.wb;
  prm.tiles_h = (g.ho + prm.hb - 1) / prm.hb;
  // ---- N tile: 256 halves the A re-reads, but a layer with few M tiles (7x7 maps: 98 pairs x N tiles at batch 256) then
  // leaves most of the 74 clusters idle in its last round; pick the width with the shorter schedule at the planned batch
  if (g.cout % 256 == 0 && !p->ds && debug_env("SPK_PAIR_BN_AUTO")) {  // measured (512->512 @7x7, batch 256): 0.063 ms at N = 128 vs 0.058-0.062 at N = 256; off
    const long long m_tiles = (long long)prm.tiles_w * prm.tiles_h * ((g.n + prm.nb - 1) / prm.nb);
    const long long pairs = (m_tiles + 1) / 2, clusters = std::max(1, ctx->sm_count / 2);
    auto rounds = [&](int bn) { return (pairs * (g.cout / bn) + clusters - 1) / clusters; };
    const double t256 = (double)rounds(256) * 256.0, t128 = (double)rounds(128) * 128.0 * 1.06;  // N = 128 streams B twice as often
    p->bn = t128 < t256 ? 128 : 256;
  }
  prm.tiles_n = g.cout / p->bn;

  // ---- taps -> (tensor map, box shift)
  p->n_maps = 0;
  for (int r = 0; r < g.kh; ++r)
    for (int s = 0; s < g.kw; ++s) {
      const int t = r * g.kw + s;
      const int th = r - g.pad, tw = s - g.pad;
      int hp = 0, wp = 0, dh = th, dw = tw;
      if (g.stride == 2) {
        hp = ((th % 2) + 2) % 2;
        wp = ((tw % 2) + 2) % 2;
        dh = (th - hp) / 2;
        dw = (tw - wp) / 2;
      }
      int mi = -1;
      for (int i = 0; i < p->n_maps; ++i)
        if (p->map_hp[i] == hp && p->map_wp[i] == wp) mi = i;
      if (mi < 0) {
        mi = p->n_maps++;
        p->map_hp[mi] = hp;
        p->map_wp[mi] = wp;
      }
      prm.tap_map[t] = (signed char)mi;
      prm.tap_dh[t] = (signed char)dh;
      prm.tap_dw[t] = (signed char)dw;
    }

  // ---- shared memory: stages under the budget
  const size_t stage = (size_t)kABytes + (size_t)(p->bn / 2) * kBK * 2 * (p->ds ? 2 : 1);
  // residual ring: 4 slabs where the layer is HBM-bound by construction (a 1x1 expansion: a few k blocks of math per 64 KB of
  // residual + 64 KB of output per tile; with 2 slabs = 32 KB in flight per SM the residual stream could not cover the
  // HBM latency), 2 for the 3x3 layers (tensor-bound, the stages matter more)
  static const char* rs_env = debug_env("SPK_PAIR_RES_SLOTS");  // A/B switch (debug build)
  prm.res_slots = g.ldres ? (rs_env ? std::max(2, std::min(kMaxResSlots, atoi(rs_env))) : (g.kh * g.kw == 1 ? 4 : 2)) : 0;
  const size_t fixed = 1024 + (size_t)(2 + prm.res_slots) * kIoSlot + (size_t)g.cout * 4 + 16 + 8 * (2 * kMaxStages + 2 * kMaxResSlots + 6) + 16;
  prm.stages = (int)std::min<size_t>(kMaxStages, (kSmemMax - fixed) / stage);
  if (prm.stages < 2) {
    delete p;
    return fail(ctx, SPK_ERR_UNSUPPORTED, "pair convolution: shared memory budget");
  }
  p->smem = fixed + (size_t)prm.stages * stage;

  // ---- weights: bf16 [Cout][taps][cin_pad], zero padded, round to nearest
  const size_t kk = (size_t)prm.taps * prm.cin_pad;
  std::vector<__nv_bfloat16> wb16((size_t)g.cout * kk, __float2bfloat16(0.f));
  for (int o = 0; o < g.cout; ++o)
    for (int t = 0; t < prm.taps; ++t)
      for (int c = 0; c < g.cin; ++c)
        wb16[(size_t)o * kk + (size_t)t * prm.cin_pad + c] = __float2bfloat16(w[((size_t)o * prm.taps + t) * g.cin + c]);
  cudaError_t e = cudaMalloc(&p->d_w, wb16.size() * 2);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_w, wb16.data(), wb16.size() * 2, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    pair_conv_plan_destroy(p);
    return fail(ctx, SPK_ERR_CUDA, "pair convolution: weight upload: %s", cudaGetErrorString(e));
  }
  p->bytes = (int64_t)wb16.size() * 2;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kk, (cuuint64_t)g.cout};
    cuuint64_t strides[1] = {(cuuint64_t)kk * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)(p->bn / 2)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&prm.map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->d_w, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pair_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "pair convolution: cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    }
  }
  if (p->ds) {
    // shortcut weights: bf16 [Cout][cin_pad]
    const size_t k1 = (size_t)prm.cin_pad;
    std::vector<__nv_bfloat16> w2((size_t)g.cout * k1, __float2bfloat16(0.f));
    for (int o = 0; o < g.cout; ++o)
      for (int c = 0; c < g.cin; ++c) w2[(size_t)o * k1 + c] = __float2bfloat16(w_ds[(size_t)o * g.cin + c]);
    cudaError_t e2 = cudaMalloc(&p->d_w2, w2.size() * 2);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(p->d_w2, w2.data(), w2.size() * 2, cudaMemcpyHostToDevice);
    if (e2 != cudaSuccess) {
      pair_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "pair convolution: shortcut weight upload: %s", cudaGetErrorString(e2));
    }
    p->bytes += (int64_t)w2.size() * 2;
    cuuint64_t dims[2] = {(cuuint64_t)k1, (cuuint64_t)g.cout};
    cuuint64_t strides[1] = {(cuuint64_t)k1 * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)(p->bn / 2)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&prm.map_b2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->d_w2, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pair_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "pair convolution: cuTensorMapEncodeTiled(W shortcut) failed: %d", (int)r);
    }
  }
  cudaError_t ea = p->ds ? cudaFuncSetAttribute(conv_pair_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax)
                   : p->bn == 256
                       ? cudaFuncSetAttribute(conv_pair_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax)
                       : cudaFuncSetAttribute(conv_pair_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
  if (ea != cudaSuccess) {
    pair_conv_plan_destroy(p);
    return fail(ctx, SPK_ERR_CUDA, "pair convolution: cudaFuncSetAttribute: %s", cudaGetErrorString(ea));
  }
  *out = p;
  return SPK_OK;
}

void pair_conv_plan_destroy(PairConvPlan* p) {
  if (!p) return;
  if (p->d_w) cudaFree(p->d_w);
  if (p->d_w2) cudaFree(p->d_w2);
  delete p;
}

int64_t pair_conv_plan_bytes(const PairConvPlan* p) { return p ? p->bytes : 0; }

void pair_conv_plan_set_reverse(PairConvPlan* p, int reverse) {
  if (p) p->prm.reverse = reverse ? 1 : 0;
}

int pair_conv_launch(spk_ctx* ctx, PairConvPlan* p, int n, const void* x, const void* res, void* y, void* y_ds) {
  if (n <= 0) return SPK_OK;
  const ConvGeom& g = p->g;
  if (n > g.n) return fail(ctx, SPK_ERR_CAPACITY, "pair convolution: batch %d > planned %d", n, g.n);
  PairParams& prm = p->prm;
  if (x != p->x_ptr) {
    for (int i = 0; i < p->n_maps; ++i) {
      const int hp = p->map_hp[i], wp = p->map_wp[i], s = g.stride;
      const char* b = (const char*)x + ((size_t)hp * g.w + wp) * g.ldx * 2;
      cuuint64_t dims[4] = {(cuuint64_t)g.cin, (cuuint64_t)((g.w - wp + s - 1) / s), (cuuint64_t)((g.h - hp + s - 1) / s), (cuuint64_t)g.n};
      cuuint64_t strides[3] = {(cuuint64_t)s * g.ldx * 2, (cuuint64_t)s * g.w * g.ldx * 2, (cuuint64_t)g.h * g.w * g.ldx * 2};
      cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)prm.wb, (cuuint32_t)prm.hb, (cuuint32_t)prm.nb};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = encode_fn()(&prm.map_a[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)b, dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "pair convolution: tensor map (x) failed: %d", (int)r);
    }
    for (int i = p->n_maps; i < 4; ++i) prm.map_a[i] = prm.map_a[0];
    p->x_ptr = x;
  }
  if (y != p->y_ptr) {
    CUresult r = encode_nhwc_bf16(&prm.map_y, y, g.cout, g.wo, g.ho, g.n, g.ldy, 64, prm.wb, prm.hb, prm.nb);
    if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "pair convolution: tensor map (y) failed: %d", (int)r);
    p->y_ptr = y;
  }
  if (p->ds) {
    if (!y_ds) return fail(ctx, SPK_ERR_INVALID, "pair convolution: fused shortcut without an output");
    if (y_ds != p->y2_ptr) {
      CUresult r = encode_nhwc_bf16(&prm.map_y2, y_ds, g.cout, g.wo, g.ho, g.n, p->ldy2, 64, prm.wb, prm.hb, prm.nb);
      if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "pair convolution: tensor map (shortcut output) failed: %d", (int)r);
      p->y2_ptr = y_ds;
    }
  }
  if (res && res != p->res_ptr) {
    CUresult r = encode_nhwc_bf16(&prm.map_res, res, g.cout, g.wo, g.ho, g.n, g.ldres, 64, prm.wb, prm.hb, prm.nb);
    if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "pair convolution: tensor map (residual) failed: %d", (int)r);
    p->res_ptr = res;
  }
  prm.has_res = res ? 1 : 0;
  if (res && !prm.res_slots) return fail(ctx, SPK_ERR_STATE, "pair convolution: residual given to a plan made without one");
  prm.n = n;
  prm.tiles_img = (n + prm.nb - 1) / prm.nb;
  prm.m_tiles = prm.tiles_w * prm.tiles_h * prm.tiles_img;
  prm.units = ((prm.m_tiles + 1) / 2) * prm.tiles_n;
  const int clusters = std::min(prm.units, ctx->sm_count / 2);
  prm.stamp = ctx->cur_stamp;
  static const bool want_trace = debug_env("SPK_PAIR_TRACE") != nullptr;
  static long long* d_trace = nullptr;
  static int trace_left = 4;
  prm.trace = nullptr;
  if (want_trace && trace_left > 0 && prm.kchunks * prm.taps <= 2) {
    if (!d_trace) cudaMalloc(&d_trace, 256 * sizeof(long long));
    cudaMemsetAsync(d_trace, 0, 256 * sizeof(long long), ctx->stream);
    prm.trace = d_trace;
  }
  if (p->ds)
    SPK_CUDA_OK(ctx, launch_pdl(conv_pair_kernel<128, true>, dim3(2 * clusters), dim3(kThreads), p->smem, ctx->stream, prm));
  else if (p->bn == 256)
    SPK_CUDA_OK(ctx, launch_pdl(conv_pair_kernel<256, false>, dim3(2 * clusters), dim3(kThreads), p->smem, ctx->stream, prm));
  else
    SPK_CUDA_OK(ctx, launch_pdl(conv_pair_kernel<128, false>, dim3(2 * clusters), dim3(kThreads), p->smem, ctx->stream, prm));
  SPK_LAUNCH_CHECK(ctx);
  if (prm.trace) {
    --trace_left;
    long long h[256];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(h, d_trace, sizeof h, cudaMemcpyDeviceToHost);
    fprintf(stderr, "pair trace BN=%d cin=%d cout=%d res=%d stages=%d units=%d (per unit: start, t_full; per slab: ld done, store-read wait, r_full, bar1, math done, bar2):\n ",
            p->bn, g.cin, g.cout, prm.has_res, prm.stages, prm.units);
    for (int i = 0; i < 250 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
    fprintf(stderr, "\n");
  }
  return SPK_OK;
}

}  // namespace spk

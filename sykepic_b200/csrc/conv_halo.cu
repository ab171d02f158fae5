// K2 (3x3 / stride 1 / pad 1): halo-resident implicit-GEMM convolution on tcgen05.
//
// Replaces the 3x3 conv / BatchNorm / ReLU / residual-add sequence of torchvision's ResNet blocks as run
// by TorchVisionNet.forward (sykepic/train/network.py:66-68); in the reference these are cuDNN / oneDNN
// library calls, there is no kernel source to follow.
//
// Why a second convolution kernel: the tap-per-TMA kernel (conv_tc.cu) re-loads the A operand for each
// of the nine filter taps and the weights for every tile; on the large maps (56x56, 28x28) it is bound
// by L2 -> SM bandwidth (ncu: 1.39 GB moved for a 103 MB input, tensor pipe 21 % active).  Here
//   * an M tile is `hb` whole image rows; its input HALO, (hb + 2) x (W + 2) pixels x 64 channels, is
//     loaded ONCE per 64-channel chunk by one 4-D TMA box (out-of-image pixels zero-filled = padding)
//     into a SWIZZLE_128B tile whose row index is the halo's linear pixel index (pitch W + 2);
//   * filter tap (r, s) is the SAME tile read through a shared-memory descriptor that starts
//     (r * pitch + s) rows = 128-byte steps further in: output pixel j (linear, pitch W + 2) needs input
//     row j + r * pitch + s.  MMA rows that fall on the two halo columns per image row are computed and
//     dropped (W * hb of 128 rows are useful: 87.5 % at W = 56 and W = 28);
//   * weights are either resident in shared memory for the whole kernel (Cin = Cout = 64: 72 KB) or
//     streamed per tap through a ring and shared by MT = 2 M tiles per CTA;
//   * the epilogue goes TMEM -> registers -> (+bias, +residual, ReLU, bf16) -> swizzled shared staging
//     -> one TMA store per 64-channel slab (full 128-byte lines); the residual tile arrives by TMA too,
//     prefetched by its own producer warp while the MMAs run.
// Persistent CTAs (one per SM), warp-specialised: warp 0 = A/B TMA producer, warp 1 = TMEM allocation +
// MMA issuer, warp 2 = residual producer, warps 3-6 = epilogue.  Accumulators are double-buffered in TMEM.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "spk_internal.h"
#include "tc_common.cuh"

namespace spk {
namespace {
using namespace tc;

constexpr int kThreads = 224;
constexpr int kMaxRing = 8;

struct alignas(64) HaloParams {
  CUtensorMap map_x, map_w, map_y, map_res;
  const float* bias;
  int n, h, w, cin, cout, relu, has_res;
  int pitch, hb, tiles_h, kchunks, tiles_n;
  int m_tiles;      // n * tiles_h
  int units;        // ceil(m_tiles / MT) * tiles_n
  int a_stage;      // bytes per A stage (multiple of 1024)
  int a_box_bytes;  // (hb + 2) * (w + 2) * 128
  int io_bytes;     // w * hb * 128: box bytes of an output / residual slab
  int io_slot;      // io_bytes rounded up to 1024
  int na, nb;       // ring depths
  long long* trace; // debug (SPK_HALO_TRACE=1): clock64 stamps of CTA 5's MMA issuer, else nullptr
};

template <int BN, int MT, bool BRES>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_halo_kernel(const __grid_constant__ HaloParams p) {
  constexpr int kAccCols = MT * BN;
  constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;
  constexpr int kSlabs = BN / 64;
  constexpr uint32_t kIdesc = idesc_bf16(128, BN);
  constexpr uint32_t kBTile = BN * 128;

  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gen = smem_raw + (base - raw);
  const uint32_t size_b = BRES ? 9u * p.kchunks * kBTile : (uint32_t)p.nb * kBTile;
  const uint32_t b_s = base;
  const uint32_t a_s = b_s + size_b;
  const uint32_t res_off = size_b + (uint32_t)p.na * p.a_stage;
  const uint32_t out_off = res_off + 2u * p.io_slot;
  const uint32_t bias_off = out_off + 2u * p.io_slot;
  const uint32_t bar0 = base + ((bias_off + (uint32_t)p.cout * 4u + 15u) & ~15u);
  float* bias_sm = reinterpret_cast<float*>(gen + bias_off);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kMaxRing + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * kMaxRing + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (3 * kMaxRing + s); };
  auto r_full = [&](int s) { return bar0 + 8u * (4 * kMaxRing + s); };
  auto r_empty = [&](int s) { return bar0 + 8u * (4 * kMaxRing + 2 + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (4 * kMaxRing + 4 + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (4 * kMaxRing + 6 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + (bar0 - base) + 8 * (4 * kMaxRing + 8));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(r_full(s), 1);
      mbar_init(r_empty(s), 4);  // one arrival per epilogue warp
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), 4);
    }
    mbar_init_fence();
    tma_prefetch_desc(&p.map_x);
    tma_prefetch_desc(&p.map_w);
    tma_prefetch_desc(&p.map_y);
    if (p.has_res) tma_prefetch_desc(&p.map_res);
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), kTmemCols);
  for (int i = threadIdx.x; i < p.cout; i += kThreads) bias_sm[i] = __ldg(p.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int m, int& img, int& h0) {
    img = m / p.tiles_h;
    h0 = (m - img * p.tiles_h) * p.hb;
  };

  if (warp == 0) {
    // ===== A / B producer (one thread) =====
    if (lane == 0) {
      if (BRES) {
        mbar_expect_tx(b_full(0), 9u * p.kchunks * kBTile);
        for (int ch = 0; ch < p.kchunks; ++ch)
          for (int tap = 0; tap < 9; ++tap)
            tma_load_2d(b_s + (uint32_t)(ch * 9 + tap) * kBTile, &p.map_w, b_full(0), tap * p.cin + ch * 64, 0);
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int nt = u % p.tiles_n;
        const int m0 = (u / p.tiles_n) * MT;
        for (int ch = 0; ch < p.kchunks; ++ch) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int m = m0 + mt;
            if (m >= p.m_tiles) continue;
            int img, h0;
            tile_coords(m, img, h0);
            mbar_wait(a_empty(as), aph ^ 1u);
            mbar_expect_tx(a_full(as), (uint32_t)p.a_box_bytes);
            tma_load_4d(a_s + (uint32_t)as * p.a_stage, &p.map_x, a_full(as), ch * 64, -1, h0 - 1, img);
            if (++as == p.na) {
              as = 0;
              aph ^= 1u;
            }
          }
          if (!BRES) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(b_empty(bs), bph ^ 1u);
              mbar_expect_tx(b_full(bs), kBTile);
              tma_load_2d(b_s + (uint32_t)bs * kBTile, &p.map_w, b_full(bs), tap * p.cin + ch * 64, nt * BN);
              if (++bs == p.nb) {
                bs = 0;
                bph ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop with warp-uniform values, one elected lane issues =====
    {
      if (BRES) {
        mbar_wait(b_full(0), 0);
        tc_fence_after();
      }
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, accph = 0;
      int tr = 0;
      const bool tracing = p.trace != nullptr && blockIdx.x == 5 && lane == 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int m0 = (u / p.tiles_n) * MT;
        const int nvalid = min(MT, p.m_tiles - m0);
        if (tracing && tr < 250) p.trace[tr++] = clock64();
        mbar_wait(t_empty(acc), accph ^ 1u);  // the epilogue has drained this accumulator buffer
        tc_fence_after();
        if (tracing && tr < 250) p.trace[tr++] = clock64();
        for (int ch = 0; ch < p.kchunks; ++ch) {
          int slot[MT];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            slot[mt] = as;
            if (mt < nvalid) {
              mbar_wait(a_full(as), aph);
              if (++as == p.na) {
                as = 0;
                aph ^= 1u;
              }
            }
          }
          tc_fence_after();
          if (tracing && tr < 250) p.trace[tr++] = clock64();
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap - 3 * r;
            uint32_t b_addr;
            if (BRES) {
              b_addr = b_s + (uint32_t)(ch * 9 + tap) * kBTile;
            } else {
              mbar_wait(b_full(bs), bph);
              tc_fence_after();
              if (tracing && tr < 250) p.trace[tr++] = clock64();
              b_addr = b_s + (uint32_t)bs * kBTile;
            }
            const uint64_t b_desc = smem_desc_sw128(b_addr);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              if (mt < nvalid) {
                // tap (r, s) = the halo tile read (r * pitch + s) rows further in
                const uint64_t a_desc = smem_desc_sw128(a_s + (uint32_t)slot[mt] * p.a_stage + (uint32_t)(r * p.pitch + s) * 128u);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols + mt * BN);
#pragma unroll
                for (int k = 0; k < 4; ++k)  // 16 bf16 = 32 bytes along K: +2 in the (addr >> 4) field
                  tc_mma_w(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), kIdesc, (ch | tap | k) != 0 ? 1u : 0u);
              }
            }
            if (!BRES) {
              tc_commit_w(b_empty(bs));
              if (++bs == p.nb) {
                bs = 0;
                bph ^= 1u;
              }
            }
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            if (mt < nvalid) tc_commit_w(a_empty(slot[mt]));  // frees the halo tile once these MMAs have read it
        }
        tc_commit_w(t_full(acc));
        if (++acc == 2) {
          acc = 0;
          accph ^= 1u;
        }
      }
    }
  } else if (warp == 2) {
    // ===== residual producer (one thread) =====
    if (lane == 0 && p.has_res) {
      int rs = 0;
      uint32_t rph = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int nt = u % p.tiles_n;
        const int m0 = (u / p.tiles_n) * MT;
        const int nvalid = min(MT, p.m_tiles - m0);
        for (int mt = 0; mt < nvalid; ++mt) {
          int img, h0;
          tile_coords(m0 + mt, img, h0);
          for (int slab = 0; slab < kSlabs; ++slab) {
            mbar_wait(r_empty(rs), rph ^ 1u);
            mbar_expect_tx(r_full(rs), (uint32_t)p.io_bytes);
            tma_load_4d(base + res_off + (uint32_t)rs * p.io_slot, &p.map_res, r_full(rs), nt * BN + slab * 64, 0, h0, img);
            if (++rs == 2) {
              rs = 0;
              rph ^= 1u;
            }
          }
        }
      }
    }
  } else {
    // ===== epilogue: warps 3-6; warp w may touch TMEM lanes [32 * (w % 4), +32) =====
    const int q = warp & 3;
    const int et = threadIdx.x - 96;  // 0..127 among the epilogue threads
    const int j = q * 32 + lane;      // TMEM lane == linear halo-pitch pixel of the tile
    const int jr = j / p.pitch, jc = j - jr * p.pitch;
    const bool inside = (jc < p.w) && (jr < p.hb);
    const int orow = jr * p.w + jc;  // row of the compact [hb][W] staging tile
    const uint32_t sw = (uint32_t)(orow & 7);
    int acc = 0, rs = 0, os = 0;
    uint32_t accph = 0, rph = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int nt = u % p.tiles_n;
      const int m0 = (u / p.tiles_n) * MT;
      const int nvalid = min(MT, p.m_tiles - m0);
      mbar_wait(t_full(acc), accph);
      tc_fence_after();
      for (int mt = 0; mt < nvalid; ++mt) {
        int img, h0;
        tile_coords(m0 + mt, img, h0);
#pragma unroll 1
        for (int slab = 0; slab < kSlabs; ++slab) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccCols + mt * BN + slab * 64);
          uint32_t v0[32], v1[32];
          tmem_ld32(taddr, v0);
          tmem_ld32(taddr + 32, v1);
          tmem_ld_wait();
          if (mt == nvalid - 1 && slab == kSlabs - 1) {  // accumulator buffer drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty(acc));
          }
          // staging slot `os` was last read by the TMA store issued two slabs ago
          if (et == 0) tma_store_wait_read<1>();
          if (p.has_res) mbar_wait(r_full(rs), rph);
          named_bar_sync(1, 128);
          if (inside) {
            const float* bsm = bias_sm + nt * BN + slab * 64;
            unsigned char* orow_p = gen + out_off + (uint32_t)os * p.io_slot + (uint32_t)orow * 128u;
            const unsigned char* rrow_p = gen + res_off + (uint32_t)rs * p.io_slot + (uint32_t)orow * 128u;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int col = c8 * 8 + e;
                f[e] = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]) + bsm[col];
              }
              const uint32_t chunk = ((uint32_t)c8 ^ sw) << 4;
              if (p.has_res) {
                const uint4 r4 = *reinterpret_cast<const uint4*>(rrow_p + chunk);
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 rf = __bfloat1622float2(r2[t]);
                  f[2 * t] += rf.x;
                  f[2 * t + 1] += rf.y;
                }
              }
              if (p.relu) {
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
          if (p.has_res) {
            __syncwarp();
            if (lane == 0) mbar_arrive(r_empty(rs));
            if (++rs == 2) {
              rs = 0;
              rph ^= 1u;
            }
          }
          fence_async_smem();
          named_bar_sync(2, 128);
          if (et == 0) {
            tma_store_4d(&p.map_y, base + out_off + (uint32_t)os * p.io_slot, nt * BN + slab * 64, 0, h0, img);
            tma_store_commit();
          }
          os ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        accph ^= 1u;
      }
    }
    if (et == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

constexpr size_t kSmemMax = 232448;  // 227 KB opt-in maximum per CTA on sm_100

struct Choice {
  int bn, mt;
  bool bres;
};

Choice pick(const ConvGeom& g) {
  if (g.cout == 64 && g.cin == 64) return {64, 1, true};
  return {g.cout % 128 == 0 ? 128 : 64, 2, false};
}

}  // namespace

struct HaloConvPlan {
  ConvGeom g;
  HaloParams prm;
  Choice ch;
  __nv_bfloat16* d_w = nullptr;
  int64_t bytes = 0;
  size_t smem = 0;
  const void *x_ptr = nullptr, *y_ptr = nullptr, *res_ptr = nullptr;
};

static int halo_rows(const ConvGeom& g, int* hb_out) {
  const int pitch = g.w + 2;
  int hb = 128 / pitch;
  if (hb > g.ho) hb = g.ho;
  *hb_out = hb;
  return pitch;
}

bool halo_conv_supported(const ConvGeom& g) {
  if (g.kh != 3 || g.kw != 3 || g.stride != 1 || g.pad != 1) return false;
  if (g.cin % 64 != 0 || g.cout % 64 != 0 || g.cout > 2048) return false;
  if (g.ldx % 8 != 0 || g.ldy % 8 != 0 || g.ldres % 8 != 0) return false;
  if (g.w + 2 > 128 || g.ho != g.h || g.wo != g.w) return false;
  int hb;
  halo_rows(g, &hb);
  if (hb < 1) return false;
  const int tiles_h = (g.h + hb - 1) / hb;
  const double eff = (double)(g.w * hb) / 128.0 * (double)g.h / (double)(tiles_h * hb);
  if (eff < 0.8) return false;
  return encode_fn() != nullptr;
}

template <int BN, int MT, bool BRES>
static int halo_launch_t(spk_ctx* ctx, HaloConvPlan* p) {
  static bool attr_done[64] = {};
  if (!attr_done[ctx->device & 63]) {
    SPK_CUDA_OK(ctx, cudaFuncSetAttribute(conv3x3_halo_kernel<BN, MT, BRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    attr_done[ctx->device & 63] = true;
  }
  const int grid = std::min(p->prm.units, ctx->sm_count);
  static const bool want_trace = getenv("SPK_HALO_TRACE") != nullptr;
  static long long* d_trace = nullptr;
  static int trace_left = 6;
  p->prm.trace = nullptr;
  if (want_trace && trace_left > 0) {
    if (!d_trace) cudaMalloc(&d_trace, 256 * sizeof(long long));
    cudaMemsetAsync(d_trace, 0, 256 * sizeof(long long), ctx->stream);
    p->prm.trace = d_trace;
  }
  conv3x3_halo_kernel<BN, MT, BRES><<<grid, kThreads, p->smem, ctx->stream>>>(p->prm);
  SPK_LAUNCH_CHECK(ctx);
  if (p->prm.trace) {
    --trace_left;
    long long h[256];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(h, d_trace, sizeof h, cudaMemcpyDeviceToHost);
    fprintf(stderr, "halo trace BN=%d MT=%d BRES=%d cin=%d w=%d units=%d na=%d nb=%d (cycles since the first stamp; per unit: start, t_empty, then per chunk: a_full%s):\n ",
            BN, MT, (int)BRES, p->prm.cin, p->prm.w, p->prm.units, p->prm.na, p->prm.nb, BRES ? "" : ", b_full x9");
    for (int i = 0; i < 250 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
    fprintf(stderr, "\n");
  }
  return SPK_OK;
}

int halo_conv_plan_create(spk_ctx* ctx, const ConvGeom& g_max, const float* w, const float* d_bias, HaloConvPlan** out) {
  if (!halo_conv_supported(g_max)) return fail(ctx, SPK_ERR_UNSUPPORTED, "halo convolution: unsupported geometry");
  HaloConvPlan* p = new HaloConvPlan;
  p->g = g_max;
  const ConvGeom& g = p->g;
  memset(&p->prm, 0, sizeof p->prm);
  HaloParams& prm = p->prm;
  p->ch = pick(g);
  prm.bias = d_bias;
  prm.h = g.h;
  prm.w = g.w;
  prm.cin = g.cin;
  prm.cout = g.cout;
  prm.relu = g.relu;
  prm.pitch = halo_rows(g, &prm.hb);
  prm.tiles_h = (g.h + prm.hb - 1) / prm.hb;
  prm.kchunks = g.cin / 64;
  prm.tiles_n = g.cout / p->ch.bn;
  prm.a_box_bytes = (prm.hb + 2) * prm.pitch * 128;
  prm.a_stage = ((2 * prm.pitch + 2 + 128) * 128 + 1023) & ~1023;
  if (prm.a_stage < prm.a_box_bytes) prm.a_stage = (prm.a_box_bytes + 1023) & ~1023;
  prm.io_bytes = g.w * prm.hb * 128;
  prm.io_slot = (prm.io_bytes + 1023) & ~1023;
  // ring depths under the shared-memory budget
  const size_t fixed = 1024 /*alignment*/ + 4 * (size_t)prm.io_slot + (size_t)g.cout * 4 + 16 + 8 * (4 * kMaxRing + 8) + 16;
  const size_t b_tile = (size_t)p->ch.bn * 128;
  size_t b_bytes;
  if (p->ch.bres) {
    b_bytes = 9 * (size_t)prm.kchunks * b_tile;
    prm.nb = 1;
  } else {
    prm.nb = 4;
    b_bytes = prm.nb * b_tile;
  }
  prm.na = (int)std::min<size_t>(kMaxRing, (kSmemMax - fixed - b_bytes) / (size_t)prm.a_stage);
  if (prm.na > 2 * p->ch.mt * 2) prm.na = 2 * p->ch.mt * 2;
  if (prm.na < p->ch.mt) {
    delete p;
    return fail(ctx, SPK_ERR_UNSUPPORTED, "halo convolution: shared memory budget");
  }
  p->smem = fixed + b_bytes + (size_t)prm.na * prm.a_stage;

  // ---- weights: bf16 [Cout][tap][Cin], round to nearest
  const size_t kk = (size_t)9 * g.cin;
  std::vector<__nv_bfloat16> wb16((size_t)g.cout * kk);
  for (size_t i = 0; i < wb16.size(); ++i) wb16[i] = __float2bfloat16(w[i]);
  cudaError_t e = cudaMalloc(&p->d_w, wb16.size() * 2);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_w, wb16.data(), wb16.size() * 2, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    halo_conv_plan_destroy(p);
    return fail(ctx, SPK_ERR_CUDA, "halo convolution: weight upload: %s", cudaGetErrorString(e));
  }
  p->bytes = (int64_t)wb16.size() * 2;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kk, (cuuint64_t)g.cout};
    cuuint64_t strides[1] = {(cuuint64_t)kk * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)p->ch.bn};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&prm.map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->d_w, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      halo_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "halo convolution: cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    }
  }
  *out = p;
  return SPK_OK;
}

void halo_conv_plan_destroy(HaloConvPlan* p) {
  if (!p) return;
  if (p->d_w) cudaFree(p->d_w);
  delete p;
}

int64_t halo_conv_plan_bytes(const HaloConvPlan* p) { return p ? p->bytes : 0; }

int halo_conv_launch(spk_ctx* ctx, HaloConvPlan* p, int n, const void* x, const void* res, void* y) {
  if (n <= 0) return SPK_OK;
  const ConvGeom& g = p->g;
  if (n > g.n) return fail(ctx, SPK_ERR_CAPACITY, "halo convolution: batch %d > planned %d", n, g.n);
  HaloParams& prm = p->prm;
  if (x != p->x_ptr) {
    CUresult r = encode_nhwc_bf16(&prm.map_x, x, g.cin, g.w, g.h, g.n, g.ldx, 64, g.w + 2, prm.hb + 2, 1);
    if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "halo convolution: tensor map (x) failed: %d", (int)r);
    p->x_ptr = x;
  }
  if (y != p->y_ptr) {
    CUresult r = encode_nhwc_bf16(&prm.map_y, y, g.cout, g.w, g.h, g.n, g.ldy, 64, g.w, prm.hb, 1);
    if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "halo convolution: tensor map (y) failed: %d", (int)r);
    p->y_ptr = y;
  }
  if (res && res != p->res_ptr) {
    CUresult r = encode_nhwc_bf16(&prm.map_res, res, g.cout, g.w, g.h, g.n, g.ldres, 64, g.w, prm.hb, 1);
    if (r != CUDA_SUCCESS) return fail(ctx, SPK_ERR_CUDA, "halo convolution: tensor map (residual) failed: %d", (int)r);
    p->res_ptr = res;
  }
  prm.has_res = res ? 1 : 0;
  prm.n = n;
  prm.m_tiles = n * prm.tiles_h;
  prm.units = ((prm.m_tiles + p->ch.mt - 1) / p->ch.mt) * prm.tiles_n;
  if (p->ch.bres) return halo_launch_t<64, 1, true>(ctx, p);
  if (p->ch.bn == 128) return halo_launch_t<128, 2, false>(ctx, p);
  return halo_launch_t<64, 2, false>(ctx, p);
}

}  // namespace spk

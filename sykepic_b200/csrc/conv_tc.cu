// K2: implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM, operands staged by TMA), with the folded-BatchNorm bias, the residual add and the ReLU in
// the epilogue.  bf16 NHWC activations, fp32 accumulation.
//
// Replaces the conv / BatchNorm / ReLU / add sequence of torchvision's ResNet blocks as run by
// TorchVisionNet.forward (sykepic/train/network.py:66-68) -- in the reference these are cuDNN /
// oneDNN library calls, there is no kernel source to follow.
//
// GEMM view:  D[M = pixels, N = Cout] = sum over taps (r,s) and 64-channel chunks of
//             A_tap[M, 64] * W_tap[N, 64]^T.
//   * An M tile is a BOX of 128 output pixels: wb x hb pixels of nb images (wb*hb*nb = 128, chosen
//     per layer so that no MMA row is wasted).  For filter tap (r,s) the A operand of the tile is
//     the same box of the INPUT tensor shifted by (r-pad, s-pad): one 4-D TMA tiled load
//     {64 channels, wb, hb, nb}; out-of-image coordinates are zero-filled by TMA, which is exactly
//     the convolution's zero padding.  No im2col buffer exists anywhere.
//   * stride 2: the input is addressed through one tensor map per (row, column) parity, whose W/H
//     strides are doubled; tap (r,s) selects the map and a shifted box in it.
//   * The box lands in shared memory as 128 rows of 128 bytes with the 128-byte swizzle, i.e. the
//     canonical K-major SWIZZLE_128B operand layout of tcgen05.mma; weights [Cout][taps*Cin] are
//     loaded the same way (2-D map).
//   * Persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer
//     (one thread) + TMEM allocation, warps 2-5 = epilogue (tcgen05.ld -> bias/residual/ReLU ->
//     bf16 stores).  The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps
//     the MMAs of tile i+1.
#include <cuda.h>

#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "spk_internal.h"
#include "tc_common.cuh"  // only tc::pdl_* / tc::launch_pdl are used here; this file keeps its own PTX wrappers

namespace spk {
namespace {

constexpr int kBM = 128;          // pixels per tile (UMMA M)
constexpr int kBK = 64;           // channels per k block (128 bytes of bf16 = one swizzle row)
constexpr int kEpiWarps = 8;      // two per TMEM lane quarter, each half of the tile's columns (one warp per quarter walked a
                                  // 128-column tile in ~3600 cycles of dependent tcgen05.ld -> math -> store: the DenseNet 1x1 layers were epilogue-bound)
constexpr int kThreads = 64 + 32 * kEpiWarps;  // TMA producer, MMA issuer, epilogue warps
constexpr int kMaxTaps = 9;
constexpr int kABytes = kBM * kBK * 2;

struct alignas(64) TcParams {
  CUtensorMap map_a[4];
  CUtensorMap map_b;
  CUtensorMap map_b2;  // fused 1x1 stride-2 downsample: weights [Cout][cin_pad] applied to the centre tap's A tiles
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* y;
  const float* bias2;
  __nv_bfloat16* y2;
  int ldy2, ds_tap;
  int n, ho, wo, cout, ldy, ldres, relu;
  int wb, hb, nb;
  int tiles_w, tiles_h, tiles_img, tiles_n;
  int taps, kchunks, cin_pad;
  int total_tiles;
  int st32;         // bf16 epilogue: the output rows allow 32-byte stores (pointer and channel stride multiples of 32 bytes)
  int split_chunk;  // kModeSplit: k blocks per accumulation chunk (kSplitChunk)
  const float* wscale;         // kModeSplit: 2^-k(o) per output channel, undoes the power-of-two scaling of the fp16 weights
  unsigned long long* faults;  // kModeSplit: counts tiles with an output beyond the fp16 range of the SplitF format
  int reverse;  // walk the tiles last-to-first (see conv_hp.cu)
  // kModePre: y = conv(relu(x * pre_scale[c] + pre_shift[c])): DenseNet's pre-activation BatchNorm + ReLU applied to the A
  // tiles in shared memory (every consumer of a concatenation has its own BatchNorm, so it cannot be folded into a producer)
  const float* pre_scale;
  const float* pre_shift;
  int pre_c, pre_relu;
  unsigned long long* stamp;  // profiling (stamp mode): global-timer slot of this launch, else nullptr
  signed char tap_map[kMaxTaps + 3], tap_dh[kMaxTaps + 3], tap_dw[kMaxTaps + 3];
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
    // a pipeline bug must not hang the GPU: give up after ~2 s and raise a launch failure instead
    if (spin == 64) t0 = clock64();
    if (spin > 64 && (spin & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// whole-warp producer variants (see tc_mma_w): all lanes compute the uniform operands, one elected lane issues
__device__ __forceinline__ void mbar_expect_tx_w(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_w(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_w(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 in, fp32 accumulate, M = 128, N from the instruction descriptor
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Whole-warp variants: every lane executes the (warp-uniform) descriptor arithmetic, so ptxas keeps the operands in
// uniform registers, and one elected lane issues.  Issuing from inside an `if (lane == 0)` region instead costs
// ~70 cycles per MMA (SASS: an ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY loop per instruction) -- more than an
// N <= 128 MMA takes to execute.
__device__ __forceinline__ void tc_mma_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_w(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
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

// K-major, SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor): start address
// >> 4 in bits [0,14), leading byte offset (unused for swizzled K-major, 1) in [16,30), stride byte
// offset = 1024 B between 8-row groups in [32,46), descriptor version 1 in [46,48), layout type
// SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// MODE 0: plain.  MODE 1 (kModeDs): a second weight tile on the centre tap feeds a second accumulator (fused 1x1 / stride-2
// shortcut).  MODE 2 (kModeSplit): FP32-accurate convolution on the 16-bit tensor cores.  Activations are stored as SplitF
// words -- fp16 hi in the low half, fp16 lo = fp16(x - hi) in the high half -- which the MMA sees as an fp16 tensor with
// twice the channels, K index 2c = hi_c, 2c + 1 = lo_c.  Every k block is multiplied by TWO weight tiles into the same
// accumulator: [w_hi, w_hi] (gives x_hi*w_hi + x_lo*w_hi) and [w_lo, w_lo] (x_hi*w_lo + x_lo*w_lo): the full 16-bit x
// 16-bit product with fp32 accumulation (torch emulation of ResNet-18: |dp| <= 5e-6, experiments/emul_bf16x3_split.py).
constexpr int kModePlain = 0, kModeDs = 1, kModeSplit = 2, kModePre = 3;
// kModeSplit: k blocks accumulated in TMEM before the epilogue takes the partial sum over.  The tensor core TRUNCATES every
// fp32 accumulation (one per K = 16 MMA): a bias of ~half an ulp of the running sum per addition, which grows linearly with
// K (rms error 4e-6 at 72 additions, 1e-5 at 576).  With 4 k blocks = 16 additions per accumulator the bias stays below
// ~8 ulp of a PARTIAL sum, and the partial sums are added in registers with round-to-nearest fp32 adds.  Measured on
// the three benchmark checkpoints (max |dp| against the reference, ResNet-18 / -50 / DenseNet-121, fp16 hi + lo operands):
// no chunking 7.1e-5 / 7.0e-6 / 4.4e-5, 8 blocks 1.4e-5 / 4.5e-6 / 2.1e-5, 2 blocks 6.1e-6 / 4.8e-6 / 1.4e-5
// at 67 / 64 / 54 k ROI/s (ResNet-18); with bf16 hi + lo operands (16 significant bits) the floor was 5e-5 .. 1e-4
// whatever the chunk.
constexpr int kSplitChunk = 4;
constexpr int kPreWarps = 8;                      // kModePre: transform warps 6..13
constexpr int kThreadsPre = kThreads + 32 * kPreWarps;
constexpr int kPreTable = 2 * 1024 * 4;           // scale | shift for up to 1024 input channels

template <int BN, int MODE = 0>
struct Cfg {
  static constexpr bool DS = MODE == kModeDs;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr bool kTwoB = MODE == kModeDs || MODE == kModeSplit;  // two weight tiles per stage, two accumulators
  static constexpr int kStageBytes = kABytes + kBBytes * (kTwoB ? 2 : 1);
  static constexpr int kStages = kTwoB ? (BN >= 128 ? 4 : 6) : (BN >= 256 ? 4 : (BN >= 128 ? 6 : 8));
  static constexpr int kAccCols = kTwoB ? 2 * BN : BN;         // main accumulator (+ the downsample's / the lo pass's beside it)
  static constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;  // double-buffered (power of two)
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + (MODE == kModePre ? kPreTable : 0);
  // instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (1 << 4), A = B = bf16 (1 << 7, 1 << 10),
  // both K-major, N >> 3 in [17,23), M >> 4 in [24,29)
  // (kModeSplit: A = B = fp16, format 0)
  static constexpr uint32_t kIdesc = (1u << 4) | (MODE == kModeSplit ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
  static constexpr uint32_t kIdescSplit2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);  // fp16 x fp16, N = 2 BN
};

template <int BN, int MODE>
__global__ void __launch_bounds__(MODE == kModePre ? kThreadsPre : kThreads, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  using C = Cfg<BN, MODE>;
  constexpr bool DS = MODE == kModeDs;
  constexpr bool SPLIT = MODE == kModeSplit;
  constexpr bool PRE = MODE == kModePre;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gen_base = smem_raw + (base - raw);
  const uint32_t bar_base = base + C::kStages * C::kStageBytes;
  // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen_base + C::kStages * C::kStageBytes + 8 * (2 * C::kStages + 4));
  auto ready_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 5 + s); };  // kModePre: the A tile has been transformed
  float* pre_tab = reinterpret_cast<float*>(gen_base + C::kStages * C::kStageBytes + 256);  // kModePre: scale[1024] | shift[1024]
  (void)ready_bar;
  (void)pre_tab;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  tc::pdl_trigger();
  tc::stamp_begin(p.stamp);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiWarps);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&p.map_b);
    tma_prefetch_desc(&p.map_a[0]);
    if (DS || SPLIT) tma_prefetch_desc(&p.map_b2);
    if (PRE)
      for (int s = 0; s < C::kStages; ++s) mbar_init(ready_bar(s), kPreWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (PRE) {  // the BatchNorm table, zero beyond the real channels (TMA zero-fills those and they must stay zero)
    for (int i = threadIdx.x; i < 2 * 1024; i += blockDim.x) {
      const int c = i & 1023;
      pre_tab[i] = c < p.pre_c ? __ldg((i < 1024 ? p.pre_scale : p.pre_shift) + c) : 0.f;
    }
  }
  if (warp == 1) {  // TMEM allocation is warp-collective; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  tc::pdl_wait();  // everything below reads / writes tensors the previous kernel may still be using

  const int kblocks = p.taps * p.kchunks;

  if (warp == 0) {
    // ===== TMA producer: whole warp, one elected lane issues =====
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int tl = p.reverse ? p.total_tiles - 1 - tile : tile;
        const int nt = tl % p.tiles_n;
        int m = tl / p.tiles_n;
        const int twi = m % p.tiles_w;
        m /= p.tiles_w;
        const int thi = m % p.tiles_h;
        const int ng = m / p.tiles_h;
        const int w0 = twi * p.wb, h0 = thi * p.hb, n0 = ng * p.nb;
        for (int kb = 0; kb < kblocks; ++kb) {
          const int tap = kb / p.kchunks;
          const int c0 = (kb - tap * p.kchunks) * kBK;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * C::kStageBytes;
          const bool ds = DS && tap == p.ds_tap;
          mbar_expect_tx_w(full_bar(stage), (uint32_t)(kABytes + C::kBBytes * ((ds || SPLIT) ? 2 : 1)));
          tma_load_4d_w(sa, &p.map_a[p.tap_map[tap]], full_bar(stage), c0, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], n0);
          tma_load_2d_w(sa + kABytes, &p.map_b, full_bar(stage), tap * p.cin_pad + c0, nt * BN);
          if (ds) tma_load_2d_w(sa + kABytes + C::kBBytes, &p.map_b2, full_bar(stage), c0, nt * BN);
          if (SPLIT) tma_load_2d_w(sa + kABytes + C::kBBytes, &p.map_b2, full_bar(stage), tap * p.cin_pad + c0, nt * BN);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop with warp-uniform values, one elected lane issues =====
    {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        uint32_t d_tmem = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          // an accumulation chunk: the whole tile, or kSplitChunk k blocks in split mode
          const bool chunk_start = SPLIT ? (kb % p.split_chunk == 0) : (kb == 0);
          const bool chunk_end = kb + 1 == kblocks || (SPLIT && (kb % p.split_chunk == p.split_chunk - 1));
          if (chunk_start) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
            tc_fence_after();
            d_tmem = tmem_base + (uint32_t)(acc * C::kAccCols);
          }
          mbar_wait(PRE ? ready_bar(stage) : full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * C::kStageBytes;
          const uint64_t a_desc = smem_desc_sw128(sa);
          const uint64_t b_desc = smem_desc_sw128(sa + kABytes);
          if constexpr (SPLIT) {
            // ONE MMA of N = 2 BN per k step: the hi-weight tile and the lo-weight tile lie back to back in the stage (two TMA
            // boxes of BN rows x 128 bytes, each a multiple of 1024 bytes: together a valid 2 BN-row SWIZZLE_128B tile), the
            // accumulator columns [0, BN) get x * w_hi and [BN, 2 BN) get x * w_lo as before (own accumulator for the small
            // terms: they would double the number of truncating additions into the large sum).  Against two N = BN MMAs the
            // activation tile is read once instead of twice: the split mode was shared-memory bound (BN = 64: 12 KB of operand
            // reads per 64 tensor cycles, now 8 KB)
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              tc_mma_w(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), C::kIdescSplit2, (chunk_start && k == 0) ? 0u : 1u);
          } else {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
              tc_mma_w(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), C::kIdesc, (chunk_start && k == 0) ? 0u : 1u);
            }
          }
          if (DS) {
            // the centre tap of a 3x3 / stride 2 / pad 1 filter samples exactly the pixels a 1x1 / stride 2
            // convolution reads: the block's downsample branch rides on the same A tiles, second accumulator
            const int tap = kb / p.kchunks;
            if (tap == p.ds_tap) {
              const int c = kb - tap * p.kchunks;
              const uint64_t b2_desc = smem_desc_sw128(sa + kABytes + C::kBBytes);
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                tc_mma_w(d_tmem + (uint32_t)BN, a_desc + (uint64_t)(2 * k), b2_desc + (uint64_t)(2 * k), C::kIdesc, (c | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit_w(empty_bar(stage));  // frees the smem stage once these MMAs have read it
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1u;
          }
          if (chunk_end) {
            tc_commit_w(tfull_bar(acc));  // accumulator (chunk) complete
            if (++acc == 2) {
              acc = 0;
              acc_phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (PRE && warp >= 2 + kEpiWarps) {
    // ===== kModePre transform: 8 warps, thread -> (pixel row, four of its eight 16-byte channel chunks).  SWIZZLE_128B
    // puts logical chunk j of row r at physical chunk j ^ (r & 7).  Same arithmetic as the stand-alone affine_relu kernel
    // (fp32 fma, ReLU, round to bf16), so the fused result is bit-identical to the unfused one. =====
    const int tt = threadIdx.x - kThreads;
    const int r = tt & 127, jh = (tt >> 7) * 4;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int c0 = (kb % p.kchunks) * kBK;
        mbar_wait(full_bar(stage), phase);
        unsigned char* row = gen_base + stage * C::kStageBytes + r * 128;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = jh + jj;
          uint4* cp = reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4));
          uint4 q = *cp;
          __nv_bfloat162* b2 = reinterpret_cast<__nv_bfloat162*>(&q);
          // the chunk's 8 scales and 8 shifts as four 16-byte loads (all lanes of a warp read the same words: broadcasts; with
          // sixteen 4-byte loads this line was 14 % of the kernel's samples)
          const float4 s0 = *reinterpret_cast<const float4*>(pre_tab + c0 + 8 * j), s1 = *reinterpret_cast<const float4*>(pre_tab + c0 + 8 * j + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(pre_tab + 1024 + c0 + 8 * j), h1 = *reinterpret_cast<const float4*>(pre_tab + 1024 + c0 + 8 * j + 4);
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(b2[e]);
            float a = fmaf(f.x, sc[2 * e], sh[2 * e]), b = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
            if (p.pre_relu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            b2[e] = __floats2bfloat162_rn(a, b);
          }
          *cp = q;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) mbar_arrive(ready_bar(stage));
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 2 && warp < 2 + kEpiWarps) {
    // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32); the two warps of a quarter take half of the
    // tile's columns each (BN = 32: the second one only keeps the barrier counts) =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kHalfCols = BN >= 64 ? BN / 2 : BN;
    const int c_lo = BN >= 64 ? half * kHalfCols : 0;
    const bool works = BN >= 64 || half == 0;
    const int row = q * 32 + lane;  // pixel of the tile == TMEM lane
    const int wbi = row % p.wb;
    const int hbi = (row / p.wb) % p.hb;
    const int nbi = row / (p.wb * p.hb);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int tl = p.reverse ? p.total_tiles - 1 - tile : tile;
      const int nt = tl % p.tiles_n;
      int m = tl / p.tiles_n;
      const int twi = m % p.tiles_w;
      m /= p.tiles_w;
      const int thi = m % p.tiles_h;
      const int ng = m / p.tiles_h;
      const int wo = twi * p.wb + wbi, ho = thi * p.hb + hbi, img = ng * p.nb + nbi;
      const bool valid = (wo < p.wo) && (ho < p.ho) && (img < p.n);
      const long long pix = ((long long)img * p.ho + ho) * p.wo + wo;
      if constexpr (!SPLIT) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
      }
      // one pass over BN accumulator columns: + bias (+ residual) (ReLU) -> bf16 -> 16-byte stores of the pixel's row
      auto drain = [&](uint32_t taddr, const float* brow, const __nv_bfloat16* rrow, __nv_bfloat16* yrow, bool relu) {
#pragma unroll 1
        for (int c = c_lo; c < c_lo + kHalfCols; c += 32) {
          if (!works) break;
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c, v);
          tmem_ld_wait();
          if (valid) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + c + j));
              f[j] = __uint_as_float(v[j]) + b4.x;
              f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
              f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
              f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
            }
            if (rrow) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(rrow + c + j));
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 rf = __bfloat1622float2(r2[t]);
                  f[j + 2 * t] += rf.x;
                  f[j + 2 * t + 1] += rf.y;
                }
              }
            }
            if (relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const __nv_bfloat162 t = __floats2bfloat162_rn(f[j], f[j + 1]);
              o[j >> 1] = *reinterpret_cast<const uint32_t*>(&t);
            }
            if (p.st32) {  // 256-bit stores: whole 32-byte sectors (16-byte stores at a pixel stride write every sector in two halves)
#pragma unroll
              for (int j = 0; j < 32; j += 16)
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yrow + c + j), "r"(o[j / 2]), "r"(o[j / 2 + 1]),
                             "r"(o[j / 2 + 2]), "r"(o[j / 2 + 3]), "r"(o[j / 2 + 4]), "r"(o[j / 2 + 5]), "r"(o[j / 2 + 6]), "r"(o[j / 2 + 7])
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(yrow + c + j) = make_uint4(o[j / 2], o[j / 2 + 1], o[j / 2 + 2], o[j / 2 + 3]);
            }
          }
          __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the predicated stores
        }
      };
      if constexpr (SPLIT) {
        // SplitF in / out: one 32-bit word per channel (ldy / ldres count words).  The accumulation chunks of the tile arrive
        // one after the other; their partial sums (hi-weight + lo-weight accumulator) are added up here in registers
        const uint32_t* rrow = p.res ? reinterpret_cast<const uint32_t*>(p.res) + pix * p.ldres + nt * BN : nullptr;
        uint32_t* yrow = reinterpret_cast<uint32_t*>(p.y) + pix * p.ldy + nt * BN;
        const float* brow = p.bias + nt * BN;
        float part[kHalfCols];  // this warp's columns [c_lo, c_lo + kHalfCols)
#pragma unroll
        for (int c = 0; c < kHalfCols; ++c) part[c] = 0.f;
        const int nchunks = (kblocks + p.split_chunk - 1) / p.split_chunk;
        for (int ch = 0; ch < nchunks; ++ch) {
          mbar_wait(tfull_bar(acc), acc_phase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::kAccCols);
          if (works) {
#pragma unroll
            for (int c = 0; c < kHalfCols; c += 16) {
              uint32_t v[16], vl[16];
              tc::tmem_ld16(taddr + (uint32_t)(c_lo + c), v);
              tc::tmem_ld16(taddr + (uint32_t)(BN + c_lo + c), vl);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) part[c + j] += __uint_as_float(v[j]) + __uint_as_float(vl[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1u;
          }
        }
        bool ovf = false;
        if (valid && works) {
          const float* srow = p.wscale + nt * BN + c_lo;
          brow += c_lo;
          yrow += c_lo;
          if (rrow) rrow += c_lo;
#pragma unroll
          for (int c = 0; c < kHalfCols; c += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(brow + c));
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(srow + c));  // powers of two: exact
            float f[4] = {fmaf(part[c], s4.x, b4.x), fmaf(part[c + 1], s4.y, b4.y), fmaf(part[c + 2], s4.z, b4.z), fmaf(part[c + 3], s4.w, b4.w)};
            if (rrow) {
              const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(rrow + c));
              f[0] += split_load(r4.x);
              f[1] += split_load(r4.y);
              f[2] += split_load(r4.z);
              f[3] += split_load(r4.w);
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            ovf |= !(fmaxf(fmaxf(fabsf(f[0]), fabsf(f[1])), fmaxf(fabsf(f[2]), fabsf(f[3]))) <= kSplitMax);  // (also catches NaN)
            *reinterpret_cast<uint4*>(yrow + c) = make_uint4(split_store(f[0]), split_store(f[1]), split_store(f[2]), split_store(f[3]));
          }
        }
        __syncwarp();
        if (__any_sync(0xffffffffu, ovf) && lane == 0) atomicAdd(p.faults, 1ULL);
        continue;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C::kAccCols);
      drain(taddr, p.bias + nt * BN, p.res ? p.res + pix * p.ldres + nt * BN : nullptr, p.y + pix * p.ldy + nt * BN, p.relu != 0);
      if (DS) drain(taddr + (uint32_t)BN, p.bias2 + nt * BN, nullptr, p.y2 + pix * p.ldy2 + nt * BN, false);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc::stamp_end(p.stamp);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

int pick_bn(int cout) {
  for (int bn : {256, 128, 64, 32})
    if (cout % bn == 0) return bn;
  return 0;
}

}  // namespace

struct TcConvPlan {
  ConvGeom g;  // split plans: the bf16 VIEW of the input (cin, ldx doubled); cout, ldy, ldres stay in SplitF words
  int bn = 0;
  bool ds = false;
  bool split = false;
  bool pre = false;
  TcParams prm;
  __nv_bfloat16* d_w = nullptr;
  __nv_bfloat16* d_w2 = nullptr;
  float* d_wscale = nullptr;  // split plans
  int64_t bytes = 0;
  const void* x_ptr = nullptr;
  int n_maps = 0;
  int map_hp[4], map_wp[4];
};

bool tc_conv_supported(const ConvGeom& g) {
  if (g.kh * g.kw > kMaxTaps) return false;
  if (g.stride != 1 && g.stride != 2) return false;
  if (g.cin % 8 != 0 || g.cin < 16 || g.ldx % 8 != 0) return false;
  if (pick_bn(g.cout) == 0) return false;
  if (g.ldy % 8 != 0 || (g.ldres % 8) != 0) return false;
  return encode_fn() != nullptr;
}

static int encode_a_maps(spk_ctx* ctx, TcConvPlan* p, const void* x) {
  const ConvGeom& g = p->g;
  EncodeTiledFn enc = encode_fn();
  for (int i = 0; i < p->n_maps; ++i) {
    const int hp = p->map_hp[i], wp = p->map_wp[i], s = g.stride;
    const char* base = (const char*)x + ((size_t)hp * g.w + wp) * g.ldx * 2;
    cuuint64_t dims[4] = {(cuuint64_t)g.cin, (cuuint64_t)((g.w - wp + s - 1) / s), (cuuint64_t)((g.h - hp + s - 1) / s),
                          (cuuint64_t)g.n};
    cuuint64_t strides[3] = {(cuuint64_t)s * g.ldx * 2, (cuuint64_t)s * g.w * g.ldx * 2, (cuuint64_t)g.h * g.w * g.ldx * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)p->prm.wb, (cuuint32_t)p->prm.hb, (cuuint32_t)p->prm.nb};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&p->prm.map_a[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(ctx, SPK_ERR_CUDA, "cuTensorMapEncodeTiled(A, %dx%d s%d cin %d, box %dx%dx%d) failed: %d", g.kh, g.kw, s,
                  g.cin, p->prm.wb, p->prm.hb, p->prm.nb, (int)r);
  }
  for (int i = p->n_maps; i < 4; ++i) p->prm.map_a[i] = p->prm.map_a[0];
  p->x_ptr = x;
  return SPK_OK;
}

bool tc_conv_ds_fusable(const ConvGeom& g3, const ConvGeom& g1) {
  return g3.kh == 3 && g3.kw == 3 && g3.stride == 2 && g3.pad == 1 && g1.kh == 1 && g1.kw == 1 && g1.stride == 2 && g1.pad == 0 &&
         g3.cin == g1.cin && g3.cout == g1.cout && g3.h == g1.h && g3.w == g1.w && g3.ldx == g1.ldx && g3.ho == g1.ho &&
         g3.wo == g1.wo && g1.relu == 0 && g3.cout % 64 == 0 && g1.ldy % 8 == 0 && tc_conv_supported(g3) && tc_conv_supported(g1);
}

bool tc_conv_split_supported(const ConvGeom& g) {
  ConvGeom v = g;
  v.cin = 2 * g.cin;
  v.ldx = 2 * g.ldx;
  if (g.cin % 4 != 0 || g.ldy % 4 != 0 || g.ldres % 4 != 0) return false;  // 16-byte SplitF accesses in the epilogue
  return tc_conv_supported(v) && pick_bn(g.cout) != 0;
}

int tc_conv_plan_set_prologue(spk_ctx* ctx, TcConvPlan* p, const float* d_scale, const float* d_shift, int channels, int relu) {
  if (!p || p->ds || p->split) return fail(ctx, SPK_ERR_STATE, "tcgen05 convolution: prologue on a fused / split plan");
  if (channels > 1024 || channels > p->g.cin) return fail(ctx, SPK_ERR_UNSUPPORTED, "tcgen05 convolution: prologue over %d channels", channels);
  p->pre = true;
  p->prm.pre_scale = d_scale;
  p->prm.pre_shift = d_shift;
  p->prm.pre_c = channels;
  p->prm.pre_relu = relu;
  cudaError_t ea = cudaSuccess;
  switch (p->bn) {
    case 256: ea = cudaFuncSetAttribute(conv_tc_kernel<256, kModePre>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<256, kModePre>::kSmem); break;
    case 128: ea = cudaFuncSetAttribute(conv_tc_kernel<128, kModePre>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128, kModePre>::kSmem); break;
    case 64: ea = cudaFuncSetAttribute(conv_tc_kernel<64, kModePre>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<64, kModePre>::kSmem); break;
    case 32: ea = cudaFuncSetAttribute(conv_tc_kernel<32, kModePre>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<32, kModePre>::kSmem); break;
  }
  if (ea != cudaSuccess) return fail(ctx, SPK_ERR_CUDA, "tcgen05 convolution: cudaFuncSetAttribute: %s", cudaGetErrorString(ea));
  return SPK_OK;
}

int tc_conv_plan_create(spk_ctx* ctx, const ConvGeom& g_max, const float* w, const float* d_bias, TcConvPlan** out,
                        const float* w_ds, const float* d_bias_ds, int ldy_ds, bool split) {
  if (split ? (!tc_conv_split_supported(g_max) || w_ds != nullptr) : !tc_conv_supported(g_max))
    return fail(ctx, SPK_ERR_UNSUPPORTED, "tcgen05 convolution: unsupported geometry");
  TcConvPlan* p = new TcConvPlan;
  p->g = g_max;
  p->split = split;
  const int cin_logical = g_max.cin;
  if (split) {
    p->g.cin = 2 * g_max.cin;
    p->g.ldx = 2 * g_max.ldx;
  }
  const ConvGeom& g = p->g;
  memset(&p->prm, 0, sizeof p->prm);
  TcParams& prm = p->prm;
  p->bn = pick_bn(g.cout);
  p->ds = w_ds != nullptr;
  if (p->ds) p->bn = g.cout % 128 == 0 ? 128 : 64;  // two accumulators per buffer: 4 * BN TMEM columns
  if (p->split && p->bn == 256) p->bn = 128;        // two weight tiles per stage
  prm.ds_tap = p->ds ? 4 : -1;                       // (r, s) = (1, 1)
  prm.bias2 = d_bias_ds;
  prm.ldy2 = ldy_ds;
  prm.bias = d_bias;
  prm.ho = g.ho;
  prm.wo = g.wo;
  prm.cout = g.cout;
  prm.ldy = g.ldy;
  prm.ldres = g.ldres;
  prm.relu = g.relu;
  prm.taps = g.kh * g.kw;
  prm.kchunks = (g.cin + kBK - 1) / kBK;
  prm.split_chunk = debug_env("SPK_SPLIT_CHUNK") ? std::max(1, atoi(debug_env("SPK_SPLIT_CHUNK"))) : kSplitChunk;
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

  // ---- weights: bf16 [Cout][taps][cin_pad], zero padded
  const size_t kk = (size_t)prm.taps * prm.cin_pad;
  std::vector<__nv_bfloat16> wb16((size_t)g.cout * kk, __float2bfloat16(0.f));
  // plain round-to-nearest: error-diffusion rounding along K was tried and measured (torch emulation of the
  // whole network, ResNet-18 and -50 checkpoints): no robust gain, worse for 1x1 layers.
  std::vector<__nv_bfloat16> wlo;  // (split plans: both buffers hold fp16 bit patterns)
  std::vector<float> wscale;
  if (p->split) {
    // [w_hi, w_hi] / [w_lo, w_lo] over the interleaved K index 2c + e, fp16 of w * 2^k(o): 2^k(o) brings the channel's largest
    // weight into [2^13, 2^14), so that hi AND lo are normal fp16 numbers for every weight above 1e-5 of the largest
    // (22 significant bits); the epilogue multiplies the accumulator by 2^-k(o)
    wlo.assign(wb16.size(), __float2bfloat16(0.f));
    wscale.assign((size_t)g.cout, 1.f);
    uint16_t* whi_bits = reinterpret_cast<uint16_t*>(wb16.data());
    uint16_t* wlo_bits = reinterpret_cast<uint16_t*>(wlo.data());
    for (int o = 0; o < g.cout; ++o) {
      double amax = 0.0;
      for (size_t i = 0; i < (size_t)prm.taps * cin_logical; ++i) amax = std::max(amax, std::fabs((double)w[(size_t)o * prm.taps * cin_logical + i]));
      int k = 0;
      if (amax > 0.0 && std::isfinite(amax)) {
        int e;
        std::frexp(amax, &e);  // amax = m * 2^e, m in [0.5, 1)
        k = std::max(-100, std::min(100, 14 - e));
      }
      wscale[(size_t)o] = (float)std::ldexp(1.0, -k);
      for (int t = 0; t < prm.taps; ++t)
        for (int c = 0; c < cin_logical; ++c) {
          const float v = (float)std::ldexp((double)w[((size_t)o * prm.taps + t) * cin_logical + c], k);
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          uint16_t hb, lb;
          memcpy(&hb, &hi, 2);
          memcpy(&lb, &lo, 2);
          const size_t at = (size_t)o * kk + (size_t)t * prm.cin_pad + 2 * c;
          whi_bits[at] = whi_bits[at + 1] = hb;
          wlo_bits[at] = wlo_bits[at + 1] = lb;
        }
    }
  } else
  for (int o = 0; o < g.cout; ++o)
    for (int t = 0; t < prm.taps; ++t)
      for (int c = 0; c < g.cin; ++c)
        wb16[(size_t)o * kk + (size_t)t * prm.cin_pad + c] = __float2bfloat16(w[((size_t)o * prm.taps + t) * g.cin + c]);
  cudaError_t e = cudaMalloc(&p->d_w, wb16.size() * 2);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_w, wb16.data(), wb16.size() * 2, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    tc_conv_plan_destroy(p);
    return fail(ctx, SPK_ERR_CUDA, "tcgen05 convolution: weight upload: %s", cudaGetErrorString(e));
  }
  p->bytes = (int64_t)wb16.size() * 2;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kk, (cuuint64_t)g.cout};
    cuuint64_t strides[1] = {(cuuint64_t)kk * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)p->bn};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&prm.map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->d_w, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      tc_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    }
  }
  if (p->split) {
    cudaError_t e2 = cudaMalloc(&p->d_w2, wlo.size() * 2);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(p->d_w2, wlo.data(), wlo.size() * 2, cudaMemcpyHostToDevice);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&p->d_wscale, wscale.size() * 4);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(p->d_wscale, wscale.data(), wscale.size() * 4, cudaMemcpyHostToDevice);
    prm.wscale = p->d_wscale;
    prm.faults = ctx->d_faults;
    if (e2 != cudaSuccess) {
      tc_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "tcgen05 convolution: lo weight upload: %s", cudaGetErrorString(e2));
    }
    p->bytes += (int64_t)wlo.size() * 2;
    cuuint64_t dims[2] = {(cuuint64_t)kk, (cuuint64_t)g.cout};
    cuuint64_t strides[1] = {(cuuint64_t)kk * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)p->bn};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&prm.map_b2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->d_w2, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      tc_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "cuTensorMapEncodeTiled(W lo) failed: %d", (int)r);
    }
  }
  if (p->ds) {
    // downsample weights: bf16 [Cout][cin_pad]
    const size_t k1 = (size_t)prm.cin_pad;
    std::vector<__nv_bfloat16> w2((size_t)g.cout * k1, __float2bfloat16(0.f));
    for (int o = 0; o < g.cout; ++o)
      for (int c = 0; c < g.cin; ++c) w2[(size_t)o * k1 + c] = __float2bfloat16(w_ds[(size_t)o * g.cin + c]);
    cudaError_t e2 = cudaMalloc(&p->d_w2, w2.size() * 2);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(p->d_w2, w2.data(), w2.size() * 2, cudaMemcpyHostToDevice);
    if (e2 != cudaSuccess) {
      tc_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "tcgen05 convolution: downsample weight upload: %s", cudaGetErrorString(e2));
    }
    p->bytes += (int64_t)w2.size() * 2;
    cuuint64_t dims[2] = {(cuuint64_t)k1, (cuuint64_t)g.cout};
    cuuint64_t strides[1] = {(cuuint64_t)k1 * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)p->bn};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&prm.map_b2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p->d_w2, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      tc_conv_plan_destroy(p);
      return fail(ctx, SPK_ERR_CUDA, "cuTensorMapEncodeTiled(W downsample) failed: %d", (int)r);
    }
  }
  // opt in to the large dynamic shared memory on THIS device (the attribute is per device)
  cudaError_t ea = cudaSuccess;
  if (p->split) {
    if (p->bn == 128) ea = cudaFuncSetAttribute(conv_tc_kernel<128, kModeSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128, kModeSplit>::kSmem);
    else if (p->bn == 64) ea = cudaFuncSetAttribute(conv_tc_kernel<64, kModeSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<64, kModeSplit>::kSmem);
    else ea = cudaFuncSetAttribute(conv_tc_kernel<32, kModeSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<32, kModeSplit>::kSmem);
  } else if (p->ds) {
    if (p->bn == 128) ea = cudaFuncSetAttribute(conv_tc_kernel<128, kModeDs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128, kModeDs>::kSmem);
    else ea = cudaFuncSetAttribute(conv_tc_kernel<64, kModeDs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<64, kModeDs>::kSmem);
  } else switch (p->bn) {
    case 256: ea = cudaFuncSetAttribute(conv_tc_kernel<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<256>::kSmem); break;
    case 128: ea = cudaFuncSetAttribute(conv_tc_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<128>::kSmem); break;
    case 64: ea = cudaFuncSetAttribute(conv_tc_kernel<64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<64>::kSmem); break;
    case 32: ea = cudaFuncSetAttribute(conv_tc_kernel<32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<32>::kSmem); break;
  }
  if (ea != cudaSuccess) {
    tc_conv_plan_destroy(p);
    return fail(ctx, SPK_ERR_CUDA, "tcgen05 convolution: cudaFuncSetAttribute: %s", cudaGetErrorString(ea));
  }
  *out = p;
  return SPK_OK;
}

void tc_conv_plan_destroy(TcConvPlan* p) {
  if (!p) return;
  if (p->d_w) cudaFree(p->d_w);
  if (p->d_w2) cudaFree(p->d_w2);
  if (p->d_wscale) cudaFree(p->d_wscale);
  delete p;
}

int64_t tc_conv_plan_bytes(const TcConvPlan* p) { return p ? p->bytes : 0; }

void tc_conv_plan_set_reverse(TcConvPlan* p, int reverse) {
  if (p) p->prm.reverse = reverse ? 1 : 0;
}

template <int BN, int MODE = 0>
static int launch_bn(spk_ctx* ctx, const TcParams& prm) {
  using C = Cfg<BN, MODE>;
  const int grid = std::min(prm.total_tiles, ctx->sm_count);
  SPK_CUDA_OK(ctx, tc::launch_pdl(conv_tc_kernel<BN, MODE>, dim3(grid), dim3(MODE == kModePre ? kThreadsPre : kThreads), C::kSmem, ctx->stream, prm));
  SPK_LAUNCH_CHECK(ctx);
  return SPK_OK;
}

int tc_conv_launch(spk_ctx* ctx, TcConvPlan* p, int n, const void* x, const void* res, void* y, void* y_ds) {
  if (n <= 0) return SPK_OK;
  if (n > p->g.n) return fail(ctx, SPK_ERR_CAPACITY, "tcgen05 convolution: batch %d > planned %d", n, p->g.n);
  if (x != p->x_ptr) {
    int rc = encode_a_maps(ctx, p, x);
    if (rc) return rc;
  }
  TcParams& prm = p->prm;
  prm.n = n;
  prm.res = (const __nv_bfloat16*)res;
  prm.y = (__nv_bfloat16*)y;
  prm.st32 = (((uintptr_t)y & 31) == 0 && p->g.ldy % 16 == 0 && (!y_ds || (((uintptr_t)y_ds & 31) == 0 && prm.ldy2 % 16 == 0))) ? 1 : 0;
  prm.y2 = (__nv_bfloat16*)y_ds;
  prm.stamp = ctx->cur_stamp;
  if (p->ds && !y_ds) return fail(ctx, SPK_ERR_INVALID, "tcgen05 convolution: fused downsample without an output");
  prm.tiles_img = (n + prm.nb - 1) / prm.nb;
  prm.total_tiles = prm.tiles_w * prm.tiles_h * prm.tiles_img * prm.tiles_n;
  if (p->pre) switch (p->bn) {
      case 256: return launch_bn<256, kModePre>(ctx, prm);
      case 128: return launch_bn<128, kModePre>(ctx, prm);
      case 64: return launch_bn<64, kModePre>(ctx, prm);
      case 32: return launch_bn<32, kModePre>(ctx, prm);
    }
  if (p->split) return p->bn == 128 ? launch_bn<128, kModeSplit>(ctx, prm) : p->bn == 64 ? launch_bn<64, kModeSplit>(ctx, prm) : launch_bn<32, kModeSplit>(ctx, prm);
  if (p->ds) return p->bn == 128 ? launch_bn<128, kModeDs>(ctx, prm) : launch_bn<64, kModeDs>(ctx, prm);
  switch (p->bn) {
    case 256: return launch_bn<256>(ctx, prm);
    case 128: return launch_bn<128>(ctx, prm);
    case 64: return launch_bn<64>(ctx, prm);
    case 32: return launch_bn<32>(ctx, prm);
  }
  return fail(ctx, SPK_ERR_STATE, "tcgen05 convolution: bad plan");
}

}  // namespace spk

// K2 (bf16 mode): fused per-sample NeRF forward on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces points_3D -> positional encoding -> 8-layer trunk -> rgb head -> transient head of
// layers/nerf_static_transient_light.py:76-145 (~46 aten kernels and ~30 HBM round trips per chunk in the
// reference) with ONE persistent kernel: activations never leave the SM.
//
// Work decomposition
//   * super-tile = 256 consecutive samples = two M=128 MMA tiles (T0, T1) that march through the 17 GEMM
//     stages in lock-step so every streamed weight chunk is used twice;
//   * per stage and tile:  D[128 x 256 fp32, TMEM] = A[128 x K bf16, SMEM] * W^T[K x 256 bf16, SMEM]
//     issued as tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256 or 16, K=16) by ONE thread;
//   * weights are pre-packed (tp_tc_pack_weights) into the exact SMEM image of each 16 KB K-chunk and
//     streamed through a 4-stage ring with cp.async.bulk (TMA bulk copy) + mbarrier complete_tx;
//   * 8 epilogue warps (one TMEM lane quarter each, 4 per tile): tcgen05.ld -> +bias -> ReLU -> bf16 ->
//     st.shared straight into the K-major core-matrix layout the next stage's A descriptor reads;
//   * per-ray constants (view-direction encoding, light latent) and per-image constants (transient latent)
//     are folded into bias tables by two tiny fp32 pre-kernels, so they cost no MMA work per sample;
//   * the trunk feature (needed by both heads) is parked in an L2-resident scratch with a bulk store after
//     the last trunk layer and bulk-loaded back before the transient head.
//
// SMEM (bytes): A0 64K | A1 64K | E0 16K | E1 16K | ring 4x16K | barriers = 229 504 (<= 227 KB opt-in).
// TMEM: 512 columns = two 128x256 fp32 accumulators.
//
// Operand layout (no swizzle, K-major "interleave" canonical layout): element (row r, col k) of a tile with
// `rows` rows lives at  (k/8)*rows*16 + r*16 + (k%8)*2  bytes -> 8x8 core matrices of 128 contiguous bytes,
// LBO (K-direction core-matrix stride) = rows*16, SBO (8-row-group stride) = 128.
#include "common.cuh"
#include "../../include/texpose_b200.h"

namespace tc {

constexpr int kThreads = 320;          // warps 0-7 epilogue/encode, warp 8 TMA producer, warp 9 MMA issuer
constexpr int kStages = 4;
constexpr uint32_t kChunkBytes = 16384;
constexpr uint32_t kABytes = 65536;    // 128 x 256 bf16
constexpr uint32_t kEBytes = 16384;    // 128 x 64 bf16
constexpr uint32_t kOffA = 0, kOffE = 2 * kABytes, kOffRing = kOffE + 2 * kEBytes;
constexpr uint32_t kOffBar = kOffRing + kStages * kChunkBytes;
constexpr uint32_t kSmemBytes = kOffBar + 128;
constexpr int kNumLayers = 17;
constexpr int kNumChunks = 122;

// stage table: chunks read from A (K=32 each), chunks read from E, small (N=16, one chunk spans K=256),
// kind of epilogue, bias handling, needs the feature reload first.
// Static biases of the 256-wide stages ride on the tensor cores: column 63 of the encoding tile is a constant 1 and
// the bias sits in the matching weight column -- for stages that read E anyway (trunk 0 and 4) inside their last E
// chunk, for the others as one extra K=16 step on E columns 48..63 with an 8 KB weight chunk that is zero except for
// that column (BIAS_MMA).  Their epilogue is then a pure convert.  Per-ray / per-image biases (fp32 tables) and the
// three N=16 output stages add their bias in the epilogue.
enum Epi : int { EPI_HIDDEN = 0, EPI_DENSITY = 1, EPI_RGB_OUT = 2, EPI_TRANS_OUT = 3 };
enum BiasKind : int { BIAS_MMA = 0, BIAS_RAY = 1, BIAS_IMAGE = 2, BIAS_SMALL = 3 };
struct Layer {
  int a_chunks, e_chunks, small, epi, bias_kind, bias_chunk, reload;
};
__constant__ Layer kLayers[kNumLayers] = {
    {0, 2, 0, EPI_HIDDEN, BIAS_MMA, 0, 0},       // trunk 0  (63 -> 256); bias in E column 63 of its own chunk
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // trunk 1
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // trunk 2
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // trunk 3
    {8, 2, 0, EPI_HIDDEN, BIAS_MMA, 0, 0},       // trunk 4  (skip: [feat | enc]); bias in E column 63
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // trunk 5
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // trunk 6
    {8, 0, 1, EPI_DENSITY, BIAS_SMALL, 0, 0},    // trunk 7 row 0     -> sigma_static (softplus)
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // trunk 7 rows 1..  -> feature (relu); parked to L2 afterwards
    {8, 1, 0, EPI_HIDDEN, BIAS_RAY, 0, 0},       // rgb 0    ([feat | xyz]; view+light folded into the ray bias)
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // rgb 1
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // rgb 2
    {8, 0, 1, EPI_RGB_OUT, BIAS_SMALL, 0, 0},    // rgb 3    -> sigmoid
    {8, 0, 0, EPI_HIDDEN, BIAS_IMAGE, 0, 1},     // trans 0  (feature reloaded; transient latent in the image bias)
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // trans 1
    {8, 0, 0, EPI_HIDDEN, BIAS_MMA, 1, 0},       // trans 2
    {8, 0, 1, EPI_TRANS_OUT, BIAS_SMALL, 0, 0}   // trans 3  -> sigmoid x3, softplus x2
};
constexpr int kSpillLayer = 8, kFirstHeadLayer = 9, kReloadIssueLayer = 12;
constexpr int kSmallBiasOffset = 0;            // biasbuf: [density b, rgb3 b(3), trans3 b(5)] (fp32, 16 floats)

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// One lane of a converged warp (elect.sync): the surrounding control flow stays warp-uniform, so descriptor and barrier
// operands live in uniform registers and UTCHMMA / UBLKCP / UTCBAR need no per-instruction R2UR waterfall loop.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) = 0 (SWIZZLE_NONE / interleave).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D=F32 (bits 4-5 = 1), A=B=BF16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Same instruction with the two 64-bit descriptors assembled from 32-bit halves inside the asm block: the issuing
// thread then spends one integer add per MMA on descriptor upkeep (the low word carries start>>4 and LBO>>4, the high
// word SBO>>4 and the version bit -- both constant per operand).
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TP_TMEM_LD32(taddr, v)                                                                                        \
  asm volatile(                                                                                                       \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                       \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "    \
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                          \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),        \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),       \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                     \
      : "r"(taddr)                                                                                                    \
      : "memory")
#define TP_TMEM_LD8(taddr, v)                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"      \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
               : "r"(taddr)                                                                         \
               : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait::ld that also names the destination registers, so no consumer of v[] can be scheduled above the wait
#define TP_TMEM_WAIT32(v)                                                                                              \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                        \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),       \
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), \
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),            \
                 "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),            \
                 "+r"(v[30]), "+r"(v[31])::"memory")
#define TP_TMEM_WAIT8(v)                                                                                         \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                  \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]):: "memory")

// {hi, lo} fp32 -> packed bf16x2 with ReLU (lo in the low half = the lower column index)
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  // no "memory" clobber: asm volatile statements keep their relative order (fences, barrier arrives), while plain
  // loads (the bias LDS of the next group) may be scheduled above the store
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}

struct Params {
  const float* center;       // [rays,3]
  const float* ray;          // [rays,3]
  const float* depth;        // [S]
  long long S;
  int N;                     // samples per ray
  long long per_image;       // samples per image
  const uint8_t* packed;     // kNumChunks x 16 KB weight image
  const float* biasbuf;      // 12 x 256 static biases + 16 small
  const float* raybias;      // [rays,256]  rgb-0 bias incl. view encoding + light latent
  const float* imgbias;      // [images,256] trans-0 bias incl. transient latent
  float* rgb;                // [S,3,2]
  float* density;            // [S,2]
  float* uncert;             // [S]
  uint8_t* scratch;          // gridDim.x x 2 x 64 KB (parked features)
  int dbg_layer;
  float* dbg_out;            // [S,256] post-activation of stage dbg_layer (debug only)
  int swap_lbo_sbo;          // debug: exchange the two descriptor strides
};

// positional encoding of one sample into the E tile (bf16): [x,y,z, per coord sin(2^k pi x) k<10, cos(...) k<10, 1].
// sincospif gives the exact-argument octave 0; higher octaves by the double-angle recurrence (error << bf16 ulp).
__device__ __forceinline__ void encode_sample(const Params& p, long long s, uint32_t e_smem, int row) {
  const long long r = s / p.N;
  const float d = p.depth[s];
  float v[64];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float x = __fadd_rn(p.center[r * 3 + j], __fmul_rn(p.ray[r * 3 + j], d));
    v[j] = x;
    float sn, cs;
    sincospif(x, &sn, &cs);
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      v[3 + j * 20 + k] = sn;
      v[3 + j * 20 + 10 + k] = cs;
      const float s2 = 2.f * sn * cs, c2 = (cs - sn) * (cs + sn);
      sn = s2;
      cs = c2;
    }
  }
  v[63] = 1.f;   // constant-1 column: carries the static biases through the MMA
#pragma unroll
  for (int k8 = 0; k8 < 8; ++k8)
    st_shared_v4(e_smem + k8 * 2048 + row * 16, pack_bf16(v[k8 * 8 + 0], v[k8 * 8 + 1]), pack_bf16(v[k8 * 8 + 2], v[k8 * 8 + 3]),
                 pack_bf16(v[k8 * 8 + 4], v[k8 * 8 + 5]), pack_bf16(v[k8 * 8 + 6], v[k8 * 8 + 7]));
}

// One 32-column slab of a hidden stage: (+fp32 bias,) ReLU, bf16, store as 4 core-matrix rows of the next A operand.
template <bool kBias>
__device__ __forceinline__ void hidden_slab(const uint32_t (&v)[32], const float* bias, uint32_t a_dst, float* dbg_row) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[i + e]);
    if (kBias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + i));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + i + 4));
      x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
      x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
    }
    st_shared_v4(a_dst + (i >> 3) * 2048, pack_relu_bf16(x[0], x[1]), pack_relu_bf16(x[2], x[3]), pack_relu_bf16(x[4], x[5]),
                 pack_relu_bf16(x[6], x[7]));
    if (dbg_row) {
#pragma unroll
      for (int e = 0; e < 8; ++e) dbg_row[i + e] = fmaxf(x[e], 0.f);
    }
  }
}

// Whole 128x256 accumulator row of one thread: TMEM loads are software-pipelined (slab j+1 in flight while slab j is
// converted), ping-ponging two register slabs.
template <bool kBias>
__device__ __forceinline__ void hidden_epilogue(uint32_t tmem_d, const float* bias, uint32_t a_row, float* dbg_row) {
  uint32_t va[32], vb[32];
  TP_TMEM_LD32(tmem_d, va);
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    TP_TMEM_WAIT32(va);
    TP_TMEM_LD32(tmem_d + (j + 1) * 32, vb);
    hidden_slab<kBias>(va, bias + j * 32, a_row + j * 4 * 2048, dbg_row ? dbg_row + j * 32 : nullptr);
    TP_TMEM_WAIT32(vb);
    if (j + 2 < 8) TP_TMEM_LD32(tmem_d + (j + 2) * 32, va);
    hidden_slab<kBias>(vb, bias + (j + 1) * 32, a_row + (j + 1) * 4 * 2048, dbg_row ? dbg_row + (j + 1) * 32 : nullptr);
  }
}

__global__ void __launch_bounds__(kThreads, 1) nerf_stl_forward_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform role index
  const int lane = threadIdx.x & 31;
  // barriers
  const uint32_t bar0 = sbase + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (kStages + s); };
  auto bar_acc = [&](int t) { return bar0 + 8 * (2 * kStages + t); };
  auto bar_ready = [&](int t) { return bar0 + 8 * (2 * kStages + 2 + t); };
  auto bar_reload = [&](int t) { return bar0 + 8 * (2 * kStages + 4 + t); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * (2 * kStages + 6));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_acc(t), 1);
      mbar_init(bar_ready(t), 128);
      mbar_init(bar_reload(t), 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {   // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_super = (p.S + 255) / 256;

  if (warp == 8) {
    // ================================================================ weight producer (converged warp, one lane issues)
    {
      uint32_t stage = 0, phase = 0;
      for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
        int c = 0;
        for (int L = 0; L < kNumLayers; ++L) {
          const Layer ly = kLayers[L];
          const int nch = ly.small ? 1 : ly.a_chunks + ly.e_chunks + ly.bias_chunk;
          for (int j = 0; j < nch; ++j, ++c) {
            // full 16 KB K=32 chunks; the N=16 chunk and the bias chunk only carry 8 KB
            const uint32_t bytes = (ly.small || j >= ly.a_chunks + ly.e_chunks) ? kChunkBytes / 2 : kChunkBytes;
            mbar_wait(bar_empty(stage), phase ^ 1);
            if (elect_one_sync()) {
              mbar_expect_tx(bar_full(stage), bytes);
              bulk_g2s(sbase + kOffRing + stage * kChunkBytes, p.packed + (size_t)c * kChunkBytes, bytes, bar_full(stage));
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ================================================================ MMA issuer (converged warp, one lane issues)
    {
      uint32_t stage = 0, phase = 0, ready_ph[2] = {0, 0}, reload_ph[2] = {0, 0};
      const uint32_t idesc256 = umma_idesc(128, 256), idesc16 = umma_idesc(128, 16);
      for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
        for (int L = 0; L < kNumLayers; ++L) {
          const Layer ly = kLayers[L];
          const int nch = ly.small ? 1 : ly.a_chunks + ly.e_chunks + ly.bias_chunk;
          for (int c = 0; c < nch; ++c) {
            mbar_wait(bar_full(stage), phase);
            tc_fence_after();
            const uint32_t wsm = sbase + kOffRing + stage * kChunkBytes;
            for (int t = 0; t < 2; ++t) {
              if (c == 0) {
                mbar_wait(bar_ready(t), ready_ph[t]);
                ready_ph[t] ^= 1;
                if (ly.reload) {
                  mbar_wait(bar_reload(t), reload_ph[t]);
                  reload_ph[t] ^= 1;
                }
                tc_fence_after();
              }
              const uint32_t d_tmem = tmem_base + t * 256;
              if (elect_one_sync()) {
              // descriptor words: lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version(1)<<14; SBO = 128 B everywhere
              constexpr uint32_t kHi = (128u >> 4) | (1u << 14);
              if (ly.small) {
                // one chunk = [32 k8][16 rows][8]: 16 K-steps over the full K=256 of A_t
                uint32_t a_lo = ((sbase + kOffA + t * kABytes) >> 4) | ((2048u >> 4) << 16);
                uint32_t b_lo = (wsm >> 4) | ((256u >> 4) << 16);
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                  umma_bf16_lohi(d_tmem, a_lo, kHi, b_lo, kHi, idesc16, ks > 0 ? 1u : 0u);
                  a_lo += 4096u >> 4;
                  b_lo += 512u >> 4;
                }
              } else if (c >= ly.a_chunks + ly.e_chunks) {
                // bias step: A = E columns 48..63 (column 63 == 1), B = [2 k8][256 rows][8], zero except the bias column
                const uint32_t a_lo = ((sbase + kOffE + t * kEBytes + 6 * 2048) >> 4) | ((2048u >> 4) << 16);
                const uint32_t b_lo = (wsm >> 4) | ((4096u >> 4) << 16);
                umma_bf16_lohi(d_tmem, a_lo, kHi, b_lo, kHi, idesc256, 1u);
              } else {
                const bool from_e = c >= ly.a_chunks;
                const uint32_t a0 = from_e ? sbase + kOffE + t * kEBytes + (c - ly.a_chunks) * 4 * 2048
                                           : sbase + kOffA + t * kABytes + c * 4 * 2048;
                const uint32_t a_lo = (a0 >> 4) | ((2048u >> 4) << 16);
                const uint32_t b_lo = (wsm >> 4) | ((4096u >> 4) << 16);
                umma_bf16_lohi(d_tmem, a_lo, kHi, b_lo, kHi, idesc256, c > 0 ? 1u : 0u);
                umma_bf16_lohi(d_tmem, a_lo + (4096u >> 4), kHi, b_lo + (8192u >> 4), kHi, idesc256, 1u);
              }
              if (c == nch - 1) umma_commit(bar_acc(t));   // accumulator of tile t complete
              }
              __syncwarp();
            }
            if (elect_one_sync()) umma_commit(bar_empty(stage));   // ring slot reusable once these MMAs retire
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ================================================================ encode + epilogue warps (0..7)
    const int t = warp >> 2, q = warp & 3, row = q * 32 + lane;
    const uint32_t a_smem = sbase + kOffA + t * kABytes, e_smem = sbase + kOffE + t * kEBytes;
    const uint32_t tmem_d = tmem_base + ((uint32_t)(q * 32) << 16) + t * 256;
    uint8_t* my_scratch = p.scratch + ((size_t)blockIdx.x * 2 + t) * kABytes;
    uint32_t acc_ph = 0;
    for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
      const long long s_raw = (st * 2 + t) * 128 + row;
      const bool live = s_raw < p.S;
      const long long s = live ? s_raw : p.S - 1;
      encode_sample(p, s, e_smem, row);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_ready(t));

      float sigma_s = 0.f, rgb_s[3] = {0.f, 0.f, 0.f};
      for (int L = 0; L < kNumLayers; ++L) {
        const Layer ly = kLayers[L];
        mbar_wait(bar_acc(t), acc_ph);
        acc_ph ^= 1;
        tc_fence_after();
        if (L == kFirstHeadLayer) {            // the parked-feature store must have finished reading A_t
          if (row == 0) bulk_wait_read();
          named_bar_sync(1 + t, 128);
        }
        if (L == kReloadIssueLayer && row == 0) {
          // every MMA that reads A_t has retired (acc barrier) -> bring the trunk feature back for the transient head
          bulk_wait_all();
          fence_proxy_async_all();
          mbar_expect_tx(bar_reload(t), kABytes);
          bulk_g2s(a_smem, my_scratch, kABytes, bar_reload(t));
        }
        if (ly.epi == EPI_HIDDEN) {
          float* dbg_row = ((L == p.dbg_layer) && live && p.dbg_out) ? p.dbg_out + s * 256 : nullptr;
          if (ly.bias_kind == BIAS_MMA) {
            hidden_epilogue<false>(tmem_d, nullptr, a_smem + row * 16, dbg_row);
          } else {
            const float* bias = ly.bias_kind == BIAS_RAY ? p.raybias + (s / p.N) * 256 : p.imgbias + (s / p.per_image) * 256;
            hidden_epilogue<true>(tmem_d, bias, a_smem + row * 16, dbg_row);
          }
          fence_proxy_async_smem();
          if (L == kSpillLayer) {               // park the trunk feature (bf16 tile image) in the L2 scratch
            named_bar_sync(1 + t, 128);
            if (row == 0) {
              bulk_s2g(my_scratch, a_smem, kABytes);
              bulk_commit();
            }
          }
        } else {
          uint32_t v[8];
          TP_TMEM_LD8(tmem_d, v);
          TP_TMEM_WAIT8(v);
          const float* sb = p.biasbuf + kSmallBiasOffset;
          if (ly.epi == EPI_DENSITY) {
            sigma_s = tp_softplus(__uint_as_float(v[0]) + sb[0]);
          } else if (ly.epi == EPI_RGB_OUT) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb_s[c] = tp_sigmoid(__uint_as_float(v[c]) + sb[1 + c]);
          } else {
            float rgb_t[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb_t[c] = tp_sigmoid(__uint_as_float(v[c]) + sb[4 + c]);
            const float sigma_t = tp_softplus(__uint_as_float(v[3]) + sb[7]);
            const float unc = tp_softplus(__uint_as_float(v[4]) + sb[8]);
            if (live) {
#pragma unroll
              for (int c = 0; c < 3; ++c)
                *reinterpret_cast<float2*>(p.rgb + s * 6 + c * 2) = make_float2(rgb_s[c], rgb_t[c]);
              *reinterpret_cast<float2*>(p.density + s * 2) = make_float2(sigma_s, sigma_t);
              p.uncert[s] = unc;
            }
          }
        }
        if (L != kNumLayers - 1) {   // the next super-tile's encode arrival covers the last stage
          tc_fence_before();
          mbar_arrive(bar_ready(t));
        }
      }
    }
    if (row == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ helper kernels

// desc row: [w_ptr, ld, row0, rows_valid, col0, cols_valid, n_layout, bias_ptr, bias_klocal, 0]
// (bias_ptr != 0: element (n, bias_klocal) of the chunk is bias[n] -- the column multiplied by the constant-1 of E)
__global__ void pack_weights_kernel(const long long* __restrict__ desc, __nv_bfloat16* __restrict__ out) {
  const long long* d = desc + (long long)blockIdx.x * 10;
  const float* W = reinterpret_cast<const float*>(d[0]);
  const long long ld = d[1], row0 = d[2], rows_valid = d[3], col0 = d[4], cols_valid = d[5], n_layout = d[6];
  const float* bias = reinterpret_cast<const float*>(d[7]);
  const long long bias_k = d[8];
  __nv_bfloat16* o = out + (long long)blockIdx.x * (kChunkBytes / 2);
  for (int e = threadIdx.x; e < (int)(kChunkBytes / 2); e += blockDim.x) {
    int n, kl;
    bool in_layout = true;
    if (n_layout == 256) {
      kl = (e >> 11) * 8 + (e & 7);
      n = (e >> 3) & 255;
    } else {
      kl = (e >> 7) * 8 + (e & 7);
      n = (e >> 3) & 15;
      in_layout = e < 4096;
    }
    float v = 0.f;
    if (in_layout && W && n < rows_valid && kl < cols_valid) v = W[(row0 + n) * ld + col0 + kl];
    if (in_layout && bias && kl == bias_k && n < rows_valid) v = bias[n];
    o[e] = __float2bfloat16_rn(v);
  }
}

// out[b, n] = bias[n] + sum_j W[n, col0 + j] * latent[b, j]
__global__ void image_bias_kernel(const float* __restrict__ W, long long ldw, int col0, int ncols,
                                  const float* __restrict__ bias, const float* __restrict__ latent, int nout,
                                  float* __restrict__ out) {
  const int b = blockIdx.x;
  for (int n = threadIdx.x; n < nout; n += blockDim.x) {
    float acc = bias ? bias[n] : 0.f;
    for (int j = 0; j < ncols; ++j) acc = fmaf(W[n * ldw + col0 + j], latent[(long long)b * ncols + j], acc);
    out[(long long)b * nout + n] = acc;
  }
}

// out[r, n] = imgbias[r / rays_per_image, n] + sum_j W[n, col0 + j] * viewenc(r)[j];  viewenc = [u, enc(u)], u = ray/|ray|
__global__ void __launch_bounds__(256) ray_bias_kernel(const float* __restrict__ ray, long long R, long long rays_per_image,
                                                       int L, const float* __restrict__ W, long long ldw, int col0,
                                                       const float* __restrict__ imgbias, float* __restrict__ out,
                                                       int rays_per_block) {
  extern __shared__ float sm[];      // Wv[vc][256] then enc[vc]
  const int vc = 3 + 6 * L;
  float* Wv = sm;
  float* enc = sm + vc * 256;
  for (int i = threadIdx.x; i < vc * 256; i += blockDim.x) {
    const int j = i / 256, n = i % 256;
    Wv[i] = W[n * ldw + col0 + j];
  }
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rays_per_block;
  for (long long r = r0; r < r0 + rays_per_block && r < R; ++r) {
    if (threadIdx.x < 3 * (1 + 2 * L)) {
      // thread -> one output column of the encoding
      const float x = ray[r * 3], y = ray[r * 3 + 1], z = ray[r * 3 + 2];
      const float len = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
      const float u[3] = {x / len, y / len, z / len};
      const int i = threadIdx.x;
      float val;
      if (i < 3) val = u[i];
      else {
        const int e = i - 3, c = e / (2 * L), k = e % (2 * L);
        const float arg = __fmul_rn(u[c], ldexpf(3.14159265358979323846f, k < L ? k : k - L));
        val = k < L ? sinf(arg) : cosf(arg);
      }
      enc[i] = val;
    }
    __syncthreads();
    const int n = threadIdx.x;
    float acc = imgbias[(r / rays_per_image) * 256 + n];
    for (int j = 0; j < vc; ++j) acc = fmaf(Wv[j * 256 + n], enc[j], acc);
    out[r * 256 + n] = acc;
    __syncthreads();
  }
}

}  // namespace tc

TP_API int tp_tc_num_chunks(void) { return tc::kNumChunks; }
TP_API int64_t tp_tc_chunk_bytes(void) { return tc::kChunkBytes; }
TP_API int64_t tp_tc_scratch_bytes(void) { return (int64_t)tp_num_sms() * 2 * tc::kABytes; }

TP_API int tp_tc_pack_weights(const int64_t* chunk_desc, int n_chunks, void* packed, void* stream) {
  if (!chunk_desc || !packed) return TP_ERR_BAD_ARG;
  if (n_chunks != tc::kNumChunks) return TP_ERR_BAD_SHAPE;
  tc::pack_weights_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(chunk_desc),
                                                                    reinterpret_cast<__nv_bfloat16*>(packed));
  return tp_launch_status();
}

TP_API int tp_tc_image_bias(const float* W, int64_t ldw, int col0, int ncols, const float* bias, const float* latent,
                            int B, int nout, float* out, void* stream) {
  if (!W || !latent || !out) return TP_ERR_BAD_ARG;
  if (B < 1 || nout < 1 || ncols < 0) return TP_ERR_BAD_SHAPE;
  tc::image_bias_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(W, ldw, col0, ncols, bias, latent, nout, out);
  return tp_launch_status();
}

TP_API int tp_tc_ray_bias(const float* ray, int64_t R, int64_t rays_per_image, int L_view, const float* W, int64_t ldw,
                          int col0, const float* imgbias, float* out, void* stream) {
  if (!ray || !W || !imgbias || !out) return TP_ERR_BAD_ARG;
  if (R < 0 || rays_per_image < 1 || L_view < 0 || L_view > 16) return TP_ERR_BAD_SHAPE;
  if (R == 0) return TP_OK;
  const int vc = 3 + 6 * L_view;
  const int rpb = 64;
  const size_t smem = (size_t)(vc * 256 + vc) * sizeof(float);
  cudaFuncSetAttribute(tc::ray_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc::ray_bias_kernel<<<(unsigned)((R + rpb - 1) / rpb), 256, smem, (cudaStream_t)stream>>>(
      ray, R, rays_per_image, L_view, W, ldw, col0, imgbias, out, rpb);
  return tp_launch_status();
}

TP_API int tp_tc_nerf_stl_forward(const float* center, const float* ray, const float* depth, int64_t S, int N,
                                  int64_t per_image, const void* packed, const float* biasbuf, const float* raybias,
                                  const float* imgbias, float* rgb, float* density, float* uncert, void* scratch,
                                  int64_t scratch_bytes, int dbg_layer, float* dbg_out, int flags, void* stream) {
  if (!center || !ray || !depth || !packed || !biasbuf || !raybias || !imgbias || !rgb || !density || !uncert || !scratch)
    return TP_ERR_BAD_ARG;
  if (S < 0 || N < 1 || per_image < 1) return TP_ERR_BAD_SHAPE;
  if (((uintptr_t)packed & 15) || ((uintptr_t)scratch & 15) || ((uintptr_t)biasbuf & 15) || ((uintptr_t)raybias & 15) ||
      ((uintptr_t)imgbias & 15) || ((uintptr_t)rgb & 7) || ((uintptr_t)density & 7))
    return TP_ERR_ALIGN;
  if (!tp_device_is_sm100()) return TP_ERR_ARCH;
  if (S == 0) return TP_OK;
  const long long n_super = (S + 255) / 256;
  int grid = tp_num_sms();
  if (n_super < grid) grid = (int)n_super;
  if (scratch_bytes < (int64_t)grid * 2 * tc::kABytes) return TP_ERR_WORKSPACE;
  tc::Params p;
  p.center = center; p.ray = ray; p.depth = depth; p.S = S; p.N = N; p.per_image = per_image;
  p.packed = reinterpret_cast<const uint8_t*>(packed); p.biasbuf = biasbuf; p.raybias = raybias; p.imgbias = imgbias;
  p.rgb = rgb; p.density = density; p.uncert = uncert; p.scratch = reinterpret_cast<uint8_t*>(scratch);
  p.dbg_layer = dbg_layer; p.dbg_out = dbg_out; p.swap_lbo_sbo = flags & 1;
  cudaError_t e = cudaFuncSetAttribute(tc::nerf_stl_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  tc::nerf_stl_forward_kernel<<<grid, tc::kThreads, tc::kSmemBytes, (cudaStream_t)stream>>>(p);
  return tp_launch_status();
}

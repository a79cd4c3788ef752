// K2 (fp32-parity mode on the tensor cores): fused per-sample NeRF forward with SPLIT-bf16 operands, sm_100a.
//
// The <= 1e-4 contract of NeRF.forward (layers/nerf_static_transient_light.py:76-145, layers/nerf.py:61-99) needs ~16
// mantissa bits per operand; bf16 has 8.  Every operand therefore exists twice -- x = hi + lo with hi = bf16(x),
// lo = bf16(x - hi) -- and every K = 16 step is three tcgen05.mma passes into one fp32 TMEM accumulator:
//
//        D += A_hi W_hi^T  +  A_hi W_lo^T  +  A_lo W_hi^T            (the dropped A_lo W_lo^T term is 2^-18 relative)
//
// so the fp32 parity mode runs at a third of the bf16 tensor rate instead of at SIMT FFMA rate (mlp_simt.cu).
//
// Work decomposition (one persistent 320-thread CTA per SM, one 128-sample tile at a time):
//   * activations live in SMEM as two tile images (A_hi, A_lo: [32 k8][128 rows][8] bf16, K-major core-matrix layout), the
//     fp32 positional encoding (reference bits: sin(fl(x * fl(2^k pi)))) as E_hi, E_lo; nothing but the parked trunk feature
//     and the per-sample outputs ever leaves the SM;
//   * weights stream through a 4-slot ring of 16 KB slots, one slot per K = 16 step: [W_hi 8 KB | W_lo 8 KB], each
//     [2 k8][256 rows][8]; the N = 16 output stages stream one 8 KB slot [32 k8][16 rows][8] whose rows 0..7 hold W_hi and
//     rows 8..15 W_lo (two passes, A_hi and A_lo; the epilogue adds accumulator columns c and c + 8);
//   * BOTH 256-column TMEM accumulators are used by ONE tile, alternating by stage, and the drain of stage L publishes the next
//     A operand in eight 32-column groups (one mbarrier each): the MMAs of stage L+1 start as soon as the first group is
//     written and overlap the rest of the drain -- the tensor pipe only idles for the first slab of every drain;
//   * biases are added in fp32 by the drain (static vectors, the per-ray row of tp_tc_ray_bias, the per-image row of
//     tp_tc_image_bias): no bf16 rounding of any bias;
//   * the stage list is DATA (tp_tc32_forward's `stages`): trunk depth, skip positions and head depths come from the caller's
//     architecture, not from a table compiled into the kernel.
//
// SMEM: A_hi 64K | A_lo 64K | E_hi 16K | E_lo 16K | ring 4 x 16K | barriers = 229 632 B.  TMEM: 512 columns.
#include <cuda_fp16.h>
#include "tc_common.cuh"
#include "../../include/texpose_b200.h"

namespace tcs {
using namespace tc;

constexpr int kThreads = 320;                  // warps 0-7 drain / encode, warp 8 weight producer, warp 9 MMA issuer
constexpr int kProducerWarp = 8, kMmaWarp = 9;
constexpr int kRing = 4;
constexpr uint32_t kSlotBytes = 16384, kHalfSlot = 8192;
constexpr uint32_t kEBytes = 16384;            // 128 x 64 bf16
constexpr uint32_t kOffAhi = 0, kOffAlo = kABytes, kOffEhi = 2 * kABytes, kOffElo = kOffEhi + kEBytes;
constexpr uint32_t kOffRing = kOffElo + kEBytes;
constexpr uint32_t kOffBar = kOffRing + kRing * kSlotBytes;
constexpr uint32_t kSmemBytes = kOffBar + 256;
static_assert(kSmemBytes <= 232448, "227 KB of shared memory per CTA");
constexpr int kMaxStages = 24;
constexpr int kEncL = 10;                      // frequencies of the point encoding (the reference's yaml files: L_3D = 10)

enum Kind : int { KIND_HIDDEN = 0, KIND_DENSITY = 1, KIND_RGB_OUT = 2, KIND_TRANS_OUT = 3 };
enum BiasKind : int { BIAS_STATIC = 0, BIAS_RAY = 1, BIAS_IMAGE = 2 };
enum Flags : int {
  F_WAIT_READY = 1,    // the stage reads the A operand the previous hidden stage's drain writes
  F_RELOAD = 2,        // the stage reads the parked trunk feature (bulk-loaded back into A)
  F_PARK = 4,          // the stage's output is the trunk feature: the drain also stores it to the CTA's scratch
  F_E_LAST = 8         // last stage of the tile that reads E: the next tile's encoding may be written once its MMAs retired
};
struct Stage {
  int a_steps, e_steps, kind, bias_kind, bias_off, flags, save_slot;      // save_slot: -1, or the tile-image slot the stage's output is kept in
};
struct Params {
  const float* center;       // [rays,3]
  const float* ray;          // [rays,3]
  const float* depth;        // [S]
  long long S;
  int N;
  long long per_image;
  const uint8_t* image;      // weight slots in consumption order (n_slots x 16 KB)
  const float* bias;         // static biases (fp32), addressed by Stage::bias_off
  const float* raybias;      // [rays,256]
  const float* imgbias;      // [images,256]
  float* rgb;                // [S,3,2]
  float* density;            // [S,2]
  float* uncert;             // [S]
  uint8_t* scratch;          // gridDim.x x 128 KB (parked feature, hi | lo), then one 32-bit status word
  unsigned int* status;      // bit 0 is set when a hidden activation left the fp16 range of the split mode (the launch's results are
                             // then not the <= 1e-4 ones): read back by the caller off the hot path (mlp_tc32.check_range)
  uint8_t* save;             // single-pass mode, training: [tiles][n_save][64 KB] bf16 tile images of the stages with a save slot
  int n_save;
  uint32_t* save_bits;       // behind the images: [tiles][n_save][8 planes][128 rows] ReLU bitmasks (plane = 32-column slab, bit = column
                             // is positive): what the backward chain masks with instead of re-reading the 64 KB tiles
  int enc_slot;              // -1, or the save slot that receives the encoding tile [x, enc(x), 1] (zero beyond column 63): the B
                             // operand of the weight-gradient GEMMs of the layers that read the encoding (column 63 = 1: their bias sums)
  int n_stages;
  Stage st[kMaxStages];
};

// hi / lo fp16 words of two fp32 values (x0 in the low half): hi = fp16(x), lo = fp16(x - hi) -> x to ~22 mantissa bits
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16(x0, x1);
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack_f16(x0 - h.x, x1 - h.y);
}
// instruction descriptor, kind::f16 with FP16 operands (format fields 0), fp32 accumulate, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma3(uint32_t d, uint32_t ah, uint32_t al, uint32_t a_hi, uint32_t bh, uint32_t bl, uint32_t b_hi,
                                      uint32_t idesc, uint32_t acc) {
  umma_bf16_lohi(d, ah, a_hi, bh, b_hi, idesc, acc);
  umma_bf16_lohi(d, ah, a_hi, bl, b_hi, idesc, 1u);
  umma_bf16_lohi(d, al, a_hi, bh, b_hi, idesc, 1u);
}

// kSingle = false: the fp32-parity mode (fp16 hi + lo, three passes).  kSingle = true: ONE pass with bf16 operands -- the general-
// architecture bf16 forward (any stage list; <= 1e-2 like mlp_tc.cu, which stays the fast path for the yaml's architecture), whose
// drain can also keep every hidden activation as a bf16 tile image for the tensor-core backward (training of layers/nerf.py).
template <bool kSingle>
__global__ void __launch_bounds__(kThreads, 1) nerf_forward_split_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t bar0 = sbase + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (kRing + s); };
  auto bar_acc = [&](int i) { return bar0 + 8 * (2 * kRing + i); };
  auto bar_ready = [&](int g) { return bar0 + 8 * (2 * kRing + 2 + g); };          // 8: A columns [32g, 32g+32) written
  auto bar_reload = [&](int g) { return bar0 + 8 * (2 * kRing + 10 + g); };        // 8: parked feature columns back in A
  const uint32_t bar_enc = bar0 + 8 * (2 * kRing + 18), bar_out = bar_enc + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * (2 * kRing + 20));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRing; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_acc(0), 1);
    mbar_init(bar_acc(1), 1);
    for (int g = 0; g < 8; ++g) {
      mbar_init(bar_ready(g), 4);        // the four lane-quarter warps of the column half
      mbar_init(bar_reload(g), 1);
    }
    mbar_init(bar_enc, 4);
    mbar_init(bar_out, 4);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long n_tiles = (p.S + 127) / 128;
  const int n_stages = p.n_stages;

  if (warp == kProducerWarp) {
    // ================================================================ weight producer
    uint32_t slot = 0, phase = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      long long c = 0;
      for (int L = 0; L < n_stages; ++L) {
        const Stage sg = p.st[L];
        const bool small = sg.kind != KIND_HIDDEN;
        const int n = small ? 1 : sg.e_steps + sg.a_steps;
        const uint32_t bytes = (small || kSingle) ? kHalfSlot : kSlotBytes;      // single pass: only the hi half of a slot exists
        for (int j = 0; j < n; ++j, ++c) {
          mbar_wait(bar_empty(slot), phase ^ 1);
          if (elect_one_sync()) {
            mbar_expect_tx(bar_full(slot), bytes);
            bulk_g2s_hint(sbase + kOffRing + slot * kSlotBytes, p.image + (size_t)c * kSlotBytes, bytes, bar_full(slot),
                          l2_policy_evict_last());
          }
          __syncwarp();
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================ MMA issuer
    uint32_t slot = 0, phase = 0, ready_ph = 0, reload_ph = 0, enc_ph = 0, out_ph = 0;
    const uint32_t idesc256 = kSingle ? umma_idesc(128, 256) : umma_idesc_f16(128, 256);
    const uint32_t idesc16 = kSingle ? umma_idesc(128, 16) : umma_idesc_f16(128, 16);
    constexpr uint32_t kHi = (128u >> 4) | (1u << 14);      // SBO = 128 B, descriptor version 1
    bool first_tile = true;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(bar_enc, enc_ph);            // the tile's encoding is in E
      enc_ph ^= 1;
      if (!first_tile) {                     // the previous tile's last accumulator has been read
        mbar_wait(bar_out, out_ph);
        out_ph ^= 1;
      }
      first_tile = false;
      for (int L = 0; L < n_stages; ++L) {
        const Stage sg = p.st[L];
        const uint32_t d_tmem = tmem_base + (uint32_t)(L & 1) * 256;
        const bool wait_ready = (sg.flags & F_WAIT_READY) != 0, reload = (sg.flags & F_RELOAD) != 0;
        if (sg.kind == KIND_HIDDEN) {
          const int n = sg.e_steps + sg.a_steps;
          for (int step = 0; step < n; ++step) {
            const bool from_e = step < sg.e_steps;
            const int ks = from_e ? step : step - sg.e_steps;
            if (!from_e && (ks & 1) == 0) {
              if (reload) mbar_wait(bar_reload(ks >> 1), reload_ph);
              else if (wait_ready) mbar_wait(bar_ready(ks >> 1), ready_ph);
            }
            mbar_wait(bar_full(slot), phase);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t wsm = sbase + kOffRing + slot * kSlotBytes;
              const uint32_t bh = (wsm >> 4) | ((4096u >> 4) << 16), bl = ((wsm + kHalfSlot) >> 4) | ((4096u >> 4) << 16);
              const uint32_t a0 = (from_e ? kOffEhi : kOffAhi) + (uint32_t)ks * 4096u;
              const uint32_t a1 = (from_e ? kOffElo : kOffAlo) + (uint32_t)ks * 4096u;
              const uint32_t ah = ((sbase + a0) >> 4) | ((2048u >> 4) << 16), al = ((sbase + a1) >> 4) | ((2048u >> 4) << 16);
              if (kSingle) umma_bf16_lohi(d_tmem, ah, kHi, bh, kHi, idesc256, step > 0 ? 1u : 0u);
              else umma3(d_tmem, ah, al, kHi, bh, bl, kHi, idesc256, step > 0 ? 1u : 0u);
              umma_commit(bar_empty(slot));
              if (step == n - 1) umma_commit(bar_acc(L & 1));
            }
            __syncwarp();
            if (++slot == kRing) { slot = 0; phase ^= 1; }
          }
        } else {
          // N = 16 output stage: one 8 KB slot [32 k8][16 rows][8] spans K = 256; pass 0 reads A_hi, pass 1 A_lo
          mbar_wait(bar_full(slot), phase);
          const uint32_t wsm = sbase + kOffRing + slot * kSlotBytes;
          for (int ks = 0; ks < 16; ++ks) {
            if ((ks & 1) == 0) {
              if (reload) mbar_wait(bar_reload(ks >> 1), reload_ph);
              else if (wait_ready) mbar_wait(bar_ready(ks >> 1), ready_ph);
            }
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t ah = ((sbase + kOffAhi + (uint32_t)ks * 4096u) >> 4) | ((2048u >> 4) << 16);
              const uint32_t b = ((wsm + (uint32_t)ks * 512u) >> 4) | ((256u >> 4) << 16);
              umma_bf16_lohi(d_tmem, ah, kHi, b, kHi, idesc16, ks > 0 ? 1u : 0u);
            }
            __syncwarp();
          }
          if (elect_one_sync()) {
            if (!kSingle) {
#pragma unroll
              for (int ks = 0; ks < 16; ++ks) {
                const uint32_t al = ((sbase + kOffAlo + (uint32_t)ks * 4096u) >> 4) | ((2048u >> 4) << 16);
                const uint32_t b = ((wsm + (uint32_t)ks * 512u) >> 4) | ((256u >> 4) << 16);
                umma_bf16_lohi(d_tmem, al, kHi, b, kHi, idesc16, 1u);
              }
            }
            umma_commit(bar_empty(slot));
            umma_commit(bar_acc(L & 1));
          }
          __syncwarp();
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        }
        if (reload) reload_ph ^= 1;
        else if (wait_ready) ready_ph ^= 1;
      }
    }
  } else {
    // ================================================================ drain / encode warps
    // warp -> (TMEM lane quarter q, column half): each thread converts 128 accumulator columns of its row
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* park = p.scratch + (size_t)blockIdx.x * 2 * kABytes;
    uint32_t acc_ph = 0;      // bit i: phase of bar_acc(i)
    float act_max = 0.f;      // split mode: largest hidden activation this thread converted (fp16 hi saturates at 65 504)

    // fp32 encoding of one sample (layers/nerf_static_transient_light.py:217-234: the reference multiplies by fl(2^k pi) and
    // takes sin / cos of the ROUNDED product -- its own encoding differs from the exact one by up to 6e-5 at k = 9, so the
    // parity mode must round where it rounds), split into E_hi / E_lo.  Done by the column-half-1 warps (row = sample).
    auto encode_tile = [&](long long tile) {
      const long long s_raw = tile * 128 + row, s = s_raw < p.S ? s_raw : p.S - 1, r = s / p.N;
      const float d = p.depth[s];
      float v[64];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float x = __fadd_rn(p.center[r * 3 + j], __fmul_rn(p.ray[r * 3 + j], d));      // camera.py:317-322
        v[j] = x;
#pragma unroll
        for (int k = 0; k < kEncL; ++k) {
          float sn, cs;
          sincosf(__fmul_rn(x, ldexpf(3.14159265358979323846f, k)), &sn, &cs);
          v[3 + j * 2 * kEncL + k] = sn;
          v[3 + j * 2 * kEncL + kEncL + k] = cs;
        }
      }
      uint8_t* const g_enc = (kSingle && p.save && p.enc_slot >= 0) ? p.save + ((size_t)tile * p.n_save + p.enc_slot) * kABytes + row * 16 : nullptr;
      v[63] = g_enc ? 1.f : 0.f;      // (no layer has a weight column 63: the packed images hold zeros there)
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (kSingle) h[e] = pack_bf16(v[k8 * 8 + 2 * e], v[k8 * 8 + 2 * e + 1]);
          else split2(v[k8 * 8 + 2 * e], v[k8 * 8 + 2 * e + 1], h[e], l[e]);
        }
        st_shared_v4(sbase + kOffEhi + k8 * 2048 + row * 16, h[0], h[1], h[2], h[3]);
        if (!kSingle) st_shared_v4(sbase + kOffElo + k8 * 2048 + row * 16, l[0], l[1], l[2], l[3]);
        if (g_enc) st_global_cs_v4(g_enc + k8 * 2048, h[0], h[1], h[2], h[3]);
      }
      if (g_enc) {
#pragma unroll 4
        for (int k8 = 8; k8 < 32; ++k8) st_global_cs_v4(g_enc + k8 * 2048, 0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_enc);
    };
    if (half == 1 && (long long)blockIdx.x < n_tiles) encode_tile(blockIdx.x);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long s_raw = tile * 128 + row;
      const bool live = s_raw < p.S;
      const long long s = live ? s_raw : p.S - 1;
      float sigma_s = 0.f, rgb_s[3] = {0.f, 0.f, 0.f}, rgb_t[3] = {0.f, 0.f, 0.f}, sigma_t = 0.f, unc = 0.f;
      for (int L = 0; L < n_stages; ++L) {
        const Stage sg = p.st[L];
        mbar_wait(bar_acc(L & 1), (acc_ph >> (L & 1)) & 1u);
        acc_ph ^= 1u << (L & 1);
        tc_fence_after();
        if (L + 1 < n_stages && (p.st[L + 1].flags & F_RELOAD) && threadIdx.x == 0) {
          // every MMA that read A has retired: bring the parked feature back, one 32-column group (hi + lo) per barrier
          fence_proxy_async_all();
#pragma unroll 1
          for (int g = 0; g < 8; ++g) {
            mbar_expect_tx(bar_reload(g), kSingle ? kHalfSlot : 2 * kHalfSlot);
            bulk_g2s(sbase + kOffAhi + g * kHalfSlot, park + g * kHalfSlot, kHalfSlot, bar_reload(g));
            if (!kSingle) bulk_g2s(sbase + kOffAlo + g * kHalfSlot, park + kABytes + g * kHalfSlot, kHalfSlot, bar_reload(g));
          }
        }
        if (sg.kind == KIND_HIDDEN) {
          const float* brow = (sg.bias_kind == BIAS_STATIC ? p.bias + sg.bias_off
                               : sg.bias_kind == BIAS_RAY  ? p.raybias + (s / p.N) * 256
                                                           : p.imgbias + (s / p.per_image) * 256) + half * 128;
          const uint32_t tmem_d = tmem_row + (uint32_t)(L & 1) * 256 + half * 128;
          const bool do_park = (sg.flags & F_PARK) != 0;
          uint8_t* const g_save = (kSingle && p.save && sg.save_slot >= 0) ? p.save + ((size_t)tile * p.n_save + sg.save_slot) * kABytes : nullptr;
          uint32_t* const g_bits = g_save ? p.save_bits + ((size_t)tile * p.n_save + sg.save_slot) * 1024 + half * 4 * 128 + row : nullptr;
          // one 32-column slab: + bias, ReLU, hi / lo split, 4 core-matrix rows of A_hi and A_lo (and of the parked copy)
          // the slab's 32 bias values are fetched one slab ahead (beside the TMEM load): loads issued inside the conversion loop
          // wait out their full latency behind the stores of the previous group (16 exposed load-use stalls per stage, ncu source page)
          auto load_bias = [&](float4 (&bq)[8], int j) {
#pragma unroll
            for (int i = 0; i < 8; ++i) bq[i] = __ldg(reinterpret_cast<const float4*>(brow + j * 32 + i * 4));
          };
          auto slab = [&](const uint32_t (&v)[32], const float4 (&bq)[8], int j) {
            uint32_t positive = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b0 = bq[2 * i], b1 = bq[2 * i + 1];
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x0 = __uint_as_float(v[i * 8 + 2 * e]) + bb[2 * e], x1 = __uint_as_float(v[i * 8 + 2 * e + 1]) + bb[2 * e + 1];
                if (kSingle) {
                  h[e] = pack_relu_bf16(x0, x1);
                  if (g_save) {      // sign bits, first column shifted in first (one funnel shift per element; reversed below)
                    positive = __funnelshift_l(__float_as_uint(x0), positive, 1);
                    positive = __funnelshift_l(__float_as_uint(x1), positive, 1);
                  }
                } else {
                  split2(fmaxf(x0, 0.f), fmaxf(x1, 0.f), h[e], l[e]);
                  act_max = fmaxf(act_max, fmaxf(x0, x1));
                }
              }
              const uint32_t off = (uint32_t)(half * 16 + j * 4 + i) * 2048 + row * 16;
              st_shared_v4(sbase + kOffAhi + off, h[0], h[1], h[2], h[3]);
              if (!kSingle) st_shared_v4(sbase + kOffAlo + off, l[0], l[1], l[2], l[3]);
              if (do_park) {
                st_global_v4(park + off, h[0], h[1], h[2], h[3]);
                if (!kSingle) st_global_v4(park + kABytes + off, l[0], l[1], l[2], l[3]);
              }
              if (g_save) st_global_cs_v4(g_save + off, h[0], h[1], h[2], h[3]);
            }
            if (g_bits) g_bits[j * 128] = __brev(~positive);      // bit c = column c of the slab is positive (sign bit clear)
            if (do_park) __threadfence();      // the parked copy is read back through the async proxy (bulk load) four stages later
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready(half * 4 + j));
          };
          uint32_t va[32], vb[32];
          float4 ba[8], bb2[8];
          TP_TMEM_LD32(tmem_d, va);
          load_bias(ba, 0);
          TP_TMEM_WAIT32(va);
          TP_TMEM_LD32(tmem_d + 32, vb);
          load_bias(bb2, 1);
          slab(va, ba, 0);
          TP_TMEM_WAIT32(vb);
          TP_TMEM_LD32(tmem_d + 64, va);
          load_bias(ba, 2);
          slab(vb, bb2, 1);
          TP_TMEM_WAIT32(va);
          TP_TMEM_LD32(tmem_d + 96, vb);
          load_bias(bb2, 3);
          slab(va, ba, 2);
          TP_TMEM_WAIT32(vb);
          slab(vb, bb2, 3);
          if ((sg.flags & F_E_LAST) && half == 1 && tile + gridDim.x < n_tiles) encode_tile(tile + gridDim.x);
        } else if (half == 0) {
          uint32_t v[16];
          TP_TMEM_LD16(tmem_row + (uint32_t)(L & 1) * 256, v);
          TP_TMEM_WAIT16(v);
          if (L == n_stages - 1) {           // the next tile's first stage may overwrite this accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_out);
          }
          const float* sb = p.bias + sg.bias_off;
          auto out = [&](int c) { return __uint_as_float(v[c]) + __uint_as_float(v[c + 8]) + sb[c]; };
          if (sg.kind == KIND_DENSITY) {
            sigma_s = tp_softplus(out(0));
          } else if (sg.kind == KIND_RGB_OUT) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb_s[c] = tp_sigmoid(out(c));
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb_t[c] = tp_sigmoid(out(c));
            sigma_t = tp_softplus(out(3));
            unc = tp_softplus(out(4));
          }
        }
      }
      if (!kSingle && !(act_max <= 6.0e4f)) atomicOr(p.status, 1u);      // (also catches NaN)
      if (half == 0 && live) {
#pragma unroll
        for (int c = 0; c < 3; ++c) *reinterpret_cast<float2*>(p.rgb + s * 6 + c * 2) = make_float2(rgb_s[c], rgb_t[c]);
        *reinterpret_cast<float2*>(p.density + s * 2) = make_float2(sigma_s, sigma_t);
        p.uncert[s] = unc;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// slot_desc row: [W ptr, ld, row0, rows_valid, col0, cols_valid, kind (256 | 16), unused]
//   kind 256: slot = [hi | lo], each [2 k8][256 rows][8]: element (n, kl) = W[row0 + n][col0 + kl], kl < 16
//   kind 16:  first 8 KB = [32 k8][16 rows][8]: rows 0..7 hi, rows 8..15 lo of W[row0 + (n & 7)][col0 + kl], kl < 256
// bf16 != 0: the single-pass image -- bf16(W) in the hi half (the lo half stays zero and is never streamed); output layers keep
// their hi / lo rows, as bf16
__global__ void pack_split_kernel(const long long* __restrict__ desc, uint16_t* __restrict__ out, int bf16) {
  const long long* d = desc + (long long)blockIdx.x * 8;
  const float* W = reinterpret_cast<const float*>(d[0]);
  const long long ld = d[1], row0 = d[2], rows_valid = d[3], col0 = d[4], cols_valid = d[5], kind = d[6];
  uint16_t* o = out + (long long)blockIdx.x * (kSlotBytes / 2);
  for (int e = threadIdx.x; e < (int)(kSlotBytes / 2); e += blockDim.x) {
    int n, kl;
    bool lo_part, in_layout = true;
    if (kind == 256) {
      lo_part = e >= 4096;
      const int f = e & 4095;
      kl = (f >> 11) * 8 + (f & 7);
      n = (f >> 3) & 255;
    } else {
      in_layout = e < 4096;
      kl = (e >> 7) * 8 + (e & 7);
      n = (e >> 3) & 15;
      lo_part = n >= 8;
      n &= 7;
    }
    float v = 0.f;
    if (in_layout && W && n < rows_valid && kl < cols_valid) v = W[(row0 + n) * ld + col0 + kl];
    if (bf16) {
      if (lo_part) v = kind == 256 ? 0.f : v - __bfloat162float(__float2bfloat16_rn(v));
      o[e] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    } else {
      if (lo_part) v -= __half2float(__float2half_rn(v));
      o[e] = __half_as_ushort(__float2half_rn(v));
    }
  }
}

}  // namespace tcs

TP_API int64_t tp_tc32_slot_bytes(void) { return tcs::kSlotBytes; }
/* bytes of the save buffer of a single-pass launch: [tiles][n_save][64 KB] tile images, then [tiles][n_save][4 KB] ReLU bitmasks */
TP_API int64_t tp_tc32_save_bytes(int64_t S, int n_save) { return ((S + 127) / 128) * (int64_t)n_save * (tc::kABytes + 4096); }
TP_API int64_t tp_tc32_scratch_bytes(void) { return (int64_t)tp_num_sms() * 2 * tc::kABytes + 16; }
TP_API int64_t tp_tc32_status_offset(void) { return (int64_t)tp_num_sms() * 2 * tc::kABytes; }
TP_API int tp_tc32_max_stages(void) { return tcs::kMaxStages; }

TP_API int tp_tc32_pack_weights(const int64_t* slot_desc, int n_slots, int precision, void* image, void* stream) {
  if (!slot_desc || !image) return TP_ERR_BAD_ARG;
  if (n_slots < 1) return TP_ERR_BAD_SHAPE;
  if (precision != 0 && precision != 1) return TP_ERR_BAD_ARG;
  tcs::pack_split_kernel<<<n_slots, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(slot_desc),
                                                                   reinterpret_cast<uint16_t*>(image), precision);
  return tp_launch_status();
}

TP_API int tp_tc32_forward(const float* center, const float* ray, const float* depth, int64_t S, int N, int64_t per_image,
                           const void* image, int n_slots, const int32_t* stages, int n_stages, const float* bias,
                           const float* raybias, const float* imgbias, float* rgb, float* density, float* uncert,
                           void* scratch, int64_t scratch_bytes, int precision, void* save, int n_save, int enc_slot, void* stream) {
  if (!center || !ray || !depth || !image || !stages || !bias || !rgb || !density || !uncert || !scratch) return TP_ERR_BAD_ARG;
  if (S < 0 || N < 1 || per_image < 1 || n_stages < 1 || n_stages > tcs::kMaxStages) return TP_ERR_BAD_SHAPE;
  if ((precision != 0 && precision != 1) || (save && (precision != 1 || n_save < 1)) || ((uintptr_t)save & 15)) return TP_ERR_BAD_ARG;
  if (enc_slot < -1 || (enc_slot >= 0 && (!save || enc_slot >= n_save))) return TP_ERR_BAD_ARG;
  if (((uintptr_t)image & 15) || ((uintptr_t)scratch & 15) || ((uintptr_t)bias & 15) || ((uintptr_t)raybias & 15) ||
      ((uintptr_t)imgbias & 15) || ((uintptr_t)rgb & 7) || ((uintptr_t)density & 7))
    return TP_ERR_ALIGN;
  if (!tp_device_is_sm100()) return TP_ERR_ARCH;
  tcs::Params p = {};
  // validate the stage list: it is the caller's architecture, and a malformed list would dead-lock the barrier protocol
  int slots = 0, pending_ready = 0, n_e_last = 0;
  bool parked = false;
  for (int L = 0; L < n_stages; ++L) {
    tcs::Stage& sg = p.st[L];
    const int32_t* r = stages + L * 7;
    sg.a_steps = r[0]; sg.e_steps = r[1]; sg.kind = r[2]; sg.bias_kind = r[3]; sg.bias_off = r[4]; sg.flags = r[5]; sg.save_slot = r[6];
    if (sg.save_slot < -1 || (save && sg.save_slot >= n_save) || (sg.save_slot >= 0 && sg.kind != tcs::KIND_HIDDEN)) return TP_ERR_BAD_ARG;
    if (sg.save_slot >= 0 && sg.save_slot == enc_slot) return TP_ERR_BAD_ARG;
    if (sg.kind < 0 || sg.kind > 3 || sg.bias_kind < 0 || sg.bias_kind > 2 || sg.bias_off < 0 || (sg.bias_off & 3)) return TP_ERR_BAD_ARG;
    if (sg.e_steps < 0 || sg.e_steps > 4 || (sg.a_steps != 0 && sg.a_steps != 16)) return TP_ERR_BAD_SHAPE;
    const bool hidden = sg.kind == tcs::KIND_HIDDEN, wait = sg.flags & tcs::F_WAIT_READY, reload = sg.flags & tcs::F_RELOAD;
    if (!hidden && (sg.a_steps != 16 || sg.e_steps != 0 || sg.bias_kind != tcs::BIAS_STATIC)) return TP_ERR_BAD_SHAPE;
    if (hidden && sg.a_steps + sg.e_steps == 0) return TP_ERR_BAD_SHAPE;
    if ((wait || reload) && sg.a_steps != 16) return TP_ERR_BAD_ARG;
    if (wait && reload) return TP_ERR_BAD_ARG;
    if (wait && pending_ready != 1) return TP_ERR_BAD_ARG;              // exactly one drain's arrivals are outstanding
    if (wait) pending_ready = 0;
    if (reload && (!parked || L == 0 || p.st[L - 1].kind == tcs::KIND_HIDDEN || pending_ready)) return TP_ERR_BAD_ARG;
    if (sg.a_steps && !wait && !reload && (L == 0 || p.st[L - 1].kind == tcs::KIND_HIDDEN)) return TP_ERR_BAD_ARG;   // A would be stale
    if (sg.bias_kind == tcs::BIAS_RAY && !raybias) return TP_ERR_BAD_ARG;
    if (sg.bias_kind == tcs::BIAS_IMAGE && !imgbias) return TP_ERR_BAD_ARG;
    if (hidden) {
      if (pending_ready) return TP_ERR_BAD_ARG;                          // the previous drain's arrivals were never consumed
      pending_ready = 1;
      if (sg.flags & tcs::F_PARK) parked = true;
    }
    if (sg.flags & tcs::F_E_LAST) ++n_e_last;
    slots += hidden ? sg.a_steps + sg.e_steps : 1;
  }
  for (int L = 0, seen = 0; L < n_stages; ++L) {      // E_LAST marks the last stage that reads E, and a hidden one
    if (p.st[L].flags & tcs::F_E_LAST) seen = 1;
    else if (seen && p.st[L].e_steps) return TP_ERR_BAD_ARG;
    if ((p.st[L].flags & tcs::F_E_LAST) && p.st[L].kind != tcs::KIND_HIDDEN) return TP_ERR_BAD_ARG;
  }
  if (n_e_last != 1 || pending_ready || p.st[0].e_steps == 0 || p.st[n_stages - 1].kind == tcs::KIND_HIDDEN) return TP_ERR_BAD_ARG;
  if (slots != n_slots) return TP_ERR_BAD_SHAPE;
  if (S == 0) return TP_OK;
  const long long n_tiles = (S + 127) / 128;
  int grid = tp_num_sms();
  if (n_tiles < grid) grid = (int)n_tiles;
  if (scratch_bytes < tp_tc32_scratch_bytes()) return TP_ERR_WORKSPACE;
  p.center = center; p.ray = ray; p.depth = depth; p.S = S; p.N = N; p.per_image = per_image;
  p.image = reinterpret_cast<const uint8_t*>(image); p.bias = bias; p.raybias = raybias; p.imgbias = imgbias;
  p.rgb = rgb; p.density = density; p.uncert = uncert; p.scratch = reinterpret_cast<uint8_t*>(scratch);
  p.status = reinterpret_cast<unsigned int*>(p.scratch + tp_tc32_status_offset());
  p.n_stages = n_stages;
  p.save = reinterpret_cast<uint8_t*>(save); p.n_save = n_save; p.enc_slot = save ? enc_slot : -1;
  p.save_bits = save ? reinterpret_cast<uint32_t*>(p.save + (size_t)((S + 127) / 128) * n_save * tc::kABytes) : nullptr;
  void (*kern)(const tcs::Params) = precision ? tcs::nerf_forward_split_kernel<true> : tcs::nerf_forward_split_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcs::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, tcs::kThreads, tcs::kSmemBytes, (cudaStream_t)stream>>>(p);
  return tp_launch_status();
}

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
#include "mlp_tc_shared.cuh"
#include "mlp_tc_epilogue.cuh"
#include "../../include/texpose_b200.h"

namespace tc {

// kHalves epilogue warps per (tile, TMEM lane quarter): kHalves = 1 -> 8 epilogue warps (256 columns per thread),
// kHalves = 2 -> 16 epilogue warps (128 columns per thread).  Then one TMA producer warp and one MMA issuer warp.
template <int kHalves> constexpr int num_threads() { return (8 * kHalves + 2) * 32; }
constexpr int kStages = 4;
constexpr uint32_t kOffA = 0, kOffE = 2 * kABytes, kOffRing = kOffE + 2 * kEBytes;
constexpr uint32_t kOffBar = kOffRing + kStages * kChunkBytes;
constexpr uint32_t kSmemBytes = kOffBar + 128;
// kNL: number of stages run (kNumLayers = all; kStaticLayers = static-only rendering).  A template parameter, not a field of
// Params: a run-time stage count costs 0.3 ms per C2 frame (same-box A/B).
// kMode: 0 = general (debug taps, timing experiments, any tile skew); 1 = the plain inference launch (no activation save);
// 2 = the plain training launch (activation save + ReLU bitmasks).  Modes 1 and 2 have every debug branch compiled out and the
// default tile skew as a constant: the same code with those decisions left to run time is 6 % slower (same-box A/B).
template <int kHalves, int kNL = kNumLayers, int kMode = 0>
__global__ void __launch_bounds__(num_threads<kHalves>(), 1) nerf_stl_forward_kernel(const Params p) {
  const int p_skew = kMode ? 1 : p.skew;
  const int p_dbg_drain = kMode ? 0 : p.dbg_drain;
  const int p_dbg_save = kMode ? 0 : p.dbg_save;
  uint8_t* const p_save = kMode == 1 ? nullptr : p.save;
  constexpr int kEpiWarps = 8 * kHalves, kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1;
  constexpr int kTileThreads = 128 * kHalves;     // epilogue threads working on one tile
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
      mbar_init(bar_ready(t), kTileThreads / 32);     // one arrive per epilogue warp of the tile
      mbar_init(bar_reload(t), 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {   // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_super = (p.S + 255) / 256;

  if (warp == kProducerWarp) {
    // ================================================================ weight producer (converged warp, one lane issues)
    {
      uint32_t stage = 0, phase = 0;
      for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
        int c = 0;
        for (int L = 0; L < kNL; ++L) {
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
  } else if (warp == kMmaWarp) {
    // ================================================================ MMA issuer (converged warp, one lane issues)
    // Inside a stage tile 0 runs `skew` weight chunks ahead of tile 1 (both still consume every chunk from the same ring
    // slot): T0's accumulator completes 256*(skew+1) cycles before T1's, so T0's epilogue -- and T0's first MMAs of the next
    // stage -- overlap T1's MMAs / epilogue instead of leaving the tensor pipe idle.
    {
      uint32_t chunk_base = 0, ready_ph = 0, reload_ph = 0;   // per-tile phase bits (bit t)
      const uint32_t idesc256 = umma_idesc(128, 256), idesc16 = umma_idesc(128, 16);
      constexpr uint32_t kHi = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
      for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
        for (int L = 0; L < kNL; ++L) {
          const Layer ly = kLayers[L];
          const int nch = ly.small ? 1 : ly.a_chunks + ly.e_chunks + ly.bias_chunk;
          const int skew = p_skew < nch ? p_skew : nch;
          for (int step = 0; step < nch + skew; ++step) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int c = t == 0 ? step : step - skew;
              if (c < 0 || c >= nch) continue;
              const uint32_t abs_chunk = chunk_base + c;
              const uint32_t stage = abs_chunk % kStages, phase = (abs_chunk / kStages) & 1u;
              if (t == 0) mbar_wait(bar_full(stage), phase);          // first touch of this ring slot
              if (c == 0) {
                mbar_wait(bar_ready(t), (ready_ph >> t) & 1u);
                ready_ph ^= 1u << t;
                if (ly.reload) {
                  mbar_wait(bar_reload(t), (reload_ph >> t) & 1u);
                  reload_ph ^= 1u << t;
                }
              }
              tc_fence_after();
              const uint32_t wsm = sbase + kOffRing + stage * kChunkBytes;
              const uint32_t d_tmem = tmem_base + t * 256;
              if (elect_one_sync()) {
                // descriptor words: lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version(1)<<14
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
                if (c == nch - 1) umma_commit(bar_acc(t));       // accumulator of tile t complete
                if (t == 1) umma_commit(bar_empty(stage));       // ring slot reusable once both tiles' MMAs retire
              }
              __syncwarp();
            }
          }
          chunk_base += nch;
        }
      }
    }
  } else {
    // ================================================================ encode + epilogue warps
    // warp -> (TMEM lane quarter q = warp % 4, tile t, column half); half 0 also encodes and owns the per-sample outputs
    const int q = warp & 3, t = (warp >> 2) & 1, half = warp >> 3, row = q * 32 + lane;
    constexpr int kCols = 256 / kHalves;            // accumulator columns converted per thread
    const uint32_t a_smem = sbase + kOffA + t * kABytes, e_smem = sbase + kOffE + t * kEBytes;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + t * 256;
    const uint32_t tmem_d = tmem_row + half * kCols;
    uint8_t* my_scratch = p.scratch + ((size_t)blockIdx.x * 2 + t) * kABytes;
    uint32_t acc_ph = 0;
    // all 32 rows of a warp share the ray (and image) when N is a multiple of 32; tail rows are clamped to the last
    // sample, which then belongs to the same ray as the warp's live rows
    const bool warp_bias = kMode ? true : (p.N % 32 == 0) && !(p_dbg_drain & 4);      // modes 1 / 2 are launched only when N % 32 == 0
    bool store_pending = false;      // a bulk store of A_t (feature park / activation save) may still be reading it
    for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
      const long long s_raw = (st * 2 + t) * 128 + row;
      const bool live = s_raw < p.S;
      const long long s = live ? s_raw : p.S - 1;
      if (half == 0) encode_sample(p, s, e_smem, row);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();                    // every lane's st.shared + proxy fence precede the warp's single arrive
      if (lane == 0) mbar_arrive(bar_ready(t));

      float sigma_s = 0.f, rgb_s[3] = {0.f, 0.f, 0.f};
      for (int L = 0; L < kNL; ++L) {
        const Layer ly = kLayers[L];
        // table biases (per ray / per image): when every row of this warp shares the bias row, each lane fetches its
        // float4 slice(s) now -- the L2 latency hides behind the MMAs of this stage
        float4 wb[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        const bool table_bias = ly.epi == EPI_HIDDEN && ly.bias_kind != BIAS_MMA;
        if (table_bias && warp_bias) {
          const float* brow = (ly.bias_kind == BIAS_RAY ? p.raybias + (s / p.N) * 256 : p.imgbias + (s / p.per_image) * 256) +
                              half * kCols;
          wb[0] = __ldg(reinterpret_cast<const float4*>(brow) + lane);
          if (kCols == 256) wb[1] = __ldg(reinterpret_cast<const float4*>(brow + 128) + lane);
        }
        mbar_wait(bar_acc(t), acc_ph);
        acc_ph ^= 1;
        tc_fence_after();
        if (ly.epi == EPI_HIDDEN && store_pending) {   // the previous bulk store must have finished reading A_t
          if (row == 0 && half == 0) bulk_wait_read();
          named_bar_sync(1 + t, kTileThreads);
          store_pending = false;
        }
        if (L == kReloadIssueLayer && kNL > kStaticLayers && row == 0 && half == 0) {
          // every MMA that reads A_t has retired (acc barrier) -> bring the trunk feature back for the transient head
          bulk_wait_all();
          fence_proxy_async_all();
          mbar_expect_tx(bar_reload(t), kABytes);
          bulk_g2s(a_smem, p_save ? p_save + ((size_t)(st * 2 + t) * kSaveSlots) * kABytes : my_scratch, kABytes, bar_reload(t));
        }
        if (ly.epi == EPI_HIDDEN) {
          float* dbg_row = (kMode == 0 && (L == p.dbg_layer) && live && p.dbg_out) ? p.dbg_out + s * 256 + half * kCols : nullptr;
          const uint32_t a_row = a_smem + half * (kCols / 8) * 2048 + row * 16;
          const bool train_stage = p_save && L != kSpillLayer && kSaveSlot[L] >= 0;
          if (train_stage) {
            // training: h1 / h2 of either head also leave a ReLU bitmask [128 rows][256 bits] for the backward chain (word
            // planes [8][128 rows]: plane = 32-column slab, so a warp's store of one slab is 128 contiguous bytes), and the
            // activations go straight from the registers to their saved tile image (the backward's operands): a warp's
            // 16-byte groups cover 512 contiguous bytes and nothing re-reads the A tile
            const int mslot = (p_dbg_save & 2) ? -1 : kMaskBitSlot[L];
            uint32_t* words = mslot < 0 ? nullptr
                                        : reinterpret_cast<uint32_t*>(p.bits + ((size_t)(st * 2 + t) * 4 + mslot) * kMaskBitBytes) +
                                              half * (kCols / 32) * 128 + row;
            uint8_t* g_row = (p_dbg_save & 1) ? nullptr
                                              : p_save + ((size_t)(st * 2 + t) * kSaveSlots + kSaveSlot[L]) * kABytes +
                                                    half * (kCols / 8) * 2048 + row * 16;
            if (ly.bias_kind == BIAS_MMA) {
              hidden_epilogue<false, kCols / 32, true>(tmem_d, nullptr, a_row, dbg_row, words, g_row);
            } else if (warp_bias) {
              hidden_epilogue_wbias<kCols / 32, true>(tmem_d, wb, a_row, dbg_row, words, g_row);
            } else {
              const float* bias = (ly.bias_kind == BIAS_RAY ? p.raybias + (s / p.N) * 256 : p.imgbias + (s / p.per_image) * 256) +
                                  half * kCols;
              hidden_epilogue<true, kCols / 32, true>(tmem_d, bias, a_row, dbg_row, words, g_row);
            }
          } else if (p_dbg_drain == 1 || p_dbg_drain == 2) {
            hidden_epilogue_experiment<kCols / 32>(tmem_d, a_row, p_dbg_drain);
          } else if (ly.bias_kind == BIAS_MMA || p_dbg_drain == 3) {   // 3: timing experiment, bias tables ignored
            hidden_epilogue<false, kCols / 32>(tmem_d, nullptr, a_row, dbg_row);
          } else if (warp_bias) {
            hidden_epilogue_wbias<kCols / 32>(tmem_d, wb, a_row, dbg_row);
          } else {
            const float* bias = (ly.bias_kind == BIAS_RAY ? p.raybias + (s / p.N) * 256 : p.imgbias + (s / p.per_image) * 256) +
                                half * kCols;
            hidden_epilogue<true, kCols / 32>(tmem_d, bias, a_row, dbg_row);
          }
          fence_proxy_async_smem();
          if (L == kSpillLayer && kNL > kStaticLayers) {      // static only: no second head, nothing to park
            // park the trunk feature (bf16 tile image) in the L2 scratch (training: in its slot of the save buffer, where
            // the backward also reads it) -- the bulk store overlaps the next stage's MMAs
            named_bar_sync(1 + t, kTileThreads);
            if (row == 0 && half == 0) {
              uint8_t* dst = p_save ? p_save + ((size_t)(st * 2 + t) * kSaveSlots + kSaveSlot[L]) * kABytes : my_scratch;
              bulk_s2g(dst, a_smem, kABytes);
              bulk_commit();
            }
            store_pending = true;
          }
        } else if (half == 0) {
          uint32_t v[8];
          TP_TMEM_LD8(tmem_row, v);
          TP_TMEM_WAIT8(v);
          const float* sb = p.biasbuf + kSmallBiasOffset;
          if (ly.epi == EPI_DENSITY) {
            sigma_s = tp_softplus(__uint_as_float(v[0]) + sb[0]);
          } else if (ly.epi == EPI_RGB_OUT) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb_s[c] = tp_sigmoid(__uint_as_float(v[c]) + sb[1 + c]);
            if (kNL == kStaticLayers && live) {      // static only: this is the last stage; transient outputs are zeros
#pragma unroll
              for (int c = 0; c < 3; ++c) *reinterpret_cast<float2*>(p.rgb + s * 6 + c * 2) = make_float2(rgb_s[c], 0.f);
              *reinterpret_cast<float2*>(p.density + s * 2) = make_float2(sigma_s, 0.f);
              p.uncert[s] = 0.f;
            }
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
        if (L != kNL - 1) {   // the next super-tile's encode arrival covers the last stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_ready(t));
        }
      }
    }
    if (row == 0 && half == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ helper kernels

// desc row: [w_ptr, ld, row0, rows_valid, col0, cols_valid, n_layout, bias_ptr, bias_klocal, transpose]
// transpose != 0: chunk element (n, kl) = W[col0 + kl][row0 + n]  (B operand of the input-gradient GEMMs, dX = dZ W)
// (bias_ptr != 0: element (n, bias_klocal) of the chunk is bias[n] -- the column multiplied by the constant-1 of E)
__global__ void pack_weights_kernel(const long long* __restrict__ desc, __nv_bfloat16* __restrict__ out) {
  const long long* d = desc + (long long)blockIdx.x * 10;
  const float* W = reinterpret_cast<const float*>(d[0]);
  const long long ld = d[1], row0 = d[2], rows_valid = d[3], col0 = d[4], cols_valid = d[5], n_layout = d[6];
  const float* bias = reinterpret_cast<const float*>(d[7]);
  const long long bias_k = d[8], transpose = d[9];
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
    if (in_layout && W && n < rows_valid && kl < cols_valid)
      v = transpose ? W[(col0 + kl) * ld + row0 + n] : W[(row0 + n) * ld + col0 + kl];
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

// out[r, n] = imgbias[r / rays_per_image, n] + sum_j W[n, col0 + j] * viewenc(r)[j];  viewenc = [u, enc(u)], u = ray/|ray|.
// Thread n keeps its weight row in registers; a CTA strides over groups of kRaysPerBlock rays: all encodings of the group first
// (one thread per ray and coordinate), then every thread streams the group's rays; stores are 1 KB-coalesced rows.
constexpr int kRaysPerBlock = 64;
__global__ void __launch_bounds__(256) ray_bias_kernel(const float* __restrict__ ray, long long R, long long rays_per_image,
                                                       int L, const float* __restrict__ W, long long ldw, int col0,
                                                       const float* __restrict__ imgbias, float* __restrict__ out) {
  __shared__ __align__(16) float enc[kRaysPerBlock][28];      // rows of 7 float4: the inner product reads them as broadcast LDS.128
  const int vc = 3 + 6 * L;     // <= 27 (L_view <= 4)
  // the thread's weight row (27 strided, uncoalesced loads) is fetched ONCE: the CTA then strides over groups of rays
  float w[27];
  const int n = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 27; ++j) w[j] = j < vc ? W[n * ldw + col0 + j] : 0.f;
  const long long n_groups = (R + kRaysPerBlock - 1) / kRaysPerBlock;
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long r0 = grp * kRaysPerBlock;
    __syncthreads();      // the previous group's encodings are no longer read
    if (threadIdx.x < kRaysPerBlock * 3) {      // one thread per (ray, coordinate): unit component, then its 2L encodings
      const int rr = threadIdx.x / 3, cc = threadIdx.x - rr * 3;
      const long long r = r0 + rr;
      const bool live = r < R;
      float uc = 0.f;
      if (live) {
        const float x = ray[r * 3], y = ray[r * 3 + 1], z = ray[r * 3 + 2];
        const float len = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
        uc = (cc == 0 ? x : (cc == 1 ? y : z)) / len;
      }
      enc[rr][cc] = uc;
      for (int k = 0; k < L; ++k) {
        const float arg = __fmul_rn(uc, ldexpf(3.14159265358979323846f, k));
        enc[rr][3 + cc * 2 * L + k] = live ? sinf(arg) : 0.f;
        enc[rr][3 + cc * 2 * L + L + k] = live ? cosf(arg) : 0.f;
      }
    } else {      // the other threads clear the pad columns (vc .. 27)
      for (int i = threadIdx.x - kRaysPerBlock * 3; i < kRaysPerBlock * (28 - vc); i += blockDim.x - kRaysPerBlock * 3)
        enc[i / (28 - vc)][vc + i % (28 - vc)] = 0.f;
    }
    __syncthreads();
    // four rays per iteration: four independent fma chains per thread (the 27-long chain of one ray is latency-bound)
    const long long img_first = r0 / rays_per_image;
    for (int rr = 0; rr < kRaysPerBlock; rr += 4) {
      float acc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long r = r0 + rr + q;
        long long img = img_first;
        while (r >= (img + 1) * rays_per_image) ++img;
        acc[q] = r < R ? imgbias[img * 256 + n] : 0.f;
      }
#pragma unroll
      for (int j4 = 0; j4 < 7; ++j4) {      // per ray the same fma order as column by column (the pad column is skipped)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 e = reinterpret_cast<const float4*>(enc[rr + q])[j4];
          acc[q] = fmaf(w[4 * j4], e.x, acc[q]);
          acc[q] = fmaf(w[4 * j4 + 1], e.y, acc[q]);
          acc[q] = fmaf(w[4 * j4 + 2], e.z, acc[q]);
          if (j4 < 6) acc[q] = fmaf(w[4 * j4 + 3], e.w, acc[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (r0 + rr + q < R) __stcs(out + (r0 + rr + q) * 256 + n, acc[q]);      // 315 MB per frame, read once by the fused kernel
    }
  }
}


// tile image [tiles][n_slots][k8=32][128 rows][8] bf16 -> row-major fp32 [S,256] (consumers: the fp32 SIMT kernels)
__global__ void unpack_images_kernel(const uint8_t* __restrict__ images, int slot, int n_slots, long long S,
                                     float* __restrict__ out) {
  const long long total = ((S + 127) / 128) * 4096;     // 16-byte vectors
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
    const long long tile = v >> 12;
    const int k8 = (int)((v >> 7) & 31), r = (int)(v & 127);
    const long long s = tile * 128 + r;
    if (s >= S) continue;
    const uint4 q = *reinterpret_cast<const uint4*>(images + ((size_t)tile * n_slots + slot) * kABytes + k8 * 2048 + r * 16);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    float f[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      f[2 * e] = __uint_as_float(w[e] << 16);
      f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
    }
    float4* o = reinterpret_cast<float4*>(out + s * 256 + k8 * 8);
    o[0] = make_float4(f[0], f[1], f[2], f[3]);
    o[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

}  // namespace tc

int tp_tc_v2_launch(const tc::Params& p, int flags, cudaStream_t stream);   // mlp_tc_v2.cu
int tp_tc_pair_launch(const tc::Params& p, cudaStream_t stream);            // mlp_tc_pair.cu

TP_API int tp_tc_num_chunks(void) { return tc::kNumChunks; }
TP_API int64_t tp_tc_chunk_bytes(void) { return tc::kChunkBytes; }
TP_API int64_t tp_tc_scratch_bytes(void) { return (int64_t)tp_num_sms() * 2 * tc::kABytes; }

// save buffer: [tiles][7][64 KB] tile images, then [tiles][4][4 KB] ReLU bitmasks (tiles rounded up to whole super-tiles)
TP_API int64_t tp_tc_save_bytes(int64_t S) {
  return ((S + 255) / 256) * 2 * (tc::kSaveSlots * (int64_t)tc::kABytes + tc::kMaskBitSlots * (int64_t)tc::kMaskBitBytes);
}

TP_API int tp_tc_unpack_images(const void* images, int slot, int n_slots, int64_t S, float* out, void* stream) {
  if (!images || !out) return TP_ERR_BAD_ARG;
  if (slot < 0 || slot >= n_slots || S < 0) return TP_ERR_BAD_SHAPE;
  if (S == 0) return TP_OK;
  tc::unpack_images_kernel<<<tp_grid_for(((S + 127) / 128) * 4096, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint8_t*>(images), slot, n_slots, S, out);
  return tp_launch_status();
}

TP_API int tp_tc_pack_weights(const int64_t* chunk_desc, int n_chunks, void* packed, void* stream) {
  if (!chunk_desc || !packed) return TP_ERR_BAD_ARG;
  if (n_chunks < 1) return TP_ERR_BAD_SHAPE;
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
  if (R < 0 || rays_per_image < 1 || L_view < 0 || L_view > 4) return TP_ERR_BAD_SHAPE;
  if (R == 0) return TP_OK;
  const long long n_groups = (R + tc::kRaysPerBlock - 1) / tc::kRaysPerBlock, cap = (long long)tp_num_sms() * 6;
  tc::ray_bias_kernel<<<(unsigned)(n_groups < cap ? n_groups : cap), 256, 0, (cudaStream_t)stream>>>(
      ray, R, rays_per_image, L_view, W, ldw, col0, imgbias, out);
  return tp_launch_status();
}

TP_API int tp_tc_nerf_stl_forward(const float* center, const float* ray, const float* depth, int64_t S, int N,
                                  int64_t per_image, const void* packed, const float* biasbuf, const float* raybias,
                                  const float* imgbias, float* rgb, float* density, float* uncert, void* scratch,
                                  int64_t scratch_bytes, void* save, int dbg_layer, float* dbg_out, int flags, void* stream) {
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
  p.save = reinterpret_cast<uint8_t*>(save);
  if (((uintptr_t)save & 15)) return TP_ERR_ALIGN;
  p.bits = save ? p.save + ((S + 255) / 256) * 2 * tc::kSaveSlots * (size_t)tc::kABytes : nullptr;
  if (save && (flags & 128)) return TP_ERR_BAD_ARG;           // the experimental kernel does not write the ReLU bitmasks
  p.dbg_layer = dbg_layer; p.dbg_out = dbg_out; p.dbg_drain = (flags >> 2) & 7;
  p.dbg_save = (flags >> 14) & 3;
  p.n_layers = (flags & (1 << 17)) ? tc::kStaticLayers : tc::kNumLayers;      // flags bit 17: static-only rendering
  if ((flags & (1 << 17)) && (save || (flags & (128 | 1024)))) return TP_ERR_BAD_ARG;      // inference launch of the default kernel only
  p.skew = ((flags >> 5) & 3) ? ((flags >> 5) & 3) - 1 : 1;      // default skew 1; flags bits 5-6 = skew+1 override (A/B)
  if (flags & 128) return tp_tc_v2_launch(p, flags, (cudaStream_t)stream);   // experimental single-tile / cluster kernel
  if (flags & 1024) {     // CTA-pair kernel (cta_group::2); `packed` must be the pair image; flags bits 11-13 = skew + 1 (default 4)
    p.skew = ((flags >> 11) & 7) ? ((flags >> 11) & 7) - 1 : 4;
    return tp_tc_pair_launch(p, (cudaStream_t)stream);
  }
  // drain width: 8 epilogue warps (256 accumulator columns per thread, 168 registers, 320 threads per CTA) by default -- same
  // tensor-pipe time as 16 warps, but the narrower CTA draws less power and the capped clock settles higher: 2 % faster for
  // the C2 frame, 1.7 % for the C3 training step (same-box A/B, profiles/r01f_summary.md section 9).  flags bit 1 selects 16 warps.
  const bool wide = (flags & 2) != 0;
  const bool stat = p.n_layers == tc::kStaticLayers;
  const bool plain = dbg_layer < 0 && !dbg_out && p.dbg_drain == 0 && p.dbg_save == 0 && p.skew == 1 && N % 32 == 0;
  void (*kern)(const tc::Params) =
      plain && !save && !wide ? (stat ? tc::nerf_stl_forward_kernel<1, tc::kStaticLayers, 1> : tc::nerf_stl_forward_kernel<1, tc::kNumLayers, 1>)
      : plain && !save        ? (stat ? tc::nerf_stl_forward_kernel<2, tc::kStaticLayers, 1> : tc::nerf_stl_forward_kernel<2, tc::kNumLayers, 1>)
      : plain && wide         ? tc::nerf_stl_forward_kernel<2, tc::kNumLayers, 2>
      : plain                 ? tc::nerf_stl_forward_kernel<1, tc::kNumLayers, 2>
      : wide ? (stat ? tc::nerf_stl_forward_kernel<2, tc::kStaticLayers> : tc::nerf_stl_forward_kernel<2, tc::kNumLayers>)
             : (stat ? tc::nerf_stl_forward_kernel<1, tc::kStaticLayers> : tc::nerf_stl_forward_kernel<1, tc::kNumLayers>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, wide ? tc::num_threads<2>() : tc::num_threads<1>(), tc::kSmemBytes, (cudaStream_t)stream>>>(p);
  return tp_launch_status();
}

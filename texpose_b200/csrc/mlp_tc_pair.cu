// K2, CTA-pair variant (flags bit 10 of tp_tc_nerf_stl_forward): the fused forward of mlp_tc.cu issued as
// tcgen05.mma.cta_group::2 over a cluster of two CTAs (the two SMs of a TPC).  Experimental: bit-identical to the default
// kernel, measured 8-12 % slower (profiles/r01b_summary.md 5); kept as the measured alternative.
//
// Why it was built: the single-CTA kernel is limited by energy and SMEM bytes per FLOP (profiles/r01b_summary.md 2).  With
// cta_group::2 one MMA covers M = 256 rows -- the 128-sample tile of the leader CTA and the 128-sample tile of its peer --
// against ONE B operand of which each SM holds half (N/2 rows):
//   * SMEM operand reads per SM and MMA drop from 12 KB (A 4 + B 8) to 8 KB (A 4 + B/2 4);
//   * every CTA streams only its half of each weight chunk, so the two tiles of a CTA can take their MMA passes over a stage
//     back to back (weights streamed once per tile, same L2 traffic per SM as the default kernel) and one tile's accumulator
//     drain -- by all 16 epilogue warps -- hides behind the other tile's MMAs.
//
// Roles per CTA: 16 epilogue warps (both tiles, alternating), one weight producer warp, one MMA warp (leader CTA only).
// Cross-CTA signalling without a forwarding hop:
//   * weights: cp.async.bulk.tensor ... cta_group::2 -- each CTA's TMA writes its half into its own ring slot and
//     complete_tx's on the LEADER's full barrier, which therefore counts the bytes of both halves;
//   * accumulator-full and ring-slot-empty: tcgen05.commit ... multicast::cluster to the barrier at the same offset in both CTAs;
//   * "the peer's A operand is written": remote mbarrier arrives from the peer's epilogue warps (one per warp) on a leader
//     barrier, with CTA-scope semantics (cluster-scope acquire / release compile to MEMBAR.ALL.GPU + CCTL.IVALL: +40 %).
//
// Weight image: tp_tc_pair_weights re-orders the standard image so that rank r's half of chunk c is contiguous at
// c * 16 KB + r * 8 KB  ([k8][128 rows][8] for the 256-row chunks, [32 k8][8 rows][8] for the N=16 chunks); the kernel sees it
// through a 2-D tensor map of 2 KB rows.
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include "mlp_tc_shared.cuh"
#include "mlp_tc_epilogue.cuh"
#include "../../include/texpose_b200.h"

// cycle-counter instrumentation (scripts/pair_prof.py): compiled in only with -DTP_PAIR_PROF
#ifdef TP_PAIR_PROF
#define PP(...) __VA_ARGS__
// event trace of one iteration of CTA pair 0: trace[cta][stage][tile][event] (clock64), behind the per-CTA counters
#define PTRACE(L, t, ev)                                                                                             \
  if (p.dbg_layer == 100 && p.dbg_out && it == 3 && blockIdx.x < 2)                                                  \
  reinterpret_cast<long long*>(p.dbg_out)[148 * 16 + ((blockIdx.x * 17 + (L)) * 2 + (t)) * 8 + (ev)] = clock64()
#else
#define PP(...)
#define PTRACE(L, t, ev)
#endif

namespace tc3 {
using namespace tc;

constexpr int kThreads = 18 * 32;              // 16 epilogue warps, producer warp, MMA warp (leader only)
constexpr int kRing = 4;
constexpr uint32_t kHalfBytes = 8192;          // one chunk half
constexpr uint32_t kSlotBytes = 2 * kHalfBytes; // a ring slot holds the halves of two consecutive chunks of a pass
constexpr uint32_t kOffA = 0, kOffE = 2 * kABytes, kOffRing = kOffE + 2 * kEBytes;
constexpr uint32_t kOffBar = kOffRing + kRing * kSlotBytes;
constexpr uint32_t kSmemBytes = kOffBar + 512;
static_assert(kSmemBytes <= 232448, "pair kernel exceeds 227 KB of shared memory");

// ------------------------------------------------------------------------------------------ cluster / cta_group::2 PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Remote arrive with the default (.release / .acquire at CTA scope) semantics, as CUTLASS' ClusterBarrier does: the
// operands the signal stands for never cross CTAs through the generic proxy -- each SM's tensor core reads its OWN shared
// memory, and the peer's writers ordered their st.shared before the local barrier with fence.proxy.async.  Cluster-scope
// acquire / release would compile to MEMBAR.ALL.GPU + CCTL.IVALL on every chunk (measured: +60 % kernel time).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load whose completion is signalled on a barrier of the LEADER CTA of the pair (cta_group::2): both CTAs' weight
// halves complete_tx on the one barrier the MMA warp waits on -- no forwarding hop for the weights
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst_smem, const CUtensorMap* map, int x, int y, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
      "l"(map), "r"(leader_bar), "r"(x), "r"(y)
      : "memory");
}
// low descriptor word; inside a cluster a shared::cta address carries the CTA rank above bit 18: mask it off
__device__ __forceinline__ uint32_t dlo(uint32_t addr, uint32_t lbo) { return ((addr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16); }
__device__ __forceinline__ void umma2_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
// completion of all prior MMAs of this thread -> the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) nerf_stl_forward_pair_kernel(const Params p, int iters, const __grid_constant__ CUtensorMap wmap8,
                                                                                                   const __grid_constant__ CUtensorMap wmap4) {
  constexpr int kProducerWarp = 16, kCtrlWarp = 17;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long long n_super = (p.S + 255) / 256;
  const long long n_pairs = gridDim.x / 2, pair = blockIdx.x / 2;
  // barriers (same offsets in both CTAs)
  const uint32_t bar0 = sbase + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (kRing + s); };
  auto bar_acc = [&](int t) { return bar0 + 8 * (2 * kRing + t); };
  auto bar_ready = [&](int t) { return bar0 + 8 * (2 * kRing + 2 + t); };
  auto bar_reload = [&](int t) { return bar0 + 8 * (2 * kRing + 4 + t); };
  auto bar_pready = [&](int t) { return bar0 + 8 * (2 * kRing + 6 + t); };    // leader: the peer's A operand of tile t is written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * (2 * kRing + 8));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRing; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_acc(t), 1);
      mbar_init(bar_ready(t), 16);         // one arrive per epilogue warp
      mbar_init(bar_reload(t), 1);
      mbar_init(bar_pready(t), 16);        // the peer's 16 epilogue warps arrive remotely
    }
    fence_barrier_init();
  }
  if (warp == kCtrlWarp) {   // TMEM: all 512 columns in both CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised and TMEM is allocated before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ================================================================ weight producer: this CTA's half of every chunk,
    // streamed once per tile (tile 0's pass over a stage, then tile 1's): the two tiles do not share ring slots, so a whole
    // stage of MMAs of one tile is contiguous and the other tile's accumulator drain hides behind it
    uint32_t cnt = 0;
    const uint32_t full_leader = map_to_rank(bar_full(0), 0);      // the leader's full barriers (own address in the leader)
    for (int it = 0; it < iters; ++it) {
      int c0 = 0;
      for (int L = 0; L < kNumLayers; ++L) {
        const Layer ly = kLayers[L];
        const int nch = ly.small ? 1 : ly.a_chunks + ly.e_chunks + ly.bias_chunk;
        const int nsl = (nch + 1) / 2;
        for (int t = 0; t < 2; ++t) {
          for (int sl = 0; sl < nsl; ++sl, ++cnt) {
            const uint32_t stage = cnt % kRing, phase = (cnt / kRing) & 1u;
            uint32_t bytes[2] = {0u, 0u};
            for (int k = 0; k < 2; ++k) {
              const int j = 2 * sl + k;
              if (j < nch) bytes[k] = (ly.small || j >= ly.a_chunks + ly.e_chunks) ? kHalfBytes / 2 : kHalfBytes;
            }
            mbar_wait(bar_empty(stage), phase ^ 1);
            if (elect_one_sync()) {
              // the leader's barrier counts the bytes of BOTH CTAs' halves (the tx-count may dip below zero until it arrives)
              if (rank == 0) mbar_expect_tx(bar_full(stage), 2 * (bytes[0] + bytes[1]));
              for (int k = 0; k < 2; ++k)
                if (bytes[k]) {
                  const int row = ((c0 + 2 * sl + k) * 2 + (int)rank) * 4;     // rows of 2 KB (256 x u64) in the pair image
                  tma_load_2d_pair(sbase + kOffRing + stage * kSlotBytes + k * kHalfBytes, bytes[k] == kHalfBytes ? &wmap8 : &wmap4,
                                   0, row, full_leader + 8 * stage);
                }
            }
            __syncwarp();
          }
        }
        c0 += nch;
      }
    }
  } else if (warp == kCtrlWarp && rank != 0) {
    // the peer's control warp has nothing to do: its weight halves signal the leader's barriers directly (TMA, cta_group::2),
    // its A-operand readiness arrives remotely from its epilogue warps
  } else if (warp == kCtrlWarp) {
    // ================================================================ MMA issuer (leader only): one wait per ring slot
    uint32_t cnt = 0, ready_ph = 0, reload_ph = 0;
    const uint32_t idesc256 = umma_idesc(256, 256), idesc16 = umma_idesc(256, 16);
    constexpr uint32_t kHi = (128u >> 4) | (1u << 14);   // SBO = 128 B, descriptor version 1
    PP(long long pf_full = 0, pf_ready[2] = {0, 0}, pf_peer[2] = {0, 0}, pf_issue = 0; const long long pf_t0 = clock64();)
    for (int it = 0; it < iters; ++it) {
      for (int L = 0; L < kNumLayers; ++L) {
        const Layer ly = kLayers[L];
        const int nch = ly.small ? 1 : ly.a_chunks + ly.e_chunks + ly.bias_chunk;
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int nsl = (nch + 1) / 2;
          for (int sl = 0; sl < nsl; ++sl, ++cnt) {
            const uint32_t stage = cnt % kRing;
            PP(long long pa = clock64();)
            mbar_wait(bar_full(stage), (cnt / kRing) & 1u);      // both CTAs' halves of this slot landed
            PP(const long long pw = clock64(); pf_peer[0] += pw - pa;)
            if (sl == 0) {                                        // the A operand of tile t is written in both CTAs
              mbar_wait(bar_ready(t), (ready_ph >> t) & 1u);
              PP(const long long pr = clock64(); pf_peer[1] += pr - pw; PTRACE(L, t, 6);)
              mbar_wait(bar_pready(t), (ready_ph >> t) & 1u);
              PP(PTRACE(L, t, 4); PTRACE(L, t, 5);)
              ready_ph ^= 1u << t;
              if (ly.reload) {
                mbar_wait(bar_reload(t), (reload_ph >> t) & 1u);
                reload_ph ^= 1u << t;
              }
            }
            tc_fence_after();
            PP(long long pb = clock64(); pf_full += pb - pa;)
            PP(if (lane == 0 && sl == 0) { PTRACE(L, t, 0); })
            const uint32_t d_tmem = tmem_base + t * 256;
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int c = 2 * sl + k;
                if (c >= nch) break;
                const uint32_t wsm = sbase + kOffRing + stage * kSlotBytes + k * kHalfBytes;
                if (ly.small) {
                  // half chunk = [32 k8][8 rows][8]: 16 K-steps over the full K=256 of A_t
                  uint32_t a_lo = dlo(sbase + kOffA + t * kABytes, 2048u);
                  uint32_t b_lo = dlo(wsm, 128u);
#pragma unroll
                  for (int ks = 0; ks < 16; ++ks) {
                    umma2_bf16_lohi(d_tmem, a_lo, kHi, b_lo, kHi, idesc16, ks > 0 ? 1u : 0u);
                    a_lo += 4096u >> 4;
                    b_lo += 256u >> 4;
                  }
                } else if (c >= ly.a_chunks + ly.e_chunks) {
                  // bias step: A = E columns 48..63 (column 63 == 1), B half = [2 k8][128 rows][8]
                  const uint32_t a_lo = dlo(sbase + kOffE + t * kEBytes + 6 * 2048, 2048u);
                  const uint32_t b_lo = dlo(wsm, 2048u);
                  umma2_bf16_lohi(d_tmem, a_lo, kHi, b_lo, kHi, idesc256, 1u);
                } else {
                  const bool from_e = c >= ly.a_chunks;
                  const uint32_t a0 = from_e ? sbase + kOffE + t * kEBytes + (c - ly.a_chunks) * 4 * 2048
                                             : sbase + kOffA + t * kABytes + c * 4 * 2048;
                  const uint32_t a_lo = dlo(a0, 2048u);
                  const uint32_t b_lo = dlo(wsm, 2048u);
                  umma2_bf16_lohi(d_tmem, a_lo, kHi, b_lo, kHi, idesc256, c > 0 ? 1u : 0u);
                  umma2_bf16_lohi(d_tmem, a_lo + (4096u >> 4), kHi, b_lo + (4096u >> 4), kHi, idesc256, 1u);
                }
              }
              PP(const long long pm = clock64(); pf_ready[0] += pm - pb;)
              if (sl == nsl - 1) umma2_commit_both(bar_acc(t));      // accumulators of tile t complete in both CTAs
              umma2_commit_both(bar_empty(stage));                   // ring slot reusable in both CTAs
              PP(pf_ready[1] += clock64() - pm;)
            }
            __syncwarp();
            PP(pf_issue += clock64() - pb;)
            PP(if (lane == 0 && sl == nsl - 1) { PTRACE(L, t, 1); })
          }
        }
      }
    }
    PP(if (p.dbg_layer == 100 && p.dbg_out && lane == 0) {
      long long* o = reinterpret_cast<long long*>(p.dbg_out) + (size_t)blockIdx.x * 16;
      o[0] = clock64() - pf_t0; o[1] = pf_full; o[2] = pf_ready[0]; o[3] = pf_ready[1]; o[4] = pf_peer[0]; o[5] = pf_peer[1];
      o[6] = pf_issue; o[7] = iters;
    })
  } else {
    // ================================================================ encode + epilogue warps
    // All 16 warps drain ONE tile at a time (tile 0's accumulator, then tile 1's, alternating with the MMA passes): warp ->
    // (TMEM lane quarter q, column quarter cq), 64 accumulator columns per thread.  With the tiles' MMA passes back to back, a
    // drain by all 16 warps (4 per scheduler) is short enough to hide behind the other tile's MMAs.
    const int q = warp & 3, cq = warp >> 2, row = q * 32 + lane;
    constexpr int kCols = 64;
    const bool warp_bias = (p.N % 32 == 0);
    uint32_t acc_ph = 0;                       // bit t
    bool store_pending[2] = {false, false};
    const uint32_t pready_remote = map_to_rank(bar_pready(0), 0);     // leader's barrier for the peer's A-operand readiness
    PP(long long pe_acc = 0, pe_work = 0;)
    for (int it = 0; it < iters; ++it) {
      const long long st = ((long long)it * n_pairs + pair) * 2 + rank;     // super-tile of this CTA (may lie beyond the end)
      const bool tile_live = st < n_super;
      long long s_t[2];
      bool live_t[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const long long s_raw = (st * 2 + t) * 128 + row;
        live_t[t] = s_raw < p.S;
        s_t[t] = live_t[t] ? s_raw : p.S - 1;
      }
      if (cq < 2) encode_sample(p, s_t[cq], sbase + kOffE + cq * kEBytes, row);      // column quarter 0 / 1 encodes tile 0 / 1
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) {
          mbar_arrive(bar_ready(0));
          mbar_arrive(bar_ready(1));
        } else {
          mbar_arrive_remote(pready_remote);
          mbar_arrive_remote(pready_remote + 8);
        }
      }

      float sigma_s[2] = {0.f, 0.f}, rgb_s[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
      for (int L = 0; L < kNumLayers; ++L) {
        const Layer ly = kLayers[L];
        const bool table_bias = ly.epi == EPI_HIDDEN && ly.bias_kind != BIAS_MMA;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const long long s = s_t[t];
          const bool live = live_t[t];
          const uint32_t a_smem = sbase + kOffA + t * kABytes;
          const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + t * 256;
          const uint32_t tmem_d = tmem_row + cq * kCols;
          uint8_t* my_scratch = p.scratch + ((size_t)blockIdx.x * 2 + t) * kABytes;
          float4 wb[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
          if (table_bias && warp_bias && lane < 16) {     // 64 columns = 16 float4 slices, redistributed by shuffles in the drain
            const float* brow = (ly.bias_kind == BIAS_RAY ? p.raybias + (s / p.N) * 256 : p.imgbias + (s / p.per_image) * 256) +
                                cq * kCols;
            wb[0] = __ldg(reinterpret_cast<const float4*>(brow) + lane);
          }
          PP(const long long ea = clock64();)
          mbar_wait(bar_acc(t), (acc_ph >> t) & 1u);
          acc_ph ^= 1u << t;
          tc_fence_after();
          PP(const long long eb = clock64(); pe_acc += eb - ea;)
          PP(if (threadIdx.x == 0) { PTRACE(L, t, 2); })
          if (ly.epi == EPI_HIDDEN && store_pending[t]) {   // the previous bulk store must have finished reading A_t
            if (threadIdx.x == 0) bulk_wait_read();
            named_bar_sync(1, 512);
            store_pending[t] = false;
          }
          if (L == kReloadIssueLayer && threadIdx.x == 0) {
            bulk_wait_all();
            fence_proxy_async_all();
            mbar_expect_tx(bar_reload(t), kABytes);
            bulk_g2s(a_smem, (p.save && tile_live) ? p.save + ((size_t)(st * 2 + t) * kSaveSlots) * kABytes : my_scratch, kABytes,
                     bar_reload(t));
          }
          if (ly.epi == EPI_HIDDEN) {
            float* dbg_row = ((L == p.dbg_layer) && live && p.dbg_out) ? p.dbg_out + s * 256 + cq * kCols : nullptr;
            const uint32_t a_row = a_smem + cq * (kCols / 8) * 2048 + row * 16;
            const int mslot = (p.bits && tile_live) ? kMaskBitSlot[L] : -1;
            if (mslot >= 0) {
              uint32_t* words = reinterpret_cast<uint32_t*>(p.bits + ((size_t)(st * 2 + t) * 4 + mslot) * kMaskBitBytes) +
                                cq * (kCols / 32) * 128 + row;
              if (ly.bias_kind == BIAS_MMA) {
                hidden_epilogue<false, kCols / 32, true>(tmem_d, nullptr, a_row, dbg_row, words);
              } else if (warp_bias) {
                hidden_epilogue_wbias<kCols / 32, true>(tmem_d, wb, a_row, dbg_row, words);
              } else {
                const float* bias = (ly.bias_kind == BIAS_RAY ? p.raybias + (s / p.N) * 256 : p.imgbias + (s / p.per_image) * 256) +
                                    cq * kCols;
                hidden_epilogue<true, kCols / 32, true>(tmem_d, bias, a_row, dbg_row, words);
              }
            } else if (ly.bias_kind == BIAS_MMA) {
              hidden_epilogue<false, kCols / 32>(tmem_d, nullptr, a_row, dbg_row);
            } else if (warp_bias) {
              hidden_epilogue_wbias<kCols / 32>(tmem_d, wb, a_row, dbg_row);
            } else {
              const float* bias = (ly.bias_kind == BIAS_RAY ? p.raybias + (s / p.N) * 256 : p.imgbias + (s / p.per_image) * 256) +
                                  cq * kCols;
              hidden_epilogue<true, kCols / 32>(tmem_d, bias, a_row, dbg_row);
            }
            fence_proxy_async_smem();
            if (L == kSpillLayer || (p.save && kSaveSlot[L] >= 0)) {
              named_bar_sync(1, 512);
              if (threadIdx.x == 0) {
                uint8_t* dst = (p.save && tile_live) ? p.save + ((size_t)(st * 2 + t) * kSaveSlots + kSaveSlot[L]) * kABytes : my_scratch;
                if (L == kSpillLayer || tile_live) {
                  bulk_s2g(dst, a_smem, kABytes);
                  bulk_commit();
                }
              }
              store_pending[t] = true;
            }
          } else if (cq == 0) {
            uint32_t v[8];
            TP_TMEM_LD8(tmem_row, v);
            TP_TMEM_WAIT8(v);
            const float* sb = p.biasbuf + kSmallBiasOffset;
            if (ly.epi == EPI_DENSITY) {
              sigma_s[t] = tp_softplus(__uint_as_float(v[0]) + sb[0]);
            } else if (ly.epi == EPI_RGB_OUT) {
#pragma unroll
              for (int c = 0; c < 3; ++c) rgb_s[t][c] = tp_sigmoid(__uint_as_float(v[c]) + sb[1 + c]);
            } else {
              float rgb_t[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) rgb_t[c] = tp_sigmoid(__uint_as_float(v[c]) + sb[4 + c]);
              const float sigma_t = tp_softplus(__uint_as_float(v[3]) + sb[7]);
              const float unc = tp_softplus(__uint_as_float(v[4]) + sb[8]);
              if (live) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                  *reinterpret_cast<float2*>(p.rgb + s * 6 + c * 2) = make_float2(rgb_s[t][c], rgb_t[c]);
                *reinterpret_cast<float2*>(p.density + s * 2) = make_float2(sigma_s[t], sigma_t);
                p.uncert[s] = unc;
              }
            }
          }
          PP(if (threadIdx.x == 0) { PTRACE(L, t, 3); })
          if (L != kNumLayers - 1) {   // the next super-tile's encode arrival covers the last stage
            tc_fence_before();
            __syncwarp();                // every lane's st.shared + proxy fence precede the warp's single arrive
            if (lane == 0) {
              if (rank == 0) mbar_arrive(bar_ready(t));
              else {
                // the next stage's A operand is the feature tile coming back by TMA: the peer vouches for its own reload
                if (kLayers[L + 1].reload) mbar_wait(bar_reload(t), it & 1u);
                mbar_arrive_remote(pready_remote + 8 * t);
              }
            }
          }
          PP(pe_work += clock64() - eb;)
        }
      }
    }
    if (threadIdx.x == 0) bulk_wait_all();
    PP(if (p.dbg_layer == 100 && p.dbg_out && threadIdx.x == 0) {
      long long* o = reinterpret_cast<long long*>(p.dbg_out) + (size_t)blockIdx.x * 16 + 8;
      o[0] = pe_acc; o[1] = pe_work;
    })
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody exits (or frees TMEM) while the leader's MMAs may still read the peer's SMEM / TMEM
  if (warp == kCtrlWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// standard weight image -> pair image: rank r's half of chunk c contiguous at c * 16 KB + r * 8 KB
__global__ void pair_weights_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n_chunks) {
  // kind of every chunk in consumption order (0 = 256-row K=32, 1 = 256-row K=16 bias chunk, 2 = 16-row K=256)
  __shared__ int kind[kNumChunks];
  if (threadIdx.x == 0) {
    int c = 0;
    for (int L = 0; L < kNumLayers; ++L) {
      const Layer ly = kLayers[L];
      if (ly.small) { kind[c++] = 2; continue; }
      for (int j = 0; j < ly.a_chunks + ly.e_chunks; ++j) kind[c++] = 0;
      if (ly.bias_chunk) kind[c++] = 1;
    }
  }
  __syncthreads();
  const int c = blockIdx.x;
  if (c >= n_chunks) return;
  const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)c * kChunkBytes);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)c * kChunkBytes);
  const int units = kind[c] == 0 ? 1024 : 512;             // 16-byte units carrying data
  for (int u = threadIdx.x; u < 1024; u += blockDim.x) dst[u] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int u = threadIdx.x; u < units; u += blockDim.x) {
    int d;
    if (kind[c] == 2) {                                    // [32 k8][16 rows] -> [r][32 k8][8 rows]
      const int k8 = u >> 4, row = u & 15;
      d = (row >> 3) * 512 + k8 * 8 + (row & 7);
    } else {                                               // [k8][256 rows] -> [r][k8][128 rows]
      const int k8 = u >> 8, row = u & 255;
      d = (row >> 7) * 512 + k8 * 128 + (row & 127);
    }
    dst[d] = src[u];
  }
}

}  // namespace tc3

TP_API int tp_tc_pair_weights(const void* packed, void* pair_packed, void* stream) {
  if (!packed || !pair_packed) return TP_ERR_BAD_ARG;
  if (((uintptr_t)packed & 15) || ((uintptr_t)pair_packed & 15)) return TP_ERR_ALIGN;
  tc3::pair_weights_kernel<<<tc::kNumChunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(packed),
                                                                            reinterpret_cast<uint8_t*>(pair_packed),
                                                                            tc::kNumChunks);
  return tp_launch_status();
}

// Tensor map over the pair image seen as rows of 2 KB (256 x u64): a chunk half is a box of 4 rows (8 KB) or 2 rows (4 KB).
// cuTensorMapEncodeTiled comes from the driver through the runtime (no link against libcuda).
static int make_weight_map(CUtensorMap* map, const void* base, int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess) return (int)e;
  if (!fn || qres != cudaDriverEntryPointSuccess) return (int)cudaErrorNotSupported;
  const cuuint64_t dims[2] = {256, (cuuint64_t)tc::kNumChunks * 8};       // 8 rows of 2 KB per 16 KB chunk
  const cuuint64_t strides[1] = {2048};
  const cuuint32_t box[2] = {256, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = reinterpret_cast<EncodeFn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), dims, strides, box,
                                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : TP_ERR_BAD_ARG;
}

// dispatched from tp_tc_nerf_stl_forward (flags bit 10); p.packed must be the pair image (tp_tc_pair_weights)
int tp_tc_pair_launch(const tc::Params& p, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(tc3::nerf_stl_forward_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc3::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  alignas(64) CUtensorMap wmap8, wmap4;
  if (int rc = make_weight_map(&wmap8, p.packed, 4)) return rc;
  if (int rc = make_weight_map(&wmap4, p.packed, 2)) return rc;
  const long long n_super = (p.S + 255) / 256;
  int grid = (tp_num_sms() / 2) * 2;
  const long long need = ((n_super + 1) / 2) * 2;
  if (need < grid) grid = (int)need;
  const long long per_pass = grid;                          // super-tiles per pass = CTAs
  const int iters = (int)((n_super + per_pass - 1) / per_pass);
  tc3::nerf_stl_forward_pair_kernel<<<grid, tc3::kThreads, tc3::kSmemBytes, stream>>>(p, iters, wmap8, wmap4);
  return tp_launch_status();
}

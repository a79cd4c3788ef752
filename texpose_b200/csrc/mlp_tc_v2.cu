// K2 v2 (experimental, selected with flags bit 7 of tp_tc_nerf_stl_forward): same fused forward as mlp_tc.cu, restructured so
// the tensor pipe does not wait for the accumulator drain.
//
//   * ONE 128-sample tile per CTA iteration; the two 256-column TMEM accumulators are a DOUBLE BUFFER: stage g accumulates
//     into D[g & 1] while the epilogue drains D[(g-1) & 1];
//   * the drain writes the next stage's A operand IN PLACE, 32 columns (one K-chunk) at a time, and signals one mbarrier per
//     slab: the next stage's MMAs on K-chunk j start as soon as slab j is converted, so MMA(g+1) trails drain(g) by one slab;
//   * weight traffic per sample would double with one tile per pass, so every 16 KB chunk is fetched ONCE PER CLUSTER: each
//     CTA of a kCluster-CTA cluster bulk-copies 1/kCluster of the chunk with .multicast::cluster into all CTAs' ring slots;
//     ring slots are released by tcgen05.commit multicast to every CTA's empty barrier (count = kCluster);
//   * 8-slot (128 KB) weight ring -- the SMEM freed by the second tile.
//
// SMEM: A 64K | E[2] 2x16K | ring 8x16K | barriers = 229 632 B.   TMEM: 512 columns = D[0], D[1].
#include "mlp_tc_shared.cuh"
#include "../../include/texpose_b200.h"

namespace tc2 {
using namespace tc;

constexpr int kThreads = 320;                 // warps 0-7 encode + drain, warp 8 weight producer, warp 9 MMA issuer
constexpr int kRing = 8;
constexpr uint32_t kOffA = 0, kOffE = kABytes, kOffRing = kOffE + 2 * kEBytes;
constexpr uint32_t kOffBar = kOffRing + kRing * kChunkBytes;
constexpr uint32_t kSmemBytes = kOffBar + 256;
constexpr int kEncodeAtStage = 1;             // the epilogue warps encode the next tile after draining this stage
// does stage L read an A operand whose slabs were written by the immediately preceding drain?
static __constant__ int kAFresh[kNumLayers] = {0, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1};

#define TP_TMEM_LD16(taddr, v)                                                                                       \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                     \
      : "r"(taddr)                                                                                                   \
      : "memory")
#define TP_TMEM_WAIT16(v)                                                                                            \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                      \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),     \
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])::"memory")

// low descriptor word: start>>4 in [0,14), LBO>>4 in [16,30).  Inside a cluster a shared::cta address carries the CTA rank
// above bit 18 (rank 1 -> bit 24 set): it must be masked off or it lands in the LBO field.
__device__ __forceinline__ uint32_t dlo(uint32_t addr, uint32_t lbo) { return ((addr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// 16 accumulator columns of one row: (+table bias) -> ReLU -> bf16 -> two 16-byte core-matrix rows of the A operand
__device__ __forceinline__ void convert16(const uint32_t (&v)[16], const float* bias16, bool use_wb, const float4 (&wb)[2], int col0,
                                          uint32_t a_dst, float* dbg) {
  float x[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) x[e] = __uint_as_float(v[e]);
  if (use_wb) {
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      const int c = col0 + gq * 4;
      const float4 src = wb[(c >> 7) & 1];
      const int l = (c & 127) >> 2;
      x[gq * 4 + 0] += __shfl_sync(0xffffffffu, src.x, l);
      x[gq * 4 + 1] += __shfl_sync(0xffffffffu, src.y, l);
      x[gq * 4 + 2] += __shfl_sync(0xffffffffu, src.z, l);
      x[gq * 4 + 3] += __shfl_sync(0xffffffffu, src.w, l);
    }
  } else if (bias16) {
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias16) + gq);
      x[gq * 4 + 0] += b.x; x[gq * 4 + 1] += b.y; x[gq * 4 + 2] += b.z; x[gq * 4 + 3] += b.w;
    }
  }
  st_shared_v4(a_dst, pack_relu_bf16(x[0], x[1]), pack_relu_bf16(x[2], x[3]), pack_relu_bf16(x[4], x[5]), pack_relu_bf16(x[6], x[7]));
  st_shared_v4(a_dst + 2048, pack_relu_bf16(x[8], x[9]), pack_relu_bf16(x[10], x[11]), pack_relu_bf16(x[12], x[13]),
               pack_relu_bf16(x[14], x[15]));
  if (dbg) {
#pragma unroll
    for (int e = 0; e < 16; ++e) dbg[e] = fmaxf(x[e], 0.f);
  }
}

template <int kCluster>
__global__ void __launch_bounds__(kThreads, 1) nerf_stl_forward_v2_kernel(const Params p, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t bar0 = sbase + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (kRing + s); };
  auto bar_acc = [&](int b) { return bar0 + 8 * (2 * kRing + b); };
  auto bar_dfree = [&](int b) { return bar0 + 8 * (2 * kRing + 2 + b); };
  auto bar_slab = [&](int j) { return bar0 + 8 * (2 * kRing + 4 + j); };
  auto bar_eready = [&](int b) { return bar0 + 8 * (2 * kRing + 12 + b); };
  const uint32_t bar_reload = bar0 + 8 * (2 * kRing + 14);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * (2 * kRing + 15));
  const uint32_t rank = kCluster > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << kCluster) - 1u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRing; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), p.dbg_drain == 6 ? 1 : kCluster);   // 6: debug, CTAs of the cluster run independently
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc(b), 1);
      mbar_init(bar_dfree(b), 256);
      mbar_init(bar_eready(b), 256);
    }
    for (int j = 0; j < 8; ++j) mbar_init(bar_slab(j), 256);
    mbar_init(bar_reload, 1);
    fence_barrier_init();
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();      // every CTA's barriers are initialised before any remote signal / copy
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long n_tiles = (p.S + 127) / 128;

  if (warp == 8) {
    // ================================================================ weight producer: my 1/kCluster of every chunk, multicast
    uint32_t cnt = 0;
    for (int it = 0; it < iters; ++it) {
      int c = 0;
      for (int L = 0; L < kNumLayers; ++L) {
        const Layer ly = kLayers[L];
        const int nch = ly.small ? 1 : ly.a_chunks + ly.e_chunks + ly.bias_chunk;
        for (int j = 0; j < nch; ++j, ++c, ++cnt) {
          const uint32_t bytes = (ly.small || j >= ly.a_chunks + ly.e_chunks) ? kChunkBytes / 2 : kChunkBytes;
          const uint32_t slot = cnt % kRing, phase = (cnt / kRing) & 1u;
          mbar_wait(bar_empty(slot), phase ^ 1);          // all kCluster CTAs have retired their MMAs on this slot
          if (elect_one_sync()) {
            mbar_expect_tx(bar_full(slot), bytes);
            if (kCluster > 1 && p.dbg_drain != 5 && p.dbg_drain != 6) {
              const uint32_t piece = bytes / kCluster;
              const uint32_t dst = sbase + kOffRing + slot * kChunkBytes + rank * piece;
              const uint8_t* src = p.packed + (size_t)c * kChunkBytes + rank * piece;
              bulk_g2s_multicast(dst, src, piece, bar_full(slot), kMask);
            } else {   // kCluster == 1, or debug: every CTA fetches the whole chunk itself (no multicast)
              bulk_g2s(sbase + kOffRing + slot * kChunkBytes, p.packed + (size_t)c * kChunkBytes, bytes, bar_full(slot));
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 9) {
    // ================================================================ MMA issuer
    uint32_t cnt = 0, g = 0, slab_ph = 0;
    const uint32_t idesc256 = umma_idesc(128, 256), idesc16 = umma_idesc(128, 16);
    constexpr uint32_t kHi = (128u >> 4) | (1u << 14);
    for (int it = 0; it < iters; ++it) {
      const uint32_t e_base = sbase + kOffE + (it & 1) * kEBytes;
      for (int L = 0; L < kNumLayers; ++L, ++g) {
        const Layer ly = kLayers[L];
        const int nch = ly.small ? 1 : ly.a_chunks + ly.e_chunks + ly.bias_chunk;
        const uint32_t buf = g & 1u, d_tmem = tmem_base + buf * 256;
        if (g >= 2) mbar_wait(bar_dfree(buf), ((g >> 1) - 1) & 1u);      // D[buf] drained by stage g-2
        if (L == 0) mbar_wait(bar_eready(it & 1), (it >> 1) & 1u);
        if (ly.reload) mbar_wait(bar_reload, it & 1u);
        const bool fresh = kAFresh[L] != 0;
        for (int c = 0; c < nch; ++c, ++cnt) {
          const uint32_t slot = cnt % kRing, phase = (cnt / kRing) & 1u;
          mbar_wait(bar_full(slot), phase);
          if (fresh) {
            if (ly.small) {
              for (int j = 0; j < 8; ++j) mbar_wait(bar_slab(j), slab_ph);
            } else if (c < ly.a_chunks) {
              mbar_wait(bar_slab(c), slab_ph);
            }
          }
          tc_fence_after();
          const uint32_t wsm = sbase + kOffRing + slot * kChunkBytes;
          if (elect_one_sync()) {
            if (ly.small) {
              uint32_t a_lo = dlo(sbase + kOffA, 2048u);
              uint32_t b_lo = dlo(wsm, 256u);
#pragma unroll
              for (int ks = 0; ks < 16; ++ks) {
                umma_bf16_lohi(d_tmem, a_lo, kHi, b_lo, kHi, idesc16, ks > 0 ? 1u : 0u);
                a_lo += 4096u >> 4;
                b_lo += 512u >> 4;
              }
            } else if (c >= ly.a_chunks + ly.e_chunks) {
              const uint32_t a_lo = dlo(e_base + 6 * 2048, 2048u);
              const uint32_t b_lo = dlo(wsm, 4096u);
              umma_bf16_lohi(d_tmem, a_lo, kHi, b_lo, kHi, idesc256, 1u);
            } else {
              const bool from_e = c >= ly.a_chunks;
              const uint32_t a0 = from_e ? e_base + (c - ly.a_chunks) * 4 * 2048 : sbase + kOffA + c * 4 * 2048;
              const uint32_t a_lo = dlo(a0, 2048u);
              const uint32_t b_lo = dlo(wsm, 4096u);
              umma_bf16_lohi(d_tmem, a_lo, kHi, b_lo, kHi, idesc256, c > 0 ? 1u : 0u);
              umma_bf16_lohi(d_tmem, a_lo + (4096u >> 4), kHi, b_lo + (8192u >> 4), kHi, idesc256, 1u);
            }
            if (c == nch - 1) umma_commit(bar_acc(buf));
            if (kCluster > 1 && p.dbg_drain != 6) umma_commit_multicast(bar_empty(slot), kMask);
            else umma_commit(bar_empty(slot));
          }
          __syncwarp();
        }
        if (fresh) slab_ph ^= 1u;
      }
    }
  } else {
    // ================================================================ encode + drain warps (all eight work on the one tile)
    const int q = warp & 3, h = warp >> 2, row = q * 32 + lane;
    const uint32_t a_smem = sbase + kOffA;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* my_scratch = p.scratch + (size_t)blockIdx.x * kABytes;
    const bool warp_bias = (p.N % 32 == 0);
    bool store_pending = false;
    uint32_t g = 0;
    auto tile_of = [&](int it) { return (long long)blockIdx.x + (long long)it * gridDim.x; };
    auto sample_of = [&](int it, bool& live) {
      const long long s_raw = tile_of(it) * 128 + row;
      live = s_raw < p.S;
      return live ? s_raw : p.S - 1;
    };
    {   // first tile of this CTA
      bool live;
      const long long s0 = sample_of(0, live);
      if (h == 0) encode_sample(p, s0, sbase + kOffE, row);
      fence_proxy_async_smem();
      mbar_arrive(bar_eready(0));
    }
    for (int it = 0; it < iters; ++it) {
      bool live;
      const long long s = sample_of(it, live);
      const long long tile = tile_of(it);
      const bool tile_live = tile < n_tiles;
      float sigma_s = 0.f, rgb_s[3] = {0.f, 0.f, 0.f};
      for (int L = 0; L < kNumLayers; ++L, ++g) {
        const Layer ly = kLayers[L];
        const uint32_t buf = g & 1u;
        const uint32_t tmem_d = tmem_row + buf * 256;
        float4 wb[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        const bool table_bias = ly.epi == EPI_HIDDEN && ly.bias_kind != BIAS_MMA;
        const float* brow = nullptr;
        if (table_bias) {
          brow = ly.bias_kind == BIAS_RAY ? p.raybias + (s / p.N) * 256 : p.imgbias + (s / p.per_image) * 256;
          if (warp_bias) {
            wb[0] = __ldg(reinterpret_cast<const float4*>(brow) + lane);
            wb[1] = __ldg(reinterpret_cast<const float4*>(brow + 128) + lane);
          }
        }
        mbar_wait(bar_acc(buf), (g >> 1) & 1u);
        tc_fence_after();
        if (L == kReloadIssueLayer && threadIdx.x == 0) {
          bulk_wait_all();
          fence_proxy_async_all();
          mbar_expect_tx(bar_reload, kABytes);
          bulk_g2s(a_smem, (p.save && tile_live) ? p.save + ((size_t)tile * kSaveSlots) * kABytes : my_scratch, kABytes, bar_reload);
        }
        if (ly.epi == EPI_HIDDEN) {
          if (store_pending) {
            if (threadIdx.x == 0) bulk_wait_read();
            named_bar_sync(1, 256);
            store_pending = false;
          }
          float* dbg_row = ((L == p.dbg_layer) && live && p.dbg_out) ? p.dbg_out + s * 256 : nullptr;
          uint32_t va[16], vb[16];
          TP_TMEM_LD16(tmem_d + h * 16, va);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            TP_TMEM_WAIT16(va);
            TP_TMEM_LD16(tmem_d + (j + 1) * 32 + h * 16, vb);
            {
              const int col0 = j * 32 + h * 16;
              convert16(va, (table_bias && !warp_bias) ? brow + col0 : nullptr, table_bias && warp_bias, wb, col0,
                        a_smem + (4 * j + 2 * h) * 2048 + row * 16, dbg_row ? dbg_row + col0 : nullptr);
              fence_proxy_async_smem();
              mbar_arrive(bar_slab(j));
            }
            TP_TMEM_WAIT16(vb);
            if (j + 2 < 8) TP_TMEM_LD16(tmem_d + (j + 2) * 32 + h * 16, va);
            {
              const int col0 = (j + 1) * 32 + h * 16;
              convert16(vb, (table_bias && !warp_bias) ? brow + col0 : nullptr, table_bias && warp_bias, wb, col0,
                        a_smem + (4 * (j + 1) + 2 * h) * 2048 + row * 16, dbg_row ? dbg_row + col0 : nullptr);
              fence_proxy_async_smem();
              mbar_arrive(bar_slab(j + 1));
            }
          }
          tc_fence_before();
          mbar_arrive(bar_dfree(buf));
          if (L == kSpillLayer || (p.save && kSaveSlot[L] >= 0)) {
            named_bar_sync(1, 256);
            if (threadIdx.x == 0) {
              uint8_t* dst = (p.save && tile_live) ? p.save + ((size_t)tile * kSaveSlots + kSaveSlot[L]) * kABytes : my_scratch;
              if (L == kSpillLayer || tile_live) {
                bulk_s2g(dst, a_smem, kABytes);
                bulk_commit();
              }
            }
            store_pending = true;
          }
        } else {
          if (h == 0) {
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
                for (int c = 0; c < 3; ++c) *reinterpret_cast<float2*>(p.rgb + s * 6 + c * 2) = make_float2(rgb_s[c], rgb_t[c]);
                *reinterpret_cast<float2*>(p.density + s * 2) = make_float2(sigma_s, sigma_t);
                p.uncert[s] = unc;
              }
            }
          }
          tc_fence_before();
          mbar_arrive(bar_dfree(buf));
        }
        if (L == kEncodeAtStage && it + 1 < iters) {     // encoding tile of the next iteration, while the MMAs of this one run
          bool live_n;
          const long long sn = sample_of(it + 1, live_n);
          if (h == 0) encode_sample(p, sn, sbase + kOffE + ((it + 1) & 1) * kEBytes, row);
          fence_proxy_async_smem();
          mbar_arrive(bar_eready((it + 1) & 1));
        }
      }
    }
    if (threadIdx.x == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();      // nobody exits while a peer may still multicast into its SMEM / barriers
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int kCluster>
int launch(const Params& p, cudaStream_t stream) {
  auto kern = nerf_stl_forward_v2_kernel<kCluster>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  if (kCluster > 1) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return (int)e;
  }
  const long long n_tiles = (p.S + 127) / 128;
  int sms = tp_num_sms();
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int grid = (sms / kCluster) * kCluster;
  if (kCluster > 1) {
    cfg.gridDim = dim3(grid);
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0 &&
        max_clusters * kCluster < grid)
      grid = max_clusters * kCluster;          // persistent kernel: only as many clusters as are co-resident
  }
  long long need = ((n_tiles + kCluster - 1) / kCluster) * kCluster;
  if (need < grid) grid = (int)need;
  cfg.gridDim = dim3(grid);
  const int iters = (int)((n_tiles + grid - 1) / grid);
  e = cudaLaunchKernelEx(&cfg, kern, p, iters);
  if (e != cudaSuccess) return (int)e;
  return tp_launch_status();
}

}  // namespace tc2

// dispatched from tp_tc_nerf_stl_forward (flags bit 7); cluster size from flags bits 8-9: 0 -> 1, 1 -> 2, 2 -> 4
int tp_tc_v2_launch(const tc::Params& p, int flags, cudaStream_t stream) {
  const int code = (flags >> 8) & 3;
  if (code == 0) return tc2::launch<1>(p, stream);
  if (code == 1) return tc2::launch<2>(p, stream);
  return tc2::launch<4>(p, stream);
}

// K3: volume-render compositing, forward and backward.
//
//   static/transient/joint 3-chain form   layers/nerf_static_transient_light.py:168-212
//   single-chain form                     layers/nerf.py:117-136
//
// One warp per ray; lanes own consecutive samples of a 32-sample chunk so every per-sample tensor is
// read/written as contiguous 128 B..768 B warp transactions.  The exclusive-cumsum transmittance is a
// warp shuffle scan with a carry across chunks (any N >= 1).  HBM-bound: 40 B/sample in, 12 B/sample
// + 56 B/ray out (forward); see DESIGN.md for the roofline.
// The backward re-runs the scan and uses total-minus-prefix suffix sums:
//   dL/ds^s_k = T_k e^{-s^s_k} V^ps_k + T_k e^{-s_k} V^p_k + T^s_k e^{-s^s_k} V^qs_k + g_as e^{-s^s_k}
//               - sum_{i>k}(V^ps_i p^s_i + V^pt_i p^t_i + V^p_i p_i) - sum_{i>k} V^qs_i q^s_i      (same for t).
#include "common.cuh"
#include "../../include/texpose_b200.h"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// inclusive scan across the warp
__device__ __forceinline__ float warp_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += n;
  }
  return v;
}
// exclusive prefix given the inclusive scan and the running carry; updates the carry.
__device__ __forceinline__ float excl_with_carry(float incl, float& carry, int lane) {
  float prev = __shfl_up_sync(kFull, incl, 1);
  const float ex = carry + (lane == 0 ? 0.f : prev);
  carry += __shfl_sync(kFull, incl, 31);
  return ex;
}

__device__ __forceinline__ float ray_len(const float* __restrict__ ray, long long r) {
  const float x = ray[r * 3], y = ray[r * 3 + 1], z = ray[r * 3 + 2];
  return sqrtf(x * x + y * y + z * z);
}

// distance of sample i to the next one (1e10 tail) times |ray|
__device__ __forceinline__ float interval(const float* __restrict__ depth, int i, int N, float di, float len) {
  const float gap = (i + 1 < N) ? __fsub_rn(depth[i + 1], di) : 1e10f;
  return __fmul_rn(gap, len);
}

struct StlSample {
  float d, dist, Es, Et, E, T, Ts, Tt;   // E* = exp(-sd*), T* = transmittance before the sample
  float ps, pt, p, qs, qt;               // weights
  float cs[3], ct[3], u;
};

// Loads sample i of ray r and computes its weights; `c*` are the running exclusive sums (carries).
__device__ __forceinline__ StlSample stl_sample(const float* __restrict__ rgb, const float* __restrict__ density,
                                                const float* __restrict__ depth, const float* __restrict__ uncert,
                                                int i, int N, float len, float& cj, float& cs, float& ct, int lane) {
  StlSample s;
  const bool in = i < N;
  float sig_s = 0.f, sig_t = 0.f;
  s.d = 0.f;
  s.u = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) s.cs[c] = s.ct[c] = 0.f;
  s.dist = 0.f;
  if (in) {
    const float2 sg = *reinterpret_cast<const float2*>(density + 2 * (long long)i);
    sig_s = sg.x;
    sig_t = sg.y;
    s.d = depth[i];
    s.u = uncert ? uncert[i] : 0.f;
    s.dist = interval(depth, i, N, s.d, len);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float2 v = *reinterpret_cast<const float2*>(rgb + 6 * (long long)i + 2 * c);
      s.cs[c] = v.x;
      s.ct[c] = v.y;
    }
  }
  const float sd_s = in ? __fmul_rn(sig_s, s.dist) : 0.f;
  const float sd_t = in ? __fmul_rn(sig_t, s.dist) : 0.f;
  const float sd = __fadd_rn(sd_s, sd_t);
  const float ex_j = excl_with_carry(warp_scan(sd, lane), cj, lane);
  const float ex_s = excl_with_carry(warp_scan(sd_s, lane), cs, lane);
  const float ex_t = excl_with_carry(warp_scan(sd_t, lane), ct, lane);
  s.Es = expf(-sd_s);
  s.Et = expf(-sd_t);
  s.E = expf(-sd);
  s.T = expf(-ex_j);
  s.Ts = expf(-ex_s);
  s.Tt = expf(-ex_t);
  const float as = 1.f - s.Es, at = 1.f - s.Et, a = 1.f - s.E;
  s.ps = in ? s.T * as : 0.f;
  s.pt = in ? s.T * at : 0.f;
  s.p = in ? s.T * a : 0.f;
  s.qs = in ? s.Ts * as : 0.f;
  s.qt = in ? s.Tt * at : 0.f;
  return s;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_stl_fwd_kernel(const float* __restrict__ ray, const float* __restrict__ rgb,
                         const float* __restrict__ density, const float* __restrict__ depth,
                         const float* __restrict__ uncert, long long R, int N, float min_uncert,
                         float* __restrict__ o_rgb, float* __restrict__ o_rgb_s, float* __restrict__ o_rgb_t,
                         float* __restrict__ o_depth, float* __restrict__ o_op, float* __restrict__ o_op_s,
                         float* __restrict__ o_op_t, float* __restrict__ o_prob, float* __restrict__ o_unc,
                         float* __restrict__ o_as, float* __restrict__ o_at) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < R; r += nwarps) {
    const float len = ray_len(ray, r);
    const float* rgb_r = rgb + r * N * 6;
    const float* den_r = density + r * N * 2;
    const float* dep_r = depth + r * N;
    const float* unc_r = uncert + r * N;
    float cj = 0.f, cs = 0.f, ct = 0.f;
    float acc[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) acc[k] = 0.f;
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const StlSample s = stl_sample(rgb_r, den_r, dep_r, unc_r, i, N, len, cj, cs, ct, lane);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc[c] += s.cs[c] * s.ps + s.ct[c] * s.pt;
        acc[3 + c] += s.qs * s.cs[c];
        acc[6 + c] += s.qt * s.ct[c];
      }
      acc[9] += s.d * s.qs;
      acc[10] += s.p;
      acc[11] += s.qs;
      acc[12] += s.qt;
      acc[13] += s.u * s.pt;
      if (i < N) {
        o_prob[r * N + i] = s.p;
        o_as[r * N + i] = 1.f - s.Es;
        o_at[r * N + i] = 1.f - s.Et;
      }
    }
#pragma unroll
    for (int k = 0; k < 14; ++k) acc[k] = warp_sum(acc[k]);
    if (lane < 3) {
      o_rgb[r * 3 + lane] = acc[lane];
      o_rgb_s[r * 3 + lane] = acc[3 + lane];
      o_rgb_t[r * 3 + lane] = acc[6 + lane];
    }
    if (lane == 0) {
      o_depth[r] = acc[9];
      o_op[r] = acc[10];
      o_op_s[r] = acc[11];
      o_op_t[r] = acc[12];
      o_unc[r] = acc[13] + min_uncert;
    }
  }
}

// Vectorised forward for N = 32 * K (K = 2, 4: the yaml's 64 / 128 samples): lane l owns the K CONSECUTIVE samples
// l*K .. l*K+K-1, so every per-sample tensor moves as 128-bit (K = 2: 64-bit for the 4-byte streams) loads / stores that a
// warp issues as one contiguous 256 B .. 3 KB block, all of them in flight before the first use; the three exclusive
// cumulative sums are a K-step serial prefix inside the lane plus ONE warp scan of the lane totals per chain (3 scans per
// ray instead of 3 per 32-sample chunk).
template <int K>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_stl_fwd_vec_kernel(const float* __restrict__ ray, const float* __restrict__ rgb,
                             const float* __restrict__ density, const float* __restrict__ depth,
                             const float* __restrict__ uncert, long long R, float min_uncert,
                             float* __restrict__ o_rgb, float* __restrict__ o_rgb_s, float* __restrict__ o_rgb_t,
                             float* __restrict__ o_depth, float* __restrict__ o_op, float* __restrict__ o_op_s,
                             float* __restrict__ o_op_t, float* __restrict__ o_prob, float* __restrict__ o_unc,
                             float* __restrict__ o_as, float* __restrict__ o_at) {
  constexpr int N = 32 * K;
  static_assert(K == 2 || K == 4, "K");
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < R; r += nwarps) {
    const long long s0 = r * N + lane * K;      // first sample of this lane
    float c[6 * K], sg[2 * K], d[K], u[K];
    if (K == 4) {
#pragma unroll
      for (int v = 0; v < 6; ++v) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(rgb + s0 * 6) + v);
        c[4 * v] = q.x; c[4 * v + 1] = q.y; c[4 * v + 2] = q.z; c[4 * v + 3] = q.w;
      }
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(density + s0 * 2) + v);
        sg[4 * v] = q.x; sg[4 * v + 1] = q.y; sg[4 * v + 2] = q.z; sg[4 * v + 3] = q.w;
      }
      const float4 qd = __ldcs(reinterpret_cast<const float4*>(depth + s0));
      d[0] = qd.x; d[1] = qd.y; d[K - 2] = qd.z; d[K - 1] = qd.w;
      const float4 qu = __ldcs(reinterpret_cast<const float4*>(uncert + s0));
      u[0] = qu.x; u[1] = qu.y; u[K - 2] = qu.z; u[K - 1] = qu.w;
    } else {
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(rgb + s0 * 6) + v);
        c[4 * v] = q.x; c[4 * v + 1] = q.y; c[4 * v + 2] = q.z; c[4 * v + 3] = q.w;
      }
      const float4 q = __ldcs(reinterpret_cast<const float4*>(density + s0 * 2));
      sg[0] = q.x; sg[1] = q.y; sg[2] = q.z; sg[3] = q.w;
      const float2 qd = __ldcs(reinterpret_cast<const float2*>(depth + s0));
      d[0] = qd.x; d[1] = qd.y;
      const float2 qu = __ldcs(reinterpret_cast<const float2*>(uncert + s0));
      u[0] = qu.x; u[1] = qu.y;
    }
    const float len = ray_len(ray, r);
    const float d_next = __shfl_down_sync(kFull, d[0], 1);      // first depth of the next lane's run
    float sds[K], sdt[K], sdj[K];
    float tj = 0.f, ts = 0.f, tt = 0.f;      // lane totals
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float nxt = k + 1 < K ? d[k + 1] : d_next;
      const float gap = (k + 1 < K || lane < 31) ? __fsub_rn(nxt, d[k]) : 1e10f;
      const float dist = __fmul_rn(gap, len);
      sds[k] = __fmul_rn(sg[2 * k], dist);
      sdt[k] = __fmul_rn(sg[2 * k + 1], dist);
      sdj[k] = __fadd_rn(sds[k], sdt[k]);
      tj += sdj[k];
      ts += sds[k];
      tt += sdt[k];
    }
    // exclusive prefix of the lane totals
    float ej = __shfl_up_sync(kFull, warp_scan(tj, lane), 1), es = __shfl_up_sync(kFull, warp_scan(ts, lane), 1),
          et = __shfl_up_sync(kFull, warp_scan(tt, lane), 1);
    if (lane == 0) ej = es = et = 0.f;
    float acc[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) acc[k] = 0.f;
    float pr[K], as_[K], at_[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float Es = expf(-sds[k]), Et = expf(-sdt[k]), E = expf(-sdj[k]);
      const float T = expf(-ej), Ts = expf(-es), Tt = expf(-et);
      ej += sdj[k];
      es += sds[k];
      et += sdt[k];
      const float as = 1.f - Es, at = 1.f - Et, a = 1.f - E;
      const float ps = T * as, pt = T * at, p = T * a, qs = Ts * as, qt = Tt * at;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float cs = c[6 * k + 2 * ch], ct = c[6 * k + 2 * ch + 1];
        acc[ch] += cs * ps + ct * pt;
        acc[3 + ch] += qs * cs;
        acc[6 + ch] += qt * ct;
      }
      acc[9] += d[k] * qs;
      acc[10] += p;
      acc[11] += qs;
      acc[12] += qt;
      acc[13] += u[k] * pt;
      pr[k] = p;
      as_[k] = as;
      at_[k] = at;
    }
    if (K == 4) {
      __stcs(reinterpret_cast<float4*>(o_prob + s0), make_float4(pr[0], pr[1], pr[K - 2], pr[K - 1]));
      __stcs(reinterpret_cast<float4*>(o_as + s0), make_float4(as_[0], as_[1], as_[K - 2], as_[K - 1]));
      __stcs(reinterpret_cast<float4*>(o_at + s0), make_float4(at_[0], at_[1], at_[K - 2], at_[K - 1]));
    } else {
      __stcs(reinterpret_cast<float2*>(o_prob + s0), make_float2(pr[0], pr[1]));
      __stcs(reinterpret_cast<float2*>(o_as + s0), make_float2(as_[0], as_[1]));
      __stcs(reinterpret_cast<float2*>(o_at + s0), make_float2(at_[0], at_[1]));
    }
#pragma unroll
    for (int k = 0; k < 14; ++k) acc[k] = warp_sum(acc[k]);
    if (lane < 3) {
      o_rgb[r * 3 + lane] = acc[lane];
      o_rgb_s[r * 3 + lane] = acc[3 + lane];
      o_rgb_t[r * 3 + lane] = acc[6 + lane];
    }
    if (lane == 0) {
      o_depth[r] = acc[9];
      o_op[r] = acc[10];
      o_op_s[r] = acc[11];
      o_op_t[r] = acc[12];
      o_unc[r] = acc[13] + min_uncert;
    }
  }
}

struct StlGrads {   // upstream gradients of one ray (zeros where the caller passed no tensor)
  float rgb[3], rs[3], rt[3], d, o, os, ot, u;
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_stl_bwd_kernel(const float* __restrict__ ray, const float* __restrict__ rgb,
                         const float* __restrict__ density, const float* __restrict__ depth,
                         const float* __restrict__ uncert, long long R, int N,
                         const float* __restrict__ g_rgb, const float* __restrict__ g_rgb_s,
                         const float* __restrict__ g_rgb_t, const float* __restrict__ g_depth,
                         const float* __restrict__ g_op, const float* __restrict__ g_op_s,
                         const float* __restrict__ g_op_t, const float* __restrict__ g_prob,
                         const float* __restrict__ g_unc, const float* __restrict__ g_as,
                         const float* __restrict__ g_at, float* __restrict__ d_rgb, float* __restrict__ d_density,
                         float* __restrict__ d_uncert) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < R; r += nwarps) {
    const float len = ray_len(ray, r);
    const float* rgb_r = rgb + r * N * 6;
    const float* den_r = density + r * N * 2;
    const float* dep_r = depth + r * N;
    const float* unc_r = uncert + r * N;
    StlGrads g;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      g.rgb[c] = g_rgb ? g_rgb[r * 3 + c] : 0.f;
      g.rs[c] = g_rgb_s ? g_rgb_s[r * 3 + c] : 0.f;
      g.rt[c] = g_rgb_t ? g_rgb_t[r * 3 + c] : 0.f;
    }
    g.d = g_depth ? g_depth[r] : 0.f;
    g.o = g_op ? g_op[r] : 0.f;
    g.os = g_op_s ? g_op_s[r] : 0.f;
    g.ot = g_op_t ? g_op_t[r] : 0.f;
    g.u = g_unc ? g_unc[r] : 0.f;

    // pass 1: totals of the three weighted series
    float cj = 0.f, cs = 0.f, ct = 0.f;
    float totJ = 0.f, totS = 0.f, totT = 0.f;
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const StlSample s = stl_sample(rgb_r, den_r, dep_r, unc_r, i, N, len, cj, cs, ct, lane);
      const float gp = (g_prob && i < N) ? g_prob[r * N + i] : 0.f;
      const float Vps = g.rgb[0] * s.cs[0] + g.rgb[1] * s.cs[1] + g.rgb[2] * s.cs[2];
      const float Vpt = g.rgb[0] * s.ct[0] + g.rgb[1] * s.ct[1] + g.rgb[2] * s.ct[2] + g.u * s.u;
      const float Vp = g.o + gp;
      const float Vqs = g.rs[0] * s.cs[0] + g.rs[1] * s.cs[1] + g.rs[2] * s.cs[2] + g.d * s.d + g.os;
      const float Vqt = g.rt[0] * s.ct[0] + g.rt[1] * s.ct[1] + g.rt[2] * s.ct[2] + g.ot;
      totJ += Vps * s.ps + Vpt * s.pt + Vp * s.p;
      totS += Vqs * s.qs;
      totT += Vqt * s.qt;
    }
    totJ = warp_sum(totJ);
    totS = warp_sum(totS);
    totT = warp_sum(totT);

    // pass 2: per-sample gradients with inclusive prefixes of the same series
    cj = cs = ct = 0.f;
    float pj = 0.f, psum = 0.f, ptsum = 0.f;   // carries of the series prefixes
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const StlSample s = stl_sample(rgb_r, den_r, dep_r, unc_r, i, N, len, cj, cs, ct, lane);
      const float gp = (g_prob && i < N) ? g_prob[r * N + i] : 0.f;
      const float Vps = g.rgb[0] * s.cs[0] + g.rgb[1] * s.cs[1] + g.rgb[2] * s.cs[2];
      const float Vpt = g.rgb[0] * s.ct[0] + g.rgb[1] * s.ct[1] + g.rgb[2] * s.ct[2] + g.u * s.u;
      const float Vp = g.o + gp;
      const float Vqs = g.rs[0] * s.cs[0] + g.rs[1] * s.cs[1] + g.rs[2] * s.cs[2] + g.d * s.d + g.os;
      const float Vqt = g.rt[0] * s.ct[0] + g.rt[1] * s.ct[1] + g.rt[2] * s.ct[2] + g.ot;
      const float inJ = warp_scan(Vps * s.ps + Vpt * s.pt + Vp * s.p, lane);
      const float inS = warp_scan(Vqs * s.qs, lane);
      const float inT = warp_scan(Vqt * s.qt, lane);
      // suffix sums over i>k; exactly empty for the last sample (its interval is 1e10*|ray|, so any
      // rounding residue of total-minus-prefix would be amplified)
      const bool tail = (i == N - 1);
      const float sufJ = tail ? 0.f : totJ - (pj + inJ);
      const float sufS = tail ? 0.f : totS - (psum + inS);
      const float sufT = tail ? 0.f : totT - (ptsum + inT);
      pj += __shfl_sync(kFull, inJ, 31);
      psum += __shfl_sync(kFull, inS, 31);
      ptsum += __shfl_sync(kFull, inT, 31);
      if (i < N) {
        const float gas = g_as ? g_as[r * N + i] : 0.f;
        const float gat = g_at ? g_at[r * N + i] : 0.f;
        const float dA = s.T * s.Es * Vps + s.T * s.E * Vp + s.Ts * s.Es * Vqs + gas * s.Es - sufJ - sufS;
        const float dB = s.T * s.Et * Vpt + s.T * s.E * Vp + s.Tt * s.Et * Vqt + gat * s.Et - sufJ - sufT;
        float2 dd;
        dd.x = dA * s.dist;
        dd.y = dB * s.dist;
        *reinterpret_cast<float2*>(d_density + (r * N + i) * 2) = dd;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float2 v;
          v.x = g.rgb[c] * s.ps + g.rs[c] * s.qs;
          v.y = g.rgb[c] * s.pt + g.rt[c] * s.qt;
          *reinterpret_cast<float2*>(d_rgb + (r * N + i) * 6 + 2 * c) = v;
        }
        d_uncert[r * N + i] = g.u * s.pt;
      }
    }
  }
}

// ---------------------------------------------------------------- single-chain (layers/nerf.py)

struct PlainSample {
  float d, dist, E, T, p, c[3];
};
__device__ __forceinline__ PlainSample plain_sample(const float* __restrict__ rgb, const float* __restrict__ density,
                                                    const float* __restrict__ depth, int i, int N, float len,
                                                    float& carry, int lane) {
  PlainSample s;
  const bool in = i < N;
  float sig = 0.f;
  s.d = 0.f;
  s.dist = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) s.c[c] = 0.f;
  if (in) {
    sig = density[i];
    s.d = depth[i];
    s.dist = interval(depth, i, N, s.d, len);
#pragma unroll
    for (int c = 0; c < 3; ++c) s.c[c] = rgb[3 * (long long)i + c];
  }
  const float sd = in ? __fmul_rn(sig, s.dist) : 0.f;
  const float ex = excl_with_carry(warp_scan(sd, lane), carry, lane);
  s.E = expf(-sd);
  s.T = expf(-ex);
  s.p = in ? s.T * (1.f - s.E) : 0.f;
  return s;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_plain_fwd_kernel(const float* __restrict__ ray, const float* __restrict__ rgb,
                           const float* __restrict__ density, const float* __restrict__ depth, long long R, int N,
                           int use_bg, float bgcolor, float* __restrict__ o_rgb, float* __restrict__ o_depth,
                           float* __restrict__ o_op, float* __restrict__ o_prob) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < R; r += nwarps) {
    const float len = ray_len(ray, r);
    float carry = 0.f, acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const PlainSample s = plain_sample(rgb + r * N * 3, density + r * N, depth + r * N, i, N, len, carry, lane);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += s.c[c] * s.p;
      acc[3] += s.d * s.p;
      acc[4] += s.p;
      if (i < N) o_prob[r * N + i] = s.p;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[k] = warp_sum(acc[k]);
    if (lane < 3) o_rgb[r * 3 + lane] = use_bg ? acc[lane] + bgcolor * (1.f - acc[4]) : acc[lane];
    if (lane == 0) {
      o_depth[r] = acc[3];
      o_op[r] = acc[4];
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_plain_bwd_kernel(const float* __restrict__ ray, const float* __restrict__ rgb,
                           const float* __restrict__ density, const float* __restrict__ depth, long long R, int N,
                           int use_bg, float bgcolor, const float* __restrict__ g_rgb,
                           const float* __restrict__ g_depth, const float* __restrict__ g_op,
                           const float* __restrict__ g_prob, float* __restrict__ d_rgb,
                           float* __restrict__ d_density) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < R; r += nwarps) {
    const float len = ray_len(ray, r);
    float gr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) gr[c] = g_rgb ? g_rgb[r * 3 + c] : 0.f;
    const float gd = g_depth ? g_depth[r] : 0.f;
    float go = g_op ? g_op[r] : 0.f;
    if (use_bg) go -= bgcolor * (gr[0] + gr[1] + gr[2]);   // rgb += bg*(1-opacity), layers/nerf.py:134-135
    float carry = 0.f, tot = 0.f;
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const PlainSample s = plain_sample(rgb + r * N * 3, density + r * N, depth + r * N, i, N, len, carry, lane);
      const float gp = (g_prob && i < N) ? g_prob[r * N + i] : 0.f;
      const float V = gr[0] * s.c[0] + gr[1] * s.c[1] + gr[2] * s.c[2] + gd * s.d + go + gp;
      tot += V * s.p;
    }
    tot = warp_sum(tot);
    carry = 0.f;
    float pre = 0.f;
    for (int base = 0; base < N; base += 32) {
      const int i = base + lane;
      const PlainSample s = plain_sample(rgb + r * N * 3, density + r * N, depth + r * N, i, N, len, carry, lane);
      const float gp = (g_prob && i < N) ? g_prob[r * N + i] : 0.f;
      const float V = gr[0] * s.c[0] + gr[1] * s.c[1] + gr[2] * s.c[2] + gd * s.d + go + gp;
      const float inc = warp_scan(V * s.p, lane);
      const float suf = (i == N - 1) ? 0.f : tot - (pre + inc);
      pre += __shfl_sync(kFull, inc, 31);
      if (i < N) {
        const float dS = s.T * s.E * V - suf;
        d_density[r * N + i] = dS * s.dist;
#pragma unroll
        for (int c = 0; c < 3; ++c) d_rgb[(r * N + i) * 3 + c] = gr[c] * s.p;
      }
    }
  }
}

inline int grid_for_rays(long long R) {
  long long need = (R + kWarpsPerBlock - 1) / kWarpsPerBlock;
  long long cap = (long long)tp_num_sms() * 8;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace

TP_API int tp_composite_stl_forward(const float* ray, const float* rgb, const float* density, const float* depth,
                                    const float* uncert, int64_t R, int N, float min_uncert, float* o_rgb,
                                    float* o_rgb_static, float* o_rgb_transient, float* o_depth, float* o_opacity,
                                    float* o_opacity_static, float* o_opacity_transient, float* o_prob,
                                    float* o_uncert, float* o_alpha_static, float* o_alpha_transient, void* stream) {
  if (!ray || !rgb || !density || !depth || !uncert || !o_rgb || !o_rgb_static || !o_rgb_transient || !o_depth ||
      !o_opacity || !o_opacity_static || !o_opacity_transient || !o_prob || !o_uncert || !o_alpha_static ||
      !o_alpha_transient)
    return TP_ERR_BAD_ARG;
  if (R < 0 || N < 1) return TP_ERR_BAD_SHAPE;
  if (R == 0) return TP_OK;
  const bool aligned = (((uintptr_t)rgb | (uintptr_t)density | (uintptr_t)depth | (uintptr_t)uncert | (uintptr_t)o_prob |
                         (uintptr_t)o_alpha_static | (uintptr_t)o_alpha_transient) & 15) == 0;
  if (aligned && N == 128)
    composite_stl_fwd_vec_kernel<4><<<grid_for_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        ray, rgb, density, depth, uncert, R, min_uncert, o_rgb, o_rgb_static, o_rgb_transient, o_depth, o_opacity,
        o_opacity_static, o_opacity_transient, o_prob, o_uncert, o_alpha_static, o_alpha_transient);
  else if (aligned && N == 64)
    composite_stl_fwd_vec_kernel<2><<<grid_for_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        ray, rgb, density, depth, uncert, R, min_uncert, o_rgb, o_rgb_static, o_rgb_transient, o_depth, o_opacity,
        o_opacity_static, o_opacity_transient, o_prob, o_uncert, o_alpha_static, o_alpha_transient);
  else
    composite_stl_fwd_kernel<<<grid_for_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        ray, rgb, density, depth, uncert, R, N, min_uncert, o_rgb, o_rgb_static, o_rgb_transient, o_depth, o_opacity,
        o_opacity_static, o_opacity_transient, o_prob, o_uncert, o_alpha_static, o_alpha_transient);
  return tp_launch_status();
}

TP_API int tp_composite_stl_backward(const float* ray, const float* rgb, const float* density, const float* depth,
                                     const float* uncert, int64_t R, int N, const float* g_rgb,
                                     const float* g_rgb_static, const float* g_rgb_transient, const float* g_depth,
                                     const float* g_opacity, const float* g_opacity_static,
                                     const float* g_opacity_transient, const float* g_prob, const float* g_uncert,
                                     const float* g_alpha_static, const float* g_alpha_transient, float* d_rgb,
                                     float* d_density, float* d_uncert, void* stream) {
  if (!ray || !rgb || !density || !depth || !uncert || !d_rgb || !d_density || !d_uncert) return TP_ERR_BAD_ARG;
  if (R < 0 || N < 1) return TP_ERR_BAD_SHAPE;
  if (R == 0) return TP_OK;
  composite_stl_bwd_kernel<<<grid_for_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      ray, rgb, density, depth, uncert, R, N, g_rgb, g_rgb_static, g_rgb_transient, g_depth, g_opacity,
      g_opacity_static, g_opacity_transient, g_prob, g_uncert, g_alpha_static, g_alpha_transient, d_rgb, d_density,
      d_uncert);
  return tp_launch_status();
}

TP_API int tp_composite_plain_forward(const float* ray, const float* rgb, const float* density, const float* depth,
                                      int64_t R, int N, int use_bg, float bgcolor, float* o_rgb, float* o_depth,
                                      float* o_opacity, float* o_prob, void* stream) {
  if (!ray || !rgb || !density || !depth || !o_rgb || !o_depth || !o_opacity || !o_prob) return TP_ERR_BAD_ARG;
  if (R < 0 || N < 1) return TP_ERR_BAD_SHAPE;
  if (R == 0) return TP_OK;
  composite_plain_fwd_kernel<<<grid_for_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      ray, rgb, density, depth, R, N, use_bg, bgcolor, o_rgb, o_depth, o_opacity, o_prob);
  return tp_launch_status();
}

TP_API int tp_composite_plain_backward(const float* ray, const float* rgb, const float* density, const float* depth,
                                       int64_t R, int N, int use_bg, float bgcolor, const float* g_rgb,
                                       const float* g_depth, const float* g_opacity, const float* g_prob,
                                       float* d_rgb, float* d_density, void* stream) {
  if (!ray || !rgb || !density || !depth || !d_rgb || !d_density) return TP_ERR_BAD_ARG;
  if (R < 0 || N < 1) return TP_ERR_BAD_SHAPE;
  if (R == 0) return TP_OK;
  composite_plain_bwd_kernel<<<grid_for_rays(R), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      ray, rgb, density, depth, R, N, use_bg, bgcolor, g_rgb, g_depth, g_opacity, g_prob, d_rgb, d_density);
  return tp_launch_status();
}

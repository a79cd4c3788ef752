// fp32 SIMT path of the NeRF MLPs (the 1e-4 parity mode) and the generic layer backward.
//
// The reference evaluates every layer as aten::addmm + elementwise kernels and materialises each
// torch.cat (layers/nerf_static_transient_light.py:76-145, layers/nerf.py:61-99).  Here one SGEMM family
// with a *segmented* X operand reads the concatenation in place: a row of X is the concatenation of up
// to 4 segments, each `ptr[(row / group) * ld + col]` -- group==1 for per-sample tensors, N for per-ray
// (view encoding) and R*N for per-image (latent) segments, so broadcast inputs are never expanded.
//   forward      Y = act(X W^T + b)                      M=S,    N=Nout, Kred=Ktot
//   input grad   dX = (dY W) * [Xact > 0]                M=S,    N=K1,   Kred=Nout
//   weight grad  dW = dY^T X, db = colsum(dY)            M=Nout, N=Ktot, Kred=S (split-K, 2-stage reduce)
// All accumulate in fp32 FFMA; fixed reduction order -> deterministic.
#include "common.cuh"
#include "../../include/texpose_b200.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

struct Segs {
  const float* ptr[4];
  long long ld[4];
  long long group[4];
  int begin[5];   // column range of segment i is [begin[i], begin[i+1])
  int n;
  // group == 1 (per-sample tensors, the common case) needs no division; the others divide in 32 bits when the row fits
  __device__ __forceinline__ float at(long long row, int col) const {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < n && col < begin[i + 1]) {
        long long r = row;
        if (group[i] != 1)
          r = (row < 0x7fffffffLL && group[i] < 0x7fffffffLL) ? (long long)((unsigned)row / (unsigned)group[i]) : row / group[i];
        return ptr[i][r * ld[i] + (col - begin[i])];
      }
    return 0.f;
  }
};

// ---- operand functors: element (m,k) of A / (k,n) of B, zero outside the problem
struct A_X {   // forward: A(m,k) = X[m,k]
  Segs x; long long M; int K;
  __device__ __forceinline__ float operator()(long long m, long long k) const { return (m < M && k < K) ? x.at(m, (int)k) : 0.f; }
};
struct B_Wt {  // forward: B(k,n) = W[n,k]
  const float* W; long long ldw; int K, N;
  __device__ __forceinline__ float operator()(long long k, long long n) const { return (k < K && n < N) ? W[n * ldw + k] : 0.f; }
};
struct A_dY {  // input grad: A(m,k) = dY[m,k]
  const float* dY; long long ld; long long M; int K;
  __device__ __forceinline__ float operator()(long long m, long long k) const { return (m < M && k < K) ? dY[m * ld + k] : 0.f; }
};
struct B_W {   // input grad: B(k,n) = W[k,n]
  const float* W; long long ldw; int K, N;
  __device__ __forceinline__ float operator()(long long k, long long n) const { return (k < K && n < N) ? W[k * ldw + n] : 0.f; }
};
struct A_dYt { // weight grad: A(m,k) = dY[k,m]
  const float* dY; long long ld; int M; long long K;
  __device__ __forceinline__ float operator()(long long m, long long k) const { return (m < M && k < K) ? dY[k * ld + m] : 0.f; }
};
struct B_X {   // weight grad: B(k,n) = X[k,n]
  Segs x; long long K; int N;
  __device__ __forceinline__ float operator()(long long k, long long n) const { return (k < K && n < N) ? x.at(k, (int)n) : 0.f; }
};

// ---- epilogues
struct EpiForward {
  const float* bias; int act; float* Y; long long ldy; float* aux0; float* aux1; long long M; int N;
  __device__ __forceinline__ void operator()(long long m, int n, float acc) const {
    if (m >= M || n >= N) return;
    float v = acc + (bias ? bias[n] : 0.f);
    switch (act) {
      case TP_ACT_NONE: Y[m * ldy + n] = v; break;
      case TP_ACT_RELU: Y[m * ldy + n] = fmaxf(v, 0.f); break;
      case TP_ACT_TRUNK_LAST_STL:   // row 0 -> softplus -> density[m,0] of [S,2]; rows 1.. -> relu feature
        if (n == 0) aux0[m * 2] = tp_softplus(v);
        else Y[m * ldy + (n - 1)] = fmaxf(v, 0.f);
        break;
      case TP_ACT_TRUNK_LAST_PLAIN: // same with density [S]
        if (n == 0) aux0[m] = tp_softplus(v);
        else Y[m * ldy + (n - 1)] = fmaxf(v, 0.f);
        break;
      case TP_ACT_RGB_STATIC:     // sigmoid -> rgb[m,n,0] of [S,3,2]
        Y[m * 6 + n * 2] = tp_sigmoid(v); break;
      case TP_ACT_TRANS_OUT:      // rows 0-2 sigmoid -> rgb[m,n,1]; row 3 softplus -> density[m,1]; row 4 -> uncert[m]
        if (n < 3) Y[m * 6 + n * 2 + 1] = tp_sigmoid(v);
        else if (n == 3) aux0[m * 2 + 1] = tp_softplus(v);
        else aux1[m] = tp_softplus(v);
        break;
      case TP_ACT_SIGMOID: Y[m * ldy + n] = tp_sigmoid(v); break;
    }
  }
};
struct EpiInputGrad {
  const float* xact; long long ldx; float* dX; long long lddx; long long M; int N;
  __device__ __forceinline__ void operator()(long long m, int n, float acc) const {
    if (m >= M || n >= N) return;
    if (xact && !(xact[m * ldx + n] > 0.f)) acc = 0.f;
    dX[m * lddx + n] = acc;
  }
};
struct EpiPartial {
  float* out; int M, N;   // out[split][M][N]
  __device__ __forceinline__ void operator()(long long m, int n, float acc) const {
    if (m >= M || n >= N) return;
    out[((long long)blockIdx.z * M + m) * N + n] = acc;
  }
};

// C[M,N] tile kernel: 256 threads, 128x128x16 tiles, 8x8 register micro-tile split 4+4 in both dims.  The operands of
// k-tile i+1 are fetched into registers before the FMAs of k-tile i and stored to shared memory after them (software
// pipelining: one global-memory round trip per k-tile is hidden behind 1024 FMAs per thread).
// A_KFAST / B_KFAST select which tile dimension consecutive threads walk so that global reads are coalesced.
// The k order of every accumulator is ascending and independent of the tile size -> results are bit-identical to a plain loop.
template <bool A_KFAST, bool B_KFAST, class FA, class FB, class Epi>
__global__ void __launch_bounds__(NT) sgemm_kernel(FA fa, FB fb, Epi epi, long long Kred, long long k_per_split) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  constexpr int kLoadsA = (BM * BK) / NT, kLoadsB = (BN * BK) / NT;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM, n0 = (long long)blockIdx.y * BN;
  const long long kb = (long long)blockIdx.z * k_per_split;
  const long long ke = (kb + k_per_split < Kred) ? kb + k_per_split : Kred;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[kLoadsA], rb[kLoadsB];
  auto fetch = [&](long long k0) {
#pragma unroll
    for (int i = 0; i < kLoadsA; ++i) {
      int mm, kk;
      if (A_KFAST) { kk = tid & (BK - 1); mm = (tid / BK) + i * (NT / BK); }
      else { mm = tid & (BM - 1); kk = (tid / BM) + i * (NT / BM); }
      const long long k = k0 + kk;
      ra[i] = (k < ke) ? fa(m0 + mm, k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kLoadsB; ++i) {
      int nn, kk;
      if (B_KFAST) { kk = tid & (BK - 1); nn = (tid / BK) + i * (NT / BK); }
      else { nn = tid & (BN - 1); kk = (tid / BN) + i * (NT / BN); }
      const long long k = k0 + kk;
      rb[i] = (k < ke) ? fb(k, n0 + nn) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < kLoadsA; ++i) {
      int mm, kk;
      if (A_KFAST) { kk = tid & (BK - 1); mm = (tid / BK) + i * (NT / BK); }
      else { mm = tid & (BM - 1); kk = (tid / BM) + i * (NT / BM); }
      As[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < kLoadsB; ++i) {
      int nn, kk;
      if (B_KFAST) { kk = tid & (BK - 1); nn = (tid / BK) + i * (NT / BK); }
      else { nn = tid & (BN - 1); kk = (tid / BN) + i * (NT / BN); }
      Bs[kk][nn] = rb[i];
    }
  };

  if (kb < ke) fetch(kb);
  for (long long k0 = kb; k0 < ke; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < ke) fetch(k0 + BK);          // in flight during the FMAs below
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = (int)n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      epi(m, n, acc[i][j]);
    }
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, int splits, long long MN, float* __restrict__ out,
                                       int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < MN; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(long long)z * MN + i];
    out[i] = accumulate ? out[i] + s : s;
  }
}

// column sums of dY grouped by row blocks of `group` rows: out[g][n] = sum_{r in group g} dY[r][n].
// One block per (group, row-slice); fixed-order two-stage reduction.
__global__ void group_colsum_kernel(const float* __restrict__ dY, long long ld, long long S, long long group, int N,
                                    int slices, float* __restrict__ part) {
  const long long g = blockIdx.x;
  const int sl = blockIdx.y;
  const long long r0 = g * group, r1 = (r0 + group < S) ? r0 + group : S;
  const long long per = (group + slices - 1) / slices;
  const long long a = r0 + sl * per, b = (a + per < r1) ? a + per : r1;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float s = 0.f;
    for (long long r = a; r < b; ++r) s += dY[r * ld + n];
    part[((long long)sl * gridDim.x + g) * N + n] = s;
  }
}

// ---- encodings -------------------------------------------------------------------------------

// enc row = [x, y, z, per coord: sin(f0..f_{L-1}), cos(f0..f_{L-1})], f_k = 2^k * fl32(pi), s = fl32(x*f_k)
// (layers/nerf_static_transient_light.py:217-223).  Accurate sinf/cosf: arguments reach ~1.6e3 rad.
__device__ __forceinline__ void write_enc(const float v[3], int L, float* __restrict__ out) {
  const float pi = 3.14159265358979323846f;
  out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < L; ++k) {
      const float s = __fmul_rn(v[c], ldexpf(pi, k));
      out[3 + c * 2 * L + k] = sinf(s);
      out[3 + c * 2 * L + L + k] = cosf(s);
    }
}

__global__ void points_enc_kernel(const float* __restrict__ center, const float* __restrict__ ray,
                                  const float* __restrict__ depth, long long S, int N, int L, float* __restrict__ enc,
                                  long long ld) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
    const long long r = s / N;
    const float d = depth[s];
    float x[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) x[j] = __fadd_rn(center[r * 3 + j], __fmul_rn(ray[r * 3 + j], d));  // camera.py:321
    write_enc(x, L, enc + s * ld);
  }
}

__global__ void raw_enc_kernel(const float* __restrict__ pts, long long S, int L, float* __restrict__ enc, long long ld) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
    const float x[3] = {pts[s * 3], pts[s * 3 + 1], pts[s * 3 + 2]};
    write_enc(x, L, enc + s * ld);
  }
}

// F.normalize(ray) (eps 1e-12) then its encoding, once per ray (the reference recomputes it per sample,
// layers/nerf_static_transient_light.py:106-108,155-157).
__global__ void view_enc_kernel(const float* __restrict__ ray, long long R, int L, int normalize, float* __restrict__ enc,
                                long long ld) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
    float v[3] = {ray[r * 3], ray[r * 3 + 1], ray[r * 3 + 2]};
    if (normalize) {
      const float len = fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12f);
      v[0] /= len; v[1] /= len; v[2] /= len;
    }
    write_enc(v, L, enc + r * ld);
  }
}

// gradient w.r.t. the pre-activations of the output layers, from the stored outputs:
// sigmoid' = y(1-y); softplus'(z) = 1 - exp(-y).
__global__ void stl_out_grad_kernel(const float* __restrict__ rgb, const float* __restrict__ density,
                                    const float* __restrict__ uncert, const float* __restrict__ g_rgb,
                                    const float* __restrict__ g_density, const float* __restrict__ g_uncert, long long S,
                                    float* __restrict__ dz_rgb, float* __restrict__ dz_trans, float* __restrict__ dz_sigma) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ys = rgb[s * 6 + c * 2], yt = rgb[s * 6 + c * 2 + 1];
      const float gs = g_rgb ? g_rgb[s * 6 + c * 2] : 0.f, gt = g_rgb ? g_rgb[s * 6 + c * 2 + 1] : 0.f;
      dz_rgb[s * 3 + c] = gs * ys * (1.f - ys);
      dz_trans[s * 5 + c] = gt * yt * (1.f - yt);
    }
    const float gd = g_density ? g_density[s * 2 + 1] : 0.f;
    dz_trans[s * 5 + 3] = gd * (1.f - expf(-density[s * 2 + 1]));
    const float gu = g_uncert ? g_uncert[s] : 0.f;
    dz_trans[s * 5 + 4] = gu * (1.f - expf(-uncert[s]));
    if (dz_sigma) dz_sigma[s] = (g_density ? g_density[s * 2] : 0.f) * (1.f - expf(-density[s * 2]));
  }
}

__global__ void plain_out_grad_kernel(const float* __restrict__ rgb, const float* __restrict__ density,
                                      const float* __restrict__ g_rgb, const float* __restrict__ g_density, long long S,
                                      float* __restrict__ dz_rgb, float* __restrict__ dz_sigma) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float y = rgb[s * 3 + c];
      dz_rgb[s * 3 + c] = (g_rgb ? g_rgb[s * 3 + c] : 0.f) * y * (1.f - y);
    }
    dz_sigma[s] = (g_density ? g_density[s] : 0.f) * (1.f - expf(-density[s]));
  }
}

// dZ of the trunk's last layer (257 rows): row 0 = density pre-activation grad, rows 1.. = relu-masked
// feature grad (feat > 0)
__global__ void trunk_last_grad_kernel(const float* __restrict__ dz_sigma, const float* __restrict__ d_feat, long long ldd,
                                       const float* __restrict__ feat, long long ldf, long long S, int F,
                                       float* __restrict__ dz, long long ldz) {
  const long long total = S * (F + 1);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / (F + 1);
    const int n = (int)(i - s * (F + 1));
    float v;
    if (n == 0) v = dz_sigma[s];
    else v = feat[s * ldf + n - 1] > 0.f ? d_feat[s * ldd + n - 1] : 0.f;
    dz[s * ldz + n] = v;
  }
}

// split-K factor of the weight-gradient GEMM: >= 256 rows per split, at most 64 splits
inline int64_t dw_splits(int64_t S) {
  int64_t splits = (S + 255) / 256;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  return splits;
}

bool make_segs(Segs& sg, const float* const* ptr, const int64_t* ld, const int64_t* group, const int32_t* cols, int nseg) {
  if (nseg < 1 || nseg > 4 || !ptr || !ld || !group || !cols) return false;
  sg.n = nseg;
  sg.begin[0] = 0;
  for (int i = 0; i < 4; ++i) {
    if (i < nseg) {
      if (!ptr[i] || group[i] < 1 || cols[i] < 1) return false;
      sg.ptr[i] = ptr[i]; sg.ld[i] = ld[i]; sg.group[i] = group[i];
      sg.begin[i + 1] = sg.begin[i] + cols[i];
    } else {
      sg.ptr[i] = nullptr; sg.ld[i] = 0; sg.group[i] = 1; sg.begin[i + 1] = sg.begin[i];
    }
  }
  return true;
}

}  // namespace

TP_API int tp_linear_forward(const float* const* seg_ptr, const int64_t* seg_ld, const int64_t* seg_group,
                             const int32_t* seg_cols, int nseg, const float* W, int64_t ldw, const float* bias,
                             int64_t S, int Nout, int act, float* Y, int64_t ldy, float* aux0, float* aux1,
                             void* stream) {
  Segs sg;
  if (!make_segs(sg, seg_ptr, seg_ld, seg_group, seg_cols, nseg) || !W) return TP_ERR_BAD_ARG;
  if (S < 0 || Nout < 1) return TP_ERR_BAD_SHAPE;
  if (act < TP_ACT_NONE || act > TP_ACT_SIGMOID) return TP_ERR_BAD_ARG;
  if (!Y) return TP_ERR_BAD_ARG;
  if ((act == TP_ACT_TRUNK_LAST_STL || act == TP_ACT_TRUNK_LAST_PLAIN) && !aux0) return TP_ERR_BAD_ARG;
  if (act == TP_ACT_TRANS_OUT && (!aux0 || !aux1 || Nout != 5)) return TP_ERR_BAD_ARG;
  if (act == TP_ACT_RGB_STATIC && Nout != 3) return TP_ERR_BAD_ARG;
  if (S == 0) return TP_OK;
  const int K = sg.begin[nseg];
  A_X fa{sg, S, K};
  B_Wt fb{W, ldw, K, Nout};
  EpiForward epi{bias, act, Y, ldy, aux0, aux1, S, Nout};
  dim3 grid((unsigned)((S + BM - 1) / BM), (unsigned)((Nout + BN - 1) / BN), 1);
  sgemm_kernel<true, true><<<grid, NT, 0, (cudaStream_t)stream>>>(fa, fb, epi, (long long)K, (long long)K);
  return tp_launch_status();
}

TP_API int tp_linear_backward_input(const float* dY, int64_t lddy, const float* W, int64_t ldw, int64_t S, int Nout,
                                    int K1, const float* xact, int64_t ldx, float* dX, int64_t lddx, void* stream) {
  if (!dY || !W || !dX) return TP_ERR_BAD_ARG;
  if (S < 0 || Nout < 1 || K1 < 1) return TP_ERR_BAD_SHAPE;
  if (S == 0) return TP_OK;
  A_dY fa{dY, lddy, S, Nout};
  B_W fb{W, ldw, Nout, K1};
  EpiInputGrad epi{xact, ldx, dX, lddx, S, K1};
  dim3 grid((unsigned)((S + BM - 1) / BM), (unsigned)((K1 + BN - 1) / BN), 1);
  sgemm_kernel<true, false><<<grid, NT, 0, (cudaStream_t)stream>>>(fa, fb, epi, (long long)Nout, (long long)Nout);
  return tp_launch_status();
}

TP_API int64_t tp_linear_backward_weight_workspace(int64_t S, int Nout, int Ktot) {
  // floats: split-K partials of dW plus the column-sum partials of db
  return dw_splits(S) * ((int64_t)Nout * Ktot + Nout);
}

TP_API int tp_linear_backward_weight(const float* dY, int64_t lddy, const float* const* seg_ptr, const int64_t* seg_ld,
                                     const int64_t* seg_group, const int32_t* seg_cols, int nseg, int64_t S, int Nout,
                                     float* dW, float* db, int accumulate, float* workspace, int64_t workspace_floats,
                                     void* stream) {
  Segs sg;
  if (!make_segs(sg, seg_ptr, seg_ld, seg_group, seg_cols, nseg) || !dY || !dW || !workspace) return TP_ERR_BAD_ARG;
  if (S < 1 || Nout < 1) return TP_ERR_BAD_SHAPE;
  const int K = sg.begin[nseg];
  if (workspace_floats < tp_linear_backward_weight_workspace(S, Nout, K)) return TP_ERR_WORKSPACE;
  const int64_t splits = dw_splits(S);
  long long kps = (S + splits - 1) / splits;
  kps = (kps + BK - 1) / BK * BK;
  A_dYt fa{dY, lddy, Nout, S};
  B_X fb{sg, S, K};
  EpiPartial epi{workspace, Nout, K};
  dim3 grid((unsigned)((Nout + BM - 1) / BM), (unsigned)((K + BN - 1) / BN), (unsigned)splits);
  cudaStream_t st = (cudaStream_t)stream;
  sgemm_kernel<false, false><<<grid, NT, 0, st>>>(fa, fb, epi, (long long)S, kps);
  const long long MN = (long long)Nout * K;
  reduce_partials_kernel<<<tp_grid_for(MN, 256, 4), 256, 0, st>>>(workspace, (int)splits, MN, dW, accumulate);
  if (db) {
    float* part = workspace + splits * MN;
    group_colsum_kernel<<<dim3(1, (unsigned)splits), 256, 0, st>>>(dY, lddy, S, S, Nout, (int)splits, part);
    reduce_partials_kernel<<<1, 256, 0, st>>>(part, (int)splits, Nout, db, accumulate);
  }
  return tp_launch_status();
}

TP_API int tp_reduce_partials(const float* partial, int splits, int64_t count, float* out, int accumulate, void* stream) {
  if (!partial || !out) return TP_ERR_BAD_ARG;
  if (splits < 1 || count < 1) return TP_ERR_BAD_SHAPE;
  reduce_partials_kernel<<<tp_grid_for(count, 256, 4), 256, 0, (cudaStream_t)stream>>>(partial, splits, count, out, accumulate);
  return tp_launch_status();
}

TP_API int tp_group_colsum(const float* dY, int64_t lddy, int64_t S, int64_t group, int Nout, float* out,
                           float* workspace, int64_t workspace_floats, void* stream) {
  if (!dY || !out || !workspace) return TP_ERR_BAD_ARG;
  if (S < 1 || group < 1 || Nout < 1) return TP_ERR_BAD_SHAPE;
  const long long G = (S + group - 1) / group;
  int slices = (int)((group + 511) / 512);
  if (slices > 256) slices = 256;
  if (workspace_floats < (long long)slices * G * Nout) return TP_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  group_colsum_kernel<<<dim3((unsigned)G, (unsigned)slices), 256, 0, st>>>(dY, lddy, S, group, Nout, slices, workspace);
  reduce_partials_kernel<<<tp_grid_for(G * Nout, 256, 4), 256, 0, st>>>(workspace, slices, G * Nout, out, 0);
  return tp_launch_status();
}

TP_API int tp_points_encode(const float* center, const float* ray, const float* depth, int64_t S, int N, int L,
                            float* enc, int64_t ld, void* stream) {
  if (!center || !ray || !depth || !enc) return TP_ERR_BAD_ARG;
  if (S < 0 || N < 1 || L < 0 || ld < 3 + 6 * L) return TP_ERR_BAD_SHAPE;
  if (S == 0) return TP_OK;
  points_enc_kernel<<<tp_grid_for(S, 128, 16), 128, 0, (cudaStream_t)stream>>>(center, ray, depth, S, N, L, enc, ld);
  return tp_launch_status();
}

TP_API int tp_positional_encode(const float* x, int64_t S, int L, float* enc, int64_t ld, void* stream) {
  if (!x || !enc) return TP_ERR_BAD_ARG;
  if (S < 0 || L < 0 || ld < 3 + 6 * L) return TP_ERR_BAD_SHAPE;
  if (S == 0) return TP_OK;
  raw_enc_kernel<<<tp_grid_for(S, 128, 16), 128, 0, (cudaStream_t)stream>>>(x, S, L, enc, ld);
  return tp_launch_status();
}

TP_API int tp_view_encode(const float* ray, int64_t R, int L, int normalize, float* enc, int64_t ld, void* stream) {
  if (!ray || !enc) return TP_ERR_BAD_ARG;
  if (R < 0 || L < 0 || ld < 3 + 6 * L) return TP_ERR_BAD_SHAPE;
  if (R == 0) return TP_OK;
  view_enc_kernel<<<tp_grid_for(R, 128, 16), 128, 0, (cudaStream_t)stream>>>(ray, R, L, normalize, enc, ld);
  return tp_launch_status();
}

TP_API int tp_stl_output_grad(const float* rgb, const float* density, const float* uncert, const float* g_rgb,
                              const float* g_density, const float* g_uncert, int64_t S, float* dz_rgb, float* dz_trans,
                              float* dz_sigma, void* stream) {
  if (!rgb || !density || !uncert || !dz_rgb || !dz_trans) return TP_ERR_BAD_ARG;
  if (S < 0) return TP_ERR_BAD_SHAPE;
  if (S == 0) return TP_OK;
  stl_out_grad_kernel<<<tp_grid_for(S, 256, 8), 256, 0, (cudaStream_t)stream>>>(rgb, density, uncert, g_rgb, g_density,
                                                                               g_uncert, S, dz_rgb, dz_trans, dz_sigma);
  return tp_launch_status();
}

TP_API int tp_plain_output_grad(const float* rgb, const float* density, const float* g_rgb, const float* g_density,
                                int64_t S, float* dz_rgb, float* dz_sigma, void* stream) {
  if (!rgb || !density || !dz_rgb || !dz_sigma) return TP_ERR_BAD_ARG;
  if (S < 0) return TP_ERR_BAD_SHAPE;
  if (S == 0) return TP_OK;
  plain_out_grad_kernel<<<tp_grid_for(S, 256, 8), 256, 0, (cudaStream_t)stream>>>(rgb, density, g_rgb, g_density, S,
                                                                                 dz_rgb, dz_sigma);
  return tp_launch_status();
}

TP_API int tp_trunk_last_grad(const float* dz_sigma, const float* d_feat, int64_t ldd, const float* feat, int64_t ldf,
                              int64_t S, int F, float* dz, int64_t ldz, void* stream) {
  if (!dz_sigma || !d_feat || !feat || !dz) return TP_ERR_BAD_ARG;
  if (S < 0 || F < 1 || ldz < F + 1) return TP_ERR_BAD_SHAPE;
  if (S == 0) return TP_OK;
  trunk_last_grad_kernel<<<tp_grid_for(S * (F + 1), 256, 8), 256, 0, (cudaStream_t)stream>>>(dz_sigma, d_feat, ldd, feat,
                                                                                          ldf, S, F, dz, ldz);
  return tp_launch_status();
}

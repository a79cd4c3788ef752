// K1: pixel->ray generation, AABB slab test, stratified sample placement, patch (grid_sample)
// rays/bounds, ray gather, fused on-the-fly box bounds, normals-from-depth.
//
// All of these are HBM-bound elementwise/stencil kernels: grid-stride loops, grids sized as a
// multiple of the SM count, coalesced (SoA-per-thread) accesses.  Arithmetic follows the reference
// operation by operation with explicit rounding intrinsics (no FMA contraction where the reference
// has none) so ray/AABB/sample results are bit-exact:
//   camera.py:292-314 (get_center_and_ray), camera.py:415-433 (aabb_ray_intersection),
//   tools/ray_sampler.py:23-69, model/nerf_adapt_st_gan.py:682-710, compute_surfelinfo.py:37-55,
//   compute_box.py:266-271 + data/lm.py:349-356.
#include "common.cuh"
#include "bilinear.cuh"
#include "rays.cuh"
#include "../../include/texpose_b200.h"

namespace {

constexpr int kThreads = 256;

__global__ void raygen_kernel(const float* __restrict__ kinv, const float* __restrict__ pinv, int B, int H, int W,
                              float pix_offset, const long long* __restrict__ ray_idx, int R,
                              float* __restrict__ center, float* __restrict__ ray) {
  const long long total = (long long)B * R;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / R);
    const long long p = ray_idx ? ray_idx[i] : (i - (long long)b * R);
    const float u = __fadd_rn((float)(p % W), pix_offset);
    const float v = __fadd_rn((float)(p / W), pix_offset);
    float c[3], d[3];
    unproject(kinv + b * 9, pinv + b * 12, u, v, c, d);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      center[i * 3 + j] = c[j];
      ray[i * 3 + j] = d[j];
    }
  }
}

__global__ void patch_rays_kernel(const float* __restrict__ kinv, const float* __restrict__ pinv,
                                  const float* __restrict__ coords, int B, int P, int H, int W,
                                  float* __restrict__ center, float* __restrict__ ray) {
  const long long total = (long long)B * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / P);
    const Bilin s = bilin_setup(coords[i * 2 + 0], coords[i * 2 + 1], H, W);
    // bilinear lookup of the integer index ramps X[y][x]=x, Y[y][x]=y (tools/ray_sampler.py:48-56)
    const float u = bilin_apply(s, H, W, [](int, int x) { return (float)x; });
    const float v = bilin_apply(s, H, W, [](int y, int) { return (float)y; });
    float c[3], d[3];
    unproject(kinv + b * 9, pinv + b * 12, u, v, c, d);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      center[i * 3 + j] = c[j];
      ray[i * 3 + j] = d[j];
    }
  }
}

__global__ void grid_sample_kernel(const float* __restrict__ img, const float* __restrict__ coords, int B, int C,
                                   int H, int W, int P, float* __restrict__ out) {
  const long long total = (long long)B * C * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % P);
    const int c = (int)((i / P) % C);
    const int b = (int)(i / ((long long)P * C));
    const float* g = coords + ((long long)b * P + p) * 2;
    const Bilin s = bilin_setup(g[0], g[1], H, W);
    const float* plane = img + ((long long)b * C + c) * H * W;
    out[i] = bilin_apply(s, H, W, [&](int y, int x) { return plane[(long long)y * W + x]; });
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, int B,
                                   long long HW, int C, int R, float* __restrict__ out) {
  const long long total = (long long)B * R * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long br = i / C;
    const int b = (int)(br / R);
    out[i] = src[((long long)b * HW + idx[br]) * C + c];
  }
}

// torch.minimum/maximum and max/min reductions propagate NaN (0*inf for axis-parallel rays).
__device__ __forceinline__ float nan_min(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }

__device__ __forceinline__ void slab(const float lo[3], const float hi[3], const float o[3], const float d[3],
                                     float& t_near, float& t_far, bool& valid) {
  float t0[3], t1[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float inv = __frcp_rn(d[j]);
    const float a = __fmul_rn(__fsub_rn(lo[j], o[j]), inv);
    const float b = __fmul_rn(__fsub_rn(hi[j], o[j]), inv);
    t0[j] = nan_min(a, b);
    t1[j] = nan_max(a, b);
  }
  t_near = nan_max(nan_max(t0[0], t0[1]), t0[2]);
  t_far = nan_min(nan_min(t1[0], t1[1]), t1[2]);
  valid = (t_far > 0.f) && (t_far > t_near);
}

__global__ void aabb_kernel(const float* __restrict__ amin, const float* __restrict__ amax, int aabb_stride,
                            const float* __restrict__ ro, const float* __restrict__ rd, long long n_per_batch,
                            long long total, float* __restrict__ t_near, float* __restrict__ t_far,
                            unsigned char* __restrict__ valid) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / n_per_batch;
    float lo[3], hi[3], o[3], d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      lo[j] = amin[b * aabb_stride + j];
      hi[j] = amax[b * aabb_stride + j];
      o[j] = ro[i * 3 + j];
      d[j] = rd[i * 3 + j];
    }
    float tn, tf;
    bool ok;
    slab(lo, hi, o, d, tn, tf, ok);
    t_near[i] = tn;
    t_far[i] = tf;
    valid[i] = ok ? 1 : 0;
  }
}

// d_i = (u_i + i) / N * (far - near) + near, one rounding per reference op (model/nerf_adapt_st_gan.py:690-697).
// IEEE division as torch's CPU kernel does (the pinned oracle); torch's CUDA kernel multiplies by fl(1/N) instead,
// identical for the power-of-two N the yamls use (64, 128) and 1 ulp apart otherwise.
__global__ void sample_depth_kernel(const float* __restrict__ z_near, const float* __restrict__ z_far,
                                    const float* __restrict__ rand, long long n_rays, int N, int mode,
                                    float* __restrict__ out) {
  const long long total = n_rays * N;
  const float fn = (float)N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N;
    const int k = (int)(i - r * N);
    out[i] = stratified_depth(mode == 0 ? rand[i] : 0.5f, k, fn, z_near[r], z_far[r]);
  }
}

// In-kernel jitter: one Philox4x32-10 block per 4 consecutive samples (counter = sample index / 4), 16 B stores.
__global__ void sample_depth_philox_kernel(const float* __restrict__ z_near, const float* __restrict__ z_far,
                                           long long n_rays, int N, unsigned long long seed, float* __restrict__ out) {
  const long long total = n_rays * N, quads = (total + 3) >> 2;
  const float fn = (float)N;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const uint4 x = tp_philox((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
    float d[4];
    if ((N & 3) == 0) {      // the four samples share the ray: one 64-bit division and one bounds fetch per quad
      const long long r = (q * 4) / N;
      const int k0 = (int)(q * 4 - r * N);
      const float lo = z_near[r], hi = z_far[r];
#pragma unroll
      for (int e = 0; e < 4; ++e) d[e] = stratified_depth(tp_u01(w[e]), k0 + e, fn, lo, hi);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long i = q * 4 + e;
        if (i < total) {
          const long long r = i / N;
          d[e] = stratified_depth(tp_u01(w[e]), (int)(i - r * N), fn, z_near[r], z_far[r]);
        } else d[e] = 0.f;
      }
    }
    if (q * 4 + 3 < total) *reinterpret_cast<float4*>(out + q * 4) = make_float4(d[0], d[1], d[2], d[3]);
    else
      for (int e = 0; e < 4 && q * 4 + e < total; ++e) out[q * 4 + e] = d[e];
  }
}

// Fused: full-frame rays -> slab test -> zero the misses -> zeros become the background range.
__global__ void box_range_kernel(const float* __restrict__ kinv, const float* __restrict__ pinv, int B, int H, int W,
                                 const float* __restrict__ amin, const float* __restrict__ amax, int aabb_stride,
                                 float bg_near, float bg_far, float* __restrict__ z_near, float* __restrict__ z_far,
                                 unsigned char* __restrict__ valid) {
  const long long HW = (long long)H * W, total = HW * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const long long p = i - (long long)b * HW;
    float c[3], d[3], lo[3], hi[3];
    unproject(kinv + b * 9, pinv + b * 12, __fadd_rn((float)(p % W), 0.5f), __fadd_rn((float)(p / W), 0.5f), c, d);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      lo[j] = amin[b * aabb_stride + j];
      hi[j] = amax[b * aabb_stride + j];
    }
    float tn, tf;
    bool ok;
    slab(lo, hi, c, d, tn, tf, ok);
    tn = ok ? tn : 0.f;
    tf = ok ? tf : 0.f;
    z_near[i] = tn > 0.f ? tn : bg_near;
    z_far[i] = tf > 0.f ? tf : bg_far;
    if (valid) valid[i] = ok ? 1 : 0;
  }
}

__global__ void guided_range_kernel(const float* __restrict__ depth, long long n, float bg_near, float bg_far,
                                    float* __restrict__ z_near, float* __restrict__ z_far) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float zn = __fmul_rn(depth[i], 0.8f), zf = __fmul_rn(depth[i], 1.2f);
    z_near[i] = zn > 0.f ? zn : bg_near;
    z_far[i] = zf > 0.f ? zf : bg_far;
  }
}

// compute_surfelinfo.py:37-55.  One thread per pixel; the 4 neighbour points are recomputed from the
// depth map (4 B in, 12 B out per pixel; neighbour depth reads hit L1/L2).
__device__ __forceinline__ void backproject(const float* kinv, const float* pinv, const float* depth, int W, int x, int y,
                                            float P[3]) {
  float c[3], d[3];
  unproject(kinv, pinv, __fadd_rn((float)x, 0.5f), __fadd_rn((float)y, 0.5f), c, d);
  const float z = depth[(long long)y * W + x];
#pragma unroll
  for (int j = 0; j < 3; ++j) P[j] = __fadd_rn(c[j], __fmul_rn(d[j], z));
}

__global__ void normal_kernel(const float* __restrict__ kinv, const float* __restrict__ pinv,
                              const float* __restrict__ depth, int B, int H, int W, float* __restrict__ out) {
  const long long HW = (long long)H * W, total = HW * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const long long p = i - (long long)b * HW;
    const int x = (int)(p % W), y = (int)(p / W);
    const float* dm = depth + (long long)b * HW;
    float n[3] = {0.f, 0.f, 0.f};
    if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
      float pe[3], pw[3], ps[3], pn[3], tu[3], tv[3];
      backproject(kinv + b * 9, pinv + b * 12, dm, W, x + 1, y, pe);
      backproject(kinv + b * 9, pinv + b * 12, dm, W, x - 1, y, pw);
      backproject(kinv + b * 9, pinv + b * 12, dm, W, x, y + 1, ps);
      backproject(kinv + b * 9, pinv + b * 12, dm, W, x, y - 1, pn);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        tu[j] = __fsub_rn(pe[j], pw[j]);
        tv[j] = __fsub_rn(ps[j], pn[j]);
      }
      n[0] = __fsub_rn(__fmul_rn(tu[1], tv[2]), __fmul_rn(tu[2], tv[1]));
      n[1] = __fsub_rn(__fmul_rn(tu[2], tv[0]), __fmul_rn(tu[0], tv[2]));
      n[2] = __fsub_rn(__fmul_rn(tu[0], tv[1]), __fmul_rn(tu[1], tv[0]));
      const float len = fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-12f);  // F.normalize eps
      n[0] /= len;
      n[1] /= len;
      n[2] = -(n[2] / len);
    }
    const float m = dm[p] > 0.f ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) out[((long long)b * 3 + j) * HW + p] = n[j] * m;
  }
}

}  // namespace

#define TP_CHECK(cond, code) \
  do {                       \
    if (!(cond)) return (code); \
  } while (0)

// Per-view constants of every ray kernel in ONE launch: K^-1 and pose^-1 (camera.py:267 `cam_intr.inverse()`, camera.py:38-44
// Pose.invert).  torch's GPU path spends ~20 tiny launches per call on them (batched LU + two triangular solves for the 3x3
// inverse, transpose / neg / bmm / cat for the pose): ~1.3 ms per frame, a sixth of a frame's share at 8 GPUs.  K^-1 is the
// cofactor formula in fp64, rounded once to fp32 (<= 0.5 ulp from the exact inverse; LU in fp32 is within a few ulp of it, so
// the two agree to the last bits but are not bit-identical -- the fp32 parity mode keeps torch's own call, camera.view_matrices);
// pose^-1 = [R^T | (-R^T) t] with the matmul as the sequential mul, fma, fma chain torch's fp32 matmuls reduce to.
__global__ void view_matrices_kernel(const float* __restrict__ intr, const float* __restrict__ pose, int B,
                                     float* __restrict__ kinv, float* __restrict__ pinv) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* K = intr + b * 9;
  const double a = K[0], bb = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i = K[8];
  const double A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + bb * Bc + c * C, inv = 1.0 / det;
  float* o = kinv + b * 9;
  o[0] = (float)(A * inv);              o[1] = (float)(-(bb * i - c * h) * inv); o[2] = (float)((bb * f - c * e) * inv);
  o[3] = (float)(Bc * inv);             o[4] = (float)((a * i - c * g) * inv);   o[5] = (float)(-(a * f - c * d) * inv);
  o[6] = (float)(C * inv);              o[7] = (float)(-(a * h - bb * g) * inv); o[8] = (float)((a * e - bb * d) * inv);
  const float* P = pose + b * 12;       // [R | t], rows of 4
  float* q = pinv + b * 12;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float r0 = P[0 * 4 + r], r1 = P[1 * 4 + r], r2 = P[2 * 4 + r];      // row r of R^T
    q[r * 4 + 0] = r0; q[r * 4 + 1] = r1; q[r * 4 + 2] = r2;
    float acc = __fmul_rn(-r0, P[0 * 4 + 3]);
    acc = __fmaf_rn(-r1, P[1 * 4 + 3], acc);
    q[r * 4 + 3] = __fmaf_rn(-r2, P[2 * 4 + 3], acc);
  }
}

TP_API int tp_view_matrices(const float* intr, const float* pose, int B, float* kinv, float* pose_inv, void* stream) {
  TP_CHECK(intr && pose && kinv && pose_inv, TP_ERR_BAD_ARG);
  TP_CHECK(B > 0, TP_ERR_BAD_SHAPE);
  view_matrices_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(intr, pose, B, kinv, pose_inv);
  return tp_launch_status();
}

TP_API int tp_raygen(const float* kinv, const float* pose_inv, int B, int H, int W, float pix_offset,
                     const int64_t* ray_idx, int R, float* center, float* ray, void* stream) {
  TP_CHECK(kinv && pose_inv && center && ray, TP_ERR_BAD_ARG);
  TP_CHECK(B > 0 && H > 0 && W > 0 && R >= 0, TP_ERR_BAD_SHAPE);
  TP_CHECK(ray_idx || R == H * W, TP_ERR_BAD_SHAPE);
  if ((long long)B * R == 0) return TP_OK;
  raygen_kernel<<<tp_grid_for((long long)B * R, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
      kinv, pose_inv, B, H, W, pix_offset, (const long long*)ray_idx, R, center, ray);
  return tp_launch_status();
}

TP_API int tp_patch_rays(const float* kinv, const float* pose_inv, const float* coords, int B, int P, int H, int W,
                         float* center, float* ray, void* stream) {
  TP_CHECK(kinv && pose_inv && coords && center && ray, TP_ERR_BAD_ARG);
  TP_CHECK(B > 0 && H > 1 && W > 1 && P >= 0, TP_ERR_BAD_SHAPE);
  if ((long long)B * P == 0) return TP_OK;
  patch_rays_kernel<<<tp_grid_for((long long)B * P, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
      kinv, pose_inv, coords, B, P, H, W, center, ray);
  return tp_launch_status();
}

TP_API int tp_grid_sample_bilinear(const float* image, const float* coords, int B, int C, int H, int W, int P,
                                   float* out, void* stream) {
  TP_CHECK(image && coords && out, TP_ERR_BAD_ARG);
  TP_CHECK(B > 0 && C > 0 && H > 1 && W > 1 && P >= 0, TP_ERR_BAD_SHAPE);
  if (P == 0) return TP_OK;
  grid_sample_kernel<<<tp_grid_for((long long)B * C * P, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
      image, coords, B, C, H, W, P, out);
  return tp_launch_status();
}

TP_API int tp_gather_rows(const float* src, const int64_t* idx, int B, int64_t HW, int C, int R, float* out,
                          void* stream) {
  TP_CHECK(src && idx && out, TP_ERR_BAD_ARG);
  TP_CHECK(B > 0 && HW > 0 && C > 0 && R >= 0, TP_ERR_BAD_SHAPE);
  if (R == 0) return TP_OK;
  gather_rows_kernel<<<tp_grid_for((long long)B * R * C, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
      src, (const long long*)idx, B, HW, C, R, out);
  return tp_launch_status();
}

TP_API int tp_aabb_intersect(const float* aabb_min, const float* aabb_max, int aabb_batched, const float* ray_o,
                             const float* ray_d, int B, int64_t n_per_batch, float* t_near, float* t_far,
                             uint8_t* valid, void* stream) {
  TP_CHECK(aabb_min && aabb_max && ray_o && ray_d && t_near && t_far && valid, TP_ERR_BAD_ARG);
  TP_CHECK(B > 0 && n_per_batch >= 0, TP_ERR_BAD_SHAPE);
  const long long total = (long long)B * n_per_batch;
  if (total == 0) return TP_OK;
  aabb_kernel<<<tp_grid_for(total, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
      aabb_min, aabb_max, aabb_batched ? 3 : 0, ray_o, ray_d, n_per_batch, total, t_near, t_far, valid);
  return tp_launch_status();
}

TP_API int tp_sample_depth(const float* z_near, const float* z_far, const float* rand, int64_t n_rays, int N,
                           int mode, uint64_t seed, float* out, void* stream) {
  TP_CHECK(z_near && z_far && out, TP_ERR_BAD_ARG);
  TP_CHECK(n_rays >= 0 && N > 0, TP_ERR_BAD_SHAPE);
  TP_CHECK(mode >= 0 && mode <= 2 && (mode != 0 || rand), TP_ERR_BAD_ARG);
  if (n_rays == 0) return TP_OK;
  if (mode == 2)
    sample_depth_philox_kernel<<<tp_grid_for((n_rays * N + 3) / 4, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
        z_near, z_far, n_rays, N, seed, out);
  else
    sample_depth_kernel<<<tp_grid_for(n_rays * N, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
        z_near, z_far, rand, n_rays, N, mode, out);
  return tp_launch_status();
}

TP_API int tp_box_range(const float* kinv, const float* pose_inv, int B, int H, int W, const float* aabb_min,
                        const float* aabb_max, int aabb_batched, float bg_near, float bg_far, float* z_near,
                        float* z_far, uint8_t* valid, void* stream) {
  TP_CHECK(kinv && pose_inv && aabb_min && aabb_max && z_near && z_far, TP_ERR_BAD_ARG);
  TP_CHECK(B > 0 && H > 0 && W > 0, TP_ERR_BAD_SHAPE);
  box_range_kernel<<<tp_grid_for((long long)B * H * W, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
      kinv, pose_inv, B, H, W, aabb_min, aabb_max, aabb_batched ? 3 : 0, bg_near, bg_far, z_near, z_far, valid);
  return tp_launch_status();
}

TP_API int tp_depth_guided_range(const float* depth, int64_t n, float bg_near, float bg_far, float* z_near,
                                 float* z_far, void* stream) {
  TP_CHECK(depth && z_near && z_far, TP_ERR_BAD_ARG);
  TP_CHECK(n >= 0, TP_ERR_BAD_SHAPE);
  if (n == 0) return TP_OK;
  guided_range_kernel<<<tp_grid_for(n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(depth, n, bg_near, bg_far,
                                                                                       z_near, z_far);
  return tp_launch_status();
}

TP_API int tp_normal_from_depth(const float* kinv, const float* pose_inv, const float* depth, int B, int H, int W,
                                float* normal, void* stream) {
  TP_CHECK(kinv && pose_inv && depth && normal, TP_ERR_BAD_ARG);
  TP_CHECK(B > 0 && H > 2 && W > 2, TP_ERR_BAD_SHAPE);
  normal_kernel<<<tp_grid_for((long long)B * H * W, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
      kinv, pose_inv, depth, B, H, W, normal);
  return tp_launch_status();
}

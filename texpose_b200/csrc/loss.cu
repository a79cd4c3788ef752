// Fused patch gather + ray-wise loss terms + backward seeds.
//
// Reference: Graph.compute_loss(train_step='nerf') (model/nerf_adapt_st_gan.py:712-763) and Model.summarize_loss
// (model/base.py:145-157).  There it is ~30 aten kernels (4 grid_samples, permutes, elementwise chains, 4 reductions) plus
// their autograd replay; here two launches produce the sampled targets, the loss terms and d(loss.all)/d{rgb, uncert,
// density} directly:
//
//   patch_loss_partial_kernel   per ray: bilinear image sample (align_corners=True, bit-exact with torch's CPU kernel),
//                               nearest mask sample (align_corners=False: x = ((c+1)W-1)/2, round half to even, zero outside),
//                               block partials of  sum m (I-rgb)^2/u^2,  sum m,  sum log u^2;  grid-stride partials of
//                               sum density[...,1].
//   patch_loss_grad_kernel      fixed-order reduction of the partials (deterministic), the loss scalars, and the seeds
//                               g_rgb = w_r * -2 m (I-rgb) / u^2 / (M + 1e-5)
//                               g_unc = w_r * -2 m sum_c (I-rgb)^2 / u^3 / (M + 1e-5) + w_u / (u * rays)
//                               g_density[..., 1] = w_t / samples.
// HBM-bound; algorithmic bytes: 40 B/ray in + 32 B/ray out, 8 B/sample in + 8 B/sample out (the density terms).
#include "bilinear.cuh"
#include "../../include/texpose_b200.h"

namespace {

constexpr int kLossThreads = 256;

struct LossParams {
  const float* image;      // [B,3,H,W]
  const float* obj_mask;   // [B,H,W] raw map (> 0 = object)
  const float* coords;     // [B,R,2] in [-1,1]
  int B, R, H, W;
  const float* rgb;        // [B,R,3]
  const float* uncert;     // [B,R]
  const float* density;    // [B*R*N,2]
  long long S;             // B*R*N
  float w_render, w_uncert, w_trans;   // linear weights (10^w); 0 with the term bit cleared = off
  int terms;               // bit 0 render, bit 1 uncert, bit 2 trans_reg
  float* image_sample;     // [B,3,R]
  float* mask_sample;      // [B,R]
  float* losses;           // [4] render, uncert, trans_reg, all
  float* g_rgb;            // [B,R,3]
  float* g_uncert;         // [B,R]
  float* g_density;        // [S,2] or NULL
  float* partial;          // [blocks][4]
  int blocks;
  const float* g_losses;   // device [4] or NULL: upstream gradients of {render, uncert, trans_reg, all} (NULL = {0,0,0,1})
  int write_losses;        // 0 in the backward launch (tp_patch_loss_backward): the scalars were written by the forward
};

__device__ __forceinline__ float block_sum(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < kLossThreads / 32; ++i) t += sm[i];
  return t;     // valid in thread 0
}

__global__ void __launch_bounds__(kLossThreads) patch_loss_partial_kernel(const LossParams p) {
  __shared__ float sm[kLossThreads / 32];
  const long long rays = (long long)p.B * p.R;
  float a = 0.f, m_sum = 0.f, lg = 0.f, tr = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rays; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / p.R);
    const long long r = i - (long long)b * p.R;
    const float gx = p.coords[i * 2], gy = p.coords[i * 2 + 1];
    const Bilin s = bilin_setup(gx, gy, p.H, p.W);
    // nearest, align_corners=False (the default the reference's mask lookups fall into, :729-730)
    const float fx = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)p.W), 1.f), 2.f);
    const float fy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)p.H), 1.f), 2.f);
    const float rx = nearbyintf(fx), ry = nearbyintf(fy);
    float m = 0.f;
    if (rx >= 0.f && rx < (float)p.W && ry >= 0.f && ry < (float)p.H)
      m = p.obj_mask[((long long)b * p.H + (int)ry) * p.W + (int)rx] > 0.f ? 1.f : 0.f;
    p.mask_sample[i] = m;
    const float u = p.uncert[i];
    float e2 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* img = p.image + ((long long)b * 3 + c) * p.H * p.W;
      const float I = bilin_apply(s, p.H, p.W, [&](int y, int x) { return img[(long long)y * p.W + x]; });
      p.image_sample[((long long)b * 3 + c) * p.R + r] = I;
      const float d = I - p.rgb[i * 3 + c];
      e2 += d * d / (u * u);
    }
    a += m * e2;
    m_sum += m;
    lg += logf(u * u);
  }
  if (p.terms & 4)
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < p.S; s += (long long)gridDim.x * blockDim.x)
      tr += p.density[s * 2 + 1];
  const float A = block_sum(a, sm), M = block_sum(m_sum, sm), L = block_sum(lg, sm), T = block_sum(tr, sm);
  if (threadIdx.x == 0) {
    float* o = p.partial + (size_t)blockIdx.x * 4;
    o[0] = A; o[1] = M; o[2] = L; o[3] = T;
  }
}

__global__ void __launch_bounds__(kLossThreads) patch_loss_grad_kernel(const LossParams p) {
  __shared__ float tot[4];
  if (threadIdx.x < 4) {                 // every block re-reduces the (few hundred) partials in the same fixed order
    float t = 0.f;
    for (int i = 0; i < p.blocks; ++i) t += p.partial[(size_t)i * 4 + threadIdx.x];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  const long long rays = (long long)p.B * p.R;
  const float denom = tot[1] + 1e-5f;
  if (p.write_losses && blockIdx.x == 0 && threadIdx.x == 0) {
    const float l_r = tot[0] / denom, l_u = 5.f + tot[2] / (float)rays / 2.f, l_t = p.S > 0 ? tot[3] / (float)p.S : 0.f;
    float all = 0.f;
    if (p.terms & 1) all += p.w_render * l_r;
    if (p.terms & 2) all += p.w_uncert * l_u;
    if (p.terms & 4) all += p.w_trans * l_t;
    p.losses[0] = (p.terms & 1) ? l_r : 0.f;
    p.losses[1] = (p.terms & 2) ? l_u : 0.f;
    p.losses[2] = (p.terms & 4) ? l_t : 0.f;
    p.losses[3] = all;
  }
  if (!p.g_rgb) return;      // forward launch of the autograd function: scalars only, the seeds are formed in the backward
  // d(sum_k g_k * loss_k)/d{.} with g = upstream gradients of the four outputs: `all` is sum_k w_k * term_k, so term k
  // carries g_k + g_all * w_k (the reference engine backpropagates `all` built by Model.summarize_loss: g = {w_r, w_u, w_t, 0})
  const float g0 = p.g_losses ? p.g_losses[0] : 0.f, g1 = p.g_losses ? p.g_losses[1] : 0.f, g2 = p.g_losses ? p.g_losses[2] : 0.f,
              g3 = p.g_losses ? p.g_losses[3] : 1.f;
  const float kr = (p.terms & 1) ? (g0 + g3 * p.w_render) / denom : 0.f;
  const float ku = (p.terms & 2) ? (g1 + g3 * p.w_uncert) / (float)rays : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rays; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / p.R);
    const long long r = i - (long long)b * p.R;
    const float m = p.mask_sample[i], u = p.uncert[i];
    const float iu2 = 1.f / (u * u);
    float e2 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = p.image_sample[((long long)b * 3 + c) * p.R + r] - p.rgb[i * 3 + c];
      p.g_rgb[i * 3 + c] = -2.f * kr * m * d * iu2;
      e2 += d * d;
    }
    p.g_uncert[i] = -2.f * kr * m * e2 * iu2 / u + ku / u;
  }
  if (p.g_density) {
    const float gt = (p.terms & 4) && p.S > 0 ? (g2 + g3 * p.w_trans) / (float)p.S : 0.f;
    float2* g = reinterpret_cast<float2*>(p.g_density);
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < p.S; s += (long long)gridDim.x * blockDim.x)
      g[s] = make_float2(0.f, gt);
  }
}

}  // namespace

TP_API int64_t tp_patch_loss_workspace(void) { return (int64_t)tp_num_sms() * 2 * 4; }

TP_API int tp_patch_loss(const float* image, const float* obj_mask, const float* coords, int B, int R, int H, int W,
                         const float* rgb, const float* uncert, const float* density, int N, float w_render, float w_uncert,
                         float w_trans_reg, int terms, float* image_sample, float* mask_sample, float* losses, float* g_rgb,
                         float* g_uncert, float* g_density, float* workspace, int64_t workspace_floats, void* stream) {
  if (!image || !obj_mask || !coords || !rgb || !uncert || !image_sample || !mask_sample || !losses || !workspace)
    return TP_ERR_BAD_ARG;
  if ((g_rgb == nullptr) != (g_uncert == nullptr) || (!g_rgb && g_density)) return TP_ERR_BAD_ARG;      // seeds: all or none
  if ((terms & 4) && !density) return TP_ERR_BAD_ARG;
  if (B < 1 || R < 1 || H < 2 || W < 2 || N < 0 || (terms & ~7)) return TP_ERR_BAD_SHAPE;
  if (((uintptr_t)g_density & 7)) return TP_ERR_ALIGN;
  LossParams p;
  p.image = image; p.obj_mask = obj_mask; p.coords = coords; p.B = B; p.R = R; p.H = H; p.W = W;
  p.rgb = rgb; p.uncert = uncert; p.density = density; p.S = (long long)B * R * N;
  p.w_render = w_render; p.w_uncert = w_uncert; p.w_trans = w_trans_reg; p.terms = terms;
  p.image_sample = image_sample; p.mask_sample = mask_sample; p.losses = losses;
  p.g_rgb = g_rgb; p.g_uncert = g_uncert; p.g_density = g_density; p.partial = workspace;
  p.g_losses = nullptr; p.write_losses = 1;
  const long long work = (terms & 4) ? p.S : (long long)B * R;
  p.blocks = tp_grid_for(work > (long long)B * R ? work : (long long)B * R, kLossThreads, 2);
  if (workspace_floats < (int64_t)p.blocks * 4) return TP_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  patch_loss_partial_kernel<<<p.blocks, kLossThreads, 0, st>>>(p);
  if (int rc = tp_launch_status()) return rc;
  patch_loss_grad_kernel<<<g_rgb ? p.blocks : 1, kLossThreads, 0, st>>>(p);      // without seeds: one block for the scalars
  return tp_launch_status();
}

// Backward of tp_patch_loss for arbitrary upstream gradients of its four scalars (g_losses: DEVICE pointer, so no host sync):
// re-reduces the forward's partials in the same fixed order and writes the seeds.  image_sample / mask_sample / workspace are
// the forward's outputs, unchanged.
TP_API int tp_patch_loss_backward(const float* g_losses, const float* image_sample, const float* mask_sample, int B, int R,
                                  const float* rgb, const float* uncert, int N, float w_render, float w_uncert,
                                  float w_trans_reg, int terms, float* g_rgb, float* g_uncert, float* g_density,
                                  const float* workspace, int64_t workspace_floats, void* stream) {
  if (!g_losses || !image_sample || !mask_sample || !rgb || !uncert || !g_rgb || !g_uncert || !workspace) return TP_ERR_BAD_ARG;
  if (B < 1 || R < 1 || N < 0 || (terms & ~7)) return TP_ERR_BAD_SHAPE;
  if (((uintptr_t)g_density & 7)) return TP_ERR_ALIGN;
  LossParams p;
  p.image = nullptr; p.obj_mask = nullptr; p.coords = nullptr; p.B = B; p.R = R; p.H = 0; p.W = 0;
  p.rgb = rgb; p.uncert = uncert; p.density = nullptr; p.S = (long long)B * R * N;
  p.w_render = w_render; p.w_uncert = w_uncert; p.w_trans = w_trans_reg; p.terms = terms;
  p.image_sample = const_cast<float*>(image_sample); p.mask_sample = const_cast<float*>(mask_sample); p.losses = nullptr;
  p.g_rgb = g_rgb; p.g_uncert = g_uncert; p.g_density = g_density; p.partial = const_cast<float*>(workspace);
  p.g_losses = g_losses; p.write_losses = 0;
  const long long work = (terms & 4) ? p.S : (long long)B * R;      // the forward's grid: same number of partials
  p.blocks = tp_grid_for(work > (long long)B * R ? work : (long long)B * R, kLossThreads, 2);
  if (workspace_floats < (int64_t)p.blocks * 4) return TP_ERR_WORKSPACE;
  patch_loss_grad_kernel<<<p.blocks, kLossThreads, 0, (cudaStream_t)stream>>>(p);
  return tp_launch_status();
}

// ------------------------------------------------------------------------------------------ eval-frame epilogue
//
// Model.evaluate_full per frame (model/nerf_adapt_st_gan.py:341-362): rgb_static [B,HW,3] -> rgb_map [B,3,H,W]; depth ->
// depth_map / depth.scale; image * mask; PSNR = -10 log10(mean((rgb_map - image*mask)^2)) -- the reference does this with six
// permute/mul/mse kernels and a .item() sync per frame (B = 1); here one pass over the pixels for any B, per-view MSE / PSNR
// left on the device (fixed-order two-stage reduction).  HBM-bound: 32 B in + 28 B out per pixel.
namespace {

constexpr int kEvalBlocksPerView = 64;

__global__ void __launch_bounds__(256) eval_epilogue_kernel(const float* __restrict__ rgb, const float* __restrict__ depth,
                                                            const float* __restrict__ image, const float* __restrict__ mask,
                                                            long long HW, float depth_scale, float* __restrict__ rgb_map,
                                                            float* __restrict__ depth_map, float* __restrict__ image_masked,
                                                            float* __restrict__ partial) {
  __shared__ float sm[8];
  const int b = blockIdx.y;
  float acc = 0.f;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const long long i = (long long)b * HW + p;
    const float mv = mask[i];                          // evaluate_full multiplies by the 0/1 obj_mask map as it is (:343,358)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float r = rgb[i * 3 + c];
      const float im = image[((long long)b * 3 + c) * HW + p] * mv;
      rgb_map[((long long)b * 3 + c) * HW + p] = r;
      image_masked[((long long)b * 3 + c) * HW + p] = im;
      const float d = r - im;
      acc += d * d;
    }
    depth_map[i] = depth[i] / depth_scale;
  }
  const float t = block_sum(acc, sm);
  if (threadIdx.x == 0) partial[(size_t)b * gridDim.x + blockIdx.x] = t;
}

__global__ void eval_psnr_kernel(const float* __restrict__ partial, int blocks, long long HW, int B, float* __restrict__ mse,
                                 float* __restrict__ psnr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float t = 0.f;
  for (int i = 0; i < blocks; ++i) t += partial[(size_t)b * blocks + i];
  const float m = t / (float)(3 * HW);
  mse[b] = m;
  psnr[b] = -10.f * log10f(m);
}

// Latent rows of a training batch (model/nerf_adapt_st_gan.py:589-603: latent_vars_trans.weight[idx], latent_vars_light.weight[idx])
// and their gradient.  torch's index backward is ~14 launches per table (sort, arange, index_put ...): 2 % of a C3 step.
__global__ void latent_rows_kernel(const float* __restrict__ ta, int ca, const float* __restrict__ tb, int cb,
                                   const long long* __restrict__ idx, int B, float* __restrict__ oa, float* __restrict__ ob) {
  const int n = B * (ca + cb);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = i / (ca + cb), c = i - b * (ca + cb);
    const long long r = idx[b];
    if (c < ca) oa[b * ca + c] = ta[r * ca + c];
    else ob[b * cb + c - ca] = tb[r * cb + c - ca];
  }
}
// d_table[r, c] = sum over b with idx[b] == r of g[b, c], in ascending b (deterministic); rows no sample touches become zero
__global__ void latent_rows_grad_kernel(const float* __restrict__ ga, int ca, long long ra, const float* __restrict__ gb, int cb,
                                        long long rb, const long long* __restrict__ idx, int B, float* __restrict__ da,
                                        float* __restrict__ db) {
  const long long na = ra * ca, n = na + rb * cb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool first = i < na;
    const long long j = first ? i : i - na;
    const int cols = first ? ca : cb;
    const long long r = j / cols;
    const int c = (int)(j - r * cols);
    const float* g = first ? ga : gb;
    float acc = 0.f;
    for (int b = 0; b < B; ++b)
      if (idx[b] == r) acc += g[b * cols + c];
    (first ? da : db)[j] = acc;
  }
}

}  // namespace

TP_API int tp_latent_rows(const float* table_a, int cols_a, const float* table_b, int cols_b, const int64_t* idx, int B,
                          float* out_a, float* out_b, void* stream) {
  if (!table_a || !table_b || !idx || !out_a || !out_b) return TP_ERR_BAD_ARG;
  if (B < 0 || cols_a < 1 || cols_b < 1) return TP_ERR_BAD_SHAPE;
  if (B == 0) return TP_OK;
  latent_rows_kernel<<<tp_grid_for((long long)B * (cols_a + cols_b), 256, 2), 256, 0, (cudaStream_t)stream>>>(
      table_a, cols_a, table_b, cols_b, reinterpret_cast<const long long*>(idx), B, out_a, out_b);
  return tp_launch_status();
}

TP_API int tp_latent_rows_backward(const float* g_a, int cols_a, int64_t rows_a, const float* g_b, int cols_b, int64_t rows_b,
                                   const int64_t* idx, int B, float* d_table_a, float* d_table_b, void* stream) {
  if (!g_a || !g_b || !idx || !d_table_a || !d_table_b) return TP_ERR_BAD_ARG;
  if (B < 0 || cols_a < 1 || cols_b < 1 || rows_a < 1 || rows_b < 1) return TP_ERR_BAD_SHAPE;
  latent_rows_grad_kernel<<<tp_grid_for(rows_a * cols_a + rows_b * cols_b, 256, 4), 256, 0, (cudaStream_t)stream>>>(
      g_a, cols_a, rows_a, g_b, cols_b, rows_b, reinterpret_cast<const long long*>(idx), B, d_table_a, d_table_b);
  return tp_launch_status();
}

TP_API int64_t tp_eval_epilogue_workspace(int B) { return (int64_t)B * kEvalBlocksPerView; }

TP_API int tp_eval_epilogue(const float* rgb, const float* depth, const float* image, const float* mask, int B, int64_t HW,
                            float depth_scale, float* rgb_map, float* depth_map, float* image_masked, float* mse, float* psnr,
                            float* workspace, int64_t workspace_floats, void* stream) {
  if (!rgb || !depth || !image || !mask || !rgb_map || !depth_map || !image_masked || !mse || !psnr || !workspace)
    return TP_ERR_BAD_ARG;
  if (B < 1 || HW < 1 || !(depth_scale > 0.f)) return TP_ERR_BAD_SHAPE;
  if (workspace_floats < tp_eval_epilogue_workspace(B)) return TP_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  eval_epilogue_kernel<<<dim3(kEvalBlocksPerView, B), 256, 0, st>>>(rgb, depth, image, mask, HW, depth_scale, rgb_map,
                                                                    depth_map, image_masked, workspace);
  if (int rc = tp_launch_status()) return rc;
  eval_psnr_kernel<<<(B + 63) / 64, 64, 0, st>>>(workspace, kEvalBlocksPerView, HW, B, mse, psnr);
  return tp_launch_status();
}

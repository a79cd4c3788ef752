// K7: mesh depth / NOCS / colour rasteriser -- SURVEY 8 (f4).
//
// Replaces tools/mvrenderer.py:33-178 as compute_surfelinfo.py:114-116 uses it: pytorch3d's MeshRasterizer (faces_per_pixel = 1,
// blur_radius = 0, perspective-correct barycentrics, no back-face culling) followed by a vertex-attribute shader (vertex
// colours under AmbientLights, or the normalised object coordinates of SoftPhongNOCSShader, mvrenderer.py:675-727) blended with
// softmax_rgb_blend(sigma = gamma = 1e-4, black background), and fragments.zbuf as the depth map (-1 where no face covers the
// pixel).  pytorch3d is a third-party dependency the reference does not pin (README.md:16) and that is absent here: the
// arithmetic below restates its published rasterisation rules (rasterize_meshes.cu fine pass, blending.py) -- PARITY UNPINNED.
//
// B200 mapping: HBM-bound integer / fp32 work, no tensor cores.  One thread per (view, face) walks the face's pixel bounding
// box (a LineMOD CAD face covers ~1 pixel at 480 x 640) and resolves visibility with ONE 64-bit atomicMin per covered pixel on
// the key (depth bits << 32 | face index): nearest face wins, equal depths resolve to the lower face index, the result does
// not depend on the order the atomics arrive in (deterministic).  A resolve pass turns keys into depth / attribute maps.
// Traffic per view: 36 B per vertex + 12 B per face read, 8 B per pixel of key + (4 + 4 C + 4) B per pixel written.
#include "common.cuh"
#include "../../include/texpose_b200.h"

namespace {

constexpr float kEps = 1e-8f;                    // pytorch3d kEpsilon
constexpr unsigned long long kEmpty = ~0ull;

struct V3 { float x, y, z; };

// pytorch3d NDC centre of pixel index i (counted from the +NDC side) along an axis of S1 pixels (other axis S2)
__device__ __forceinline__ float pix_to_ndc(int i, int S1, int S2) {
  const float range = S1 > S2 ? __fdiv_rn(2.0f * (float)S1, (float)S2) : 2.0f;
  const float offset = __fdiv_rn(range, (float)S1);
  return __fadd_rn(__fadd_rn(-__fdiv_rn(range, 2.0f), __fdiv_rn(offset, 2.0f)), __fmul_rn(offset, (float)i));
}
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return __fsub_rn(__fmul_rn(__fsub_rn(px, ax), __fsub_rn(by, ay)), __fmul_rn(__fsub_rn(py, ay), __fsub_rn(bx, ax)));
}
__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = bax * bax + bay * bay;
  if (l2 <= kEps) return (px - bx) * (px - bx) + (py - by) * (py - by);
  float t = (bax * (px - ax) + bay * (py - ay)) / l2;
  t = fminf(fmaxf(t, 0.f), 1.f);
  const float qx = ax + t * bax, qy = ay + t * bay;
  return (px - qx) * (px - qx) + (py - qy) * (py - qy);
}

// Coverage of pixel centre (px, py) by the face (v0, v1, v2): barycentrics b0 (screen space), perspective-corrected bary w,
// interpolated depth pz.  false: the centre is outside (any b0 <= 0), the face is degenerate, or the point is behind the camera.
__device__ __forceinline__ bool cover(float px, float py, const V3& v0, const V3& v1, const V3& v2, float (&w)[3], float& pz) {
  const float area = edge_fn(v2.x, v2.y, v0.x, v0.y, v1.x, v1.y);
  if (area <= kEps && area >= -kEps) return false;
  const float a = __fadd_rn(area, kEps);
  const float b0 = __fdiv_rn(edge_fn(px, py, v1.x, v1.y, v2.x, v2.y), a);
  const float b1 = __fdiv_rn(edge_fn(px, py, v2.x, v2.y, v0.x, v0.y), a);
  const float b2 = __fdiv_rn(edge_fn(px, py, v0.x, v0.y, v1.x, v1.y), a);
  if (!(b0 > 0.f && b1 > 0.f && b2 > 0.f)) return false;
  const float t0 = __fmul_rn(__fmul_rn(b0, v1.z), v2.z), t1 = __fmul_rn(__fmul_rn(v0.z, b1), v2.z), t2 = __fmul_rn(__fmul_rn(v0.z, v1.z), b2);
  const float den = fmaxf(__fadd_rn(__fadd_rn(t0, t1), t2), kEps);
  w[0] = __fdiv_rn(t0, den);
  w[1] = __fdiv_rn(t1, den);
  w[2] = __fdiv_rn(t2, den);
  pz = __fadd_rn(__fadd_rn(__fmul_rn(w[0], v0.z), __fmul_rn(w[1], v1.z)), __fmul_rn(w[2], v2.z));
  return pz >= 0.f;
}

// view-space vertex (pytorch3d axes: +X left, +Y up = the reference's T_calib, a rotation by pi about z, applied to the OpenCV
// pose, mvrenderer.py:47-48,143-150) -> NDC xy + view depth.  Screen-space intrinsics -> NDC as PerspectiveCameras(in_ndc=False).
__global__ void mesh_project_kernel(const float* __restrict__ verts, int V, const float* __restrict__ pose, const float* __restrict__ K,
                                    int B, int H, int W, float* __restrict__ out) {
  const long long n = (long long)B * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / V), v = (int)(i - (long long)b * V);
    const float* P = pose + b * 12;      // [R | t] row-major 3 x 4, X_cam = R X + t
    const float* Kb = K + b * 9;
    const float x = verts[v * 3], y = verts[v * 3 + 1], z = verts[v * 3 + 2];
    // one rounding per operation, left to right (the CPU restatement in oracle/mesh_oracle.py does the same: equal bits)
    auto dot = [](float a, float b, float c, float d, float x, float y, float z) {
      return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y)), __fmul_rn(c, z)), d);
    };
    const float xc = dot(P[0], P[1], P[2], P[3], x, y, z);
    const float yc = dot(P[4], P[5], P[6], P[7], x, y, z);
    const float zc = dot(P[8], P[9], P[10], P[11], x, y, z);
    const float s = __fmul_rn(0.5f, (float)(H < W ? H : W));
    const float fx = __fdiv_rn(Kb[0], s), fy = __fdiv_rn(Kb[4], s);
    const float px = __fdiv_rn(-__fsub_rn(Kb[2], __fmul_rn(0.5f, (float)W)), s), py = __fdiv_rn(-__fsub_rn(Kb[5], __fmul_rn(0.5f, (float)H)), s);
    const float xv = -xc, yv = -yc;      // rotation by pi about z
    out[i * 3] = __fdiv_rn(__fadd_rn(__fmul_rn(fx, xv), __fmul_rn(px, zc)), zc);
    out[i * 3 + 1] = __fdiv_rn(__fadd_rn(__fmul_rn(fy, yv), __fmul_rn(py, zc)), zc);
    out[i * 3 + 2] = zc;
  }
}

__global__ void mesh_clear_kernel(unsigned long long* __restrict__ keys, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) keys[i] = kEmpty;
}

__device__ __forceinline__ void load_face(const float* __restrict__ ndc, const int* __restrict__ faces, int f, V3& v0, V3& v1, V3& v2) {
  const int i0 = faces[f * 3], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
  v0 = {ndc[i0 * 3], ndc[i0 * 3 + 1], ndc[i0 * 3 + 2]};
  v1 = {ndc[i1 * 3], ndc[i1 * 3 + 1], ndc[i1 * 3 + 2]};
  v2 = {ndc[i2 * 3], ndc[i2 * 3 + 1], ndc[i2 * 3 + 2]};
}

__global__ void mesh_raster_kernel(const float* __restrict__ ndc, int V, const int* __restrict__ faces, int F, int B, int H, int W,
                                   unsigned long long* __restrict__ keys) {
  const long long n = (long long)B * F;
  const float s = 0.5f * (float)(H < W ? H : W);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / F), f = (int)(i - (long long)b * F);
    V3 v0, v1, v2;
    load_face(ndc + (size_t)b * V * 3, faces, f, v0, v1, v2);
    if (!(v0.z > 0.f && v1.z > 0.f && v2.z > 0.f)) continue;      // (the reference clips such faces at z = 0; a CAD model in front of the camera has none)
    // pixel columns / rows (counted from the left / top) the face's bounding box may cover, with one pixel of margin
    const float xmax = fmaxf(v0.x, fmaxf(v1.x, v2.x)), xmin = fminf(v0.x, fminf(v1.x, v2.x));
    const float ymax = fmaxf(v0.y, fmaxf(v1.y, v2.y)), ymin = fminf(v0.y, fminf(v1.y, v2.y));
    int c0 = (int)floorf(0.5f * W - xmax * s - 1.5f), c1 = (int)ceilf(0.5f * W - xmin * s + 0.5f);
    int r0 = (int)floorf(0.5f * H - ymax * s - 1.5f), r1 = (int)ceilf(0.5f * H - ymin * s + 0.5f);
    c0 = c0 < 0 ? 0 : c0; r0 = r0 < 0 ? 0 : r0;
    c1 = c1 > W - 1 ? W - 1 : c1; r1 = r1 > H - 1 ? H - 1 : r1;
    for (int r = r0; r <= r1; ++r) {
      const float py = pix_to_ndc(H - 1 - r, H, W);
      for (int c = c0; c <= c1; ++c) {
        const float px = pix_to_ndc(W - 1 - c, W, H);
        float w[3], pz;
        if (!cover(px, py, v0, v1, v2, w, pz)) continue;
        const unsigned long long key = ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned)f;
        atomicMin(keys + ((size_t)b * H + r) * W + c, key);
      }
    }
  }
}

// attr [V,C] vertex attributes (C <= 8) -> out [B,C,H,W]; depth [B,H,W] (-1 = background); pix_to_face [B,H,W] (-1 = background)
__global__ void mesh_resolve_kernel(const float* __restrict__ ndc, int V, const int* __restrict__ faces, const float* __restrict__ attr,
                                    int C, int B, int H, int W, const unsigned long long* __restrict__ keys, float sigma,
                                    float* __restrict__ out, float* __restrict__ depth, int* __restrict__ pix_to_face) {
  const long long HW = (long long)H * W, n = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const long long p = i - (long long)b * HW;
    const int r = (int)(p / W), c = (int)(p - (long long)r * W);
    const unsigned long long key = keys[i];
    if (key == kEmpty) {
      if (depth) depth[i] = -1.f;
      if (pix_to_face) pix_to_face[i] = -1;
      if (out)
        for (int ch = 0; ch < C; ++ch) out[((size_t)b * C + ch) * HW + p] = 0.f;
      continue;
    }
    const int f = (int)(unsigned)(key & 0xffffffffull);
    V3 v0, v1, v2;
    load_face(ndc + (size_t)b * V * 3, faces, f, v0, v1, v2);
    const float px = pix_to_ndc(W - 1 - c, W, H), py = pix_to_ndc(H - 1 - r, H, W);
    float w[3], pz;
    cover(px, py, v0, v1, v2, w, pz);      // same arithmetic as the raster pass: the same values
    if (depth) depth[i] = pz;
    if (pix_to_face) pix_to_face[i] = f;
    if (!out) continue;
    // softmax_rgb_blend with one face per pixel: colour = prob c / (prob + delta), prob = sigmoid(d^2 / sigma) for a covered
    // pixel (d = distance to the nearest edge), delta = exp((eps - z_inv) / gamma) clamped to eps = 1e-10
    const float d2 = fminf(seg_dist2(px, py, v0.x, v0.y, v1.x, v1.y),
                           fminf(seg_dist2(px, py, v1.x, v1.y, v2.x, v2.y), seg_dist2(px, py, v2.x, v2.y, v0.x, v0.y)));
    const float prob = 1.f / (1.f + expf(-d2 / sigma));
    const float denom = prob + 1e-10f;
    const int i0 = faces[f * 3], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
    for (int ch = 0; ch < C; ++ch) {
      const float a = __fadd_rn(__fadd_rn(__fmul_rn(w[0], attr[i0 * C + ch]), __fmul_rn(w[1], attr[i1 * C + ch])), __fmul_rn(w[2], attr[i2 * C + ch]));
      out[((size_t)b * C + ch) * HW + p] = (prob * a) / denom;
    }
  }
}

}  // namespace

TP_API int64_t tp_mesh_render_workspace(int B, int V, int H, int W) {
  // bytes: projected vertices [B,V,3] fp32 (rounded up to 8) + one 64-bit key per pixel
  return (((int64_t)B * V * 12 + 7) & ~(int64_t)7) + (int64_t)B * H * W * 8;
}

TP_API int tp_mesh_render(const float* verts, int V, const int32_t* faces, int F, const float* attr, int C, const float* pose,
                          const float* K, int B, int H, int W, float sigma, float* out, float* depth, int32_t* pix_to_face,
                          void* workspace, int64_t workspace_bytes, void* stream) {
  if (!verts || !faces || !pose || !K || !workspace) return TP_ERR_BAD_ARG;
  if (out && !attr) return TP_ERR_BAD_ARG;
  if (V < 1 || F < 0 || B < 1 || H < 1 || W < 1 || C < 0 || C > 8 || !(sigma > 0.f)) return TP_ERR_BAD_SHAPE;
  if (workspace_bytes < tp_mesh_render_workspace(B, V, H, W)) return TP_ERR_WORKSPACE;
  if ((uintptr_t)workspace & 7) return TP_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  float* ndc = reinterpret_cast<float*>(workspace);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(workspace) + (((int64_t)B * V * 12 + 7) & ~(int64_t)7));
  const long long n_pix = (long long)B * H * W;
  mesh_project_kernel<<<tp_grid_for((long long)B * V, 256, 8), 256, 0, st>>>(verts, V, pose, K, B, H, W, ndc);
  mesh_clear_kernel<<<tp_grid_for(n_pix, 256, 8), 256, 0, st>>>(keys, n_pix);
  if (F > 0) mesh_raster_kernel<<<tp_grid_for((long long)B * F, 128, 16), 128, 0, st>>>(ndc, V, faces, F, B, H, W, keys);
  mesh_resolve_kernel<<<tp_grid_for(n_pix, 256, 8), 256, 0, st>>>(ndc, V, faces, attr, C, B, H, W, keys, sigma, out, depth, pix_to_face);
  return tp_launch_status();
}

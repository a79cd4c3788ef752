// Bit-exact bilinear grid_sample arithmetic shared by the ray (rays.cu) and loss (loss.cu) kernels.
#pragma once
#include "common.cuh"

// F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True) coordinate + weights, in the exact
// arithmetic of torch's CPU kernel (the pinned oracle): ix = (x+1)*((W-1)/2); w = ix-floor(ix), e = 1-w;
// corner weights s*e, s*w, n*e, n*w; value = fma chain over nw, ne, sw, se (bit-exact vs tests/golden/patch.npz).
struct Bilin {
  int x0, y0;
  float wnw, wne, wsw, wse;
};
__device__ __forceinline__ Bilin bilin_setup(float gx, float gy, int H, int W) {
  const float ix = __fmul_rn(__fadd_rn(gx, 1.f), __fdiv_rn((float)(W - 1), 2.f));
  const float iy = __fmul_rn(__fadd_rn(gy, 1.f), __fdiv_rn((float)(H - 1), 2.f));
  const float fx = floorf(ix), fy = floorf(iy);
  Bilin s;
  s.x0 = (int)fx;
  s.y0 = (int)fy;
  const float w = __fsub_rn(ix, fx), e = __fsub_rn(1.f, w);
  const float n = __fsub_rn(iy, fy), so = __fsub_rn(1.f, n);
  s.wnw = __fmul_rn(so, e);
  s.wne = __fmul_rn(so, w);
  s.wsw = __fmul_rn(n, e);
  s.wse = __fmul_rn(n, w);
  return s;
}
template <class F>
__device__ __forceinline__ float bilin_apply(const Bilin& s, int H, int W, F val) {
  float acc = 0.f;
  const bool xl = s.x0 >= 0 && s.x0 < W, xr = s.x0 + 1 >= 0 && s.x0 + 1 < W;
  const bool yt = s.y0 >= 0 && s.y0 < H, yb = s.y0 + 1 >= 0 && s.y0 + 1 < H;
  if (xl && yt) acc = __fmaf_rn(val(s.y0, s.x0), s.wnw, acc);
  if (xr && yt) acc = __fmaf_rn(val(s.y0, s.x0 + 1), s.wne, acc);
  if (xl && yb) acc = __fmaf_rn(val(s.y0 + 1, s.x0), s.wsw, acc);
  if (xr && yb) acc = __fmaf_rn(val(s.y0 + 1, s.x0 + 1), s.wse, acc);
  return acc;
}


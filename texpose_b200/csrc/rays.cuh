// Per-ray arithmetic shared by the standalone ray kernels (rays.cu) and the fused render kernel (mlp_tc.cu): both must
// produce the same bits for a ray and its sample depths (the parity contract of camera.py:292-314 and
// model/nerf_adapt_st_gan.py:682-700 is bit-exactness).
#pragma once
#include "common.cuh"

// [u,v,1] @ Kinv^T then [cam,1] @ pose_inv^T, ray = grid - center.  The reference evaluates both
// products with torch.matmul; its fp32 result equals a sequential mul,fma,fma(,fma) chain over k
// (verified bit-for-bit against the reference on CPU, tests/golden/rays.npz).
__device__ __forceinline__ void unproject(const float* __restrict__ kinv, const float* __restrict__ pinv,
                                          float u, float v, float c[3], float d[3]) {
  float cam[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float acc = __fmul_rn(u, kinv[j * 3 + 0]);
    acc = __fmaf_rn(v, kinv[j * 3 + 1], acc);
    cam[j] = __fmaf_rn(1.0f, kinv[j * 3 + 2], acc);
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float acc = __fmul_rn(cam[0], pinv[j * 4 + 0]);
    acc = __fmaf_rn(cam[1], pinv[j * 4 + 1], acc);
    acc = __fmaf_rn(cam[2], pinv[j * 4 + 2], acc);
    acc = __fmaf_rn(1.0f, pinv[j * 4 + 3], acc);
    c[j] = pinv[j * 4 + 3];  // 0*R + 1*t == t exactly
    d[j] = __fsub_rn(acc, c[j]);
  }
}

// model/nerf_adapt_st_gan.py:690-697: (rand + k) / N * (far - near) + near, one rounding per operation (IEEE division, as
// torch's CPU kernel divides).
__device__ __forceinline__ float stratified_depth(float u, int k, float fn, float lo, float hi) {
  const float t = __fdiv_rn(__fadd_rn(u, (float)k), fn);
  return __fadd_rn(__fmul_rn(t, __fsub_rn(hi, lo)), lo);
}


// |ray| clamped as F.normalize does (eps 1e-12), with one rounding per operation so that every kernel that normalises a view
// direction (the bias-table kernel, the render launch's in-kernel row) produces the same bits
__device__ __forceinline__ float unit_length(float x, float y, float z) {
  return fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))), 1e-12f);
}

// Shared helpers for the texpose_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define TP_OK 0
#define TP_ERR_BAD_ARG (-1)
#define TP_ERR_BAD_SHAPE (-2)
#define TP_ERR_ALIGN (-3)
#define TP_ERR_ARCH (-4)
#define TP_ERR_WORKSPACE (-5)

#define TP_API extern "C" __attribute__((visibility("default")))

// Launch-error -> C-ABI return code (positive = cudaError_t).  Never synchronises.
static inline int tp_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TP_OK : (int)e;
}

static inline int tp_num_sms() {
  static int sms = 0;  // cached, read-only after first call (device property, not mutable state)
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// Grid sized as a multiple of the SM count (grid-stride kernels).
static inline int tp_grid_for(long long work_items, int threads, int ctas_per_sm) {
  long long need = (work_items + threads - 1) / threads;
  long long cap = (long long)tp_num_sms() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float tp_softplus(float x) {
  // F.softplus defaults: beta=1, threshold=20 (layers/nerf_static_transient_light.py:98,136-137)
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float tp_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// Philox4x32-10 counter RNG for the in-kernel stratified jitter (perf mode; the parity mode
// receives torch.rand draws from the host wrapper, SURVEY.md "RNG parity").
__device__ __forceinline__ uint4 tp_philox(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float tp_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

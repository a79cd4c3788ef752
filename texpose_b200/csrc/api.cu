// Library-level entry points.
#include "common.cuh"
#include "../../include/texpose_b200.h"

TP_API int tp_version(void) { return TP_VERSION; }

TP_API int tp_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

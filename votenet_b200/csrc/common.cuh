// Shared host/device helpers for the votenet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/votenet_b200.h"

namespace vnb {

// thread-local last-error message (vnb_last_error)
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define VNB_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return ::vnb::set_err(VNB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

void count_launch();

// after a launch: count it; report (and clear) any launch error
static inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(VNB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return VNB_OK;
}

#define VNB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) return ::vnb::set_err(VNB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Squared distance with the exact contraction nvcc applies to the reference expression
// (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1) (tf_sampling_g.cu:142, tf_grouping_g.cu:24):
// FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)  — spelled with intrinsics so no compiler flag can change it.
__device__ __forceinline__ float d2_ref_gpu(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

}  // namespace vnb

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

// Debugging aid: where did a bounded wait give up?  vnb_debug_trap_buffer(p) hands every translation unit a pointer to
// host-mapped memory; a trap site stores {line, blockDim.x, blockIdx.x, threadIdx.x, gridDim.x} there before __trap().
void register_trap_setter(void (*fn)(int*));
#ifdef __CUDACC__
static __device__ int* g_trap_slot = nullptr;  // one copy per translation unit
__device__ __noinline__ static void trap_at(int line) {
  if (g_trap_slot != nullptr) {
    volatile int* t = g_trap_slot;
    t[1] = (int)blockDim.x; t[2] = (int)blockIdx.x; t[3] = (int)threadIdx.x; t[4] = (int)gridDim.x; t[0] = line;
    __threadfence_system();
  }
  __trap();
}
namespace {
struct TrapRegistration {
  TrapRegistration() {
    register_trap_setter([](int* p) { cudaMemcpyToSymbol(g_trap_slot, &p, sizeof(p)); });
  }
} g_trap_registration;
}  // namespace
#endif

#ifdef __CUDACC__
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  A thread of the tensor-core epilogues / loaders owns a ROW, so a warp-wide access
// touches 32 different lines whatever its width; with 32 bytes per lane every access fills whole 32-byte sectors — half the
// LSU transactions of the 16-byte form for the same bytes (the fp32 row stores / residual loads of the epilogues were
// LSU-bound: 11 500 cycles for a 128 KB tile).  Pointers must be 32-byte aligned.
struct __align__(32) f32x8 { float v[8]; };
__device__ __forceinline__ f32x8 ldg_f32x8(const float* p) {
  f32x8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ f32x8 ld_f32x8(const float* p) {   // coherent form (data written earlier by this kernel)
  f32x8 r;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_f32x8(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f),
               "f"(g), "f"(h)
               : "memory");
}
__device__ __forceinline__ void st_b32x8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                         uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
               "r"(g), "r"(h)
               : "memory");
}
#endif

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Squared distance with the exact contraction nvcc applies to the reference expression
// (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1) (tf_sampling_g.cu:142, tf_grouping_g.cu:24):
// FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)  — spelled with intrinsics so no compiler flag can change it.
__device__ __forceinline__ float d2_ref_gpu(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

}  // namespace vnb

// Device helpers shared by the FPS kernels (fps.cu: register/cluster kernel, fps_pruned.cu: bucket-pruned kernel).
#pragma once
#include "common.cuh"

namespace vnb {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t raddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(raddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_volatile_v2(const void* p) {
  uint2 v;
  asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_addr(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t redux_max(uint32_t v) {
  uint32_t r;
  asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}

// 16-bit tie key: bit 15 = valid, low 15 bits = 0x7FFF - (((k & 511) << 6) | (k >> 9)); larger wins.
__device__ __forceinline__ uint32_t tie_key(int k) { return 0x8000u | (0x7FFFu - ((((uint32_t)k & 511u) << 6) | ((uint32_t)k >> 9))); }
__device__ __forceinline__ int tie_key_to_index(uint32_t key) {
  uint32_t t = 0x7FFFu - (key & 0x7FFFu);
  return (int)((t >> 6) | ((t & 63u) << 9));
}

__device__ __forceinline__ uint32_t redux_add(uint32_t v) {
  uint32_t r;
  asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ int redux_max_s32(int v) {
  int r;
  asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ int redux_min_s32(int v) {
  int r;
  asm volatile("redux.sync.min.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}

}  // namespace vnb

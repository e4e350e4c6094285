// TMEM read (tcgen05.ld 32x32b.x32) latency / throughput micro-benchmark, B200 sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../votenet_b200/csrc/umma.cuh"
using namespace umma;
#define N 512
template <int PIPE>
__global__ void k(long long* out, int* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&tptr, 512);
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tmem = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = clock64();
  if (PIPE == 1) {
    for (int i = 0; i < N; ++i) {
      uint32_t v[32];
      tmem_ld_x32(tmem + (i & 7) * 32 + (warp >> 2) * 256 % 512, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
    }
  } else {
    for (int i = 0; i < N; i += 2) {
      uint32_t v[32], w[32];
      tmem_ld_x32(tmem + (i & 7) * 32, v);
      tmem_ld_x32(tmem + ((i + 1) & 7) * 32, w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j];
    }
  }
  long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}
int main() {
  long long* d; int* s; cudaMalloc(&d, 32 * 8); cudaMalloc(&s, 4);
  for (int pipe = 1; pipe <= 2; ++pipe)
    for (int w : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) { if (pipe == 1) k<1><<<1, w * 32>>>(d, s); else k<2><<<1, w * 32>>>(d, s); cudaDeviceSynchronize(); }
      long long h[32]; cudaMemcpy(h, d, sizeof(long long) * w, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < w; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("tcgen05.ld x32 (4 KB/warp) %s warps=%2d: %6.1f cycles per load  -> %6.1f B/cycle/SM  (%s)\n", pipe == 1 ? "ld+wait+32 xor" : "2 loads in flight", w,
             (double)mx / N, 4096.0 * w / ((double)mx / N), cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}

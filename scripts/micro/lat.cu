// Latency micro-benchmarks for the warp-level primitives on the FPS critical path (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run: ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 256
__device__ __forceinline__ int redux_max_s32(int v) { int r; asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ unsigned redux_max_u32(unsigned v) { unsigned r; asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v)); return r; }

template <int WHICH>
__global__ void k(long long* out, int* sink, int seed) {
  __shared__ int sm[1024];
  __shared__ unsigned long long sm64[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7 + 1) & 1023;
  if (threadIdx.x < 64) sm64[threadIdx.x] = 0;
  __syncthreads();
  int v = seed + lane * 3 + warp;
  float f = (float)v;
  long long t0 = clock64();
  if (WHICH == 0) {  // dependent redux chain
#pragma unroll 8
    for (int i = 0; i < N; ++i) v = redux_max_s32(v ^ lane) + i;
  } else if (WHICH == 1) {  // dependent ballot chain
#pragma unroll 8
    for (int i = 0; i < N; ++i) v = (int)__ballot_sync(0xffffffffu, (v + lane) & 1) + i;
  } else if (WHICH == 2) {  // dependent shfl chain
#pragma unroll 8
    for (int i = 0; i < N; ++i) v = __shfl_sync(0xffffffffu, v, (v + i) & 31) + 1;
  } else if (WHICH == 3) {  // ffs chain
#pragma unroll 8
    for (int i = 0; i < N; ++i) v = __ffs(v | 0x10000) + v;
  } else if (WHICH == 4) {  // LDS pointer chase
#pragma unroll 8
    for (int i = 0; i < N; ++i) v = sm[v & 1023];
  } else if (WHICH == 5) {  // barrier loop
#pragma unroll 8
    for (int i = 0; i < N; ++i) { __syncthreads(); }
  } else if (WHICH == 6) {  // shuffle butterfly max (5 steps)
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
      int m = v ^ (lane * 2654435761u >> 7);
#pragma unroll
      for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
      v = m + i;
    }
  } else if (WHICH == 7) {  // smem atomicMax (64-bit) + read back, one address per warp
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
      atomicMax(&sm64[warp], ((unsigned long long)(unsigned)(v + i) << 32) | (unsigned)lane);
      __syncwarp();
      v = (int)(sm64[warp] >> 32) + 1;
      __syncwarp();
    }
  } else if (WHICH == 8) {  // independent redux throughput (4 in flight)
    int a = v, b = v + 1, c = v + 2, d = v + 3;
#pragma unroll 4
    for (int i = 0; i < N / 4; ++i) { a = redux_max_s32(a ^ lane) + i; b = redux_max_s32(b ^ lane) + i; c = redux_max_s32(c ^ lane) + i; d = redux_max_s32(d ^ lane) + i; }
    v = a + b + c + d;
  } else if (WHICH == 9) {  // FMNMX/FADD dependent chain (ALU latency reference)
#pragma unroll 8
    for (int i = 0; i < N; ++i) f = fmaxf(f - 1.5f, 0.25f) + f;
    v = (int)f;
  } else if (WHICH == 10) {  // match_any chain
#pragma unroll 8
    for (int i = 0; i < N; ++i) v = (int)__match_any_sync(0xffffffffu, (v + lane) & 3) + i;
  } else if (WHICH == 11) {  // STS + barrier + LDS round trip (the publish/collect pattern)
#pragma unroll 4
    for (int i = 0; i < N; ++i) { if (lane == 0) sm[warp + (i & 1) * 32] = v; __syncthreads(); v = sm[(lane & 15) + (i & 1) * 32] + i; }
  } else if (WHICH == 12) {  // redux u32 (unsigned) dependent chain
    unsigned u = (unsigned)v;
#pragma unroll 8
    for (int i = 0; i < N; ++i) u = redux_max_u32(u ^ lane) + i;
    v = (int)u;
  } else if (WHICH == 13) {  // vote.any predicate -> uniform branch chain (ballot + popc compare)
#pragma unroll 8
    for (int i = 0; i < N; ++i) { unsigned mk = __ballot_sync(0xffffffffu, v == lane + i); v += (mk & (mk - 1)) ? 3 : 1; }
  }
  long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
  if (v == 0x7fffffff) sink[0] = v;
}

template <int W>
void run(const char* name, int warps) {
  long long* d; int* s;
  cudaMalloc(&d, 32 * 8 * sizeof(long long)); cudaMalloc(&s, 4);
  k<W><<<1, warps * 32>>>(d, s, 5); cudaDeviceSynchronize();
  k<W><<<1, warps * 32>>>(d, s, 5); cudaDeviceSynchronize();
  long long h[32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-46s warps=%2d  %7.1f cycles/op (warp0 %7.1f)\n", name, warps, (double)mx / N, (double)h[0] / N);
  cudaFree(d); cudaFree(s);
}

int main() {
  for (int w : {1, 4, 16}) {
    run<0>("redux.sync.max.s32 dependent", w);
    run<12>("redux.sync.max.u32 dependent", w);
    run<8>("redux.sync 4 independent (per op)", w);
    run<1>("ballot dependent", w);
    run<13>("ballot + uniform select", w);
    run<2>("shfl.idx dependent", w);
    run<3>("ffs dependent", w);
    run<4>("LDS pointer chase", w);
    run<5>("__syncthreads", w);
    run<6>("shfl butterfly max (5 steps)", w);
    run<7>("ATOMS.MAX.64 + LDS", w);
    run<9>("FMNMX+FADD+FADD chain (3 ALU ops)", w);
    run<10>("match.any dependent", w);
    run<11>("STS + bar + LDS", w);
  }
  return 0;
}

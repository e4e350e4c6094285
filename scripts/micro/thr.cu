// Issue-throughput micro-benchmark of the epilogue instructions (B200, sm_100a): cycles per warp-instruction per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#define ITER 256
template <int W>
__global__ void k(long long* out, float* sink, float seed) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + lane + i * 0.37f;
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = 0x3c003c00u + i + lane;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (W == 0) {  // FFMA (16 independent)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 1.0001f, 0.5f);
    } else if (W == 1) {  // FADD2 (8 independent pairs)
#pragma unroll
      for (int i = 0; i < 16; i += 2) { float2 r = __fadd2_rn(make_float2(a[i], a[i + 1]), make_float2(0.5f, 0.25f)); a[i] = r.x; a[i + 1] = r.y; }
    } else if (W == 2) {  // FFMA2
#pragma unroll
      for (int i = 0; i < 16; i += 2) { float2 r = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f)); a[i] = r.x; a[i + 1] = r.y; }
    } else if (W == 3) {  // F2FP pack (8 independent) -- result fed back through a cheap int->float move
#pragma unroll
      for (int i = 0; i < 16; i += 2) { __half2 p = __floats2half2_rn(a[i], a[i + 1]); h[i >> 1] ^= *reinterpret_cast<uint32_t*>(&p); }
    } else if (W == 4) {  // HMNMX2 (8 independent)
#pragma unroll
      for (int i = 0; i < 8; ++i) { __half2 p = __hmax2(*reinterpret_cast<__half2*>(&h[i]), __float2half2_rn(0.f)); h[i] = *reinterpret_cast<uint32_t*>(&p) + 1; }
    } else if (W == 5) {  // FMNMX3 chains (4 independent)
#pragma unroll
      for (int i = 0; i < 16; i += 4) a[i] = fmaxf(fmaxf(a[i], a[i + 1]), a[i + 2]) + 0.f * a[i + 3];
    } else if (W == 6) {  // FMNMX (16 independent)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaxf(a[i], seed * i);
    } else if (W == 7) {  // LOP3 / IADD (16 independent int ops)
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = (h[i] ^ (h[(i + 1) & 7] >> 3)) + 7;
    } else if (W == 8) {  // HADD2.F32 conversions half2 -> float2 (8 independent)
#pragma unroll
      for (int i = 0; i < 8; ++i) { float2 f = __half22float2(*reinterpret_cast<__half2*>(&h[i])); a[2 * i] += f.x; a[2 * i + 1] += f.y; }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i]; uint32_t x = 0; for (int i = 0; i < 8; ++i) x ^= h[i];
  if (s == 1.2345f || x == 77) sink[0] = s + x;
  if (lane == 0) out[warp] = t1 - t0;
}
template <int W> void run(const char* name, int nops, int warps) {
  long long* d; float* s; cudaMalloc(&d, 64 * 8); cudaMalloc(&s, 4);
  k<W><<<1, warps * 32>>>(d, s, 1.5f); cudaDeviceSynchronize(); k<W><<<1, warps * 32>>>(d, s, 1.5f); cudaDeviceSynchronize();
  long long h[64]; cudaMemcpy(h, d, 8 * warps, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
  double per_smsp_warps = warps / 4.0;
  printf("%-28s warps/SMSP=%4.1f : %6.2f cycles per warp-instr per SMSP\n", name, per_smsp_warps, (double)mx / ITER / nops / (per_smsp_warps < 1 ? 1 : per_smsp_warps));
  cudaFree(d); cudaFree(s);
}
int main() {
  for (int w : {4, 16, 32}) {
    run<0>("FFMA", 16, w); run<1>("FADD2", 8, w); run<2>("FFMA2", 8, w); run<3>("F2FP.PACK_AB (+LOP3)", 8, w); run<4>("HMNMX2 (+IADD)", 8, w);
    run<5>("FMNMX3 (+FFMA)", 4, w); run<6>("FMNMX", 16, w); run<7>("LOP3+IADD (2 ops)", 8, w); run<8>("half2->float2 (+2 FADD)", 8, w);
  }
  return 0;
}

// Skeleton of one FPS round (no point work): what do the block-wide barrier + the replicated end-of-round reduction
// cost with 16 warps on one SM, and what would a scout/owner hand-off over mbarriers cost instead?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o round round.cu ; run: ./round
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
__device__ __forceinline__ int redux_max_s32(int v) { int r; asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ unsigned redux_max_u32(unsigned v) { unsigned r; asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) { while (!mbar_try(b, par)) {} }

template <int WHICH>
__global__ void __launch_bounds__(544, 1) k(long long* out, float* sink, int seed) {
  extern __shared__ __align__(16) unsigned char dyn[];   // forces 1 CTA/SM and the big carve-out like the real kernel
  __shared__ int s_whi[2][16];
  __shared__ unsigned s_wkey[2][16];
  __shared__ float4 s_rec[2][128];
  __shared__ float4 s_box[256];
  __shared__ unsigned s_words[2][4];
  __shared__ float4 s_pick[2];
  __shared__ uint64_t bar_work[2], bar_done[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 256; i += blockDim.x) { s_box[i] = make_float4(i * 0.01f, i * 0.02f, i * 0.03f, 0.f); s_rec[i >> 7][i & 127] = make_float4(i * 0.5f, i * 0.25f, i * 0.125f, 0.f); }
  if (tid < 32) { s_whi[tid >> 4][tid & 15] = tid * 3 + seed; s_wkey[tid >> 4][tid & 15] = (unsigned)(tid * 2 + 1) << 1; }
  if (tid == 0) { mbar_init(&bar_work[0], 1); mbar_init(&bar_work[1], 1); mbar_init(&bar_done[0], 16); mbar_init(&bar_done[1], 16); }
  __syncthreads();
  float lx = 0.1f * seed, ly = 0.2f, lz = 0.3f;
  int chi = seed;
  const int tb = (warp & 15) * 8 + (lane & 7);
  long long t0 = clock64();
  if (WHICH == 0) {  // barrier only
    for (int r = 1; r < ITERS; ++r) __syncthreads();
  } else if (WHICH == 1 || WHICH == 2) {  // barrier + replicated 16-entry reduce (+ bound section when WHICH == 2)
    for (int r = 1; r < ITERS; ++r) {
      const int par = r & 1;
      if (WHICH == 2) {
        const float4 blo = s_box[tb], bhi = s_box[128 + tb];
        const float gx = fmaxf(fmaxf(blo.x - lx, lx - bhi.x), 0.f), gy = fmaxf(fmaxf(blo.y - ly, ly - bhi.y), 0.f), gz = fmaxf(fmaxf(blo.z - lz, lz - bhi.z), 0.f);
        const float bound = gz * gz + (gx * gx + gy * gy);
        const unsigned mask = __ballot_sync(0xffffffffu, lane < 8 && bound < __int_as_float(chi));
        if (mask == 0x12345u && lane == 8) s_whi[par][warp] = r;   // never true: keeps the ballot live
        __syncwarp();
      }
      __syncthreads();
      const int hw = lane < 16 ? s_whi[par][lane] : (int)0x80000000;
      const unsigned kw = lane < 16 ? s_wkey[par][lane] : 0u;
      const int whi = redux_max_s32(hw);
      const unsigned wk = redux_max_u32(hw == whi ? kw : 0u);
      const float4 rec = s_rec[par][(wk >> 1) & 127u];
      lx = rec.x + lx * 1e-9f; ly = rec.y; lz = rec.z;
    }
  } else if (WHICH == 3) {  // as 2, but only warps 0..3 (one per SM sub-partition) reduce; the others wait on a 2nd barrier
    for (int r = 1; r < ITERS; ++r) {
      const int par = r & 1;
      const float4 blo = s_box[tb], bhi = s_box[128 + tb];
      const float gx = fmaxf(fmaxf(blo.x - lx, lx - bhi.x), 0.f), gy = fmaxf(fmaxf(blo.y - ly, ly - bhi.y), 0.f), gz = fmaxf(fmaxf(blo.z - lz, lz - bhi.z), 0.f);
      const float bound = gz * gz + (gx * gx + gy * gy);
      const unsigned mask = __ballot_sync(0xffffffffu, lane < 8 && bound < __int_as_float(chi));
      if (mask == 0x12345u && lane == 8) s_whi[par][warp] = r;
      __syncwarp();
      __syncthreads();
      if (warp == 0) {
        const int hw = lane < 16 ? s_whi[par][lane] : (int)0x80000000;
        const unsigned kw = lane < 16 ? s_wkey[par][lane] : 0u;
        const int whi = redux_max_s32(hw);
        const unsigned wk = redux_max_u32(hw == whi ? kw : 0u);
        const float4 rec = s_rec[par][(wk >> 1) & 127u];
        if (lane == 0) s_pick[par] = rec;
      }
      __syncthreads();
      const float4 p = s_pick[par];
      lx = p.x + lx * 1e-9f; ly = p.y; lz = p.z;
    }
  } else if (WHICH == 4) {  // scout (warp 16) / 16 owners over mbarriers: work -> done ping-pong, scout does 128 bound tests + reduce
    if (warp == 16) {
      float4 blo[4], bhi[4];
      int ch[4];
      for (int j = 0; j < 4; ++j) { blo[j] = s_box[lane * 4 + j]; bhi[j] = s_box[128 + lane * 4 + j]; ch[j] = seed + j; }
      for (int r = 1; r < ITERS; ++r) {
        const int par = r & 1;
        const uint32_t ph = (uint32_t)(((r - 1) >> 1) & 1);
        unsigned w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float gx = fmaxf(fmaxf(blo[j].x - lx, lx - bhi[j].x), 0.f), gy = fmaxf(fmaxf(blo[j].y - ly, ly - bhi[j].y), 0.f), gz = fmaxf(fmaxf(blo[j].z - lz, lz - bhi[j].z), 0.f);
          const float bound = gz * gz + (gx * gx + gy * gy);
          w[j] = __ballot_sync(0xffffffffu, bound < __int_as_float(ch[j]));
        }
        if (lane < 4) s_words[par][lane] = lane == 0 ? w[0] : lane == 1 ? w[1] : lane == 2 ? w[2] : w[3];
        if (lane == 4) s_pick[par] = make_float4(lx, ly, lz, 0.f);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_work[par]);
        mbar_wait(&bar_done[par], ph);
        const int hw = lane < 16 ? s_whi[par][lane] : (int)0x80000000;
        const unsigned kw = lane < 16 ? s_wkey[par][lane] : 0u;
        const int whi = redux_max_s32(hw);
        const unsigned wk = redux_max_u32(hw == whi ? kw : 0u);
        const float4 rec = s_rec[par][(wk >> 1) & 127u];
        lx = rec.x + lx * 1e-9f; ly = rec.y; lz = rec.z;
      }
    } else {
      for (int r = 1; r < ITERS; ++r) {
        const int par = r & 1;
        const uint32_t ph = (uint32_t)(((r - 1) >> 1) & 1);
        mbar_wait(&bar_work[par], ph);
        const unsigned wd = s_words[par][warp >> 2];
        const float4 p = s_pick[par];
        if (((wd >> ((warp & 3) * 8)) & 0xffu) == 0x5au && lane == 8) s_whi[par][warp] = (int)p.x;   // (almost) never: no rescans
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_done[par]);
      }
    }
  }
  long long t1 = clock64();
  if (tid == 0 || tid == 512) out[tid >> 9] = (t1 - t0);
  if (lx == 123.f) sink[0] = lx + ly + lz + chi;
}

template <int W>
void run(const char* name, int threads) {
  long long* d; float* s;
  cudaMalloc(&d, 16); cudaMalloc(&s, 16);
  cudaMemset(d, 0, 16);
  cudaFuncSetAttribute(k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<W><<<1, threads, 200 * 1024>>>(d, s, 3);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-70s %8.1f cycles/round  (scout view %8.1f) %s\n", name, (double)h[0] / (ITERS - 1), (double)h[1] / (ITERS - 1), e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout); cudaFree(d); cudaFree(s);
}

int main() {
  run<0>("barrier only, 16 warps", 512);
  run<1>("barrier + replicated 16-entry reduce (2 LDS, 2 redux, LDS.128)", 512);
  run<2>("  + bound section (2 LDS.128, 12 FP, ballot, syncwarp)", 512);
  run<3>("bound + barrier + reduce by warp 0 only + 2nd barrier + LDS", 512);
  run<4>("scout warp + 16 owners over mbarriers (no rescans)", 544);
  return 0;
}

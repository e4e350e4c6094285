// Farthest point sampling for B200: one thread-block CLUSTER per cloud, everything on-chip.
//
// Reference: farthestpointsamplingKernel, tf_ops/sampling/tf_sampling_g.cu:105-170 (launch <<<32,512>>> :204):
// one 512-thread block per cloud, running min-distances `temp` in GLOBAL memory re-read and conditionally
// re-written every round, a 9-level shared-memory tree with 10 block barriers per round.
//
// Here: the cloud is spread over the registers of a cluster of CTAs (CL CTAs x T threads x P points/thread), the
// running min-distance of every point stays in a register for all m-1 rounds, and one round costs
//   scan (P fused distance updates per thread)  ->  2x redux.sync arg-max inside each warp
//   ->  each warp's champion {key, xyz} is pushed into EVERY CTA's shared memory with st.async (DSMEM) that
//       completes a transaction barrier (mbarrier) there  ->  every warp waits on its own CTA's barrier and reduces
//       the CL*W champions locally.
// No __syncthreads, no cluster barrier and no global-memory round trip inside the loop.
//
// Bit-exactness.  d = fmaf(dz,dz,fmaf(dx,dx,dy*dy)) with d* = point - last (the contraction nvcc emits for the
// reference), temp = min(d, temp) from 1e38f.  The reference's winner is the candidate with maximal
// temp, ties to the smallest (k mod 512), then the smallest k (per-thread strict '>' over k = t, t+512, ...; tree
// keeps the lower slot).  All temps are >= +0 so their bit patterns order as unsigned integers; the 64-bit key
//   (temp_bits << 32) | (0xFFFFFFFF - (((k & 511) << 20) | (k >> 9)))          (n <= 2^20)
// reduced with an integer max reproduces exactly that rule under ANY reduction topology.
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
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 24)) __trap();
}
// remote (DSMEM) stores that complete `bytes` on the destination CTA's mbarrier
__device__ __forceinline__ void st_async_v2(uint32_t raddr, uint32_t a, uint32_t b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(raddr),
               "r"(a), "r"(b), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                            uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   raddr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ uint32_t redux_max(uint32_t v) {
  uint32_t r;
  asm volatile("redux.sync.max.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}

__device__ __forceinline__ uint32_t tie_key(int k) { return 0xFFFFFFFFu - ((((uint32_t)k & 511u) << 20) | ((uint32_t)k >> 9)); }
__device__ __forceinline__ int tie_key_to_index(uint32_t key) {
  uint32_t t = 0xFFFFFFFFu - key;
  return (int)(((t & 0xFFFFFu) << 9) | (t >> 20));
}

// T threads per CTA, P points per thread, CL CTAs per cluster (launch attribute); W = T/32 warps.
template <int T, int P>
__global__ void __launch_bounds__(T) fps_cluster_kernel(int n, int m, int CL, const float* __restrict__ xyz,
                                                        int* __restrict__ out) {
  constexpr int W = T / 32;
  constexpr int MAXCH = 16 * W;  // champions per round at the largest cluster size (16)
  // per round parity: champion keys (8 B) and coordinates (16 B), one slot per (cta, warp) of the cluster
  __shared__ __align__(16) uint2 s_key[2][MAXCH];
  __shared__ __align__(16) float4 s_xyz[2][MAXCH];
  __shared__ __align__(8) uint64_t s_bar[2];
  extern __shared__ __align__(16) float4 s_pts[];  // [T*P] this CTA's points (x,y,z,-) for the champion's coordinates

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int cloud = blockIdx.x / CL;
  const float* pc = xyz + (size_t)cloud * n * 3;
  int* oc = out + (size_t)cloud * m;
  const int nch = CL * W;  // champions per round

  float px[P], py[P], pz[P], td[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    int k = (i * CL + (int)rank) * T + tid;
    if (k < n) {
      px[i] = pc[(size_t)k * 3 + 0];
      py[i] = pc[(size_t)k * 3 + 1];
      pz[i] = pc[(size_t)k * 3 + 2];
    } else {
      px[i] = py[i] = pz[i] = 0.f;
    }
    td[i] = 1e38f;  // tf_sampling_g.cu:118
    s_pts[i * T + tid] = make_float4(px[i], py[i], pz[i], 0.f);
  }
  if (tid == 0) {
    mbar_init(smem_addr(&s_bar[0]), 1);
    mbar_init(smem_addr(&s_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // round r uses barrier r&1: arm round 1 and round 2 now, round r+2 right after round r completes
    const uint32_t tx = (uint32_t)nch * 24u;
    mbar_arrive_expect_tx(smem_addr(&s_bar[1]), tx);
    mbar_arrive_expect_tx(smem_addr(&s_bar[0]), tx);
  }
  if (rank == 0 && tid == 0) oc[0] = 0;  // first sample is index 0 (:114-116)
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers are initialised before any remote st.async can land

  float lx = pc[0], ly = pc[1], lz = pc[2];  // coordinates of the last pick (index 0)
  const uint32_t my_slot = rank * W + warp;

  for (int r = 1; r < m; ++r) {
    const int par = r & 1;
    // ---- scan: update running min-distances, per-thread champion -------------------------------------------
    uint32_t bhi = 0, blo = 0;
    int bi = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      int k = (i * CL + (int)rank) * T + tid;
      float d = d2_ref_gpu(px[i] - lx, py[i] - ly, pz[i] - lz);
      td[i] = fminf(d, td[i]);
      uint32_t hi = __float_as_uint(td[i]);
      uint32_t lo = tie_key(k);
      bool valid = k < n;
      bool better = valid && (hi > bhi || (hi == bhi && lo > blo));
      if (better) { bhi = hi; blo = lo; bi = i; }
    }
    // ---- warp arg-max: two redux.sync ---------------------------------------------------------------------
    uint32_t whi = redux_max(bhi);
    uint32_t wlo = redux_max(bhi == whi ? blo : 0u);
    // the champion lane (unique: tie keys are unique per point; an all-invalid warp sends key 0 from lane 0)
    bool champ = (bhi == whi) && (blo == wlo);
    unsigned cm = __ballot_sync(0xffffffffu, champ);
    if (lane == __ffs(cm) - 1) {
      float4 c = s_pts[bi * T + tid];
      const uint32_t kaddr = smem_addr(&s_key[par][my_slot]);
      const uint32_t xaddr = smem_addr(&s_xyz[par][my_slot]);
      const uint32_t baddr = smem_addr(&s_bar[par]);
      for (int dst = 0; dst < CL; ++dst) {
        uint32_t rb = mapa(baddr, dst);
        st_async_v2(mapa(kaddr, dst), wlo, whi, rb);
        st_async_v4(mapa(xaddr, dst), __float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z), 0u, rb);
      }
    }
    // ---- wait for all champions of this round to land in OUR shared memory ---------------------------------
    const uint32_t parity = (uint32_t)(((r - 1) >> 1) & 1);
    mbar_wait(smem_addr(&s_bar[par]), parity);
    if (tid == 0 && r + 2 < m) mbar_arrive_expect_tx(smem_addr(&s_bar[par]), (uint32_t)nch * 24u);  // arm round r+2
    // ---- every warp reduces the nch champions (identical result everywhere) ---------------------------------
    uint32_t chi = 0, clo = 0;
    int cs = 0;
    for (int s = lane; s < nch; s += 32) {
      uint2 kv = s_key[par][s];
      if (kv.y > chi || (kv.y == chi && kv.x > clo)) { chi = kv.y; clo = kv.x; cs = s; }
    }
    uint32_t ghi = redux_max(chi);
    uint32_t glo = redux_max(chi == ghi ? clo : 0u);
    unsigned gm = __ballot_sync(0xffffffffu, chi == ghi && clo == glo);
    int gs = __shfl_sync(0xffffffffu, cs, __ffs(gm) - 1);
    float4 w = s_xyz[par][gs];
    lx = w.x; ly = w.y; lz = w.z;
    if (rank == 0 && tid == 0) oc[r] = tie_key_to_index(glo);
  }
  cluster_sync_all();  // nobody exits while a peer may still write into its shared memory
}

template <int T, int P>
static int launch_fps(int b, int n, int m, int CL, const float* xyz, int* out, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * CL));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = (size_t)T * P * sizeof(float4);
  VNB_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<T, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (CL > 8) VNB_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<T, P>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  VNB_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel<T, P>, n, m, CL, xyz, out));
  return check_launch("farthest_point_sample");
}

}  // namespace vnb

using namespace vnb;

extern "C" int vnb_farthest_point_sample(int b, int n, int m, const float* xyz, int* out_idx, void* stream) {
  VNB_REQUIRE(m > 0, "FarthestPointSample expects positive npoint");                     // tf_sampling.cpp:99
  VNB_REQUIRE(b >= 0 && n >= 1, "FarthestPointSample expects (batch_size,num_points,3) inp shape");  // :105
  VNB_REQUIRE(n <= 65536, "farthest_point_sample: n = %d exceeds the on-chip limit of 65536 points per cloud", n);
  if (b == 0) return VNB_OK;
  cudaStream_t st = as_stream(stream);
  // choose (cluster size, points per thread) so that CL * 256 * P >= n with the least padding
  constexpr int T = 256;
  if (n <= T * 2) return launch_fps<T, 2>(b, n, m, 1, xyz, out_idx, st);
  if (n <= T * 4) return launch_fps<T, 4>(b, n, m, 1, xyz, out_idx, st);
  if (n <= T * 8) return launch_fps<T, 8>(b, n, m, 1, xyz, out_idx, st);
  if (n <= 4 * T * 4) return launch_fps<T, 4>(b, n, m, 4, xyz, out_idx, st);
  if (n <= 8 * T * 4) return launch_fps<T, 4>(b, n, m, 8, xyz, out_idx, st);
  if (n <= 8 * T * 8) return launch_fps<T, 8>(b, n, m, 8, xyz, out_idx, st);
  if (n <= 8 * T * 10) return launch_fps<T, 10>(b, n, m, 8, xyz, out_idx, st);
  if (n <= 8 * T * 16) return launch_fps<T, 16>(b, n, m, 8, xyz, out_idx, st);
  return launch_fps<T, 16>(b, n, m, 16, xyz, out_idx, st);
}

// Farthest point sampling for B200: one thread-block CLUSTER per cloud, everything on-chip.
//
// Reference: farthestpointsamplingKernel, tf_ops/sampling/tf_sampling_g.cu:105-170 (launch <<<32,512>>> :204):
// one 512-thread block per cloud, running min-distances `temp` in GLOBAL memory re-read and conditionally
// re-written every round, a 9-level shared-memory tree with 10 block barriers per round.
//
// Here the cloud is spread over the REGISTERS of a cluster of CL CTAs x T threads x P points/thread; the running
// min-distance of every point stays in a register for all m-1 rounds.  One round =
//   scan (P fused distance updates per thread, first-max per thread)
//   -> 2x redux.sync arg-max inside each warp -> champions to shared memory, ONE __syncthreads
//   -> every warp reduces the W warp champions (2x redux.sync) to the CTA champion
//   -> (CL > 1) warp 0 pushes {key, xyz} of the CTA champion into EVERY CTA of the cluster through distributed shared
//      memory (one lane per destination), synchronised either by a cluster barrier (mode 0) or by tagged slots that the
//      receivers poll in their own shared memory (mode 1); every warp then reduces the CL champions locally.
// No global-memory traffic inside the loop except the 4-byte result of each round.
//
// Bit-exactness.  d = fmaf(dz,dz,fmaf(dx,dx,dy*dy)) with d* = point - last (the contraction nvcc emits for the
// reference), temp = min(d, temp) from 1e38f.  The reference's winner is the candidate with maximal
// temp, ties to the smallest (k mod 512), then the smallest k (per-thread strict '>' over k = t, t+512, ...; the tree
// keeps the lower slot).  All temps are >= +0 so their bit patterns order as unsigned integers; the key
//   (temp_bits, 0x8000 | (0x7FFF - (((k & 511) << 6) | (k >> 9))))                (n <= 32768)
// compared lexicographically and reduced with integer max reproduces exactly that rule under ANY reduction topology.
// A thread owns points k = (i*CL + rank)*T + tid with CL*T a multiple of 512, so all its points share (k mod 512) and
// its tie key decreases with i: "first strict maximum over i" is the reference's per-thread rule.
//
// Nested levels (SA2..SA4, proposal) run FPS on the previous level's FPS output, whose result is the identity prefix
// 0..m-1 unless exact float ties reorder it.  vnb_farthest_point_sample_nested proves that per cloud
// with a fully PARALLEL check (every round's arg-max condition is independent once the picks are hypothesised) and
// only falls back to the sequential kernel for clouds where the proof fails — the output is always bit-identical.
#include "fps_common.cuh"

#include <string_view>

namespace vnb {

extern int g_bq_variant;
extern int g_bq_grid_min_n;
extern int g_sa_variant;
extern int g_sa_sms;
extern int g_sa_split;
extern int g_sa_min_tpc;
int g_fps_mode = 1;   // 0: cluster barrier per round, 1: CTA champions in tagged slots + polling (default)
int g_fps_cl = 0;     // 0: automatic cluster size, else forced (power of two <= 16)
int g_fps_threads = 256;  // threads per CTA of the cluster kernel (256 / 512 / 1024)
int g_fps_variant = 1;    // 0: register/cluster kernel only; 1: bucket-pruned single-CTA kernel for 2048 < n <= 20480;
                          // 2: bucket-pruned kernel for every n <= 20480

int fps_pruned_capacity();
int launch_fps_pruned(int b, int n, int m, const float* xyz, int* out, const int* flags, int* tie_out, int tie_rounds,
                      cudaStream_t st);

// T threads per CTA, P points per thread, CL CTAs per cluster (launch attribute, CL*T % 512 == 0); W = T/32 warps.
// MODE 0: cluster barrier per round; MODE 1: tagged slots, receivers poll their own shared memory.
// `flags` (may be null): per-cloud "already done" flags — a cluster whose cloud is flagged exits immediately.
template <int T, int P, int MODE, bool SP, bool PROF = false>
__global__ void __launch_bounds__(T) fps_cluster_kernel(int n, int m, int CL, const float* __restrict__ xyz,
                                                        int* __restrict__ out, const int* __restrict__ flags,
                                                        long long* __restrict__ prof = nullptr) {
  constexpr int W = T / 32;
  __shared__ __align__(16) uint2 s_wkey[2][W];    // warp champions: {tie key, temp bits}
  __shared__ int s_wslot[2][W];                   //                 their slot in s_pts
  __shared__ __align__(16) uint4 s_xa[2][16];     // CTA champions of the cluster: {temp bits, tag<<16 | tie, x, y}
  __shared__ __align__(16) uint2 s_xb[2][16];     //                               {z, tag}
  extern __shared__ __align__(16) float4 s_pts[];  // [T*P] this CTA's points (x,y,z,-)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int cloud = blockIdx.x / CL;
  if (flags != nullptr && flags[cloud] != 0) return;  // whole cluster takes the same branch
  const float* pc = xyz + (size_t)cloud * n * 3;
  int* oc = out + (size_t)cloud * m;

  float px[P], py[P], pz[P], td[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    int k = (i * CL + (int)rank) * T + tid;
    if (k < n) {
      px[i] = pc[(size_t)k * 3 + 0];
      py[i] = pc[(size_t)k * 3 + 1];
      pz[i] = pc[(size_t)k * 3 + 2];
      td[i] = 1e38f;  // tf_sampling_g.cu:118
    } else {
      px[i] = py[i] = pz[i] = 0.f;
      td[i] = -1.f;  // padding: min(d,-1) stays -1 and never beats the strict '>' against best = -1
    }
    if (SP) s_pts[i * T + tid] = make_float4(px[i], py[i], pz[i], 0.f);
  }
  if (tid < 32) {
    s_xa[0][tid & 15] = make_uint4(0, 0, 0, 0); s_xa[1][tid & 15] = make_uint4(0, 0, 0, 0);
    s_xb[0][tid & 15] = make_uint2(0, 0); s_xb[1][tid & 15] = make_uint2(0, 0);
  }
  if (rank == 0 && tid == 0) oc[0] = 0;  // first sample is index 0 (:114-116)
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every CTA's slots are initialised before any remote store can land

  float lx = pc[0], ly = pc[1], lz = pc[2];  // coordinates of the last pick (index 0)
  const uint32_t key0 = tie_key((int)rank * T + tid);   // tie key of this thread's point i = 0
  const uint32_t kstep = (uint32_t)(CL * T) >> 9;       // it decreases by this much per i

  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt0 = 0;  // PROF only: cycles per phase, warp 0 of every CTA
#define FPS_TICK(i)                                 \
  if (PROF) {                                       \
    long long _t = clock64();                       \
    pacc[i] += _t - pt0;                            \
    pt0 = _t;                                       \
  }
  // destination slots of this CTA's champion in CTA `lane` of the cluster (lanes < CL), both parities
  uint32_t ra0 = 0, ra1 = 0, rb0 = 0, rb1 = 0;
  if (CL > 1 && lane < CL) {
    ra0 = mapa(smem_addr(&s_xa[0][rank]), lane); ra1 = mapa(smem_addr(&s_xa[1][rank]), lane);
    rb0 = mapa(smem_addr(&s_xb[0][rank]), lane); rb1 = mapa(smem_addr(&s_xb[1][rank]), lane);
  }
  for (int r = 1; r < m; ++r) {
    const int par = r & 1;
    if (PROF) pt0 = clock64();
    // ---- scan: update running min-distances; first strict maximum over i ------------------------------------
    float best = -1.f;
    int bi = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      float d = d2_ref_gpu(px[i] - lx, py[i] - ly, pz[i] - lz);
      td[i] = fminf(d, td[i]);
      if (td[i] > best) { best = td[i]; bi = i; }
    }
    FPS_TICK(0)  // scan
    const bool has = best >= 0.f;
    const uint32_t bhi = has ? __float_as_uint(best) : 0u;
    const uint32_t blo = has ? key0 - (uint32_t)bi * kstep : 0u;
    // ---- warp arg-max: two redux.sync -----------------------------------------------------------------------
    const uint32_t whi = redux_max(bhi);
    const uint32_t wlo = redux_max(bhi == whi ? blo : 0u);
    const unsigned cm = __ballot_sync(0xffffffffu, bhi == whi && blo == wlo);
    if (lane == __ffs(cm) - 1) {
      s_wkey[par][warp] = make_uint2(wlo, whi);
      if (SP) s_wslot[par][warp] = bi * T + tid;
    }
    FPS_TICK(1)  // warp arg-max + shared store
    __syncthreads();
    FPS_TICK(2)  // block barrier
    // ---- CTA champion (every warp computes it; no second barrier) -------------------------------------------
    uint2 kv = lane < W ? s_wkey[par][lane] : make_uint2(0u, 0u);
    uint32_t chi = redux_max(kv.y);
    uint32_t clo = redux_max(kv.y == chi ? kv.x : 0u);
    const unsigned wm = __ballot_sync(0xffffffffu, lane < W && kv.y == chi && kv.x == clo);
    const int cw = __ffs(wm) - 1;  // the warp that owns the CTA champion
    float4 cpt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (SP) cpt = s_pts[s_wslot[par][cw]];
    FPS_TICK(3)  // CTA champion
    if (CL > 1) {
      // ---- cluster exchange: the champion's own warp pushes {key, xyz} to every CTA (one lane per destination) --
      const uint32_t tag = (uint32_t)r & 0xFFFFu;
      if (warp == cw) {
        float sx = px[0], sy = py[0], sz = pz[0];  // coordinates straight from the champion lane's registers
#pragma unroll
        for (int i = 1; i < P; ++i)
          if (bi == i) { sx = px[i]; sy = py[i]; sz = pz[i]; }
        const int cl = __ffs(cm) - 1;
        sx = __shfl_sync(0xffffffffu, sx, cl); sy = __shfl_sync(0xffffffffu, sy, cl); sz = __shfl_sync(0xffffffffu, sz, cl);
        if (lane < CL) {
          st_cluster_v4(par ? ra1 : ra0, chi, (tag << 16) | clo, __float_as_uint(sx), __float_as_uint(sy));
          st_cluster_v2(par ? rb1 : rb0, __float_as_uint(sz), tag);
        }
      }
      FPS_TICK(4)  // remote stores issued (only the champion warp does work here)
      uint4 a = make_uint4(0, 0, 0, 0);
      uint2 bq = make_uint2(0, 0);
      if (MODE == 0) {
        cluster_sync_all();
        if (lane < CL) { a = s_xa[par][lane]; bq = s_xb[par][lane]; }
      } else {
        if (lane < CL) {
          uint32_t spin = 0;
          do {
            a = ld_volatile_v4(&s_xa[par][lane]);
            bq = ld_volatile_v2(&s_xb[par][lane]);
            if (++spin > (1u << 22)) trap_at(__LINE__);  // protocol bug: trap instead of hanging the GPU
          } while ((a.y >> 16) != tag || bq.y != tag);
        }
        __syncwarp();
      }
      FPS_TICK(5)  // wait for all CTA champions
      const uint32_t ahi = lane < CL ? a.x : 0u, alo = lane < CL ? (a.y & 0xFFFFu) : 0u;
      chi = redux_max(ahi);
      clo = redux_max(ahi == chi ? alo : 0u);
      const unsigned gm = __ballot_sync(0xffffffffu, lane < CL && ahi == chi && alo == clo);
      const int src = __ffs(gm) - 1;
      lx = __uint_as_float(__shfl_sync(0xffffffffu, a.z, src));
      ly = __uint_as_float(__shfl_sync(0xffffffffu, a.w, src));
      lz = __uint_as_float(__shfl_sync(0xffffffffu, bq.x, src));
    } else {
      lx = cpt.x; ly = cpt.y; lz = cpt.z;
    }
    if (rank == 0 && tid == 0) oc[r] = tie_key_to_index(clo);
    FPS_TICK(6)  // final reduce + broadcast
  }
#undef FPS_TICK
  if (PROF && prof != nullptr && warp == 0 && lane == 0 && cloud == 0)
    for (int i = 0; i < 8; ++i) prof[rank * 8 + i] = pacc[i];
  if (CL > 1) cluster_sync_all();  // nobody exits while a peer may still write into its shared memory
}

// ---------------------------------------------------------------------------------------------------------------------
// Parallel proof that FPS(xyz, m) == (0, 1, ..., m-1).  With picks 0..j-1 hypothesised, the running min-distance of
// point k at round j is T_j[k] = min_{i<j} d(k,i) (min is exact), and round j picks j iff
//   (T_j[k], tie(k)) < (T_j[j], tie(j))  for every k != j.
// Kernel A: R[j] = T_j[j] for j < m (thread j, loop over i < j; triangular, m^2/2 distance evaluations per cloud).
// Kernel B: thread k walks i = 0..m-2 keeping its running min and checks round j = i+1 against R[j].
// Induction over j makes the hypothesis true whenever all checks pass.
// Kernel B is split along the picks as well: warp s of a CTA owns segment s of the pick range for the CTA's 32 points
// (lane = point, so every shared-memory read is a broadcast).  Pass 1: minimum over the segment; the exclusive
// prefix-minimum over the earlier segments is the running min at the segment's start (min is exact and associative,
// so this equals the sequential value bit for bit); pass 2: the checked walk.  1.5x the instructions of one long walk,
// but 8x the threads and a serial chain 4x shorter — the single walk ran at ~3 warps per SM, bound by its own latency.
constexpr int VP = 32;  // points per CTA (one per lane)
constexpr int VS = 8;   // segments of the pick range (one per warp)
__global__ void __launch_bounds__(128) fps_prefix_r_kernel(int n, int m, const float* __restrict__ xyz,
                                                           float* __restrict__ R /* (b,m) */,
                                                           const int* __restrict__ hint) {
  extern __shared__ float sv[];  // xyz of picks 0 .. jend-1
  const int cloud = blockIdx.y;
  if (hint != nullptr && hint[cloud] >= m) return;  // proven by provenance (see vnb_farthest_point_sample_nested_hint)
  const float* pc = xyz + (size_t)cloud * n * 3;
  const int j0 = blockIdx.x * 128, jend = min(m, j0 + 128);
  for (int t = threadIdx.x; t < jend * 3; t += 128) sv[t] = pc[t];
  __syncthreads();
  const int j = j0 + threadIdx.x;
  if (j >= m) return;
  const float x = sv[j * 3], y = sv[j * 3 + 1], z = sv[j * 3 + 2];
  float t = 1e38f;
  for (int i = 0; i < j; ++i) t = fminf(t, d2_ref_gpu(x - sv[i * 3], y - sv[i * 3 + 1], z - sv[i * 3 + 2]));
  R[(size_t)cloud * m + j] = t;
}

__global__ void __launch_bounds__(VP * VS) fps_prefix_verify_kernel(int n, int m, const float* __restrict__ xyz,
                                                                     const float* __restrict__ R,
                                                                     int* __restrict__ fail /* per cloud, pre-zeroed */,
                                                                     const int* __restrict__ hint) {
  extern __shared__ float4 sq[];  // [m-1]: {xyz of pick i, R[i+1]} -> one broadcast LDS.128 per round
  __shared__ float s_min[VS][VP];
  const int cloud = blockIdx.y;
  if (hint != nullptr && hint[cloud] >= m) return;
  const float* pc = xyz + (size_t)cloud * n * 3;
  for (int i = threadIdx.x; i + 1 < m; i += VP * VS)
    sq[i] = make_float4(pc[(size_t)i * 3], pc[(size_t)i * 3 + 1], pc[(size_t)i * 3 + 2], R[(size_t)cloud * m + i + 1]);
  __syncthreads();
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int k = blockIdx.x * VP + lane;
  const bool live = k < n;
  const int kk = live ? k : 0;
  const float x = pc[(size_t)kk * 3], y = pc[(size_t)kk * 3 + 1], z = pc[(size_t)kk * 3 + 2];
  const int L = (m - 1 + VS - 1) / VS;
  const int i0 = seg * L, i1 = min(m - 1, i0 + L);
  // pass 1: minimum over this segment (the last segment's is never needed)
  float t = 1e38f;
  if (seg + 1 < VS) {
#pragma unroll 4
    for (int i = i0; i < i1; ++i) {
      const float4 p = sq[i];
      t = fminf(t, d2_ref_gpu(x - p.x, y - p.y, z - p.z));
    }
  }
  s_min[seg][lane] = t;
  __syncthreads();
  t = 1e38f;  // tf_sampling_g.cu:118
  for (int s2 = 0; s2 < seg; ++s2) t = fminf(t, s_min[s2][lane]);
  // pass 2: t == T_{i0}[k]; after iteration i, t == T_{i+1}[k], checked against round j = i + 1
  const uint32_t tk = tie_key(k);
  bool bad = false;
#pragma unroll 4
  for (int i = i0; i < i1; ++i) {
    const float4 p = sq[i];
    t = fminf(t, d2_ref_gpu(x - p.x, y - p.y, z - p.z));
    const int j = i + 1;
    if (k != j && (t > p.w || (t == p.w && tk > tie_key(j)))) bad = true;
  }
  if (live && bad) atomicOr(&fail[cloud], 1);
}

__global__ void fps_identity_kernel(int b, int m, const int* __restrict__ fail, int* __restrict__ out, int* __restrict__ done,
                                    int* __restrict__ proven_rounds) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < b) done[t] = fail[t] ? 0 : 1;
  if (t < b && proven_rounds != nullptr) proven_rounds[t] = fail[t] ? 0 : m;
  if (t >= b * m) return;
  if (!fail[t / m]) out[t] = t % m;
}

template <int T, int P, int MODE, bool SP>
static int launch_fps(int b, int n, int m, int CL, const float* xyz, int* out, const int* flags, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * CL));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = SP ? (size_t)T * P * sizeof(float4) : 0;
  if (cfg.dynamicSmemBytes > 48 * 1024)
    VNB_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<T, P, MODE, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)cfg.dynamicSmemBytes));
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (CL > 8)
    VNB_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<T, P, MODE, SP>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  long long* noprof = nullptr;
  VNB_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel<T, P, MODE, SP>, n, m, CL, xyz, out, flags, noprof));
  return check_launch("farthest_point_sample");
}

long long* g_fps_prof = nullptr;  // debugging: device buffer of 16*8 cycle counters (vnb_debug_fps_profile)

// Instances of the register / cluster sampler that are kept (it is the general fallback — clouds beyond the pruned
// kernel's 20 480 points, the small nested levels when their identity proof fails — not the benchmarked path):
//   one CTA of 512 threads, xyz in shared memory, for n <= 8192;  clusters of 256-thread CTAs (tagged slots) above.
template <int T, int P>
static int launch_fps_mode(int b, int n, int m, int CL, const float* xyz, int* out, const int* flags, cudaStream_t st) {
  if (CL == 1) return launch_fps<T, P, 1, true>(b, n, m, CL, xyz, out, flags, st);
  return launch_fps<T, P, 1, false>(b, n, m, CL, xyz, out, flags, st);
}

template <int T>
static int fps_dispatch_t(int b, int n, int m, int CL, const float* xyz, int* out, const int* flags, cudaStream_t st) {
  const int per = (n + CL * T - 1) / (CL * T);  // points per thread needed
  if (per <= 1) return launch_fps_mode<T, 1>(b, n, m, CL, xyz, out, flags, st);
  if (per <= 2) return launch_fps_mode<T, 2>(b, n, m, CL, xyz, out, flags, st);
  if (per <= 4) return launch_fps_mode<T, 4>(b, n, m, CL, xyz, out, flags, st);
  if (per <= 8) return launch_fps_mode<T, 8>(b, n, m, CL, xyz, out, flags, st);
  if (per <= 16) return launch_fps_mode<T, 16>(b, n, m, CL, xyz, out, flags, st);
  if constexpr (T <= 256) {
    if (per <= 20) return launch_fps_mode<T, 20>(b, n, m, CL, xyz, out, flags, st);
  }
  return set_err(VNB_ERR_INVALID, "farthest_point_sample: %d points per thread do not fit (cluster %d x %d threads)", per, CL, T);
}

// Geometry of the fallback sampler: n <= 8192 -> one CTA of 512 threads; above, clusters of 256-thread CTAs
// (g_fps_cl CTAs per cluster, or as many as ~10 points per thread need; at most 16).
static int fps_dispatch(int b, int n, int m, const float* xyz, int* out, const int* flags, cudaStream_t st,
                        int* tie_out = nullptr, int tie_rounds = 0) {
  if (n <= fps_pruned_capacity() && ((g_fps_variant == 1 && n > 2048) || g_fps_variant == 2))
    return launch_fps_pruned(b, n, m, xyz, out, flags, tie_out, tie_rounds, st);
  // the register / cluster kernels do not track ties: report "round 0" (no round is known to be tie-free)
  if (tie_out != nullptr) VNB_CUDA(cudaMemsetAsync(tie_out, 0, sizeof(int) * (size_t)b, st));
  if (n <= 8192 && g_fps_cl <= 1) return fps_dispatch_t<512>(b, n, m, 1, xyz, out, flags, st);
  int CL = g_fps_cl < 2 ? 2 : g_fps_cl;
  while (CL < 16 && (long long)CL * 256 * 20 < n) CL *= 2;
  if (g_fps_cl == 0)
    while (CL < 8 && (long long)CL * 256 * 10 < n) CL *= 2;
  return fps_dispatch_t<256>(b, n, m, CL, xyz, out, flags, st);
}

}  // namespace vnb

using namespace vnb;

namespace vnb { extern int g_fps_dispatch, g_fps_ablate, g_nms_cluster; }
extern "C" int vnb_set_tuning(const char* key, int value) {
  std::string_view k(key);
  if (k == "fps_mode") g_fps_mode = value;
  else if (k == "fps_cluster") g_fps_cl = value;
  else if (k == "fps_threads") g_fps_threads = value;
  else if (k == "fps_variant") g_fps_variant = value;
  else if (k == "fps_dispatch") vnb::g_fps_dispatch = value;
  else if (k == "fps_ablate") vnb::g_fps_ablate = value;
  else if (k == "nms_cluster") vnb::g_nms_cluster = (value == 1 || value == 2 || value == 4 || value == 8) ? value : 8;
  else if (k == "ball_query_variant") vnb::g_bq_variant = value;
  else if (k == "bq_grid_min_n") vnb::g_bq_grid_min_n = value;
  else if (k == "sa_variant") vnb::g_sa_variant = value;
  else if (k == "sa_sms") vnb::g_sa_sms = value;
  else if (k == "sa_split") vnb::g_sa_split = value < 1 ? 1 : value;
  else if (k == "sa_min_tpc") vnb::g_sa_min_tpc = value < 1 ? 1 : value;
  else return set_err(VNB_ERR_INVALID, "set_tuning: unknown key %s", key);
  return VNB_OK;
}

// debugging aid (not part of the drop-in boundary): per-phase cycle counters of the cluster FPS kernel
namespace vnb { extern long long* g_sa_trace; }
// debugging aid (not part of the drop-in boundary): stage timeline of CTA 0 of the fused SA kernels
extern "C" int vnb_debug_sa_trace(void* device_buffer_8x64x2_i64) {
  vnb::g_sa_trace = static_cast<long long*>(device_buffer_8x64x2_i64);
  return VNB_OK;
}

extern "C" int vnb_debug_fps_profile(void* device_buffer_16x8_i64) {
  vnb::g_fps_prof = static_cast<long long*>(device_buffer_16x8_i64);
  return VNB_OK;
}

static int fps_check_args(int b, int n, int m) {
  VNB_REQUIRE(m > 0, "FarthestPointSample expects positive npoint");                     // tf_sampling.cpp:99
  VNB_REQUIRE(b >= 0 && n >= 1, "FarthestPointSample expects (batch_size,num_points,3) inp shape");  // :105
  VNB_REQUIRE(n <= 32768, "farthest_point_sample: n = %d exceeds the on-chip limit of 32768 points per cloud", n);
  return VNB_OK;
}

extern "C" int vnb_farthest_point_sample(int b, int n, int m, const float* xyz, int* out_idx, void* stream) {
  if (int rc = fps_check_args(b, n, m)) return rc;
  if (b == 0) return VNB_OK;
  return fps_dispatch(b, n, m, xyz, out_idx, nullptr, as_stream(stream));
}

extern "C" size_t vnb_fps_nested_workspace_bytes(int b, int m) {
  return (size_t)(b > 0 ? b : 1) * (2 * sizeof(int) + (size_t)(m > 0 ? m : 1) * sizeof(float)) + 256;
}

__global__ void fps_hint_identity_kernel(int b, int m, const int* __restrict__ hint, int* __restrict__ out, int* __restrict__ done) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < b) done[t] = hint[t] >= m ? 1 : 0;
  if (t >= b * m) return;
  if (hint[t / m] >= m) out[t] = t % m;
}

__global__ void fps_fill_kernel(int b, int* __restrict__ v, int value) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < b) v[t] = value;
}

static int fps_nested_impl(int b, int n, int m, const float* xyz, int* out_idx, void* workspace, const int* hint,
                           cudaStream_t st, int* proven_rounds = nullptr) {
  const size_t smem = (size_t)m * sizeof(float4);
  if (m > n || smem > 200 * 1024) {  // the identity prefix needs m <= n; huge m does not fit the proof kernel
    if (proven_rounds != nullptr) {
      fps_fill_kernel<<<(b + 127) / 128, 128, 0, st>>>(b, proven_rounds, 0);
      if (int rc = check_launch("fps proven rounds")) return rc;
    }
    return fps_dispatch(b, n, m, xyz, out_idx, nullptr, st);
  }
  int* fail = static_cast<int*>(workspace);
  int* done = fail + b;
  float* R = reinterpret_cast<float*>(static_cast<char*>(workspace) + (((size_t)b * 2 * sizeof(int) + 255) / 256) * 256);
  if (hint != nullptr) {
    // Hinted level: clouds covered by the hint get the identity prefix, the others (a proof failed upstream: exact float
    // ties, rare) go straight to the sequential sampler — two launches instead of the proof's four.
    fps_hint_identity_kernel<<<(b * m + 255) / 256, 256, 0, st>>>(b, m, hint, out_idx, done);
    if (int rc = check_launch("fps identity (hinted)")) return rc;
    return fps_dispatch(b, n, m, xyz, out_idx, done, st);
  }
  VNB_CUDA(cudaMemsetAsync(fail, 0, sizeof(int) * (size_t)b, st));
  if ((size_t)m * 12 > 48 * 1024)
    VNB_CUDA(cudaFuncSetAttribute(fps_prefix_r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)m * 12)));
  fps_prefix_r_kernel<<<dim3((m + 127) / 128, b), 128, (size_t)m * 12, st>>>(n, m, xyz, R, hint);
  if (int rc = check_launch("fps prefix R")) return rc;
  if (smem > 48 * 1024)
    VNB_CUDA(cudaFuncSetAttribute(fps_prefix_verify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fps_prefix_verify_kernel<<<dim3((n + VP - 1) / VP, b), VP * VS, smem, st>>>(n, m, xyz, R, fail, hint);
  if (int rc = check_launch("fps prefix proof")) return rc;
  fps_identity_kernel<<<(b * m + 255) / 256, 256, 0, st>>>(b, m, fail, out_idx, done, proven_rounds);
  if (int rc = check_launch("fps identity")) return rc;
  // sequential kernel for the clouds whose proof failed (clusters of proven clouds exit at once)
  return fps_dispatch(b, n, m, xyz, out_idx, done, st);
}

extern "C" int vnb_farthest_point_sample_nested(int b, int n, int m, const float* xyz, int* out_idx, void* workspace,
                                                void* stream) {
  if (int rc = fps_check_args(b, n, m)) return rc;
  if (b == 0) return VNB_OK;
  return fps_nested_impl(b, n, m, xyz, out_idx, workspace, nullptr, as_stream(stream));
}

extern "C" int vnb_farthest_point_sample_nested_proof(int b, int n, int m, const float* xyz, int* out_idx, void* workspace,
                                                      int* proven_rounds, void* stream) {
  if (int rc = fps_check_args(b, n, m)) return rc;
  VNB_REQUIRE(proven_rounds != nullptr, "farthest_point_sample_nested_proof: proven_rounds buffer missing");
  if (b == 0) return VNB_OK;
  return fps_nested_impl(b, n, m, xyz, out_idx, workspace, nullptr, as_stream(stream), proven_rounds);
}

extern "C" int vnb_farthest_point_sample_ties(int b, int n, int m, const float* xyz, int* out_idx, int* first_tie_round,
                                              int track_rounds, void* stream) {
  if (int rc = fps_check_args(b, n, m)) return rc;
  VNB_REQUIRE(first_tie_round != nullptr, "farthest_point_sample_ties: first_tie_round buffer missing");
  if (b == 0) return VNB_OK;
  return fps_dispatch(b, n, m, xyz, out_idx, nullptr, as_stream(stream), first_tie_round, track_rounds > 0 ? track_rounds : m);
}

extern "C" int vnb_farthest_point_sample_nested_hint(int b, int n, int m, const float* xyz, int* out_idx, void* workspace,
                                                     const int* parent_first_tie_round, void* stream) {
  if (int rc = fps_check_args(b, n, m)) return rc;
  if (b == 0) return VNB_OK;
  return fps_nested_impl(b, n, m, xyz, out_idx, workspace, parent_first_tie_round, as_stream(stream));
}

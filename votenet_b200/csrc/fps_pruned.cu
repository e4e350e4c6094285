// Farthest point sampling, bucket-pruned: ONE CTA (one SM) per cloud, bit-identical to the reference.
//
// Reference: farthestpointsamplingKernel, tf_ops/sampling/tf_sampling_g.cu:105-170 — every round re-evaluates the
// distance of ALL n points to the last pick.  Almost all of that work is provably a no-op: once a few hundred points
// are picked, the running min-distance `temp` of a point far from the new pick cannot change.
//
// Layout.  At kernel start the cloud is counting-sorted by a 12-bit Morton cell key (4 bits per axis over the cloud's
// bounding box) and cut into sub-buckets of 160 consecutive points = 32 lanes x 5 points; warp w owns sub-buckets
// w, w+16, w+32, ... (8 per warp, neighbouring buckets land in different warps).  x,y live in shared memory
// ([40][512] float2, conflict-free), z and the running min-distance in REGISTERS (40 + 40 per thread), the 16-bit tie
// key of every point in shared memory.  512 threads x 40 points = 20480 points = the reference's POINT_NUM (config.py:1).
//
// Pruning (exact).  Lane s of a warp holds the bounding box of sub-bucket s and its cached champion (max temp).  For
// the new pick L the lane evaluates bound = d2(gap_x, gap_y, gap_z), gap = max(lo - L, L - hi, 0) per axis, with the SAME
// float expression as the point distance.  Float subtraction, multiplication and fma are monotone under round-to-
// nearest, so bound <= d2(p - L) (as computed in float) for every point p of the box; if bound >= champion temp then
// min(d, temp) == temp for every point of the sub-bucket: nothing changes, the cached champion stays valid, the
// sub-bucket is skipped.  Otherwise the warp rescans its 160 points (5 per lane) and re-elects the champion.  On the
// synthetic SUN-RGB-D-shaped clouds ~7 of 125 sub-buckets are active per round (profiles/).
//
// Election.  The reference's winner is the candidate with maximal temp, ties to the smallest (k mod 512), then the
// smallest k; the 16-bit tie key of fps_common.cuh encodes that order, so (temp, tie key) compared
// lexicographically is a TOTAL order on points and the arg-max is independent of the storage layout and of the
// reduction topology.  temp >= 0 so its bit pattern orders as a signed integer (padding points carry -1.0f and never
// win).  Every level takes the fast path "one redux.sync.max + one ballot" and consults tie keys only when two
// candidates share the maximal temp exactly.  Within a lane the 5 points of a sub-bucket are stored in descending
// tie-key order, so "first strict maximum" is the reference's per-thread rule.
//   round = bound test (8 lanes) -> rescan active sub-buckets -> lanes 0..7 publish {temp, slot, x, y, z} of their
//           sub-bucket (double-buffered) -> ONE __syncthreads -> every warp reduces the 128 champions -> next pick.
#include "fps_common.cuh"

namespace vnb {

constexpr int PT = 512;          // threads per CTA
constexpr int PW = PT / 32;      // warps
constexpr int PS = 8;            // sub-buckets per warp
constexpr int PQ = 5;            // points per lane per sub-bucket
constexpr int PP = PS * PQ;      // points per thread
constexpr int PCAP = PT * PP;    // 20480 points per cloud
constexpr int PNB = PW * PS;     // 128 sub-buckets
constexpr int PBINS = 4096;      // Morton cells (4 bits per axis)

constexpr size_t OFF_XY = 0;                                   // float2 [PP][PT]
constexpr size_t OFF_TIE = OFF_XY + (size_t)PP * PT * 8;       // uint16 [PP][PT]
constexpr size_t OFF_HI = OFF_TIE + (size_t)PP * PT * 2;       // int    [2][PNB]  champion temp bits
constexpr size_t OFF_TK = OFF_HI + 2 * PNB * 4;                // uint32 [2][PNB]  champion tie key | shared-temp flag << 16
constexpr size_t OFF_REC = OFF_TK + 2 * PNB * 4;               // float4 [2][PNB]  champion {x, y, z, -}
constexpr size_t OFF_RED = OFF_REC + 2 * PNB * 16;             // float  [PW][8]   setup reductions
constexpr size_t PRUNED_SMEM = OFF_RED + PW * 8 * 4;
// setup-only aliases (dead before the point arrays are filled)
constexpr size_t OFF_HIST = OFF_XY;                            // uint32 [PBINS]
constexpr size_t OFF_KEY = OFF_XY + PBINS * 4;                 // uint16 [PCAP] cell key of point k
constexpr size_t OFF_IDX = OFF_TIE;                            // uint16 [PCAP] point index at sorted position p
static_assert(OFF_KEY + (size_t)PCAP * 2 <= OFF_TIE, "setup aliases overflow the xy array");
static_assert(PRUNED_SMEM <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ int fmap(float f) {  // order-preserving float -> signed int (self-inverse)
  int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7FFFFFFF);
}
__device__ __forceinline__ float funmap(int b) { return __int_as_float(b ^ ((b >> 31) & 0x7FFFFFFF)); }
__device__ __forceinline__ uint32_t spread4(uint32_t v) {  // abcd -> a00b00c00d
  return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

template <bool PROF, bool TIES>
__global__ void __launch_bounds__(PT, 1) fps_pruned_kernel(int n, int m, const float* __restrict__ xyz,
                                                           int* __restrict__ out, const int* __restrict__ flags,
                                                           long long* __restrict__ prof, int* __restrict__ tie_out, int tie_rounds) {
  extern __shared__ __align__(16) unsigned char smem[];
  float2* sxy = reinterpret_cast<float2*>(smem + OFF_XY);
  uint16_t* stie = reinterpret_cast<uint16_t*>(smem + OFF_TIE);
  int* s_hi = reinterpret_cast<int*>(smem + OFF_HI);
  uint32_t* s_tk = reinterpret_cast<uint32_t*>(smem + OFF_TK);
  float4* s_rec = reinterpret_cast<float4*>(smem + OFF_REC);
  float* s_red = reinterpret_cast<float*>(smem + OFF_RED);
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + OFF_HIST);
  uint16_t* skey = reinterpret_cast<uint16_t*>(smem + OFF_KEY);
  uint16_t* sidx = reinterpret_cast<uint16_t*>(smem + OFF_IDX);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cloud = blockIdx.x;
  if (flags != nullptr && flags[cloud] != 0) return;
  const float* pc = xyz + (size_t)cloud * n * 3;
  int* oc = out + (size_t)cloud * m;
  long long t_begin = 0;
  if (PROF) t_begin = clock64();

  // ---------------------------------------------------------------- setup A: bounding box of the cloud
  for (int i = tid; i < PBINS; i += PT) hist[i] = 0u;
  float mn0 = INFINITY, mn1 = INFINITY, mn2 = INFINITY, mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
  for (int k = tid; k < n; k += PT) {
    const float x = pc[(size_t)k * 3], y = pc[(size_t)k * 3 + 1], z = pc[(size_t)k * 3 + 2];
    mn0 = fminf(mn0, x); mx0 = fmaxf(mx0, x);
    mn1 = fminf(mn1, y); mx1 = fmaxf(mx1, y);
    mn2 = fminf(mn2, z); mx2 = fmaxf(mx2, z);
  }
  mn0 = funmap(redux_min_s32(fmap(mn0))); mx0 = funmap(redux_max_s32(fmap(mx0)));
  mn1 = funmap(redux_min_s32(fmap(mn1))); mx1 = funmap(redux_max_s32(fmap(mx1)));
  mn2 = funmap(redux_min_s32(fmap(mn2))); mx2 = funmap(redux_max_s32(fmap(mx2)));
  if (lane == 0) {
    float* r = s_red + warp * 8;
    r[0] = mn0; r[1] = mn1; r[2] = mn2; r[3] = mx0; r[4] = mx1; r[5] = mx2;
  }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < PW; ++w) {
    const float* r = s_red + w * 8;
    mn0 = fminf(mn0, r[0]); mn1 = fminf(mn1, r[1]); mn2 = fminf(mn2, r[2]);
    mx0 = fmaxf(mx0, r[3]); mx1 = fmaxf(mx1, r[4]); mx2 = fmaxf(mx2, r[5]);
  }
  // ---------------------------------------------------------------- setup B: cell keys + histogram
  {
    const float iv0 = mx0 > mn0 ? 16.f / (mx0 - mn0) : 0.f;
    const float iv1 = mx1 > mn1 ? 16.f / (mx1 - mn1) : 0.f;
    const float iv2 = mx2 > mn2 ? 16.f / (mx2 - mn2) : 0.f;
    for (int k = tid; k < n; k += PT) {
      const float x = pc[(size_t)k * 3], y = pc[(size_t)k * 3 + 1], z = pc[(size_t)k * 3 + 2];
      const int c0 = min(15, max(0, __float2int_rz((x - mn0) * iv0)));
      const int c1 = min(15, max(0, __float2int_rz((y - mn1) * iv1)));
      const int c2 = min(15, max(0, __float2int_rz((z - mn2) * iv2)));
      const uint32_t key = spread4((uint32_t)c0) | (spread4((uint32_t)c1) << 1) | (spread4((uint32_t)c2) << 2);
      skey[k] = (uint16_t)key;
      atomicAdd(&hist[key], 1u);
    }
  }
  __syncthreads();
  // ---------------------------------------------------------------- setup C: exclusive scan of the histogram
  {
    constexpr int PER = PBINS / PT;  // 8 bins per thread
    uint32_t c[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { c[j] = hist[tid * PER + j]; sum += c[j]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    uint32_t* s_tot = reinterpret_cast<uint32_t*>(s_red);  // the bounding-box partials were consumed before setup B's barrier
    if (lane == 31) s_tot[warp] = inc;
    __syncthreads();
    uint32_t base = inc - sum;
    for (int w = 0; w < warp; ++w) base += s_tot[w];
#pragma unroll
    for (int j = 0; j < PER; ++j) { hist[tid * PER + j] = base; base += c[j]; }
  }
  __syncthreads();
  // ---------------------------------------------------------------- setup D: scatter point indices into cell order
  for (int k = tid; k < n; k += PT) {
    const uint32_t pos = atomicAdd(&hist[skey[k]], 1u);
    sidx[pos] = (uint16_t)k;
  }
  __syncthreads();
  // ---------------------------------------------------------------- setup E: this thread's 40 points
  uint32_t tk[PP];
#pragma unroll
  for (int i = 0; i < PP; ++i) {
    const int s = i / PQ, q = i % PQ;
    const int p = ((s * PW + warp) * PQ + q) * 32 + lane;  // sorted position: sub-bucket (s*PW + warp), slot q, lane
    tk[i] = p < n ? tie_key((int)sidx[p]) : 0u;
  }
  __syncthreads();  // sidx (aliases stie) and hist/skey (alias sxy) are dead from here on
#pragma unroll
  for (int s = 0; s < PS; ++s) {  // descending tie key inside each (lane, sub-bucket); padding (0) sinks to the end
#pragma unroll
    for (int a = 0; a < PQ - 1; ++a)
#pragma unroll
      for (int b = 0; b < PQ - 1 - a; ++b) {
        const uint32_t u = tk[s * PQ + b], v = tk[s * PQ + b + 1];
        tk[s * PQ + b] = max(u, v);
        tk[s * PQ + b + 1] = min(u, v);
      }
  }
  float z[PP], td[PP];
  float blx = 0.f, bly = 0.f, blz = 0.f, bhx = 0.f, bhy = 0.f, bhz = 0.f;  // lane s: bounding box of sub-bucket s
#pragma unroll
  for (int s = 0; s < PS; ++s) {
    float l0 = INFINITY, l1 = INFINITY, l2 = INFINITY, h0 = -INFINITY, h1 = -INFINITY, h2 = -INFINITY;
#pragma unroll
    for (int q = 0; q < PQ; ++q) {
      const int i = s * PQ + q;
      float x = 0.f, y = 0.f;
      if (tk[i] != 0u) {
        const int k = tie_key_to_index(tk[i]);
        x = pc[(size_t)k * 3]; y = pc[(size_t)k * 3 + 1]; z[i] = pc[(size_t)k * 3 + 2];
        td[i] = 1e38f;  // tf_sampling_g.cu:118
        l0 = fminf(l0, x); h0 = fmaxf(h0, x);
        l1 = fminf(l1, y); h1 = fmaxf(h1, y);
        l2 = fminf(l2, z[i]); h2 = fmaxf(h2, z[i]);
      } else {
        z[i] = 0.f;
        td[i] = -1.f;  // padding: min(d,-1) stays -1; a negative temp never wins a signed max against a real point
      }
      sxy[i * PT + tid] = make_float2(x, y);
      stie[i * PT + tid] = (uint16_t)tk[i];
    }
    l0 = funmap(redux_min_s32(fmap(l0))); h0 = funmap(redux_max_s32(fmap(h0)));
    l1 = funmap(redux_min_s32(fmap(l1))); h1 = funmap(redux_max_s32(fmap(h1)));
    l2 = funmap(redux_min_s32(fmap(l2))); h2 = funmap(redux_max_s32(fmap(h2)));
    if (lane == s) { blx = l0; bly = l1; blz = l2; bhx = h0; bhy = h1; bhz = h2; }
  }
  __syncthreads();

  // ---------------------------------------------------------------- rounds
  // Champion table (one entry per sub-bucket, two copies indexed by round parity): temp bits, tie key (+ bit 16: the
  // champion's temp is shared by another point of the sub-bucket), coordinates.  A rescan's winning lane writes the
  // entry of the CURRENT parity in place; the other copy is brought up to date one round later by lane s (after the
  // barrier that ends every read of that copy) — so warps without an active sub-bucket publish nothing at all.
  int chi = __float_as_int(-1.f);            // lane s (< PS): champion temp bits of sub-bucket s (for the bound test)
  float lx = pc[0], ly = pc[1], lz = pc[2];  // last pick: index 0 (tf_sampling_g.cu:114-116)
  if (tid == 0) oc[0] = 0;
  int first_tie = 0x7fffffff;  // first round whose arg-max was not unique (decided by the tie rule)
  unsigned prev_mask = 0u;     // sub-buckets of this warp rescanned in the previous round
  const int tb = warp * PS + (lane & (PS - 1));  // table entry of "my" sub-bucket (lanes < PS)

  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt0 = 0;
#define PF_TICK(i)                 \
  if (PROF) {                      \
    const long long _t = clock64(); \
    pacc[i] += _t - pt0;           \
    pt0 = _t;                      \
  }
  if (PROF) { pt0 = clock64(); pacc[6] = pt0 - t_begin; }

  for (int r = 1; r < m; ++r) {
    const int par = r & 1;
    int* t_hi = s_hi + par * PNB;
    uint32_t* t_tk = s_tk + par * PNB;
    float4* t_rec = s_rec + par * PNB;
    // ---- entries written last round live in the other copy only: fetch them now, store after the bound test
    const bool stale = lane < PS && ((prev_mask >> lane) & 1u);
    int c_hi = 0;
    uint32_t c_tk = 0u;
    float4 c_rec = make_float4(0.f, 0.f, 0.f, 0.f);
    if (stale) { c_hi = s_hi[(par ^ 1) * PNB + tb]; c_tk = s_tk[(par ^ 1) * PNB + tb]; c_rec = s_rec[(par ^ 1) * PNB + tb]; }
    // ---- which sub-buckets can change?  (lanes 0..7, one sub-bucket each)
    const float gx = fmaxf(fmaxf(blx - lx, lx - bhx), 0.f);
    const float gy = fmaxf(fmaxf(bly - ly, ly - bhy), 0.f);
    const float gz = fmaxf(fmaxf(blz - lz, lz - bhz), 0.f);
    const float bound = d2_ref_gpu(gx, gy, gz);
    const bool act = lane < PS && (r == 1 || bound < __int_as_float(chi));
    const unsigned mask = __ballot_sync(0xffffffffu, act);
    if (stale) { t_hi[tb] = c_hi; t_tk[tb] = c_tk; t_rec[tb] = c_rec; }
    __syncwarp();  // the copy above precedes this round's in-place writes of the same entries
    PF_TICK(0)
    if (PROF && mask) { pacc[4] += __popc(mask); pacc[5] += 1; }
    // ---- rescan the active sub-buckets of this warp
#pragma unroll
    for (int s = 0; s < PS; ++s) {
      if (mask & (1u << s)) {
        float xs[PQ], ys[PQ];
#pragma unroll
        for (int q = 0; q < PQ; ++q) {
          const int i = s * PQ + q;
          const float2 v = sxy[i * PT + tid];
          xs[q] = v.x; ys[q] = v.y;
          td[i] = fminf(d2_ref_gpu(v.x - lx, v.y - ly, z[i] - lz), td[i]);
        }
        const float best = fmaxf(fmaxf(fmaxf(td[s * PQ], td[s * PQ + 1]), fmaxf(td[s * PQ + 2], td[s * PQ + 3])), td[s * PQ + 4]);
        // first maximum in descending tie-key order = the reference's per-thread rule
        int bq = PQ - 1;
#pragma unroll
        for (int q = PQ - 2; q >= 0; --q) bq = td[s * PQ + q] == best ? q : bq;
        const uint32_t tkm = stie[(s * PQ + bq) * PT + tid];    // in flight beside the first reduction
        const int whi = redux_max_s32(__float_as_int(best));
        const bool mine = __float_as_int(best) == whi;
        const uint32_t wk = redux_max(mine ? ((tkm << 5) | (uint32_t)lane) : 0u);
        bool dup = false;
        if (TIES) {  // another point of the sub-bucket shares the champion's temp (another lane, or inside the lane)
          int same = 0;
#pragma unroll
          for (int q = 0; q < PQ; ++q) same += td[s * PQ + q] == best ? 1 : 0;
          const unsigned mm = __ballot_sync(0xffffffffu, mine);
          dup = (mm & (mm - 1u)) != 0u || __ballot_sync(0xffffffffu, mine && same > 1) != 0u;
        }
        if (mine && (wk & 31u) == (uint32_t)lane) {
          float bx = xs[0], by = ys[0], bz = z[s * PQ];
#pragma unroll
          for (int q = 1; q < PQ; ++q)
            if (bq == q) { bx = xs[q]; by = ys[q]; bz = z[s * PQ + q]; }
          t_hi[warp * PS + s] = whi;
          t_tk[warp * PS + s] = (wk >> 5) | (dup ? 0x10000u : 0u);
          t_rec[warp * PS + s] = make_float4(bx, by, bz, 0.f);
        }
        if (lane == s) chi = whi;
      }
    }
    prev_mask = mask;
    PF_TICK(1)
    __syncthreads();
    PF_TICK(2)
    // ---- every warp reduces the 128 champions (4 per lane): max temp, then max tie key among the holders
    const int4 h = reinterpret_cast<const int4*>(t_hi)[lane];
    const uint4 t = reinterpret_cast<const uint4*>(t_tk)[lane];
    const int whi = redux_max_s32(max(max(h.x, h.y), max(h.z, h.w)));
    // key: tie key (bits 8..23) | entry (bits 1..7) | shared-temp flag (bit 0); tie keys are unique per point
    const uint32_t e0 = (uint32_t)lane << 3;
    const uint32_t k0 = h.x == whi ? (((t.x & 0xffffu) << 8) | e0 | (t.x >> 16)) : 0u;
    const uint32_t k1 = h.y == whi ? (((t.y & 0xffffu) << 8) | (e0 + 2u) | (t.y >> 16)) : 0u;
    const uint32_t k2 = h.z == whi ? (((t.z & 0xffffu) << 8) | (e0 + 4u) | (t.z >> 16)) : 0u;
    const uint32_t k3 = h.w == whi ? (((t.w & 0xffffu) << 8) | (e0 + 6u) | (t.w >> 16)) : 0u;
    const uint32_t wk = redux_max(max(max(k0, k1), max(k2, k3)));
    const float4 rec = t_rec[(wk >> 1) & 127u];
    lx = rec.x; ly = rec.y; lz = rec.z;
    if (TIES && r < tie_rounds && first_tie == 0x7fffffff && whi >= 0) {
      const int cnt = (h.x == whi) + (h.y == whi) + (h.z == whi) + (h.w == whi);
      const unsigned holders = __ballot_sync(0xffffffffu, cnt > 0), multi = __ballot_sync(0xffffffffu, cnt > 1);
      if ((holders & (holders - 1u)) | multi | (wk & 1u)) first_tie = r;
    }
    if (warp == (r & (PW - 1)) && lane == 0) oc[r] = tie_key_to_index(wk >> 8);
    PF_TICK(3)
  }
#undef PF_TICK
  if (TIES && tie_out != nullptr && tid == 0) tie_out[cloud] = first_tie;
  if (PROF && prof != nullptr && lane == 0 && cloud == 0) {
    pacc[7] = clock64() - t_begin;
    for (int i = 0; i < 8; ++i) prof[warp * 8 + i] = pacc[i];
  }
}

extern long long* g_fps_prof;

int fps_pruned_capacity() { return PCAP; }

int launch_fps_pruned(int b, int n, int m, const float* xyz, int* out, const int* flags, int* tie_out, int tie_rounds,
                      cudaStream_t st) {
  if (g_fps_prof != nullptr && flags == nullptr) {
    VNB_CUDA(cudaFuncSetAttribute(fps_pruned_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PRUNED_SMEM));
    fps_pruned_kernel<true, true><<<b, PT, PRUNED_SMEM, st>>>(n, m, xyz, out, flags, g_fps_prof, tie_out, tie_rounds);
    return check_launch("farthest_point_sample (pruned, profiled)");
  }
  if (tie_out != nullptr) {
    VNB_CUDA(cudaFuncSetAttribute(fps_pruned_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PRUNED_SMEM));
    fps_pruned_kernel<false, true><<<b, PT, PRUNED_SMEM, st>>>(n, m, xyz, out, flags, nullptr, tie_out, tie_rounds);
    return check_launch("farthest_point_sample (pruned, tie tracking)");
  }
  VNB_CUDA(cudaFuncSetAttribute(fps_pruned_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PRUNED_SMEM));
  fps_pruned_kernel<false, false><<<b, PT, PRUNED_SMEM, st>>>(n, m, xyz, out, flags, nullptr, nullptr, 0);
  return check_launch("farthest_point_sample (pruned)");
}

}  // namespace vnb

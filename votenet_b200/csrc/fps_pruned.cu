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
constexpr size_t OFF_HI = OFF_TIE + (size_t)PP * PT * 2;       // int    [2][PW]   warp champion temp bits
constexpr size_t OFF_TK = OFF_HI + 2 * PW * 4;                 // uint32 [2][PW]   warp champion key (tie key << 8 | entry << 1 | shared)
constexpr size_t OFF_WREC = OFF_TK + 2 * PW * 4;               // float4 [2][PW]   warp champion {x, y, z, -}
constexpr size_t OFF_REC = OFF_WREC + 2 * PW * 16;             // float4 [PNB]     sub-bucket champion {x, y, z, -} (warp-private)
constexpr size_t OFF_RED = OFF_REC + PNB * 16;                 // float  [PW][8]   setup reductions
constexpr int PZS = 8;           // z of the LAST PZS points of a thread lives in shared memory, not in registers: 80
                                 // registers of point state leave too few for the rest of the round loop, and with the
                                 // whole L1 carved out as shared memory a compiler spill is an L2 round trip
constexpr int PZR = PP - PZS;    // z of points 0..PZR-1: registers
constexpr size_t OFF_Z = OFF_RED + PW * 8 * 4;                 // float  [PZS][PT]
constexpr size_t PRUNED_SMEM = OFF_Z + (size_t)PZS * PT * 4;
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

// Register-specific part of a rescan of sub-bucket S: new running distances of the lane's PQ points, the lane-local
// champion (first maximum in descending tie-key order = the reference's per-thread rule), its coordinates, and how many
// of the lane's points share that maximum.
template <int S, bool TIES>
__device__ __forceinline__ void rescan_head(const float2* __restrict__ sxy, const float* __restrict__ sz,
                                            const uint16_t* __restrict__ stie, int tid,
                                            const float (&zr)[PZR], float (&td)[PP],
                                            float lx, float ly, float lz, float& best, int& bq, int& same, float& bx,
                                            float& by, float& bz, uint32_t& btk) {
  float xs[PQ], ys[PQ], z[PQ];
  uint32_t tks[PQ];   // the lane's tie keys, loaded beside the coordinates so that the election never waits for them
#pragma unroll
  for (int q = 0; q < PQ; ++q) {
    const int i = S * PQ + q;
    const float2 v = sxy[i * PT + tid];
    tks[q] = stie[i * PT + tid];
    xs[q] = v.x; ys[q] = v.y;
    z[q] = i < PZR ? zr[i < PZR ? i : 0] : sz[(i - PZR) * PT + tid];
    td[i] = fminf(d2_ref_gpu(v.x - lx, v.y - ly, z[q] - lz), td[i]);
  }
  best = fmaxf(fmaxf(fmaxf(td[S * PQ], td[S * PQ + 1]), fmaxf(td[S * PQ + 2], td[S * PQ + 3])), td[S * PQ + 4]);
  bq = PQ - 1;
  bx = xs[PQ - 1]; by = ys[PQ - 1]; bz = z[PQ - 1]; btk = tks[PQ - 1];
  same = 0;
#pragma unroll
  for (int q = PQ - 2; q >= 0; --q)
    if (td[S * PQ + q] == best) { bq = q; bx = xs[q]; by = ys[q]; bz = z[q]; btk = tks[q]; }
  if (TIES) {
#pragma unroll
    for (int q = 0; q < PQ; ++q) same += td[S * PQ + q] == best ? 1 : 0;
  }
}

template <bool PROF, bool TIES, bool ABL>
__global__ void __launch_bounds__(PT, 1) fps_pruned_kernel(int n, int m, const float* __restrict__ xyz,
                                                           int* __restrict__ out, const int* __restrict__ flags,
                                                           long long* __restrict__ prof, int* __restrict__ tie_out, int tie_rounds, int abl) {
  extern __shared__ __align__(16) unsigned char smem[];
  float2* sxy = reinterpret_cast<float2*>(smem + OFF_XY);
  uint16_t* stie = reinterpret_cast<uint16_t*>(smem + OFF_TIE);
  int* s_whi = reinterpret_cast<int*>(smem + OFF_HI);
  uint32_t* s_wkey = reinterpret_cast<uint32_t*>(smem + OFF_TK);
  float4* s_rec = reinterpret_cast<float4*>(smem + OFF_REC);
  float4* s_wrec = reinterpret_cast<float4*>(smem + OFF_WREC);
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
  float zr[PZR], td[PP];
  float blx = 0.f, bly = 0.f, blz = 0.f, bhx = 0.f, bhy = 0.f, bhz = 0.f;  // lane s: bounding box of sub-bucket s
  float* sz = reinterpret_cast<float*>(smem + OFF_Z);
#pragma unroll
  for (int s = 0; s < PS; ++s) {
    float l0 = INFINITY, l1 = INFINITY, l2 = INFINITY, h0 = -INFINITY, h1 = -INFINITY, h2 = -INFINITY;
#pragma unroll
    for (int q = 0; q < PQ; ++q) {
      const int i = s * PQ + q;
      float x = 0.f, y = 0.f, zi = 0.f;
      if (tk[i] != 0u) {
        const int k = tie_key_to_index(tk[i]);
        x = pc[(size_t)k * 3]; y = pc[(size_t)k * 3 + 1]; zi = pc[(size_t)k * 3 + 2];
        td[i] = 1e38f;  // tf_sampling_g.cu:118
        l0 = fminf(l0, x); h0 = fmaxf(h0, x);
        l1 = fminf(l1, y); h1 = fmaxf(h1, y);
        l2 = fminf(l2, zi); h2 = fmaxf(h2, zi);
      } else {
        td[i] = -1.f;  // padding: min(d,-1) stays -1; a negative temp never wins a signed max against a real point
      }
      sxy[i * PT + tid] = make_float2(x, y);
      stie[i * PT + tid] = (uint16_t)tk[i];
      if (i < PZR) zr[i < PZR ? i : 0] = zi;
      else sz[(i - PZR) * PT + tid] = zi;
    }
    l0 = funmap(redux_min_s32(fmap(l0))); h0 = funmap(redux_max_s32(fmap(h0)));
    l1 = funmap(redux_min_s32(fmap(l1))); h1 = funmap(redux_max_s32(fmap(h1)));
    l2 = funmap(redux_min_s32(fmap(l2))); h2 = funmap(redux_max_s32(fmap(h2)));
    if (lane == s) { blx = l0; bly = l1; blz = l2; bhx = h0; bhy = h1; bhz = h2; }
  }
  __syncthreads();

  // ---------------------------------------------------------------- rounds
  // Two-level champion cache.  Lane s (< PS) keeps the champion of "its" sub-bucket in registers: temp bits `chi` and a
  // key `ckey` = tie key << 8 | table entry << 1 | (another point of the sub-bucket shares that temp); its coordinates
  // sit in t_rec[entry], a table only the owning warp touches.  A warp that rescanned something re-elects the best of
  // its PS champions (wwhi / wwk / wrec, uniform registers); EVERY warp publishes that warp champion every round into
  // the copy of the 16-entry table selected by the round's parity (three predicated stores from registers) — no
  // read-modify-write of shared tables, so the only hazard is write-after-read across rounds, which the parity removes.
  // The end-of-round reduction is over PW = 16 warp champions.  What all 16 warps execute between two barriers costs
  // issue slots as much as latency (4 warps per scheduler run the same code at the same time), so that part is kept
  // to ~45 instructions; the picks go to global memory as raw keys and are decoded after the loop.
  int chi = __float_as_int(-1.f);
  uint32_t ckey = 0u;
  int wwhi = (int)0x80000000;
  uint32_t wwk = 0u;
  float4 wrec = make_float4(0.f, 0.f, 0.f, 0.f);
  float lx = pc[0], ly = pc[1], lz = pc[2];  // last pick: index 0 (tf_sampling_g.cu:114-116)
  int first_tie = 0x7fffffff;  // first round whose arg-max was not unique (decided by the tie rule)
  float4* t_rec = s_rec + warp * PS;   // this warp's PS sub-bucket champions {x, y, z, -}

  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt0 = 0;
  // phase clock: the read is ordered after `dep` has been produced (a true point on the dependency chain)
#define PF_TICK(i, dep)                                                                  \
  if (PROF) {                                                                            \
    long long _t;                                                                        \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(_t) : "r"((int)(dep)) : "memory");      \
    pacc[i] += _t - pt0;                                                                 \
    pt0 = _t;                                                                            \
  }
  const long long t_setup = PROF ? clock64() - t_begin : 0;
  if (PROF) pt0 = clock64();

  for (int r = 1; r < m; ++r) {
    const int par = r & 1;
    // ---- which sub-buckets can change?  (lanes 0..7, one sub-bucket each; the other lanes compute on zeros)
    const float gx = fmaxf(fmaxf(blx - lx, lx - bhx), 0.f);
    const float gy = fmaxf(fmaxf(bly - ly, ly - bhy), 0.f);
    const float gz = fmaxf(fmaxf(blz - lz, lz - bhz), 0.f);
    const float bound = d2_ref_gpu(gx, gy, gz);
    bool act = lane < PS && (r == 1 || bound < __int_as_float(chi));
    if (ABL) act = act && (r == 1 || !(abl & 3));
    const unsigned mask = __ballot_sync(0xffffffffu, act);
    PF_TICK(0, mask)
    if (PROF && mask) { pacc[4] += __popc(mask); pacc[5] += 1; }
    // ---- rescan the active sub-buckets of this warp, then re-elect the warp's champion
    if (!ABL && mask) {   // (the ablation variant has no rescan code at all: measures the loop without its bulk)
      auto elect = [&](int s, float best, int bq, int same, float bx, float by, float bz, uint32_t tkm) {
        const int whi = redux_max_s32(__float_as_int(best));
        const bool mine = __float_as_int(best) == whi;
        const uint32_t wk = redux_max(mine ? ((tkm << 5) | (uint32_t)lane) : 0u);
        // tie tracking: how many points of the sub-bucket sit at the champion's temp (off the election's chain)
        // (votes, not redux.add: REDUX.SUM is far slower than the CREDUX min/max path)
        bool shared = false;
        if (TIES) {
          const unsigned mm = __ballot_sync(0xffffffffu, mine);
          shared = (mm & (mm - 1u)) != 0u || __ballot_sync(0xffffffffu, mine && same > 1) != 0u;
        }
        if (mine && (wk & 31u) == (uint32_t)lane) t_rec[s] = make_float4(bx, by, bz, 0.f);
        if (lane == s) {
          chi = whi;
          ckey = ((wk >> 5) << 8) | (uint32_t)((warp * PS + s) << 1) | (shared ? 1u : 0u);
        }
      };
      float best, bx, by, bz;
      int bq, same;
      uint32_t btk;
#define VNB_RESCAN(S)                                                                               \
      if (mask & (1u << S)) {                                                                         \
        rescan_head<S, TIES>(sxy, sz, stie, tid, zr, td, lx, ly, lz, best, bq, same, bx, by, bz, btk);  \
        elect(S, best, bq, same, bx, by, bz, btk);                                                    \
      }
      VNB_RESCAN(0) VNB_RESCAN(1) VNB_RESCAN(2) VNB_RESCAN(3) VNB_RESCAN(4) VNB_RESCAN(5) VNB_RESCAN(6) VNB_RESCAN(7)
#undef VNB_RESCAN
      wwhi = redux_max_s32(lane < PS ? chi : (int)0x80000000);
      const bool top = lane < PS && chi == wwhi;
      wwk = redux_max(top ? ckey : 0u);
      if (TIES) {
        const unsigned tm = __ballot_sync(0xffffffffu, top);
        if (tm & (tm - 1u)) wwk |= 1u;
      }
      __syncwarp();   // the winners' t_rec stores above are visible to the whole warp
      wrec = t_rec[(wwk >> 1) & (PS - 1)];
    }
    // ---- every warp publishes its champion (registers -> this round's copy of the table)
    if (lane == PS) { s_whi[par * PW + warp] = wwhi; s_wkey[par * PW + warp] = wwk; }
    if (lane == PS + 1) s_wrec[par * PW + warp] = wrec;
    PF_TICK(1, wwk)
    __syncthreads();
    PF_TICK(2, 0)
    // ---- every warp reduces the 16 warp champions: max temp, then max tie key among the holders
    const int hw = s_whi[par * PW + (lane & (PW - 1))];
    const uint32_t kw = s_wkey[par * PW + (lane & (PW - 1))];
    const int whi = redux_max_s32(hw);
    const uint32_t wk = redux_max(hw == whi ? kw : 0u);
    const float4 rec = s_wrec[par * PW + ((wk >> 4) & (PW - 1))];   // entry = warp * PS + s: the warp is entry >> 3
    PF_TICK(3, wk)
    if (!ABL || !(abl & 4)) { lx = rec.x; ly = rec.y; lz = rec.z; }
    // keep the pick in vector registers: the compiler would otherwise move these warp-uniform values to uniform
    // registers (R2UR) on the round's critical chain
    asm volatile("" : "+f"(lx), "+f"(ly), "+f"(lz));
    PF_TICK(6, __float_as_int(lx))
    if (TIES && r < tie_rounds) {  // unique arg-max?  (several champions at the maximal temp, or one that won on its tie key)
      const unsigned hm = __ballot_sync(0xffffffffu, hw == whi) & 0xffffu;
      if (whi >= 0 && ((hm & (hm - 1u)) || (wk & 1u))) first_tie = min(first_tie, r);
    }
    if (tid == 0) oc[r] = (int)(wk >> 8);   // raw tie key; decoded below
    PF_TICK(7, first_tie)
  }
#undef PF_TICK
  __syncthreads();   // thread 0's stores of the raw keys are visible to the block
  for (int r = tid; r < m; r += PT) oc[r] = r == 0 ? 0 : tie_key_to_index((uint32_t)oc[r]);
  if (TIES && tie_out != nullptr && tid == 0) tie_out[cloud] = first_tie;
  if (PROF && prof != nullptr && lane == 0 && cloud == 0) {
    // [0] bound test -> mask, [1] rescans, [2] barrier, [3] barrier -> winner key, [6] winner key -> pick coordinates,
    // [7] tie bookkeeping + index store; [4] rescans, [5] rounds with a rescan; then setup and total cycles
    for (int i = 0; i < 8; ++i) prof[warp * 10 + i] = pacc[i];
    prof[warp * 10 + 8] = t_setup;
    prof[warp * 10 + 9] = clock64() - t_begin;
  }
}

extern long long* g_fps_prof;
int g_fps_dispatch = 0;  // (unused; kept so that old tuning scripts do not fail)
int g_fps_ablate = 0;    // tuning "fps_ablate": timing experiments only (wrong results): 1 no rescans, 2 no bound test, 4 fixed pick

int fps_pruned_capacity() { return PCAP; }

template <bool PROF, bool TIES, bool ABL>
static int launch_pruned(int b, int n, int m, const float* xyz, int* out, const int* flags, long long* prof, int* tie_out,
                         int tie_rounds, cudaStream_t st, const char* what) {
  VNB_CUDA(cudaFuncSetAttribute(fps_pruned_kernel<PROF, TIES, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PRUNED_SMEM));
  fps_pruned_kernel<PROF, TIES, ABL><<<b, PT, PRUNED_SMEM, st>>>(n, m, xyz, out, flags, prof, tie_out, tie_rounds, g_fps_ablate);
  return check_launch(what);
}

int launch_fps_pruned(int b, int n, int m, const float* xyz, int* out, const int* flags, int* tie_out, int tie_rounds,
                      cudaStream_t st) {
  if (g_fps_prof != nullptr && flags == nullptr)
    return launch_pruned<true, true, false>(b, n, m, xyz, out, flags, g_fps_prof, tie_out, tie_rounds, st, "farthest_point_sample (pruned, profiled)");
  if (g_fps_ablate != 0)   // timing experiments only
    return tie_out != nullptr ? launch_pruned<false, true, true>(b, n, m, xyz, out, flags, nullptr, tie_out, tie_rounds, st, "farthest_point_sample (ablation)")
                              : launch_pruned<false, false, true>(b, n, m, xyz, out, flags, nullptr, nullptr, 0, st, "farthest_point_sample (ablation)");
  if (tie_out != nullptr)
    return launch_pruned<false, true, false>(b, n, m, xyz, out, flags, nullptr, tie_out, tie_rounds, st, "farthest_point_sample (pruned, tie tracking)");
  return launch_pruned<false, false, false>(b, n, m, xyz, out, flags, nullptr, nullptr, 0, st, "farthest_point_sample (pruned)");
}

}  // namespace vnb

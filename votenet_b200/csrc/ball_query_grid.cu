// Grid-accelerated ball query, bit-identical to the reference's exhaustive scan.
//
// Reference: query_ball_point_gpu, tf_ops/grouping/tf_grouping_g.cu:3-36 — every query scans ALL n points in index
// order and keeps the first `nsample` hits.  For VoteNet's SA1 (n = 20000, m = 2048, r = 0.2) a ball holds ~35 points,
// so > 99.8 % of the 41 M distance tests per cloud miss.
//
// Here, per cloud:
//   build   one CTA bins the n source points into a uniform grid with cell edge >= 1.01 r (counting sort in global
//           scratch: per-cell counts by atomics, block-wide exclusive scan, scatter of (x,y,z,index) records);
//   query   one warp per query visits only the <= 27 cells around it (9 contiguous x-runs), tests the candidates with the
//           reference's exact predicate, and records each hit as ONE BIT of an n-bit bitmap in shared memory (bit = point
//           index).  Reading the bitmap in ascending bit order yields the hits in ascending index order regardless of the
//           order in which cells were visited, i.e. exactly "the first nsample hits of the index-ordered scan"
//           (tf_grouping_g.cu:16-17); padding with the first hit (:26-29) and the count (:34) follow.
// A point within r of the query differs from it by < one cell edge per axis, so it lies in one of the 27 visited cells:
// the candidate set is a superset of the hit set and the exact test decides — the outputs are bit-identical to the scan.
#include "common.cuh"

#include <math.h>

namespace vnb {

extern int g_bq_variant;
int g_bq_grid_min_n = 1024;  // tuning: smallest source cloud that takes the grid path

struct GridHdr {      // per cloud, at the start of its workspace slice
  float ox, oy, oz;   // grid origin (min corner)
  float inv;          // 1 / cell edge
  int nx, ny, nz;     // cells per axis
  int ncell;
};

constexpr int GRID_MAX_DIM = 32;                 // <= 32768 cells per cloud
constexpr int GRID_MAX_CELLS = GRID_MAX_DIM * GRID_MAX_DIM * GRID_MAX_DIM;
constexpr int CS_INTS = GRID_MAX_CELLS + 4;           // cell_start entries (padded so the records stay 16-byte aligned)
constexpr int BT = 1024;

__device__ __forceinline__ int cell_coord(float p, float o, float inv, int nmax) {
  int c = (int)floorf((p - o) * inv);
  return min(max(c, 0), nmax - 1);
}

// workspace layout per cloud (bytes): [GridHdr 32][cell_start CS_INTS int][cursor GRID_MAX_CELLS int][rec n float4]
__host__ __device__ inline size_t grid_slice_bytes(int n) {
  size_t b = 32 + (size_t)CS_INTS * 4 + (size_t)GRID_MAX_CELLS * 4 + (size_t)n * 16;
  return (b + 255) / 256 * 256;
}

__global__ void __launch_bounds__(BT) grid_build_kernel(int n, float cell_min, const float* __restrict__ xyz,
                                                         char* __restrict__ ws, size_t slice) {
  __shared__ float s_red[6][32];
  __shared__ GridHdr s_h;
  __shared__ int s_scan[BT];
  const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* p = xyz + (size_t)cloud * n * 3;
  char* base = ws + (size_t)cloud * slice;
  GridHdr* hdr = reinterpret_cast<GridHdr*>(base);
  int* cell_start = reinterpret_cast<int*>(base + 32);
  int* cursor = cell_start + CS_INTS;
  float4* rec = reinterpret_cast<float4*>(reinterpret_cast<char*>(cursor) + (size_t)GRID_MAX_CELLS * 4);
  // ---- bounding box
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = tid; k < n; k += BT)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float v = p[(size_t)k * 3 + a];
      lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if (lane == 0) { s_red[a][warp] = lo[a]; s_red[3 + a][warp] = hi[a]; }
  }
  __syncthreads();
  if (tid == 0) {
    float l[3], h[3];
    for (int a = 0; a < 3; ++a) {
      l[a] = s_red[a][0]; h[a] = s_red[3 + a][0];
      for (int w = 1; w < BT / 32; ++w) { l[a] = fminf(l[a], s_red[a][w]); h[a] = fmaxf(h[a], s_red[3 + a][w]); }
    }
    float ext = fmaxf(fmaxf(h[0] - l[0], h[1] - l[1]), h[2] - l[2]);
    float cell = fmaxf(cell_min, ext / (float)(GRID_MAX_DIM - 1));  // never smaller than 1.01 r
    if (!(cell > 0.f)) cell = 1.f;
    GridHdr g;
    g.ox = l[0]; g.oy = l[1]; g.oz = l[2];
    g.inv = 1.0f / cell;
    g.nx = min(GRID_MAX_DIM, (int)floorf((h[0] - l[0]) * g.inv) + 1);
    g.ny = min(GRID_MAX_DIM, (int)floorf((h[1] - l[1]) * g.inv) + 1);
    g.nz = min(GRID_MAX_DIM, (int)floorf((h[2] - l[2]) * g.inv) + 1);
    g.ncell = g.nx * g.ny * g.nz;
    s_h = g;
    *hdr = g;
  }
  __syncthreads();
  const GridHdr g = s_h;
  for (int c = tid; c < g.ncell; c += BT) cursor[c] = 0;
  __syncthreads();
  // ---- counts
  for (int k = tid; k < n; k += BT) {
    int cx = cell_coord(p[(size_t)k * 3], g.ox, g.inv, g.nx), cy = cell_coord(p[(size_t)k * 3 + 1], g.oy, g.inv, g.ny),
        cz = cell_coord(p[(size_t)k * 3 + 2], g.oz, g.inv, g.nz);
    atomicAdd(&cursor[(cz * g.ny + cy) * g.nx + cx], 1);
  }
  __syncthreads();
  // ---- exclusive scan of the counts (each thread owns a contiguous chunk)
  const int per = (g.ncell + BT - 1) / BT;
  const int c0 = tid * per, c1 = min(g.ncell, c0 + per);
  int sum = 0;
  for (int c = c0; c < c1; ++c) sum += cursor[c];
  s_scan[tid] = sum;
  __syncthreads();
  for (int o = 1; o < BT; o <<= 1) {
    int v = tid >= o ? s_scan[tid - o] : 0;
    __syncthreads();
    s_scan[tid] += v;
    __syncthreads();
  }
  int run = s_scan[tid] - sum;
  for (int c = c0; c < c1; ++c) {
    int cnt = cursor[c];
    cell_start[c] = run;
    cursor[c] = run;
    run += cnt;
  }
  if (tid == BT - 1) cell_start[g.ncell] = n;
  __syncthreads();
  // ---- scatter (order inside a cell is irrelevant: the query's bitmap restores index order)
  for (int k = tid; k < n; k += BT) {
    float x = p[(size_t)k * 3], y = p[(size_t)k * 3 + 1], z = p[(size_t)k * 3 + 2];
    int cx = cell_coord(x, g.ox, g.inv, g.nx), cy = cell_coord(y, g.oy, g.inv, g.ny), cz = cell_coord(z, g.oz, g.inv, g.nz);
    int pos = atomicAdd(&cursor[(cz * g.ny + cy) * g.nx + cx], 1);
    rec[pos] = make_float4(x, y, z, __int_as_float(k));
  }
}

constexpr int GQ_WARPS = 8;

// Query kernel.  (Its first generation cleared and re-read the whole n-bit bitmap for every query — 2 x 20 shared accesses
// per lane at n = 20000 — and walked the nine x-runs through dependent loads; ncu showed it issue-bound: 77 % issue-active,
// ~1700 warp instructions per query, 31 us for sa1.)  Here
//   * the nine (start, end) pairs are loaded by lanes 0..8 at once and handed out by shuffles;
//   * the bitmap is SPARSE-AWARE: lane l owns the `per` words [l * per, (l + 1) * per), per = ceil(words / 32); with
//     per > GQ_DIRECT it also owns a summary word whose bit i says "word l * per + i is non-zero" (every hit sets it with
//     a second fire-and-forget atomic — an atomic whose result is awaited costs a shared-memory round trip per step of
//     the candidate loop).  Counting, emitting and clearing touch only non-zero words, so the bitmap is zeroed once per
//     warp and left clean by every query.
// Bit order = index order, so the emitted row is still "the first nsample hits of the index-ordered scan".
constexpr int GQ_DIRECT = 4;   // up to this many words per lane the lane simply reads its words (no summary)
template <bool SUMMARY>
__global__ void __launch_bounds__(GQ_WARPS * 32) grid_query2_kernel(int n, int m, float d2_max, int nsample, int per,
                                                                     uint32_t inv /* ceil(65536 / per) */,
                                                                     const float* __restrict__ xyz2,
                                                                     const char* __restrict__ ws, size_t slice,
                                                                     int* __restrict__ idx, int* __restrict__ pts_cnt) {
  extern __shared__ uint32_t s_bm[];  // GQ_WARPS x (32 summary words + 32 * per bitmap words)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cloud = blockIdx.y;
  uint32_t* summ = s_bm + (size_t)warp * (32 + 32 * per);
  uint32_t* bm = summ + 32;
  const char* base = ws + (size_t)cloud * slice;
  const GridHdr g = *reinterpret_cast<const GridHdr*>(base);
  const int* cell_start = reinterpret_cast<const int*>(base + 32);
  const float4* rec = reinterpret_cast<const float4*>(base + 32 + (size_t)CS_INTS * 4 + (size_t)GRID_MAX_CELLS * 4);
  for (int w = lane * 4; w < 32 + 32 * per; w += 128) *reinterpret_cast<uint4*>(summ + w) = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  uint32_t* own = bm + lane * per;
  for (int j = blockIdx.x * GQ_WARPS + warp; j < m; j += gridDim.x * GQ_WARPS) {
    const float* q = xyz2 + ((size_t)cloud * m + j) * 3;
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    // the query itself may lie outside the source points' bounding box: clamp like the builder does; a cell more than
    // one step away from the unclamped coordinate cannot hold a hit, but visiting it is harmless (exact test below)
    const int cx = cell_coord(qx, g.ox, g.inv, g.nx), cy = cell_coord(qy, g.oy, g.inv, g.ny), cz = cell_coord(qz, g.oz, g.inv, g.nz);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    int rs = 0, re = 0;  // lane r < 9: candidate records [rs, re) of x-run r (cells x0..x1 of one (y, z) are contiguous)
    {
      const int z = cz - 1 + lane / 3, y = cy - 1 + lane % 3;
      if (lane < 9 && z >= 0 && z < g.nz && y >= 0 && y < g.ny) {
        const int rowc = (z * g.ny + y) * g.nx;
        rs = __ldg(cell_start + rowc + x0);
        re = __ldg(cell_start + rowc + x1 + 1);
      }
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int s = __shfl_sync(0xffffffffu, rs, r), e = __shfl_sync(0xffffffffu, re, r);
      for (int t = s + lane; t < e; t += 32) {
        const float4 c = __ldg(rec + t);
        if (d2_ref_gpu(qx - c.x, qy - c.y, qz - c.z) <= d2_max) {
          const uint32_t k = (uint32_t)__float_as_int(c.w), w = k >> 5;
          atomicOr(&bm[w], 1u << (k & 31u));
          if (SUMMARY) {
            const uint32_t l = (w * inv) >> 16;   // == w / per for w < 1024 (per <= 32)
            atomicOr(&summ[l], 1u << (w - l * (uint32_t)per));
          }
        }
      }
    }
    __syncwarp();
    // ---- count, then emit the set bits in ascending order; only non-zero words are visited, and they are cleared on the way
    uint32_t flags = 0u;
    if (SUMMARY) {
      flags = summ[lane];
    } else {
#pragma unroll
      for (int i = 0; i < GQ_DIRECT; ++i)
        if (i < per && own[i] != 0u) flags |= 1u << i;
    }
    int mine = 0;
    for (uint32_t f = flags; f; f &= f - 1u) mine += __popc(own[__ffs(f) - 1]);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int pos = incl - mine;
    int* row = idx + ((size_t)cloud * m + j) * nsample;
    int first_local = -1;
    for (uint32_t f = flags; f; f &= f - 1u) {
      const int wi = __ffs(f) - 1;
      uint32_t bits = own[wi];
      own[wi] = 0u;
      const int kbase = (lane * per + wi) << 5;
      if (first_local < 0) first_local = kbase + __ffs(bits) - 1;
      while (bits && pos < nsample) {
        row[pos++] = kbase + __ffs(bits) - 1;
        bits &= bits - 1u;
      }
    }
    if (SUMMARY) summ[lane] = 0u;
    // first hit overall = first hit of the lowest lane that has any (lanes own ascending index ranges)
    const unsigned havem = __ballot_sync(0xffffffffu, mine > 0);
    if (total > 0) {
      const int first = __shfl_sync(0xffffffffu, first_local, __ffs(havem) - 1);
      const int cnt = min(total, nsample);
      for (int l = cnt + lane; l < nsample; l += 32) row[l] = first;   // pad with the first hit (tf_grouping_g.cu:26-29)
    }
    if (lane == 0) pts_cnt[(size_t)cloud * m + j] = min(total, nsample);
    __syncwarp();  // the cleared words / summary are visible before the next query's atomics
  }
}

float ball_d2_max(float radius);  // point_ops.cu

}  // namespace vnb

using namespace vnb;

extern "C" size_t vnb_query_ball_point_workspace_bytes(int b, int n) {
  if (b <= 0 || n <= 0) return 256;
  return (size_t)b * grid_slice_bytes(n);
}

constexpr int GQ_MAX_N = 32 * 32 * 32;   // a lane owns at most 32 bitmap words (its summary word has 32 bits)
static bool bq_grid_path(int n, float radius, const void* workspace) {
  return workspace != nullptr && g_bq_variant != 0 && n >= g_bq_grid_min_n && n <= GQ_MAX_N && radius > 1e-20f;
}

// Build half of vnb_query_ball_point_ws: needs the searched set only, so a caller can run it before the queries exist
// (the engine runs it beside the FPS that produces them).  A no-op when the _ws call would take the scan path.
extern "C" int vnb_query_ball_point_prepare(int b, int n, float radius, const float* xyz1, void* workspace, void* stream) {
  VNB_REQUIRE(radius > 0, "QueryBallPoint expects positive radius");          // tf_grouping.cpp:71
  VNB_REQUIRE(b >= 0 && n >= 0, "QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.");
  if (!bq_grid_path(n, radius, workspace) || b == 0) return VNB_OK;
  grid_build_kernel<<<b, BT, 0, as_stream(stream)>>>(n, radius * 1.01f, xyz1, static_cast<char*>(workspace), grid_slice_bytes(n));
  return check_launch("query_ball_point grid build");
}

// Query half: `workspace` was filled by vnb_query_ball_point_prepare(b, n, radius, xyz1, ...) with the same b, n, radius, xyz1.
extern "C" int vnb_query_ball_point_prepared(int b, int n, int m, float radius, int nsample, const float* xyz1,
                                             const float* xyz2, int* idx, int* pts_cnt, void* workspace, void* stream) {
  if (!bq_grid_path(n, radius, workspace))
    return vnb_query_ball_point(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, stream);
  VNB_REQUIRE(radius > 0, "QueryBallPoint expects positive radius");          // tf_grouping.cpp:71
  VNB_REQUIRE(nsample > 0, "QueryBallPoint expects positive nsample");        // tf_grouping.cpp:74
  VNB_REQUIRE(b >= 0 && n >= 0 && m >= 0, "QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.");
  if (b == 0 || m == 0) return VNB_OK;
  cudaStream_t st = as_stream(stream);
  const size_t slice = grid_slice_bytes(n);
  const int nw = (n + 31) / 32;
  const int per = (nw + 31) / 32;   // bitmap words per lane
  const size_t smem = (size_t)GQ_WARPS * (32 + 32 * per) * 4;
  VNB_REQUIRE(per <= 32 && smem <= 200 * 1024, "query_ball_point: n too large for the bitmap path");
  const uint32_t inv = (uint32_t)((65536 + per - 1) / per);
  auto kern = per > GQ_DIRECT ? grid_query2_kernel<true> : grid_query2_kernel<false>;
  if (smem > 48 * 1024)  // (never lower the limit below the default: a profiler that patches the kernel needs the headroom)
    VNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((m + GQ_WARPS - 1) / GQ_WARPS, b);   // one query per warp; the kernel's loop accepts any smaller grid
  kern<<<grid, GQ_WARPS * 32, smem, st>>>(n, m, ball_d2_max(radius), nsample, per, inv, xyz2,
                                          static_cast<const char*>(workspace), slice, idx, pts_cnt);
  return check_launch("query_ball_point grid query");
}

extern "C" int vnb_query_ball_point_ws(int b, int n, int m, float radius, int nsample, const float* xyz1,
                                       const float* xyz2, int* idx, int* pts_cnt, void* workspace, void* stream) {
  if (!bq_grid_path(n, radius, workspace))
    return vnb_query_ball_point(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, stream);
  VNB_REQUIRE(nsample > 0, "QueryBallPoint expects positive nsample");        // tf_grouping.cpp:74
  VNB_REQUIRE(b >= 0 && n >= 0 && m >= 0, "QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.");
  if (b == 0 || m == 0) return VNB_OK;
  if (int rc = vnb_query_ball_point_prepare(b, n, radius, xyz1, workspace, stream)) return rc;
  return vnb_query_ball_point_prepared(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, workspace, stream);
}

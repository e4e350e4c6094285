// Tile packing for the fused SA kernels: skip the ball query's padding rows.
//
// query_ball_point pads a row of fewer than nsample hits with copies of its FIRST hit (tf_grouping_g.cu:26-29), so a
// centroid with cnt hits contributes only cnt distinct grouped rows; the other 64 - cnt rows repeat row 0, produce the
// same activations and cannot change the max-pool (utils.py:132).  On SUN-RGB-D-shaped clouds the mean count is 25 / 11
// / 23 / 26 of 64 at the four levels.  The fused kernels therefore run centroids in SLOTS of 16, 32 or 64 rows — the
// smallest that holds all of a centroid's distinct rows — and a 128-row MMA tile carries 8, 4 or 2 centroids of one
// slot size.  This kernel classifies the centroids by pts_cnt and writes the tile table:
//   hdr[0] = total tiles, hdr[1] = tiles of 64-row slots (first), hdr[2] = tiles of 32-row slots (next); the rest are 16s
//   tile_cid[t*8 + g] = centroid (flat b*m index) in slot g of tile t, -1 = empty slot
// Per-row results do not depend on which tile a row sits in, so the outputs are bit-identical to the unpacked kernels.
// cnt == 0 (empty ball, rows never written by the reference) and pts_cnt == NULL keep the full 64-row slot.
// One CTA: counts -> block scan -> ordered placement (deterministic table).
#include "common.cuh"

namespace vnb {

constexpr int PK_T = 1024;

__device__ __forceinline__ int slot_class(int c) { return (c <= 0 || c > 32) ? 0 : (c > 16 ? 1 : 2); }  // 0: 64, 1: 32, 2: 16

__global__ void __launch_bounds__(PK_T) sa_pack_tiles_kernel(int total, const int* __restrict__ cnt, int* __restrict__ hdr,
                                                             int* __restrict__ tile_cid) {
  __shared__ int s_w[3][PK_T / 32];
  __shared__ int s_tot[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (total + PK_T - 1) / PK_T;
  const int i0 = tid * per, i1 = min(total, i0 + per);
  int c[3] = {0, 0, 0};
  for (int i = i0; i < i1; ++i) ++c[cnt ? slot_class(cnt[i]) : 0];
  int base[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // block-wide exclusive scan of c[k]
    int inc = c[k];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_w[k][warp] = inc;
    base[k] = inc - c[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int v = s_w[k][lane];
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      s_w[k][lane] = inc - v;
      if (lane == 31) s_tot[k] = inc;
    }
  }
  __syncthreads();
  const int n64 = s_tot[0], n32 = s_tot[1], n16 = s_tot[2];
  const int t64 = (n64 + 1) >> 1, t32 = (n32 + 3) >> 2, t16 = (n16 + 7) >> 3;
  if (tid == 0) { hdr[0] = t64 + t32 + t16; hdr[1] = t64; hdr[2] = t32; hdr[3] = 0; }
  int p[3] = {base[0] + s_w[0][warp], base[1] + s_w[1][warp], base[2] + s_w[2][warp]};
  for (int i = i0; i < i1; ++i) {
    const int k = cnt ? slot_class(cnt[i]) : 0;
    const int q = p[k]++;
    int slot;
    if (k == 0) slot = (q >> 1) * 8 + (q & 1);
    else if (k == 1) slot = (t64 + (q >> 2)) * 8 + (q & 3);
    else slot = (t64 + t32 + (q >> 3)) * 8 + (q & 7);
    tile_cid[slot] = i;
  }
}

// relative coordinates (utils.py:51) and flat source-row index of every grouped row: rel[(g*64+s)] = {xyz[idx]-c, row}
// pts_cnt (optional): rows beyond the centroid's slot (16 / 32 / 64 rows, sa_pack.cu) are never read and are skipped.
// Latency-bound (three dependent loads per row: count -> index -> point), so every thread carries GR_U independent rows,
// 256 * k apart: each step of the chain is issued for all of them before the first result is needed.
constexpr int GR_U = 4;
__global__ void __launch_bounds__(256) group_rel_kernel(int n, int m, long long total_rows, const float* __restrict__ xyz,
                                                        const float* __restrict__ new_xyz, const int* __restrict__ idx,
                                                        const int* __restrict__ pts_cnt, float4* __restrict__ rel) {
  const long long t0 = (long long)blockIdx.x * (256 * GR_U) + threadIdx.x;
  bool live[GR_U];
  int pid[GR_U];
#pragma unroll
  for (int u = 0; u < GR_U; ++u) {
    const long long t = t0 + 256 * u;
    live[u] = t < total_rows;
    if (live[u] && pts_cnt != nullptr) {
      const int c = __ldg(pts_cnt + (t >> 6));
      const int slot = (c <= 0 || c > 32) ? 64 : (c > 16 ? 32 : 16);
      live[u] = (int)(t & 63) < slot;
    }
  }
#pragma unroll
  for (int u = 0; u < GR_U; ++u) pid[u] = live[u] ? __ldg(idx + t0 + 256 * u) : 0;
  float px[GR_U], py[GR_U], pz[GR_U], cx[GR_U], cy[GR_U], cz[GR_U];
#pragma unroll
  for (int u = 0; u < GR_U; ++u) {
    px[u] = py[u] = pz[u] = cx[u] = cy[u] = cz[u] = 0.f;
    if (!live[u]) continue;
    const int g = (int)((t0 + 256 * u) >> 6);
    const float* pp = xyz + ((size_t)(g / m) * n + pid[u]) * 3;
    const float* cc = new_xyz + (size_t)g * 3;
    px[u] = __ldg(pp); py[u] = __ldg(pp + 1); pz[u] = __ldg(pp + 2);
    cx[u] = __ldg(cc); cy[u] = __ldg(cc + 1); cz[u] = __ldg(cc + 2);
  }
#pragma unroll
  for (int u = 0; u < GR_U; ++u) {
    if (!live[u]) continue;
    const long long t = t0 + 256 * u;
    const int bi = (int)(t >> 6) / m;
    rel[t] = make_float4(px[u] - cx[u], py[u] - cy[u], pz[u] - cz[u], __int_as_float(bi * n + pid[u]));
  }
}

void launch_group_rel(int n, int m, long long rows, const float* xyz, const float* new_xyz, const int* idx,
                      const int* pts_cnt, void* rel, cudaStream_t st) {
  group_rel_kernel<<<(unsigned)((rows + 256 * GR_U - 1) / (256 * GR_U)), 256, 0, st>>>(n, m, rows, xyz, new_xyz, idx, pts_cnt,
                                                                                     static_cast<float4*>(rel));
}

// Narrow-input variant (sa1: c <= 4 feature channels): the helper writes the FINISHED fp16 operand row of the first layer,
// a0[(g*64+s)] = {x, y, z, f0, f1, f2, f3, 0} (relative coordinates, utils.py:51, and the gathered features, utils.py:53-55),
// 16 bytes — the first eight K columns of a tile row.  A centroid's rows are contiguous, so the fused kernel stages a tile
// with one bulk copy per slot (sa1_ws2.cu) instead of gathering through registers.
__device__ __forceinline__ uint32_t a0_pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__global__ void __launch_bounds__(256) group_a0_kernel(int n, int m, int c, long long total_rows, const float* __restrict__ xyz,
                                                       const float* __restrict__ feat, const float* __restrict__ new_xyz,
                                                       const int* __restrict__ idx, const int* __restrict__ pts_cnt,
                                                       uint4* __restrict__ a0) {
  const long long t0 = (long long)blockIdx.x * (256 * GR_U) + threadIdx.x;
  bool live[GR_U];
  int pid[GR_U];
#pragma unroll
  for (int u = 0; u < GR_U; ++u) {
    const long long t = t0 + 256 * u;
    live[u] = t < total_rows;
    if (live[u] && pts_cnt != nullptr) {
      const int cn = __ldg(pts_cnt + (t >> 6));
      const int slot = (cn <= 0 || cn > 32) ? 64 : (cn > 16 ? 32 : 16);
      live[u] = (int)(t & 63) < slot;
    }
  }
#pragma unroll
  for (int u = 0; u < GR_U; ++u) pid[u] = live[u] ? __ldg(idx + t0 + 256 * u) : 0;
  float px[GR_U], py[GR_U], pz[GR_U], cx[GR_U], cy[GR_U], cz[GR_U], f[GR_U][4];
#pragma unroll
  for (int u = 0; u < GR_U; ++u) {
    px[u] = py[u] = pz[u] = cx[u] = cy[u] = cz[u] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) f[u][i] = 0.f;
    if (!live[u]) continue;
    const int g = (int)((t0 + 256 * u) >> 6);
    const size_t src = (size_t)(g / m) * n + pid[u];
    const float* pp = xyz + src * 3;
    const float* cc = new_xyz + (size_t)g * 3;
    const float* fp = feat + src * c;
    px[u] = __ldg(pp); py[u] = __ldg(pp + 1); pz[u] = __ldg(pp + 2);
    cx[u] = __ldg(cc); cy[u] = __ldg(cc + 1); cz[u] = __ldg(cc + 2);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < c) f[u][i] = __ldg(fp + i);
  }
#pragma unroll
  for (int u = 0; u < GR_U; ++u) {
    if (!live[u]) continue;
    a0[t0 + 256 * u] = make_uint4(a0_pack2(px[u] - cx[u], py[u] - cy[u]), a0_pack2(pz[u] - cz[u], f[u][0]),
                                  a0_pack2(f[u][1], f[u][2]), a0_pack2(f[u][3], 0.f));
  }
}

void launch_group_a0(int n, int m, int c, long long rows, const float* xyz, const float* feat, const float* new_xyz,
                     const int* idx, const int* pts_cnt, void* a0, cudaStream_t st) {
  group_a0_kernel<<<(unsigned)((rows + 256 * GR_U - 1) / (256 * GR_U)), 256, 0, st>>>(n, m, c, rows, xyz, feat, new_xyz, idx,
                                                                                    pts_cnt, static_cast<uint4*>(a0));
}

// workspace layout of the fused SA path: [rel: rows * 16 B][hdr: 256 B][tile_cid: (total/2 + 4) * 32 B]
size_t sa_rel_bytes(long long rows) { return ((size_t)rows * 16 + 255) / 256 * 256; }
size_t sa_tile_table_bytes(int total_centroids) { return 256 + ((size_t)total_centroids / 2 + 4) * 32; }

int launch_sa_pack(int total_centroids, const int* pts_cnt, int* hdr, int* tile_cid, cudaStream_t st) {
  VNB_CUDA(cudaMemsetAsync(tile_cid, 0xFF, ((size_t)total_centroids / 2 + 4) * 32, st));  // every slot empty (-1)
  sa_pack_tiles_kernel<<<1, PK_T, 0, st>>>(total_centroids, pts_cnt, hdr, tile_cid);
  return check_launch("sa_group_mlp_max: tile packing");
}

}  // namespace vnb

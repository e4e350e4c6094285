// 3-D NMS for B200: score-rank sort -> warp-ballot IoU bitmask -> greedy pass on the bitmask -> global ordering.
//
// Reference: tf_ops/3d_nms/tf_nms3d.cpp (single CPU thread): candidates = boxes with objectness[1] > objectness[0]
// (:230), popped from a max-heap on score over the whole batch (:222-234); a candidate is dropped if any
// already-selected box OF THE SAME CLOUD has IoU3D > thr (:248-255); IoU3D = BEV convex-polygon clip area x
// y-overlap / union (:178-192).  Greedy NMS is order-dependent only through the score order, so per cloud it is
// "walk candidates by descending score; keep iff no kept earlier candidate overlaps" — computed here as:
//   K1  per cloud: rank candidates by (score desc, box index asc)                      [O(k^2) compares, 1 CTA]
//   K2  bit (p,q), q<p, = IOUGreaterThanThreshold(candidate p, earlier candidate q)    [one warp -> one 32-bit word
//       via __ballot_sync; argument order (candidate, selected) as at :250 because the clip is not symmetric in float]
//   K3  per cloud: one warp walks p = 0..ncand-1 with the kept-set as a bitmask in registers
//   K4  rows (batch, box) of the survivors in global descending-score order (the reference's output order)
// This translation unit is compiled with -fmad=false: the reference is g++ -O2 for generic x86-64 (no FMA), so
// every float/double expression below must stay un-fused to reproduce its roundings.
#include "common.cuh"

#include <math.h>

namespace vnb {

// tf_nms3d.cpp:43-46
__device__ __forceinline__ float nms_area2d(const float* bb) {
  return sqrtf((bb[0] - bb[3]) * (bb[0] - bb[3]) + (bb[2] - bb[5]) * (bb[2] - bb[5])) *
         sqrtf((bb[3] - bb[6]) * (bb[3] - bb[6]) + (bb[5] - bb[8]) * (bb[5] - bb[8]));
}
// :48-50
__device__ __forceinline__ float nms_area3d(const float* bb) { return nms_area2d(bb) * (bb[1] - bb[13]); }

// :53-67  even-odd ray cast against corners 0..3 projected on (x,z)
__device__ __forceinline__ bool point_in_polygon(float px, float pz, const float* poly) {
  bool result = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 3) & 3;
    if ((poly[i * 3 + 2] > pz) != (poly[j * 3 + 2] > pz) &&
        (px < (poly[j * 3] - poly[i * 3]) * (pz - poly[i * 3 + 2]) / (poly[j * 3 + 2] - poly[i * 3 + 2]) + poly[i * 3]))
      result = !result;
  }
  return result;
}

#define NMS_MIN(a, b) (((a) < (b)) ? (a) : (b))
#define NMS_MAX(a, b) (((a) > (b)) ? (a) : (b))

// :69-100  segment/segment intersection in double, narrowed to float
__device__ __forceinline__ bool seg_intersect(float ax, float az, float bx, float bz, float cx, float cz, float dx,
                                              float dz, float& ox, float& oz) {
  double A1 = bz - az;
  double B1 = ax - bx;
  double C1 = A1 * ax + B1 * az;
  double A2 = dz - cz;
  double B2 = cx - dx;
  double C2 = A2 * cx + B2 * cz;
  double det = A1 * B2 - A2 * B1;
  if (fabs(det) < 1e-7) return false;
  double x = (B2 * C1 - B1 * C2) / det;
  double z = (A1 * C2 - A2 * C1) / det;
  bool on1 = (NMS_MIN(ax, bx) <= x) && (NMS_MAX(ax, bx) >= x) && (NMS_MIN(az, bz) <= z) && (NMS_MAX(az, bz) >= z);
  bool on2 = (NMS_MIN(cx, dx) <= x) && (NMS_MAX(cx, dx) >= x) && (NMS_MIN(cz, dz) <= z) && (NMS_MAX(cz, dz) >= z);
  if (on1 && on2) {
    ox = (float)x;
    oz = (float)z;
    return true;
  }
  return false;
}

// :122-175  BEV clip area of box1 against box2 (each 8x3 corners; corners 0..3 = top face)
__device__ float intersection2d(const float* b1, const float* b2) {
  float px[24], pz[24], ang[24];
  int np = 0;
  for (int i = 0; i < 4; ++i)
    if (point_in_polygon(b1[i * 3], b1[i * 3 + 2], b2)) { px[np] = b1[i * 3]; pz[np] = b1[i * 3 + 2]; ++np; }
  for (int i = 0; i < 4; ++i)
    if (point_in_polygon(b2[i * 3], b2[i * 3 + 2], b1)) { px[np] = b2[i * 3]; pz[np] = b2[i * 3 + 2]; ++np; }
  for (int i = 0; i < 4; ++i) {
    const int nx = (i + 1) & 3;
    for (int e = 0; e < 4; ++e) {
      const int en = (e + 1) & 3;
      float ox, oz;
      if (seg_intersect(b1[i * 3], b1[i * 3 + 2], b1[nx * 3], b1[nx * 3 + 2], b2[e * 3], b2[e * 3 + 2], b2[en * 3],
                        b2[en * 3 + 2], ox, oz)) {
        px[np] = ox; pz[np] = oz; ++np;
      }
    }
  }
  float mx = 0.f, mz = 0.f;
  for (int i = 0; i < np; ++i) { mx += px[i]; mz += pz[i]; }
  mx /= (float)np;  // 0/0 -> NaN when np == 0; loops below are then empty (area 0), as in the reference
  mz /= (float)np;
  for (int i = 0; i < np; ++i) ang[i] = atan2f(pz[i] - mz, px[i] - mx);
  for (int i = 1; i < np; ++i) {  // sort by angle (:164-166); keys are distinct except for coincident points
    float a = ang[i], x = px[i], z = pz[i];
    int j = i - 1;
    while (j >= 0 && a < ang[j]) { ang[j + 1] = ang[j]; px[j + 1] = px[j]; pz[j + 1] = pz[j]; --j; }
    ang[j + 1] = a; px[j + 1] = x; pz[j + 1] = z;
  }
  float area = 0.f;
  for (int i = 0, j = np - 1; i < np; j = i++)
    area += fabsf((mx * (pz[i] - pz[j]) + px[i] * (pz[j] - mz) + px[j] * (mz - pz[i])) / 2);
  return area;
}

// :178-192
// Exact early exits (same boolean as the full evaluation for every input, 0 <= thr <= 1 as validated at :287-300):
//  * h <= 0 (or NaN): inter3d = 0 * inter2d is 0 or NaN, so iou is +-0 or NaN and `iou > thr` is false;
//  * the BEV bounding rectangles of the two top faces are separated along x or z by more than `eps`: then the clip has
//    no vertex (np == 0, area 0, iou 0 or NaN -> false) —
//      - seg_intersect accepts a point only if it lies inside BOTH segments' coordinate ranges (exact double
//        comparisons, :92-95), impossible for disjoint ranges;
//      - point_in_polygon (:53-67) is false when pz is outside the polygon's z-range (no edge straddles), and for pz
//        inside it the straddling edges come in an even number and the float abscissa
//        (xj-xi)*(pz-zi)/(zj-zi)+xi lies within ~6e-7 * max|coordinate| of [min x, max x] (four roundings), so a query
//        more than eps = 1e-4 * max|coordinate| outside that range sees every comparison `px < abscissa` come out
//        the same and the parity stays even.
//    NaN coordinates fail every comparison below and take the full path.
__device__ __forceinline__ bool iou_trivially_false(const float* bi, const float* bj) {
  const float h = NMS_MIN(bi[1], bj[1]) - NMS_MAX(bi[13], bj[13]);
  if (!(h > 0.f)) return true;
  float lo1x = bi[0], hi1x = bi[0], lo1z = bi[2], hi1z = bi[2], lo2x = bj[0], hi2x = bj[0], lo2z = bj[2], hi2z = bj[2];
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    lo1x = fminf(lo1x, bi[i * 3]); hi1x = fmaxf(hi1x, bi[i * 3]);
    lo1z = fminf(lo1z, bi[i * 3 + 2]); hi1z = fmaxf(hi1z, bi[i * 3 + 2]);
    lo2x = fminf(lo2x, bj[i * 3]); hi2x = fmaxf(hi2x, bj[i * 3]);
    lo2z = fminf(lo2z, bj[i * 3 + 2]); hi2z = fmaxf(hi2z, bj[i * 3 + 2]);
  }
  const float mag = fmaxf(fmaxf(fmaxf(fabsf(lo1x), fabsf(hi1x)), fmaxf(fabsf(lo1z), fabsf(hi1z))),
                          fmaxf(fmaxf(fabsf(lo2x), fabsf(hi2x)), fmaxf(fabsf(lo2z), fabsf(hi2z))));
  const float eps = 1e-4f * mag;
  bool nan = false;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    nan = nan || bi[i * 3] != bi[i * 3] || bi[i * 3 + 2] != bi[i * 3 + 2] || bj[i * 3] != bj[i * 3] || bj[i * 3 + 2] != bj[i * 3 + 2];
  if (nan || !(mag < 1e30f)) return false;
  return hi1x + eps < lo2x || hi2x + eps < lo1x || hi1z + eps < lo2z || hi2z + eps < lo1z;
}

__device__ bool iou_greater_full(const float* bi, const float* bj, float thr) {
  float inter2d = intersection2d(bi, bj);
  float h = NMS_MIN(bi[1], bj[1]) - NMS_MAX(bi[13], bj[13]);
  float inter3d = NMS_MAX(h, 0.f) * inter2d;
  float iou = inter3d / (nms_area3d(bi) + nms_area3d(bj) - inter3d);
  return iou > thr;
}

// K1: per cloud, rank candidates by (score desc, index asc).  order[b][rank] = box, ncand[b].
__global__ void nms_rank_kernel(int k, const float* __restrict__ scores, const float* __restrict__ obj,
                                int* __restrict__ order, int* __restrict__ ncand, uint8_t* __restrict__ keep,
                                int* __restrict__ out_count) {
  extern __shared__ float s_sc[];  // k scores, then k candidate flags (as int)
  int* s_c = reinterpret_cast<int*>(s_sc + k);
  __shared__ int s_n;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s_n = 0;
  if (b == 0 && threadIdx.x == 0) *out_count = 0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    s_sc[i] = scores[(size_t)b * k + i];
    s_c[i] = obj[((size_t)b * k + i) * 2 + 1] > obj[((size_t)b * k + i) * 2] ? 1 : 0;  // :230
    keep[(size_t)b * k + i] = 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    if (!s_c[i]) continue;
    const float si = s_sc[i];
    int rank = 0;
    for (int j = 0; j < k; ++j) rank += (s_c[j] && (s_sc[j] > si || (s_sc[j] == si && j < i))) ? 1 : 0;
    order[(size_t)b * k + rank] = i;
    atomicAdd(&s_n, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) ncand[b] = s_n;
}

// K2: mask[b][p][w] bit (q & 31), q = 32 w + lane < p  <=>  candidate p is suppressed by earlier candidate q.
// The polygon clip is a long divergent chain, but almost every pair is rejected by the cheap exact test above, so the
// work is split: K2a runs the cheap test for every pair (one warp per candidate p, lanes over q) and compacts the
// survivors into one global list (warp ballot -> shared list per CTA -> one global reservation per CTA); K2b runs the
// clip on the dense list, one thread per surviving pair, and sets mask bits with atomicOr (bits are positional, so
// the list order is irrelevant).
constexpr int K2A_WARPS = 8;
__global__ void __launch_bounds__(K2A_WARPS * 32) nms_pairs_kernel(int k, const float* __restrict__ bbox,
                                                                    const int* __restrict__ order,
                                                                    const int* __restrict__ ncand,
                                                                    uint2* __restrict__ pairs, unsigned* __restrict__ npairs) {
  __shared__ uint32_t s_list[K2A_WARPS * 1024];  // k <= 1024: at most k survivors per candidate
  __shared__ unsigned s_n, s_base;
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * K2A_WARPS + warp;
  const int nc = ncand[b];
  if (blockIdx.x * K2A_WARPS >= nc) return;  // whole CTA
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  if (p < nc) {
    const float* bi = bbox + ((size_t)b * k + order[(size_t)b * k + p]) * 24;
    for (int q0 = 0; q0 < p; q0 += 32) {
      const int q = q0 + lane;
      const bool sv = q < p && !iou_trivially_false(bi, bbox + ((size_t)b * k + order[(size_t)b * k + q]) * 24);
      const unsigned m = __ballot_sync(0xffffffffu, sv);
      unsigned base = 0;
      if (lane == 0 && m) base = atomicAdd(&s_n, (unsigned)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (sv) s_list[base + __popc(m & ((1u << lane) - 1u))] = ((uint32_t)p << 16) | (uint32_t)q;
    }
  }
  __syncthreads();
  const unsigned n = s_n;
  if (threadIdx.x == 0 && n) s_base = atomicAdd(npairs, n);
  __syncthreads();
  for (unsigned t = threadIdx.x; t < n; t += blockDim.x) pairs[s_base + t] = make_uint2((unsigned)b, s_list[t]);
}

__global__ void __launch_bounds__(128) nms_clip_kernel(int k, int W, float thr, const float* __restrict__ bbox,
                                                        const int* __restrict__ order, const uint2* __restrict__ pairs,
                                                        const unsigned* __restrict__ npairs, uint32_t* __restrict__ mask) {
  const unsigned n = *npairs;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const uint2 pr = pairs[t];
    const int b = (int)pr.x, p = (int)(pr.y >> 16), q = (int)(pr.y & 0xffffu);
    const float* bi = bbox + ((size_t)b * k + order[(size_t)b * k + p]) * 24;   // argument order (candidate, selected), :250
    const float* bj = bbox + ((size_t)b * k + order[(size_t)b * k + q]) * 24;
    if (iou_greater_full(bi, bj, thr)) atomicOr(&mask[((size_t)b * k + p) * W + (q >> 5)], 1u << (q & 31));
  }
}

// K3: greedy pass.  One CTA per cloud stages the cloud's bitmask rows in shared memory; warp 0 walks them with the
// kept-set held as W<=32 words, one per lane.
__global__ void __launch_bounds__(256) nms_greedy_kernel(int k, int W, const int* __restrict__ order,
                                                          const int* __restrict__ ncand,
                                                          const uint32_t* __restrict__ mask, uint8_t* __restrict__ keep,
                                                          uint8_t* __restrict__ kept_pos, int* __restrict__ out_count) {
  extern __shared__ uint32_t s_mask[];
  const int b = blockIdx.x;
  const int nc = ncand[b];
  for (int t = threadIdx.x; t < nc * W; t += blockDim.x) s_mask[t] = mask[(size_t)b * k * W + t];
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  uint32_t kept = 0;
  int nkept = 0;
  for (int p = 0; p < nc; ++p) {
    uint32_t v = (lane < W) ? (s_mask[p * W + lane] & kept) : 0u;
    if (!__any_sync(0xffffffffu, v != 0u)) {
      if (lane == (p >> 5)) kept |= 1u << (p & 31);
      if (lane == 0) keep[(size_t)b * k + order[(size_t)b * k + p]] = 1;
      ++nkept;
    }
  }
  (void)kept_pos; (void)out_count; (void)nkept;
}

// K4: global order of the survivors: descending score, exact ties by ascending (batch, box).  Also used to merge an
// all-gathered set of per-rank records (strided fields, vnb_merge_detections).  Every CTA compacts the kept entries into
// its shared memory in flat-index order (ballot + prefix; cheap and redundant), then the CTAs split the O(nkept^2)
// ranking: CTA c ranks compacted entries c*256+tid, (c+G)*256+tid, ...
__global__ void __launch_bounds__(256) rank_emit_kernel(int world, int per_rank, int k, const char* __restrict__ sc_base,
                                                          size_t sc_stride, const char* __restrict__ kp_base,
                                                          size_t kp_stride, int* __restrict__ out_idx,
                                                          int* __restrict__ out_count) {
  extern __shared__ __align__(16) char s_dyn[];
  const int total = world * per_rank;
  float* s_sc = reinterpret_cast<float*>(s_dyn);
  int* s_id = reinterpret_cast<int*>(s_sc + total);
  __shared__ int s_wcnt[8];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < total; c0 += 256) {
    const int e = c0 + tid;
    bool kept = false;
    float sc = 0.f;
    if (e < total) {
      const int r = e / per_rank, l = e % per_rank;
      kept = *reinterpret_cast<const uint8_t*>(kp_base + r * kp_stride + l) != 0;
      sc = *reinterpret_cast<const float*>(sc_base + r * sc_stride + (size_t)l * 4);
    }
    const unsigned bm = __ballot_sync(0xffffffffu, kept);
    if (lane == 0) s_wcnt[warp] = __popc(bm);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_wcnt[w];
    if (kept) {
      const int pos = off + __popc(bm & ((1u << lane) - 1u));
      s_sc[pos] = sc;
      s_id[pos] = e;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += s_wcnt[w];
      s_base += t;
    }
    __syncthreads();
  }
  const int nk = s_base;
  // rank of entry i = number of entries that precede it; entries are dealt to the CTAs round-robin so every CTA of the
  // grid has work; the scan reads four entries per shared-memory load (pad entries never precede anything)
  const bool vec = (total & 3) == 0;
  const int nk4 = vec ? ((nk + 3) & ~3) : nk;
  if (tid < nk4 - nk) { s_sc[nk + tid] = -INFINITY; s_id[nk + tid] = 0x7fffffff; }
  __syncthreads();
  for (int i = (int)blockIdx.x + (int)gridDim.x * tid; i < nk; i += (int)gridDim.x * 256) {
    const float se = s_sc[i];
    const int e = s_id[i];
    int rank = 0;
    if (vec) {
      for (int j = 0; j < nk4; j += 4) {
        const float4 sj = *reinterpret_cast<const float4*>(s_sc + j);
        const int4 ij = *reinterpret_cast<const int4*>(s_id + j);
        rank += ((sj.x > se) | ((sj.x == se) & (ij.x < e))) + ((sj.y > se) | ((sj.y == se) & (ij.y < e))) +
                ((sj.z > se) | ((sj.z == se) & (ij.z < e))) + ((sj.w > se) | ((sj.w == se) & (ij.w < e)));
      }
    } else {
      for (int j = 0; j < nk; ++j) rank += (s_sc[j] > se) | ((s_sc[j] == se) & (s_id[j] < e));
    }
    out_idx[rank * 2 + 0] = e / k;
    out_idx[rank * 2 + 1] = e % k;
  }
  if (tid == 0 && blockIdx.x == 0) *out_count = nk;
}

static int launch_rank_emit(int world, int per_rank, int k, const char* sc, size_t scs, const char* kp, size_t kps,
                            int* out_idx, int* out_count, cudaStream_t st) {
  const size_t smem = (size_t)world * per_rank * 8;
  if (smem > 200 * 1024) return set_err(VNB_ERR_INVALID, "nms: more than 25600 boxes in one ordering pass");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rank_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_err(VNB_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  int grid = (world * per_rank + 63) / 64;
  if (grid > 32) grid = 32;
  rank_emit_kernel<<<grid, 256, smem, st>>>(world, per_rank, k, sc, scs, kp, kps, out_idx, out_count);
  return check_launch("nms3d order");
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace vnb

using namespace vnb;

extern "C" size_t vnb_nms3d_workspace_bytes(int b, int k) {
  if (b <= 0 || k <= 0) return 256;
  size_t W = (size_t)(k + 31) / 32;
  return align256((size_t)b * k * 4) + align256((size_t)b * 4) + align256((size_t)b * k * W * 4) + 256 /* pair counter */ +
         align256((size_t)b * k * (size_t)(k > 1 ? k - 1 : 1) / 2 * 8 + 8) /* surviving (candidate, earlier) pairs */;
}

extern "C" int vnb_nms3d(int b, int k, const float* bbox, const float* scores, const float* objectiveness,
                         float iou_threshold, uint8_t* keep, int* out_idx, int* out_count, void* workspace,
                         void* stream) {
  VNB_REQUIRE(b >= 0 && k >= 0, "3D NMS expects (batch_size, nbbox, 8, 3) bbox shape.");            // tf_nms3d.cpp:287
  VNB_REQUIRE(iou_threshold >= 0 && iou_threshold <= 1, "iou_threshold must be in [0, 1]");        // :300
  VNB_REQUIRE(k <= 1024, "nms3d: at most 1024 boxes per cloud (got %d)", k);
  cudaStream_t st = as_stream(stream);
  if (b == 0 || k == 0) {
    VNB_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int), st));
    return VNB_OK;
  }
  const int W = (k + 31) / 32;
  char* ws = static_cast<char*>(workspace);
  int* order = reinterpret_cast<int*>(ws);
  int* ncand = reinterpret_cast<int*>(ws + align256((size_t)b * k * 4));
  uint32_t* mask = reinterpret_cast<uint32_t*>(ws + align256((size_t)b * k * 4) + align256((size_t)b * 4));
  nms_rank_kernel<<<b, 256, (size_t)k * 8, st>>>(k, scores, objectiveness, order, ncand, keep, out_count);
  if (int rc = check_launch("nms3d rank")) return rc;
  unsigned* npairs = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(mask) + align256((size_t)b * k * W * 4));
  uint2* pairs = reinterpret_cast<uint2*>(reinterpret_cast<char*>(npairs) + 256);
  VNB_CUDA(cudaMemsetAsync(mask, 0, align256((size_t)b * k * W * 4) + 256, st));  // mask words + pair counter
  nms_pairs_kernel<<<dim3((k + K2A_WARPS - 1) / K2A_WARPS, b), K2A_WARPS * 32, 0, st>>>(k, bbox, order, ncand, pairs, npairs);
  if (int rc = check_launch("nms3d pairs")) return rc;
  nms_clip_kernel<<<296, 128, 0, st>>>(k, W, iou_threshold, bbox, order, pairs, npairs, mask);
  if (int rc = check_launch("nms3d clip")) return rc;
  size_t smem = (size_t)k * W * 4;
  if (smem > 48 * 1024)
    VNB_CUDA(cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_greedy_kernel<<<b, 256, smem, st>>>(k, W, order, ncand, mask, keep, nullptr, out_count);
  if (int rc = check_launch("nms3d greedy")) return rc;
  return launch_rank_emit(1, b * k, k, reinterpret_cast<const char*>(scores), 0, reinterpret_cast<const char*>(keep), 0,
                          out_idx, out_count, st);
}

extern "C" int vnb_merge_detections(int world, int b, int k, const void* gathered, size_t rank_stride,
                                    size_t off_scores, size_t off_keep, int* out_idx, int* out_count, void* stream) {
  VNB_REQUIRE(world > 0 && b >= 0 && k >= 0, "merge_detections: bad shape");
  cudaStream_t st = as_stream(stream);
  if (world * b * k == 0) {
    VNB_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int), st));
    return VNB_OK;
  }
  const char* g = static_cast<const char*>(gathered);
  return launch_rank_emit(world, b * k, k, g + off_scores, rank_stride, g + off_keep, rank_stride, out_idx, out_count, st);
}

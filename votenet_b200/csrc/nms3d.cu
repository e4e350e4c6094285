// Box decode + 3-D NMS for B200: ONE kernel, one CTA per cloud — decode -> score ranking -> exact pair filter ->
// polygon clip of the surviving pairs -> greedy pass on the suppression bitmask; the last CTA to finish merges the
// per-cloud kept lists (each already in score order) into the reference's global order.
//
// Reference: tf_ops/3d_nms/tf_nms3d.cpp (single CPU thread): candidates = boxes with objectness[1] > objectness[0]
// (:230), popped from a max-heap on score over the whole batch (:222-234); a candidate is dropped if any
// already-selected box OF THE SAME CLOUD has IoU3D > thr (:248-255); IoU3D = BEV convex-polygon clip area x
// y-overlap / union (:178-192).  Greedy NMS is order-dependent only through the score order, so per cloud it is
// "walk candidates by descending score; keep iff no kept earlier candidate overlaps" — computed here as:
//   1  rank candidates by (score desc, box index asc) on an order-preserving integer key (NaN scores sort last, so
//      the order is total and every index stays in range — the reference's heap is memory-safe with NaN too)
//   2  bit (p,q), q<p, = IOUGreaterThanThreshold(candidate p, earlier candidate q); argument order (candidate,
//      selected) as at :250 because the clip is not symmetric in float.  A cheap EXACT rejection test on per-box
//      summaries in shared memory drops almost every pair; the survivors are clipped one thread per pair
//   3  one warp walks p = 0..ncand-1 with the kept-set as a bitmask in registers
//   4  rows (batch, box) of the survivors in global descending-score order (the reference's output order): a k-way
//      merge by rank — position in the own list + binary searches in the other lists, O(n log n), no re-sort
// This translation unit is compiled with -fmad=false: the reference is g++ -O2 for generic x86-64 (no FMA), so
// every float/double expression below must stay un-fused to reproduce its roundings.
#include "common.cuh"

#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

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

// `near` (optional): counts pairs whose IoU lies within 1e-5 of the threshold — the only pairs on which a last-bit
// difference between this clip and the reference's x86 build (atan2f, the double-precision intersection points) could
// flip the keep mask (hard part 6); the parity tests assert it is 0 on their inputs.
__device__ bool iou_greater_full(const float* bi, const float* bj, float thr, unsigned* near = nullptr) {
  float inter2d = intersection2d(bi, bj);
  float h = NMS_MIN(bi[1], bj[1]) - NMS_MAX(bi[13], bj[13]);
  float inter3d = NMS_MAX(h, 0.f) * inter2d;
  float iou = inter3d / (nms_area3d(bi) + nms_area3d(bj) - inter3d);
  if (near != nullptr && fabsf(iou - thr) < 1e-5f) atomicAdd(near, 1u);
  return iou > thr;
}

// ---------------------------------------------------------------------------------------------------------
// box decode — model.py:100-129.  NH=12, NS=NC=10 (config.py:2-3).  Every float operation is spelled with a
// round-to-nearest intrinsic, so the result does not depend on this translation unit's -fmad setting.
constexpr int NH = 12, NS = 10, NC = 10, PCH = 5 + 2 * NH + 4 * NS + NC;
__device__ void decode_one(int t, const float* __restrict__ pxyz, const float* __restrict__ pout,
                           const float* __restrict__ mean_size, float* __restrict__ bboxes, float* __restrict__ scores,
                           float* __restrict__ objectness, float* __restrict__ class_scores) {
  const float* po = pout + (size_t)t * PCH;
  // argmax = first maximal index (tf.argmax), model.py:115,122
  int sc = 0;
  float best = po[5 + 2 * NH];
  for (int i = 1; i < NS; ++i) {
    float v = po[5 + 2 * NH + i];
    if (v > best) { best = v; sc = i; }
  }
  int hc = 0;
  best = po[5];
  for (int i = 1; i < NH; ++i) {
    float v = po[5 + i];
    if (v > best) { best = v; hc = i; }
  }
  float size[3];
  for (int a = 0; a < 3; ++a) {
    float res = po[5 + 2 * NH + NS + sc * 3 + a];
    size[a] = __fmul_rn(mean_size[sc * 3 + a], fmaxf(__fadd_rn(1.0f, res), 1e-6f));  // :119
  }
  float cx = __fadd_rn(pxyz[t * 3 + 0], po[2]), cy = __fadd_rn(pxyz[t * 3 + 1], po[3]),
        cz = __fadd_rn(pxyz[t * 3 + 2], po[4]);  // :121
  float hres = po[5 + NH + hc];
  const float PI_F = 3.14159265358979323846f;
  float ang = __fdiv_rn(__fmul_rn(__fadd_rn(__fmul_rn((float)hc, 2.0f), hres), PI_F), (float)NH);
  const float TWO_PI = __fmul_rn(2.0f, PI_F);
  // tf.floormod: result takes the sign of the divisor
  float heading = fmodf(ang, TWO_PI);
  if (heading < 0.0f) heading = __fadd_rn(heading, TWO_PI);
  float c = cosf(heading), s = sinf(heading);
  float l = size[0], w = size[1], h = size[2];  // lwh (x,z,y) order, model.py:108
  const float sx[8] = {1, 1, -1, -1, 1, 1, -1, -1};
  const float sy[8] = {1, 1, 1, 1, -1, -1, -1, -1};
  const float sz[8] = {1, -1, -1, 1, 1, -1, -1, 1};
  float* bb = bboxes + (size_t)t * 24;
  for (int k = 0; k < 8; ++k) {
    float x = __fmul_rn(sx[k], __fmul_rn(l, 0.5f)), y = __fmul_rn(sy[k], __fmul_rn(h, 0.5f)),
          z = __fmul_rn(sz[k], __fmul_rn(w, 0.5f));
    // rotation [[c,0,s],[0,1,0],[-s,0,c]] (model.py:107), einsum 'ijkl,ijlm->ijmk'
    bb[k * 3 + 0] = __fadd_rn(__fadd_rn(__fmul_rn(c, x), __fmul_rn(s, z)), cx);
    bb[k * 3 + 1] = __fadd_rn(y, cy);
    bb[k * 3 + 2] = __fadd_rn(__fadd_rn(__fmul_rn(-s, x), __fmul_rn(c, z)), cz);
  }
  float mx = po[PCH - NC];
  for (int i = 0; i < NC; ++i) {
    float v = po[PCH - NC + i];
    class_scores[(size_t)t * NC + i] = v;
    mx = fmaxf(mx, v);
  }
  scores[t] = mx;
  objectness[t * 2 + 0] = po[0];
  objectness[t * 2 + 1] = po[1];
}

__global__ void decode_kernel(int total, const float* __restrict__ pxyz, const float* __restrict__ pout,
                              const float* __restrict__ mean_size, float* __restrict__ bboxes,
                              float* __restrict__ scores, float* __restrict__ objectness,
                              float* __restrict__ class_scores) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) decode_one(t, pxyz, pout, mean_size, bboxes, scores, objectness, class_scores);
}

// ---------------------------------------------------------------------------------------------------------
// Score key: order-preserving map float -> uint32 for non-NaN values (with -0 == +0, as `<` on floats has it);
// NaN -> 0, below every number.  Comparing keys is a TOTAL order, so the ranks below are a permutation whatever the
// scores hold (with float comparisons a NaN score left order[] entries unwritten).
__device__ __forceinline__ uint32_t score_key(float s) {
  if (s != s) return 0u;
  const uint32_t u = __float_as_uint(s == 0.f ? 0.f : s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// per-box summary for the exact rejection tests (iou_trivially_false on precomputed extrema + a separating-axis test)
struct BoxSum { float top, bot, lox, hix, loz, hiz, mag, nan; float cx[4], cz[4]; };

__device__ __forceinline__ BoxSum summarize(const float* bb) {
  BoxSum s;
  s.top = bb[1]; s.bot = bb[13];
  s.lox = bb[0]; s.hix = bb[0]; s.loz = bb[2]; s.hiz = bb[2];
  bool nan = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s.cx[i] = bb[i * 3]; s.cz[i] = bb[i * 3 + 2];
    if (i) {
      s.lox = fminf(s.lox, bb[i * 3]); s.hix = fmaxf(s.hix, bb[i * 3]);
      s.loz = fminf(s.loz, bb[i * 3 + 2]); s.hiz = fmaxf(s.hiz, bb[i * 3 + 2]);
    }
    nan = nan || bb[i * 3] != bb[i * 3] || bb[i * 3 + 2] != bb[i * 3 + 2];
  }
  s.mag = fmaxf(fmaxf(fabsf(s.lox), fabsf(s.hix)), fmaxf(fabsf(s.loz), fabsf(s.hiz)));
  s.nan = nan ? 1.f : 0.f;
  return s;
}

// Separating axis d = (dx, dz): the projections of the two top faces are disjoint with a margin.  Sufficient for "the
// clip has no vertex" by the same argument as the bounding-rectangle test above, for any direction: a gap of
// 2*eps*(|dx|+|dz|) in projected units is a Euclidean gap > eps = 1e-4 * max|coordinate| between the two quads (the
// projections are rounded to ~2e-7 * mag * (|dx|+|dz|)), so (i) a corner of one quad is more than eps away from every
// boundary point of the other, every `px < abscissa` comparison of point_in_polygon (:53-67, abscissa rounded to
// ~6e-7 * mag) comes out as in exact arithmetic and the parity says "outside"; (ii) seg_intersect's double-precision
// point (:69-100) is accepted only inside both segments' coordinate ranges, i.e. only for segments that touch to ~1e-9.
__device__ __forceinline__ bool axis_separates(const BoxSum& a, const BoxSum& b, float dx, float dz, float eps) {
  float amin = a.cx[0] * dx + a.cz[0] * dz, amax = amin, bmin = b.cx[0] * dx + b.cz[0] * dz, bmax = bmin;
#pragma unroll
  for (int i = 1; i < 4; ++i) {
    const float ta = a.cx[i] * dx + a.cz[i] * dz, tb = b.cx[i] * dx + b.cz[i] * dz;
    amin = fminf(amin, ta); amax = fmaxf(amax, ta);
    bmin = fminf(bmin, tb); bmax = fmaxf(bmax, tb);
  }
  const float margin = 2.f * eps * (fabsf(dx) + fabsf(dz));
  return amax + margin < bmin || bmax + margin < amin;
}

// same boolean as iou_trivially_false(bi, bj) for the first two tests — fminf / fmaxf are exact and associative, so the
// extrema and `mag` are the same floats whether they are folded per box or per pair — plus the separating-axis test
// along the edge directions of both top faces (yawed boxes whose bounding rectangles overlap but which do not touch)
__device__ __forceinline__ bool sum_trivially_false(const BoxSum& a, const BoxSum& b) {
  const float h = NMS_MIN(a.top, b.top) - NMS_MAX(a.bot, b.bot);
  if (!(h > 0.f)) return true;
  const float mag = fmaxf(a.mag, b.mag);
  const float eps = 1e-4f * mag;
  if (a.nan != 0.f || b.nan != 0.f || !(mag < 1e30f)) return false;
  if (a.hix + eps < b.lox || b.hix + eps < a.lox || a.hiz + eps < b.loz || b.hiz + eps < a.loz) return true;
  return axis_separates(a, b, a.cx[1] - a.cx[0], a.cz[1] - a.cz[0], eps) || axis_separates(a, b, a.cx[3] - a.cx[0], a.cz[3] - a.cz[0], eps) ||
         axis_separates(a, b, b.cx[1] - b.cx[0], b.cz[1] - b.cz[0], eps) || axis_separates(a, b, b.cx[3] - b.cx[0], b.cz[3] - b.cz[0], eps);
}

// Number of entries of a list sorted by descending key that precede an entry with key `key` coming from ANOTHER list:
// keys greater than it, plus equal keys when the other list is the earlier one (exact score ties are ordered by
// ascending (batch, box), and an earlier list holds the smaller batch ids).
template <class KeyAt>
__device__ __forceinline__ int count_preceding(KeyAt key_at, int n, uint32_t key, bool other_first) {
  int lo = 0, hi = n;  // first position whose key does not precede
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const uint32_t km = key_at(mid);
    if (km > key || (other_first && km == key)) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

constexpr int NMS_T = 512;          // threads per cloud CTA
constexpr int NMS_LIST = 4096;      // surviving pairs staged per pass

struct NmsSmem {  // dynamic shared memory layout for k boxes, W = ceil(k/32) mask words per row; every array 16-byte aligned
  static __host__ __device__ size_t al(size_t x) { return (x + 15) / 16 * 16; }
  static __host__ __device__ size_t sum_off() { return 0; }
  static __host__ __device__ size_t key_off(int k) { return al((size_t)k * sizeof(BoxSum)); }
  static __host__ __device__ size_t order_off(int k) { return key_off(k) + al((size_t)k * 4); }
  static __host__ __device__ size_t kept_off(int k) { return order_off(k) + al((size_t)k * 4); }
  static __host__ __device__ size_t mask_off(int k) { return kept_off(k) + al((size_t)k * 4); }
  static __host__ __device__ size_t list_off(int k) { return mask_off(k) + al((size_t)k * ((k + 31) / 32) * 4); }
  static __host__ __device__ size_t bytes(int k) { return list_off(k) + (size_t)NMS_LIST * 4; }
};

// workspace (global): kept_key u32 [b][k] | kept_box i32 [b][k] | nk i32 [b] | done counter u32
// A cloud is worked on by a CLUSTER of CTAs (8 for k >= 128, else 1): every CTA ranks the candidates (cheap, redundant),
// the (candidate, earlier candidate) rows are dealt round-robin to the cluster's warps, and the polygon clips — a long
// divergent chain, the bulk of the work — run one thread per surviving pair across the whole cluster, setting bits of
// the suppression mask that lives in the shared memory of the cluster's CTA 0 (distributed shared memory atomics).
template <bool DECODE>
__global__ void __launch_bounds__(NMS_T) nms_cloud_kernel(int k, float thr, const float* __restrict__ pxyz,
                                                           const float* __restrict__ pout,
                                                           const float* __restrict__ mean_size, float* bbox, float* scores,
                                                           float* obj, float* class_scores, uint8_t* __restrict__ keep,
                                                           int* __restrict__ out_idx, uint32_t* __restrict__ out_key,
                                                           int* __restrict__ out_count, float* __restrict__ bboxes_pred,
                                                           float* __restrict__ cls_pred, int* __restrict__ batch_idx,
                                                           uint32_t* ws_key, int* ws_box, int* ws_nk, unsigned* ws_done) {
  extern __shared__ __align__(16) unsigned char nsm[];
  BoxSum* s_sum = reinterpret_cast<BoxSum*>(nsm + NmsSmem::sum_off());
  uint32_t* s_key = reinterpret_cast<uint32_t*>(nsm + NmsSmem::key_off(k));
  int* s_order = reinterpret_cast<int*>(nsm + NmsSmem::order_off(k));         // rank -> box
  int* s_kept = reinterpret_cast<int*>(nsm + NmsSmem::kept_off(k));           // candidate flags, later kept rank positions
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(nsm + NmsSmem::mask_off(k));
  uint32_t* s_list = reinterpret_cast<uint32_t*>(nsm + NmsSmem::list_off(k));
  __shared__ int s_nc, s_nk, s_nlist, s_last;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
  const int b = blockIdx.x / CL, ncloud = gridDim.x / CL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = (k + 31) / 32;
  float* cb = bbox + (size_t)b * k * 24;
  if (tid == 0) { s_nc = 0; s_nk = 0; s_nlist = 0; }
#ifdef NMS_PROF
  long long pt[10]; int pi = 0;
#define NMS_TICK() do { __syncthreads(); pt[pi++] = clock64(); } while (0)
#else
#define NMS_TICK() do {} while (0)
#endif
  NMS_TICK();
  // ---- decode (fused entry point): the cluster's CTAs decode disjoint chunks of the cloud's boxes, then everybody reads
  //      all of them back through L2 for the per-box summaries, keys and candidate flags
  if (DECODE) {
    const int chunk = (k + CL - 1) / CL;
    for (int i = cr * chunk + tid; i < min(k, (cr + 1) * chunk); i += NMS_T)
      decode_one(b * k + i, pxyz, pout, mean_size, bbox, scores, obj, class_scores);
    __threadfence();
    cluster.sync();
  }
  unsigned long long* s_rk = reinterpret_cast<unsigned long long*>(s_list);   // rank keys (the pair list is not in use yet)
  for (int i = tid; i < k; i += NMS_T) {
    const int t = b * k + i;
    float bb[24];
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      const float4 q = __ldcg(reinterpret_cast<const float4*>(cb + (size_t)i * 24) + v);
      bb[v * 4] = q.x; bb[v * 4 + 1] = q.y; bb[v * 4 + 2] = q.z; bb[v * 4 + 3] = q.w;
    }
    s_sum[i] = summarize(bb);
    const float2 ob = __ldcg(reinterpret_cast<const float2*>(obj) + t);
    const bool cand = ob.y > ob.x;   // :230
    const uint32_t key = score_key(__ldcg(scores + t));
    s_key[i] = key;
    // candidates: (key, smaller index first) as one 64-bit number, never 0; everything else 0 and never counted
    s_rk[i] = cand ? (((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i)) : 0ull;
  }
  __syncthreads();
  NMS_TICK();
  // ---- 1. rank of candidate i = number of candidates with a larger rank key
  for (int i = tid; i < k; i += NMS_T) {
    const unsigned long long ki = s_rk[i];
    if (ki == 0ull) continue;
    int rank = 0;
#pragma unroll 8
    for (int j = 0; j < k; ++j) rank += s_rk[j] > ki ? 1 : 0;
    s_order[rank] = i;
    atomicAdd(&s_nc, 1);
  }
  __syncthreads();
  const int nc = s_nc;
  if (cr == 0)
    for (int t = tid; t < nc * W; t += NMS_T) s_mask[t] = 0u;
  NMS_TICK();
  cluster.sync();   // CTA 0's mask is zeroed before anyone sets a bit in it
  NMS_TICK();
  uint32_t* mask0 = cluster.map_shared_rank(s_mask, 0);
  // ---- 2. pair filter + clip.  This CTA owns rows p = 1 + cr, 1 + cr + CL, ...; passes of rows whose pairs fit the list
  for (int p0 = 1 + cr; p0 < nc;) {
    int p1 = p0, tot = 0;
    while (p1 < nc && tot + p1 <= NMS_LIST) { tot += p1; p1 += CL; }
    for (int p = p0 + warp * CL; p < p1; p += (NMS_T / 32) * CL) {
      const BoxSum sp = s_sum[s_order[p]];
      for (int q0 = 0; q0 < p; q0 += 32) {
        const int q = q0 + lane;
        const bool sv = q < p && !sum_trivially_false(sp, s_sum[s_order[q]]);
        const unsigned m = __ballot_sync(0xffffffffu, sv);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_nlist, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (sv) s_list[base + __popc(m & ((1u << lane) - 1u))] = ((uint32_t)p << 16) | (uint32_t)q;
      }
    }
    __syncthreads();
    const int nl = s_nlist;
    for (int t = tid; t < nl; t += NMS_T) {
      const int p = (int)(s_list[t] >> 16), q = (int)(s_list[t] & 0xffffu);
      // argument order (candidate, selected), :250
      if (iou_greater_full(cb + (size_t)s_order[p] * 24, cb + (size_t)s_order[q] * 24, thr, ws_done + 1))
        atomicOr(&mask0[p * W + (q >> 5)], 1u << (q & 31));
    }
    __syncthreads();
    if (tid == 0) s_nlist = 0;
    p0 = p1;
    __syncthreads();
  }
  NMS_TICK();
  cluster.sync();   // every bit is set; CTA 0 finishes the cloud, the others are done
  if (cr != 0) return;
  NMS_TICK();
  // ---- 3. greedy pass, warp 0, in blocks of 32 candidates: lane j owns candidate 32*blk + j.  Suppression by kept
  //      candidates of EARLIER blocks is a parallel AND over the kept words; inside the block, lane 0 resolves the 32
  //      candidates in order with register arithmetic (their intra-block mask words come from shared memory, loads
  //      independent of the running result) — ~15 cycles per candidate instead of a warp vote per candidate.
  if (warp == 0) {
    uint32_t* s_kw = s_list;          // kept words, one per block
    uint32_t* s_m = s_list + 32;      // intra-block mask words of the current block
    int nk = 0;
    for (int blk = 0; blk * 32 < nc; ++blk) {
      const int p = blk * 32 + lane;
      const bool valid = p < nc;
      bool sup = false;
      for (int w = 0; w < blk; ++w) sup = sup || (valid && (s_mask[p * W + w] & s_kw[w]) != 0u);
      const unsigned alive = __ballot_sync(0xffffffffu, valid && !sup);
      s_m[lane] = valid ? s_mask[p * W + blk] : 0u;
      __syncwarp();
      uint32_t kb = 0u;
      if (lane == 0) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j)
          if (((alive >> j) & 1u) && !(s_m[j] & kb)) kb |= 1u << j;
        s_kw[blk] = kb;
      }
      kb = __shfl_sync(0xffffffffu, kb, 0);
      if ((kb >> lane) & 1u) s_kept[nk + __popc(kb & ((1u << lane) - 1u))] = p;
      nk += __popc(kb);
      __syncwarp();
    }
    if (lane == 0) s_nk = nk;
  }
  __syncthreads();
  NMS_TICK();
  const int nk = s_nk;
  for (int i = tid; i < k; i += NMS_T) keep[(size_t)b * k + i] = 0;
  __syncthreads();
  for (int j = tid; j < nk; j += NMS_T) {
    const int box = s_order[s_kept[j]];
    keep[(size_t)b * k + box] = 1;
    ws_key[(size_t)b * k + j] = s_key[box];
    ws_box[(size_t)b * k + j] = box;
  }
  if (tid == 0) ws_nk[b] = nk;
  // ---- 4. the last CTA to get here merges the per-cloud kept lists into the global order
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(ws_done, 1u) == (unsigned)ncloud - 1u) ? 1 : 0;
  __syncthreads();
  NMS_TICK();
#ifdef NMS_PROF
  if (tid == 0 && b == 0) printf("nms cloud0 cta0 cycles: decode %lld rank %lld zero %lld csync %lld pairs+clip %lld csync %lld greedy %lld publish %lld\n",
                                pt[1] - pt[0], pt[2] - pt[1], pt[3] - pt[2], pt[3] - pt[2], pt[4] - pt[3], pt[5] - pt[4], pt[6] - pt[5], pt[7] - pt[6]);
#endif
  if (!s_last) return;
  __threadfence();
  const int B = ncloud;
  // (a) few survivors (the usual case): one 64-bit composite per survivor — score key, then the flat index complemented so
  //     that equal scores order by ascending (batch, box) — bitonic-sorted in shared memory by the whole CTA;
  // (b) otherwise a k-way merge by rank: list offsets in shared memory, every list's keys staged too when they fit
  //     (all of this CTA's dynamic shared memory is free now), else the searches go to L2.
  __shared__ int s_cnt_small[65];   // exclusive prefix of the list lengths when B <= 64
  if (B <= 64) {
    if (tid < B) s_cnt_small[tid + 1] = __ldcg(ws_nk + tid);
    if (tid == 0) s_cnt_small[0] = 0;
    __syncthreads();
    if (tid == 0)
      for (int c = 0; c < B; ++c) s_cnt_small[c + 1] += s_cnt_small[c];
    __syncthreads();
  }
  if (B <= 64) {
    const int tot = s_cnt_small[B];
    // every list is already sorted: lay list c out as a run of kp slots — descending for even c, ascending for odd c,
    // zero-padded — and run only the MERGE stages of the bitonic network (sizes 2*kp .. npow)
    int kp = 1, bp = 1;
    while (kp < k) kp <<= 1;
    while (bp < B) bp <<= 1;
    const int npow = kp * bp;
    if ((size_t)npow * 8 <= NmsSmem::bytes(k)) {
      unsigned long long* s_srt = reinterpret_cast<unsigned long long*>(nsm);
      for (int e = tid; e < npow; e += NMS_T) {
        const int c = e / kp, slot = e - c * kp;
        const int j = (c & 1) ? kp - 1 - slot : slot;
        unsigned long long v = 0ull;
        if (c < B && j < s_cnt_small[c + 1] - s_cnt_small[c]) {
          const uint32_t flat = (uint32_t)(c * k + __ldcg(ws_box + (size_t)c * k + j));
          v = ((unsigned long long)__ldcg(ws_key + (size_t)c * k + j) << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
        }
        s_srt[e] = v;
      }
      NMS_TICK();
      __syncthreads();
      for (int size = 2 * kp; size <= npow; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int t = tid; t < (npow >> 1); t += NMS_T) {
            const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1)), hi = lo | stride;
            const bool desc = (lo & size) == 0;   // descending runs first: the final order is descending
            const unsigned long long a = s_srt[lo], c2 = s_srt[hi];
            if ((a < c2) == desc) { s_srt[lo] = c2; s_srt[hi] = a; }
          }
          __syncthreads();
        }
      NMS_TICK();
      for (int r = tid; r < tot; r += NMS_T) {
        const unsigned long long v = s_srt[r];
        const uint32_t flat = 0xFFFFFFFFu - (uint32_t)v;
        const int c = (int)(flat / (uint32_t)k), box = (int)(flat % (uint32_t)k);
        out_idx[r * 2 + 0] = c;
        out_idx[r * 2 + 1] = box;
        if (out_key != nullptr) out_key[r] = (uint32_t)(v >> 32);
        if (batch_idx != nullptr) batch_idx[r] = c;                                   // model.py:137
        if (bboxes_pred != nullptr) {                                                    // model.py:135
          const float4* src = reinterpret_cast<const float4*>(bbox + ((size_t)c * k + box) * 24);
          float4* dst = reinterpret_cast<float4*>(bboxes_pred + (size_t)r * 24);
          float4 q[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) q[i] = __ldcg(src + i);
#pragma unroll
          for (int i = 0; i < 6; ++i) dst[i] = q[i];
        }
        if (cls_pred != nullptr) {                                                       // model.py:136
          const float2* src = reinterpret_cast<const float2*>(class_scores + ((size_t)c * k + box) * 10);
          float2* dst = reinterpret_cast<float2*>(cls_pred + (size_t)r * 10);
          float2 q[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) q[i] = __ldcg(src + i);
#pragma unroll
          for (int i = 0; i < 5; ++i) dst[i] = q[i];
        }
      }
#ifdef NMS_PROF
      NMS_TICK();
      if (tid == 0) printf("nms last cta (cloud %d): emit stage %lld sort %lld write %lld cycles, total kept %d npow %d\n", b, pt[pi - 3] - pt[pi - 4], pt[pi - 2] - pt[pi - 3], pt[pi - 1] - pt[pi - 2], tot, npow);
#endif
      if (tid == 0) { *out_count = tot; *ws_done = 0u; }
      return;
    }
  }
  int* s_off = reinterpret_cast<int*>(nsm);                      // [B + 1] exclusive prefix of the list lengths
  const size_t head = ((size_t)(B + 1) * 4 + 15) / 16 * 16;
  uint32_t* s_all = reinterpret_cast<uint32_t*>(nsm + head);
  const size_t cap = (NmsSmem::bytes(k) - head) / 4;
  const bool offs_fit = head <= NmsSmem::bytes(k);
  if (tid == 0 && offs_fit) {
    int acc = 0;
    for (int c = 0; c < B; ++c) { s_off[c] = acc; acc += __ldcg(ws_nk + c); }
    s_off[B] = acc;
  }
  __syncthreads();
  int total = 0;
  if (offs_fit) total = s_off[B];
  else for (int c = 0; c < B; ++c) total += __ldcg(ws_nk + c);
  const bool staged = offs_fit && (size_t)total <= cap;
  if (staged) {
    for (int e = tid; e < total; e += NMS_T) {
      int c = 0;
      while (s_off[c + 1] <= e) ++c;
      s_all[e] = __ldcg(ws_key + (size_t)c * k + (e - s_off[c]));
    }
  }
  __syncthreads();
  for (int e = tid; e < total; e += NMS_T) {
    int c = 0, j = e;
    if (offs_fit) { while (s_off[c + 1] <= e) ++c; j = e - s_off[c]; }
    else { int n; while (j >= (n = __ldcg(ws_nk + c))) { j -= n; ++c; } }
    const uint32_t key = staged ? s_all[e] : __ldcg(ws_key + (size_t)c * k + j);
    int rank = j;
    for (int c2 = 0; c2 < B; ++c2) {
      if (c2 == c) continue;
      if (staged) rank += count_preceding([&](int i) { return s_all[s_off[c2] + i]; }, s_off[c2 + 1] - s_off[c2], key, c2 < c);
      else rank += count_preceding([&](int i) { return __ldcg(ws_key + (size_t)c2 * k + i); }, __ldcg(ws_nk + c2), key, c2 < c);
    }
    const int box = __ldcg(ws_box + (size_t)c * k + j);
    out_idx[rank * 2 + 0] = c;
    out_idx[rank * 2 + 1] = box;
    if (out_key != nullptr) out_key[rank] = key;
    if (batch_idx != nullptr) batch_idx[rank] = c;                                  // model.py:137
    if (bboxes_pred != nullptr) {                                                   // model.py:135
      const float4* src = reinterpret_cast<const float4*>(bbox + ((size_t)c * k + box) * 24);
      float4* dst = reinterpret_cast<float4*>(bboxes_pred + (size_t)rank * 24);
      float4 v[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) v[i] = __ldcg(src + i);
#pragma unroll
      for (int i = 0; i < 6; ++i) dst[i] = v[i];
    }
    if (cls_pred != nullptr) {                                                      // model.py:136
      const float2* src = reinterpret_cast<const float2*>(class_scores + ((size_t)c * k + box) * 10);
      float2* dst = reinterpret_cast<float2*>(cls_pred + (size_t)rank * 10);
      float2 v[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) v[i] = __ldcg(src + i);
#pragma unroll
      for (int i = 0; i < 5; ++i) dst[i] = v[i];
    }
  }
#ifdef NMS_PROF
  NMS_TICK();
  if (tid == 0) printf("nms last cta (cloud %d): emit %lld cycles, total kept %d staged %d\n", b, pt[pi - 1] - pt[pi - 2], total, (int)staged);
#endif
  if (tid == 0) { *out_count = total; *ws_done = 0u; }   // counter re-armed for the next launch on this workspace
}

// Merge `L` lists, each sorted by descending key (rows (batch, box) + keys + count inside a record of `stride` bytes),
// into the global order; list l holds batches [l*b, (l+1)*b).  Optionally gathers the kept rows' boxes / class scores /
// batch ids (model.py:135-137).  One thread per entry; every CTA stages all keys in shared memory when they fit.
__global__ void __launch_bounds__(256) merge_lists_kernel(int L, int b, int k, const char* __restrict__ base, size_t stride,
                                                          size_t off_idx, size_t off_key, size_t off_count,
                                                          size_t off_bboxes, size_t off_cls, int* __restrict__ out_idx,
                                                          int* __restrict__ out_count, float* __restrict__ bboxes_pred,
                                                          float* __restrict__ cls_pred, int* __restrict__ batch_idx,
                                                          int smem_words) {
  extern __shared__ uint32_t s_all[];
  __shared__ int s_cnt[64];   // L <= 64
  const int tid = threadIdx.x;
  if (tid < L) s_cnt[tid] = min(max(*reinterpret_cast<const int*>(base + tid * stride + off_count), 0), b * k);
  __syncthreads();
  int total = 0;
  for (int l = 0; l < L; ++l) total += s_cnt[l];
  if (blockIdx.x == 0 && tid == 0) *out_count = total;
  const int e0 = blockIdx.x * 256;
  if (e0 >= total) return;
  const bool staged = total <= smem_words;
  if (staged) {
    int off = 0;
    for (int l = 0; l < L; ++l) {
      const uint32_t* kl = reinterpret_cast<const uint32_t*>(base + l * stride + off_key);
      for (int j = tid; j < s_cnt[l]; j += 256) s_all[off + j] = kl[j];
      off += s_cnt[l];
    }
    __syncthreads();
  }
  const int e = e0 + tid;
  if (e >= total) return;
  int l = 0, j = e, offl = 0;
  while (j >= s_cnt[l]) { j -= s_cnt[l]; offl += s_cnt[l]; ++l; }
  const char* rl = base + l * stride;
  const uint32_t key = staged ? s_all[offl + j] : reinterpret_cast<const uint32_t*>(rl + off_key)[j];
  int rank = j, off2 = 0;
  for (int l2 = 0; l2 < L; ++l2) {
    const int n2 = s_cnt[l2];
    if (l2 != l) {
      const uint32_t* k2 = reinterpret_cast<const uint32_t*>(base + l2 * stride + off_key);
      if (staged) rank += count_preceding([&](int i) { return s_all[off2 + i]; }, n2, key, l2 < l);
      else rank += count_preceding([&](int i) { return k2[i]; }, n2, key, l2 < l);
    }
    off2 += n2;
  }
  const int* il = reinterpret_cast<const int*>(rl + off_idx);
  const int lb = il[j * 2], box = il[j * 2 + 1];
  out_idx[rank * 2 + 0] = l * b + lb;
  out_idx[rank * 2 + 1] = box;
  if (batch_idx != nullptr) batch_idx[rank] = l * b + lb;
  if (bboxes_pred != nullptr) {
    const float4* src = reinterpret_cast<const float4*>(rl + off_bboxes + ((size_t)lb * k + box) * 96);
    float4* dst = reinterpret_cast<float4*>(bboxes_pred + (size_t)rank * 24);
#pragma unroll
    for (int i = 0; i < 6; ++i) dst[i] = src[i];
  }
  if (cls_pred != nullptr) {
    const float2* src = reinterpret_cast<const float2*>(rl + off_cls + ((size_t)lb * k + box) * 40);
    float2* dst = reinterpret_cast<float2*>(cls_pred + (size_t)rank * 10);
#pragma unroll
    for (int i = 0; i < 5; ++i) dst[i] = src[i];
  }
}

int g_nms_cluster = 8;   // tuning "nms_cluster": CTAs per cloud (1, 2, 4 or 8)

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

template <bool DECODE>
static int launch_nms_cloud(int b, int k, float thr, const float* pxyz, const float* pout, const float* mean_size,
                            float* bbox, float* scores, float* obj, float* cls, uint8_t* keep, int* out_idx,
                            uint32_t* out_key, int* out_count, float* bboxes_pred, float* cls_pred, int* batch_idx,
                            void* workspace, cudaStream_t st) {
  char* ws = static_cast<char*>(workspace);
  uint32_t* ws_key = reinterpret_cast<uint32_t*>(ws);
  int* ws_box = reinterpret_cast<int*>(ws + align256((size_t)b * k * 4));
  int* ws_nk = reinterpret_cast<int*>(ws + 2 * align256((size_t)b * k * 4));
  unsigned* ws_done = reinterpret_cast<unsigned*>(ws + 2 * align256((size_t)b * k * 4) + align256((size_t)b * 4));
  VNB_CUDA(cudaMemsetAsync(ws_done, 0, 2 * sizeof(unsigned), st));   // done counter + near-threshold pair count (a memset node)
  const size_t smem = NmsSmem::bytes(k);
  if (smem > 48 * 1024)
    VNB_CUDA(cudaFuncSetAttribute(nms_cloud_kernel<DECODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  const int CL = k >= 128 ? g_nms_cluster : 1;
  cfg.gridDim = dim3((unsigned)(b * CL));
  cfg.blockDim = dim3(NMS_T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  VNB_CUDA(cudaLaunchKernelEx(&cfg, nms_cloud_kernel<DECODE>, k, thr, pxyz, pout, mean_size, bbox, scores, obj, cls, keep,
                              out_idx, out_key, out_count, bboxes_pred, cls_pred, batch_idx, ws_key, ws_box, ws_nk, ws_done));
  return check_launch(DECODE ? "decode + nms3d" : "nms3d");
}

}  // namespace vnb

using namespace vnb;

extern "C" size_t vnb_nms3d_workspace_bytes(int b, int k) {
  if (b <= 0 || k <= 0) return 256;
  return 2 * align256((size_t)b * k * 4) + align256((size_t)b * 4) + 256;
}

extern "C" size_t vnb_nms3d_near_threshold_offset(int b, int k) {
  if (b <= 0 || k <= 0) return 0;
  return 2 * align256((size_t)b * k * 4) + align256((size_t)b * 4) + 4;
}

extern "C" int vnb_decode_boxes(int b, int k, const float* proposals_xyz, const float* proposals_output,
                                const float* class_mean_size, float* bboxes, float* scores, float* objectness,
                                float* class_scores, void* stream) {
  VNB_REQUIRE(b >= 0 && k >= 0, "decode_boxes: bad shape");
  const int total = b * k;
  if (total == 0) return VNB_OK;
  decode_kernel<<<(total + 127) / 128, 128, 0, as_stream(stream)>>>(total, proposals_xyz, proposals_output, class_mean_size,
                                                                    bboxes, scores, objectness, class_scores);
  return check_launch("decode_boxes");
}

static int nms_args_ok(int b, int k, float thr) {
  VNB_REQUIRE(b >= 0 && k >= 0, "3D NMS expects (batch_size, nbbox, 8, 3) bbox shape.");            // tf_nms3d.cpp:287
  VNB_REQUIRE(thr >= 0 && thr <= 1, "iou_threshold must be in [0, 1]");                             // :300
  VNB_REQUIRE(k <= 1024, "nms3d: at most 1024 boxes per cloud (got %d)", k);
  return VNB_OK;
}

extern "C" int vnb_nms3d(int b, int k, const float* bbox, const float* scores, const float* objectiveness,
                         float iou_threshold, uint8_t* keep, int* out_idx, int* out_count, void* workspace,
                         void* stream) {
  if (int rc = nms_args_ok(b, k, iou_threshold)) return rc;
  cudaStream_t st = as_stream(stream);
  if (b == 0 || k == 0) {
    VNB_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int), st));
    return VNB_OK;
  }
  return launch_nms_cloud<false>(b, k, iou_threshold, nullptr, nullptr, nullptr, const_cast<float*>(bbox),
                                 const_cast<float*>(scores), const_cast<float*>(objectiveness), nullptr, keep, out_idx,
                                 nullptr, out_count, nullptr, nullptr, nullptr, workspace, st);
}

extern "C" int vnb_decode_nms3d(int b, int k, const float* proposals_xyz, const float* proposals_output,
                                const float* class_mean_size, float iou_threshold, float* bboxes, float* scores,
                                float* objectness, float* class_scores, uint8_t* keep, int* out_idx, uint32_t* out_key,
                                int* out_count, float* bboxes_pred, float* class_scores_pred, int* batch_idx,
                                void* workspace, void* stream) {
  if (int rc = nms_args_ok(b, k, iou_threshold)) return rc;
  cudaStream_t st = as_stream(stream);
  if (b == 0 || k == 0) {
    VNB_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int), st));
    return VNB_OK;
  }
  return launch_nms_cloud<true>(b, k, iou_threshold, proposals_xyz, proposals_output, class_mean_size, bboxes, scores,
                                objectness, class_scores, keep, out_idx, out_key, out_count, bboxes_pred, class_scores_pred,
                                batch_idx, workspace, st);
}

extern "C" int vnb_merge_detections(int world, int b, int k, const void* gathered, size_t rank_stride, size_t off_idx,
                                    size_t off_key, size_t off_count, size_t off_bboxes, size_t off_class_scores,
                                    int* out_idx, int* out_count, float* bboxes_pred, float* class_scores_pred,
                                    int* batch_idx, void* stream) {
  VNB_REQUIRE(world > 0 && world <= 64 && b >= 0 && k >= 0, "merge_detections: bad shape (1 <= world <= 64)");
  cudaStream_t st = as_stream(stream);
  if (world * b * k == 0) {
    VNB_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int), st));
    return VNB_OK;
  }
  const long long total = (long long)world * b * k;
  size_t smem = (size_t)(total < 40960 ? total : 40960) * 4;   // stage every key when they fit in 160 KB
  if (smem > 48 * 1024)
    VNB_CUDA(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)((total + 255) / 256);
  merge_lists_kernel<<<grid, 256, smem, st>>>(world, b, k, static_cast<const char*>(gathered), rank_stride, off_idx, off_key,
                                              off_count, off_bboxes, off_class_scores, out_idx, out_count, bboxes_pred,
                                              class_scores_pred, batch_idx, (int)(smem / 4));
  return check_launch("merge_detections");
}

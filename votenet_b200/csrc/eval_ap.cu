// Detection evaluator on the device: rotated 3-D IoU, greedy detection-to-ground-truth matching and VOC average
// precision for one class — SURVEY.md §8(f) rank 3.
//
// Reference: /root/reference/evaluator.py — iou_3d (:26-39, BEV polygon intersection by shapely x height overlap),
// eval_det_cls (:77-151: detections in descending confidence; a detection is a true positive iff its best-IoU ground
// truth of the same image has IoU > ovthresh and was not claimed by an earlier detection), voc_ap (:42-74, area under
// the precision envelope).  The reference runs this in Python loops with one shapely call per (detection, ground
// truth) pair.  Here: one thread per detection clips its box against the ground truths of its image (Sutherland-Hodgman
// on the two convex top faces, double precision like shapely/GEOS), the confidence order is a rank by counting, the
// "first claimant of a ground truth" is an atomicMin over ranks, and one CTA turns the flags into cumulative
// precision / recall and the AP.
#include "common.cuh"

namespace vnb {

struct P2 { double x, z; };

__device__ __forceinline__ double cross2(P2 a, P2 b, P2 c) { return (b.x - a.x) * (c.z - a.z) - (b.z - a.z) * (c.x - a.x); }

__device__ double quad_area_signed(const P2* q) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const P2 a = q[i], b = q[(i + 1) & 3]; s += a.x * b.z - b.x * a.z; }
  return 0.5 * s;
}

// area of the intersection of two convex quadrilaterals (Sutherland-Hodgman: subject A clipped by every edge of B)
__device__ double convex_quad_intersection_area(const P2* A, const P2* Bq) {
  P2 B[4];
  const bool flipB = quad_area_signed(Bq) < 0.0;   // clip edges counter-clockwise
#pragma unroll
  for (int i = 0; i < 4; ++i) B[i] = Bq[flipB ? 3 - i : i];
  P2 poly[12], tmp[12];
  int n = 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) poly[i] = A[i];
  for (int e = 0; e < 4 && n > 0; ++e) {
    const P2 c0 = B[e], c1 = B[(e + 1) & 3];
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const P2 p = poly[i], q = poly[(i + 1) % n];
      const double sp = cross2(c0, c1, p), sq = cross2(c0, c1, q);
      const bool pin = sp >= 0.0, qin = sq >= 0.0;
      if (pin) tmp[m++] = p;
      if (pin != qin) {
        const double t = sp / (sp - sq);
        tmp[m].x = p.x + t * (q.x - p.x);
        tmp[m].z = p.z + t * (q.z - p.z);
        ++m;
      }
    }
    n = m;
    for (int i = 0; i < n; ++i) poly[i] = tmp[i];
  }
  double s = 0.0;
  for (int i = 0; i < n; ++i) { const P2 a = poly[i], b = poly[(i + 1) % n]; s += a.x * b.z - b.x * a.z; }
  return fabs(0.5 * s);
}

// evaluator.py:26-39 on (8,3) corner boxes, corners 0..3 = top face
__device__ double iou_3d_dev(const float* b1, const float* b2) {
  P2 A[4], B[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { A[i].x = b1[i * 3]; A[i].z = b1[i * 3 + 2]; B[i].x = b2[i * 3]; B[i].z = b2[i * 3 + 2]; }
  const double inter_area = convex_quad_intersection_area(A, B);
  const double t1 = b1[1], u1 = b1[13], t2 = b2[1], u2 = b2[13];
  const double inter_vol = inter_area * fmax(0.0, fmin(t1, t2) - fmax(u1, u2));
  const double a1 = fabs(quad_area_signed(A)), a2 = fabs(quad_area_signed(B));
  return inter_vol / (a1 * (t1 - u1) + a2 * (t2 - u2) - inter_vol);
}

__global__ void iou3d_pairs_kernel(int n, const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = iou_3d_dev(a + (size_t)i * 24, b + (size_t)i * 24);
}

// :124-139  best ground truth of the detection's image (first maximum), and :110-118 the confidence rank (descending,
// equal confidences in input order)
__global__ void eval_match_kernel(int nd, const float* __restrict__ det_boxes, const float* __restrict__ det_scores,
                                  const int* __restrict__ det_img, const float* __restrict__ gt_boxes,
                                  const int* __restrict__ gt_offsets, double thr, double* __restrict__ ovmax,
                                  int* __restrict__ jmax, int* __restrict__ rank, int* __restrict__ first_claim) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const int img = det_img[d];
  const int g0 = gt_offsets[img], g1 = gt_offsets[img + 1];
  double best = -INFINITY;
  int bj = -1;
  for (int g = g0; g < g1; ++g) {
    const double iou = iou_3d_dev(det_boxes + (size_t)d * 24, gt_boxes + (size_t)g * 24);
    if (iou > best) { best = iou; bj = g; }
  }
  const float s = det_scores[d];
  int r = 0;
  for (int e = 0; e < nd; ++e) {
    const float se = det_scores[e];
    r += (se > s || (se == s && e < d)) ? 1 : 0;
  }
  ovmax[d] = best;
  jmax[d] = bj;
  rank[d] = r;
  if (best > thr) atomicMin(&first_claim[bj], r);   // :140-147: the earliest detection claims the ground truth
}

// :148-151 + voc_ap (:58-74): one CTA; tp/fp in rank order -> cumulative sums -> recall, precision -> AP
__global__ void __launch_bounds__(1024) eval_ap_kernel(int nd, int npos, double thr, const double* __restrict__ ovmax,
                                                        const int* __restrict__ jmax, const int* __restrict__ rank,
                                                        const int* __restrict__ first_claim, int* __restrict__ tp_sorted,
                                                        double* __restrict__ rec, double* __restrict__ prec,
                                                        double* __restrict__ ap) {
  __shared__ int s_part[1024];
  __shared__ double s_dpart[1024];
  const int tid = threadIdx.x;
  // scatter the true-positive flags into confidence order
  for (int d = tid; d < nd; d += 1024) {
    const bool tp = ovmax[d] > thr && first_claim[jmax[d]] == rank[d];
    tp_sorted[rank[d]] = tp ? 1 : 0;
  }
  __syncthreads();
  // inclusive scan of tp over positions (chunk per thread, then a scan of the chunk sums)
  const int per = (nd + 1023) / 1024;
  const int p0 = min(nd, tid * per), p1 = min(nd, p0 + per);
  int sum = 0;
  for (int p = p0; p < p1; ++p) sum += tp_sorted[p];
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) { int acc = 0; for (int i = 0; i < 1024; ++i) { const int v = s_part[i]; s_part[i] = acc; acc += v; } }
  __syncthreads();
  int ctp = s_part[tid];
  for (int p = p0; p < p1; ++p) {
    ctp += tp_sorted[p];
    const double tpc = (double)ctp, fpc = (double)(p + 1 - ctp);
    rec[p] = tpc / (double)npos;                                   // :150 (npos == 0 gives inf/nan like the reference)
    prec[p] = tpc / fmax(tpc + fpc, 2.220446049250313e-16);        // :153 np.finfo(np.float64).eps
  }
  __syncthreads();
  // precision envelope from the right (:66-67) with the sentinel mpre[nd+1] = 0, then sum over recall steps (:71-74);
  // sentinels: mrec[0] = 0, mrec[nd+1] = 1.  suffix max by chunks.
  double mx = 0.0;
  for (int p = p1 - 1; p >= p0; --p) mx = fmax(mx, prec[p]);
  s_dpart[tid] = mx;
  __syncthreads();
  if (tid == 0) { double acc = 0.0; for (int i = 1023; i >= 0; --i) { const double v = s_dpart[i]; s_dpart[i] = acc; acc = fmax(acc, v); } }
  __syncthreads();
  double env = s_dpart[tid];   // max of prec over positions right of this chunk
  double part = 0.0;
  for (int p = p1 - 1; p >= p0; --p) {
    env = fmax(env, prec[p]);                       // mpre[p+1] after the envelope
    const double prev = p == 0 ? 0.0 : rec[p - 1];  // mrec[p]
    if (rec[p] != prev) part += (rec[p] - prev) * env;
  }
  s_dpart[tid] = part;
  __syncthreads();
  if (tid == 0) {
    double a = 0.0;
    for (int i = 0; i < 1024; ++i) a += s_dpart[i];
    // last step mrec[nd+1] = 1 vs mrec[nd]: multiplied by mpre[nd+1] = 0 -> contributes nothing
    *ap = a;
  }
}

}  // namespace vnb

using namespace vnb;

extern "C" int vnb_iou3d_pairs(int n, const float* boxes_a, const float* boxes_b, double* out, void* stream) {
  VNB_REQUIRE(n >= 0, "iou3d_pairs: n >= 0");
  if (n == 0) return VNB_OK;
  iou3d_pairs_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(n, boxes_a, boxes_b, out);
  return check_launch("iou3d_pairs");
}

extern "C" size_t vnb_eval_det_cls_workspace_bytes(int nd, int ng) {
  return ((size_t)(nd > 0 ? nd : 1) * (8 + 4 + 4 + 4) + (size_t)(ng > 0 ? ng : 1) * 4 + 1024) / 256 * 256 + 256;
}

extern "C" int vnb_eval_det_cls(int nd, int ng, int nimg, const float* det_boxes, const float* det_scores,
                                const int* det_img, const float* gt_boxes, const int* gt_offsets, double ovthresh,
                                double* rec, double* prec, double* ap, void* workspace, void* stream) {
  VNB_REQUIRE(nd >= 0 && ng >= 0 && nimg >= 0, "eval_det_cls: bad sizes");
  cudaStream_t st = as_stream(stream);
  if (nd == 0) {
    VNB_CUDA(cudaMemsetAsync(ap, 0, sizeof(double), st));
    return VNB_OK;
  }
  char* ws = static_cast<char*>(workspace);
  double* ovmax = reinterpret_cast<double*>(ws);
  int* jmax = reinterpret_cast<int*>(ws + (size_t)nd * 8);
  int* rank = jmax + nd;
  int* tp_sorted = rank + nd;
  int* first_claim = tp_sorted + nd;
  VNB_CUDA(cudaMemsetAsync(first_claim, 0x7f, (size_t)(ng > 0 ? ng : 1) * 4, st));
  eval_match_kernel<<<(nd + 127) / 128, 128, 0, st>>>(nd, det_boxes, det_scores, det_img, gt_boxes, gt_offsets, ovthresh, ovmax,
                                                     jmax, rank, first_claim);
  if (int rc = check_launch("eval match")) return rc;
  eval_ap_kernel<<<1, 1024, 0, st>>>(nd, ng, ovthresh, ovmax, jmax, rank, first_claim, tp_sorted, rec, prec, ap);
  return check_launch("eval ap");
}

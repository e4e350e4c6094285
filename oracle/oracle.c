/* TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load this library; the product path (votenet_b200/) never does.
 *
 * CPU restatement ("oracle") of the reference VoteNet op algorithms, function by function.  Every function
 * cites the reference file:line it follows (paths relative to /root/reference).  Pinned against the real
 * reference: oracle/_ref/libvotenet_ref_cpu.so (the reference's own tf_interpolate.cpp / tf_nms3d.cpp compiled
 * unmodified) in tests/test_oracle.py, and against the reference's own GPU kernels
 * (oracle/_ref/libvotenet_ref_gpu.so) on the GPU box in tests/test_gpu_ref_kernels.py.
 *
 * Build: gcc -O2 -ffp-contract=off (see oracle/Makefile); single-threaded like the reference CPU ops — callers
 * that want all cores (bench.py --impl reference) fan clouds out over a thread pool (ctypes drops the GIL).
 * -ffp-contract=off so that the only fused multiply-adds are the ones spelled fmaf() below (which mirror what
 * nvcc emits for the reference .cu files).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Squared distance exactly as nvcc 12.9 -O2 (default -fmad=true) contracts the reference expression
 * (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1)  (tf_sampling_g.cu:142, tf_grouping_g.cu:24):
 * SASS is FMUL(dy,dy) ; FFMA(dx,dx,.) ; FFMA(dz,dz,.)   (SURVEY.md fact 7 / Appendix A.1). */
static inline float d2_gpu_contracted(float x1, float y1, float z1, float x2, float y2, float z2) {
  float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* ---- farthest point sampling: tf_ops/sampling/tf_sampling_g.cu:105-170 (launch <<<32,512>>> at :204) ----
 * xyz (b,n,3) -> out (b,m) int32.  Reproduces the reference's tie rule: each of the 512 threads keeps the
 * first (lowest k) strict maximum among k = t, t+512, ... (:146-149, best starts at -1 / besti 0 :125-126);
 * the shared-memory tree keeps the LOWER slot on ties (strict '<' at :157).                               */
void vno_fps(int b, int n, int m, const float* xyz, int* out) {
  if (m <= 0) return; /* :106 */
  const int BS = 512;
  for (int i = 0; i < b; ++i) {
    const float* p = xyz + (size_t)i * n * 3;
    float* temp = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int k = 0; k < n; ++k) temp[k] = 1e38f; /* :118 */
    int old = 0;
    out[(size_t)i * m] = old; /* :115-116 */
    float tb[512];
    int tbi[512];
    for (int j = 1; j < m; ++j) {
      float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2]; /* :127-129 */
      for (int t = 0; t < BS; ++t) { tb[t] = -1.f; tbi[t] = 0; }
      for (int k = 0; k < n; ++k) {
        float d = d2_gpu_contracted(x1, y1, z1, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]); /* :142 */
        float td = temp[k];
        float d2 = fminf(d, td); /* :143 */
        if (d2 != td) temp[k] = d2; /* :144-145 */
        int t = k & (BS - 1);       /* thread that owns k (:130) */
        if (d2 > tb[t]) { tb[t] = d2; tbi[t] = k; } /* :146-149 */
      }
      /* tree reduction :153-163: lower slot wins ties == left-to-right scan with strict '>' */
      float best = tb[0];
      int besti = tbi[0];
      for (int t = 1; t < BS; ++t)
        if (best < tb[t]) { best = tb[t]; besti = tbi[t]; }
      old = besti;            /* :165 */
      out[(size_t)i * m + j] = old; /* :166-167 */
    }
    free(temp);
  }
}

/* ---- gather_point: tf_sampling_g.cu:172-181 ---- */
void vno_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j) {
      int a = idx[(size_t)i * m + j];
      for (int c = 0; c < 3; ++c) out[((size_t)i * m + j) * 3 + c] = inp[((size_t)i * n + a) * 3 + c];
    }
}

/* ---- ball query: tf_ops/grouping/tf_grouping_g.cu:3-36 ----
 * first `nsample` indices k (ascending) with max(sqrtf(d2),1e-20f) < radius (:24-25); on the first hit ALL
 * nsample slots are set to k (:26-29); pts_cnt = number of hits (:34).  Rows of empty balls are left
 * untouched (the reference never writes them).                                                            */
void vno_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                          int* idx, int* pts_cnt) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j) {
      const float* p1 = xyz1 + (size_t)i * n * 3;
      const float* q = xyz2 + ((size_t)i * m + j) * 3;
      int* row = idx + ((size_t)i * m + j) * nsample;
      int cnt = 0;
      for (int k = 0; k < n; ++k) {
        if (cnt == nsample) break; /* :16-17 */
        /* reference takes query - point (:24); the square makes the sign irrelevant */
        float dd = d2_gpu_contracted(p1[k * 3 + 0], p1[k * 3 + 1], p1[k * 3 + 2], q[0], q[1], q[2]);
        float d = fmaxf(sqrtf(dd), 1e-20f);
        if (d < radius) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) row[l] = k;
          row[cnt] = k;
          cnt += 1;
        }
      }
      pts_cnt[(size_t)i * m + j] = cnt;
    }
}

/* ---- group_point: tf_grouping_g.cu:40-57 ---- */
void vno_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j)
      for (int k = 0; k < nsample; ++k) {
        int ii = idx[((size_t)i * m + j) * nsample + k];
        memcpy(out + (((size_t)i * m + j) * nsample + k) * c, points + ((size_t)i * n + ii) * c, sizeof(float) * c);
      }
}

/* ---- three_nn: tf_ops/3d_interpolation/tf_interpolate.cpp:60-103 ----
 * d is evaluated in FLOAT, un-fused ((dx*dx + dy*dy) + dz*dz), then widened to double (:73); strict '<'
 * insertion (:74-89) -> earlier k wins ties; outputs are squared distances narrowed to float (:91-96).    */
void vno_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const float* u = xyz1 + ((size_t)i * n + j) * 3;
      const float* kn = xyz2 + (size_t)i * m * 3;
      float x1 = u[0], y1 = u[1], z1 = u[2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        float x2 = kn[k * 3 + 0], y2 = kn[k * 3 + 1], z2 = kn[k * 3 + 2];
        float df = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
        double d = df;
        if (d < best1) {
          best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2; best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      float* dd = dist + ((size_t)i * n + j) * 3;
      int* di = idx + ((size_t)i * n + j) * 3;
      dd[0] = (float)best1; di[0] = besti1;
      dd[1] = (float)best2; di[1] = besti2;
      dd[2] = (float)best3; di[2] = besti3;
    }
}

/* ---- three_interpolate: tf_interpolate.cpp:107-127 (un-fused, left to right, :119) ---- */
void vno_three_interpolate(int b, int m, int c, int n, const float* points, const int* idx, const float* weight,
                           float* out) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const float* w = weight + ((size_t)i * n + j) * 3;
      const int* id = idx + ((size_t)i * n + j) * 3;
      const float* p = points + (size_t)i * m * c;
      float* o = out + ((size_t)i * n + j) * c;
      for (int l = 0; l < c; ++l)
        o[l] = p[(size_t)id[0] * c + l] * w[0] + p[(size_t)id[1] * c + l] * w[1] + p[(size_t)id[2] * c + l] * w[2];
    }
}

/* ================= backward ops =================
 * Sequential float sums in the reference's loop order.  The reference's GPU kernels (and the product) accumulate with
 * atomics, whose order is not fixed: parity for these three is "equal up to the rounding of a reordered sum".      */

/* ---- gather_point_grad: scatteraddpointKernel, tf_sampling_g.cu:183-192 (output zeroed, tf_sampling.cpp:174) ---- */
void vno_gather_point_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g) {
  memset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j) {
      int a = idx[(size_t)i * m + j];
      for (int l = 0; l < 3; ++l) inp_g[((size_t)i * n + a) * 3 + l] += out_g[((size_t)i * m + j) * 3 + l];
    }
}

/* ---- group_point_grad: tf_grouping_g.cu:61-78 (output zeroed, tf_grouping.cpp:203) ---- */
void vno_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                          float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * n * c);
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j)
      for (int k = 0; k < nsample; ++k) {
        int ii = idx[((size_t)i * m + j) * nsample + k];
        for (int l = 0; l < c; ++l)
          grad_points[((size_t)i * n + ii) * c + l] += grad_out[(((size_t)i * m + j) * nsample + k) * c + l];
      }
}

/* ---- three_interpolate_grad: tf_interpolate.cpp:131-153 (output zeroed, :255) ---- */
void vno_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int* idx, const float* weight,
                                float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * m * c);
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const float* w = weight + ((size_t)i * n + j) * 3;
      const int* id = idx + ((size_t)i * n + j) * 3;
      const float* g = grad_out + ((size_t)i * n + j) * c;
      float* gp = grad_points + (size_t)i * m * c;
      for (int l = 0; l < c; ++l) {
        gp[(size_t)id[0] * c + l] += g[l] * w[0];
        gp[(size_t)id[1] * c + l] += g[l] * w[1];
        gp[(size_t)id[2] * c + l] += g[l] * w[2];
      }
    }
}

/* ================= 3-D NMS: tf_ops/3d_nms/tf_nms3d.cpp ================= */

/* :43-46 */
static float area2d(const float* bb) {
  return sqrtf((bb[0] - bb[3]) * (bb[0] - bb[3]) + (bb[2] - bb[5]) * (bb[2] - bb[5])) *
         sqrtf((bb[3] - bb[6]) * (bb[3] - bb[6]) + (bb[5] - bb[8]) * (bb[5] - bb[8]));
}
/* :48-50 */
static float area3d(const float* bb) { return area2d(bb) * (bb[1] - bb[4 * 3 + 1]); }

/* :53-67 even-odd ray cast against the first 4 corners (x,z) */
static int point_in_polygon(float px, float pz, const float* poly) {
  int result = 0;
  for (int i = 0, j = 3; i < 4; j = i++) {
    if ((poly[i * 3 + 2] > pz) != (poly[j * 3 + 2] > pz) &&
        (px < (poly[j * 3] - poly[i * 3]) * (pz - poly[i * 3 + 2]) / (poly[j * 3 + 2] - poly[i * 3 + 2]) + poly[i * 3]))
      result = !result;
  }
  return result;
}

#define VNO_MIN(a, b) (((a) < (b)) ? (a) : (b))
#define VNO_MAX(a, b) (((a) > (b)) ? (a) : (b))

/* :69-100 segment/segment intersection in double, narrowed to float on return */
static int seg_intersect(float ax, float az, float bx, float bz, float cx, float cz, float dx, float dz, float* ox,
                         float* oz) {
  double A1 = bz - az; /* float subtraction, then widened — as in the reference */
  double B1 = ax - bx;
  double C1 = A1 * ax + B1 * az;
  double A2 = dz - cz;
  double B2 = cx - dx;
  double C2 = A2 * cx + B2 * cz;
  double det = A1 * B2 - A2 * B1;
  if (fabs(det) < 1e-7) return 0;
  double x = (B2 * C1 - B1 * C2) / det;
  double z = (A1 * C2 - A2 * C1) / det;
  int on1 = (VNO_MIN(ax, bx) <= x) && (VNO_MAX(ax, bx) >= x) && (VNO_MIN(az, bz) <= z) && (VNO_MAX(az, bz) >= z);
  int on2 = (VNO_MIN(cx, dx) <= x) && (VNO_MAX(cx, dx) >= x) && (VNO_MIN(cz, dz) <= z) && (VNO_MAX(cz, dz) >= z);
  if (on1 && on2) {
    *ox = (float)x;
    *oz = (float)z;
    return 1;
  }
  return 0;
}

/* :122-175.  At most 4 + 4 + 16 points. Sort by atan2f about the centroid (:164-166): the reference uses
 * std::sort (order among EQUAL keys unspecified); this restatement uses a stable insertion sort — identical
 * whenever the keys are distinct, and equal keys almost always mean coincident points (area unaffected). */
float vno_intersection2d(const float* b1, const float* b2) {
  float px[24], pz[24], ang[24];
  int np = 0;
  for (int i = 0; i < 4; ++i)
    if (point_in_polygon(b1[i * 3], b1[i * 3 + 2], b2)) { px[np] = b1[i * 3]; pz[np] = b1[i * 3 + 2]; ++np; }
  for (int i = 0; i < 4; ++i)
    if (point_in_polygon(b2[i * 3], b2[i * 3 + 2], b1)) { px[np] = b2[i * 3]; pz[np] = b2[i * 3 + 2]; ++np; }
  for (int i = 0; i < 4; ++i) {
    int next = (i + 1 == 4) ? 0 : i + 1;
    for (int e = 0; e < 4; ++e) {
      int en = (e + 1 == 4) ? 0 : e + 1;
      float ox, oz;
      if (seg_intersect(b1[i * 3], b1[i * 3 + 2], b1[next * 3], b1[next * 3 + 2], b2[e * 3], b2[e * 3 + 2], b2[en * 3],
                        b2[en * 3 + 2], &ox, &oz)) {
        px[np] = ox; pz[np] = oz; ++np;
      }
    }
  }
  float mx = 0, mz = 0;
  for (int i = 0; i < np; ++i) { mx += px[i]; mz += pz[i]; }
  mx /= (float)np; /* 0/0 = NaN when np == 0; the loops below are then empty (area 0) */
  mz /= (float)np;
  for (int i = 0; i < np; ++i) ang[i] = atan2f(pz[i] - mz, px[i] - mx);
  for (int i = 1; i < np; ++i) {
    float a = ang[i], x = px[i], z = pz[i];
    int j = i - 1;
    while (j >= 0 && a < ang[j]) { ang[j + 1] = ang[j]; px[j + 1] = px[j]; pz[j + 1] = pz[j]; --j; }
    ang[j + 1] = a; px[j + 1] = x; pz[j + 1] = z;
  }
  float area = 0;
  for (int i = 0, j = np - 1; i < np; j = i++)
    area += fabsf((mx * (pz[i] - pz[j]) + px[i] * (pz[j] - mz) + px[j] * (mz - pz[i])) / 2);
  return area;
}

float vno_area2d(const float* bb) { return area2d(bb); }
float vno_area3d(const float* bb) { return area3d(bb); }

/* :178-192. Returns the 3-D IoU the reference thresholds; *gt = (iou > thr). */
float vno_iou3d(const float* bi, const float* bj) {
  float inter2d = vno_intersection2d(bi, bj);
  float ymin = VNO_MIN(bi[1], bj[1]);
  float ymax = VNO_MAX(bi[4 * 3 + 1], bj[4 * 3 + 1]);
  float h = ymin - ymax;
  float inter3d = VNO_MAX(h, 0) * inter2d;
  return inter3d / (area3d(bi) + area3d(bj) - inter3d);
}
int vno_iou_greater(const float* bi, const float* bj, float thr) { return vno_iou3d(bi, bj) > thr; }

typedef struct { int b, k; float score; } vno_cand;

/* libstdc++ std::push_heap / std::pop_heap with comp(a,b) = a.score < b.score, restated so that the pop order
 * among EXACTLY equal scores matches std::priority_queue<Candidate, std::deque<Candidate>, cmp> (:222-226). */
static void heap_push_up(vno_cand* h, int hole, int top, vno_cand v) {
  int parent = (hole - 1) / 2;
  while (hole > top && h[parent].score < v.score) {
    h[hole] = h[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  h[hole] = v;
}
static void heap_adjust(vno_cand* h, int hole, int len, vno_cand v) {
  int top = hole, child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (h[child].score < h[child - 1].score) child--;
    h[hole] = h[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h[hole] = h[child - 1];
    hole = child - 1;
  }
  heap_push_up(h, hole, top, v);
}

/* ---- NonMaxSuppression3D: tf_nms3d.cpp:202-308 ----
 * bbox (b,k,8,3), scores (b,k), objectiveness (b,k,2).  Writes rows (batch, box) in pop order into out_idx
 * (capacity b*k rows) and, if keep != NULL, a (b,k) 0/1 keep mask.  Returns the number of rows, or -1 when
 * iou_threshold is outside [0,1] (:300).                                                                   */
int vno_nms3d(int b, int k, const float* bbox, const float* scores, const float* objectiveness, float thr,
              int* out_idx, unsigned char* keep) {
  if (!(thr >= 0 && thr <= 1)) return -1;
  int total = b * k;
  vno_cand* heap = (vno_cand*)malloc(sizeof(vno_cand) * (size_t)(total > 0 ? total : 1));
  int len = 0;
  for (int i = 0; i < total; ++i) {
    if (objectiveness[i * 2 + 1] > objectiveness[i * 2]) { /* :230 */
      vno_cand c = {i / k, i % k, scores[i]};
      heap[len] = c; /* push_back + push_heap */
      ++len;
      heap_push_up(heap, len - 1, 0, c);
    }
  }
  if (keep && total > 0) memset(keep, 0, (size_t)total);
  int nsel = 0;
  while (len > 0) {
    vno_cand next = heap[0]; /* top() :240 */
    int should_select = 1;
    for (int j = nsel - 1; j >= 0; --j) { /* newest -> oldest :248 */
      if (out_idx[j * 2] == next.b &&
          vno_iou_greater(bbox + ((size_t)next.b * k + next.k) * 24, bbox + ((size_t)next.b * k + out_idx[j * 2 + 1]) * 24,
                          thr)) {
        should_select = 0;
        break;
      }
    }
    if (should_select) {
      out_idx[nsel * 2] = next.b;
      out_idx[nsel * 2 + 1] = next.k;
      if (keep) keep[(size_t)next.b * k + next.k] = 1;
      ++nsel;
    }
    /* pop(): pop_heap + pop_back */
    if (len > 1) {
      vno_cand last = heap[len - 1];
      heap[len - 1] = heap[0];
      heap_adjust(heap, 0, len - 1, last);
    }
    --len;
  }
  free(heap);
  return nsel;
}

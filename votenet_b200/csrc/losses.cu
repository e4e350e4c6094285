// Training losses + label assignment on the device — SURVEY.md §8(f) rank 2 (forward values; the reference obtains the
// gradients from TensorFlow autodiff).
//
// Reference: /root/reference/model.py:62-84 (vote loss: every seed is assigned to its nearest ground-truth centre, and
// supervised only if it lies inside at least one box) and :141-238 (proposal assignment by centre distance with the
// POSITIVE / NEGATIVE thresholds of config.py:4-5, objectness / centre (+ the dual "every box gets its nearest
// proposal" term) / heading / size / semantic losses, total_cost :231).  Dense (B, BB) ground-truth arrays exactly as
// the reference feeds them (run.py pads a batch to its biggest scene; padded rows take part, as in the reference).
// One CTA per cloud accumulates the cloud's partial sums (double) and adds them to 16 global accumulators; a one-thread
// kernel turns them into the means the reference reports.  tf.losses.huber_loss has delta = 1; a mean over an empty set
// is NaN, as tf.reduce_mean gives.
#include "common.cuh"

namespace vnb {

constexpr int LNH = 12, LNS = 10, LNC = 10, LPCH = 5 + 2 * LNH + 4 * LNS + LNC;
enum { A_VOTE = 0, A_OBJ_POS, A_OBJ_NEG, A_CENTER, A_DUAL, A_HCLS, A_HRES, A_SCLS, A_SRES, A_SEM, A_OBJ_OK, A_SEM_OK, A_NPOS, A_NNEG, A_N };

__device__ __forceinline__ double huber1(double x) { const double a = fabs(x); return a <= 1.0 ? 0.5 * x * x : a - 0.5; }

__device__ double cross_entropy(const float* logits, int n, int label) {   // sparse_softmax_cross_entropy_with_logits
  double mx = logits[0];
  for (int i = 1; i < n; ++i) mx = fmax(mx, (double)logits[i]);
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += exp((double)logits[i] - mx);
  return log(s) + mx - (double)logits[label];
}
__device__ bool in_top_1(const float* logits, int n, int label) {          // tf.nn.in_top_k(..., 1): ties count as correct
  for (int i = 0; i < n; ++i)
    if (logits[i] > logits[label]) return false;
  return true;
}

__global__ void __launch_bounds__(256) losses_kernel(int n_seed, int n_prop, int n_box, const float* __restrict__ seeds_xyz,
                                                      const float* __restrict__ votes_xyz, const float* __restrict__ prop_xyz,
                                                      const float* __restrict__ prop_out, const float* __restrict__ bxyz,
                                                      const float* __restrict__ blwh, const float* __restrict__ broty,
                                                      const int* __restrict__ sem_lab, const int* __restrict__ head_lab,
                                                      const float* __restrict__ head_res, const int* __restrict__ size_lab,
                                                      const float* __restrict__ size_res, float pos_thr, float neg_thr,
                                                      double* __restrict__ acc) {
  extern __shared__ float s_box[];   // [n_box][8]: x y z, l/2 w/2 h/2 (as lwh order), cos(-roty), sin(-roty)
  __shared__ double s_red[A_N];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < A_N) s_red[tid] = 0.0;
  for (int j = tid; j < n_box; j += 256) {
    const float* c = bxyz + ((size_t)b * n_box + j) * 3;
    const float* l = blwh + ((size_t)b * n_box + j) * 3;
    const float a = -broty[(size_t)b * n_box + j];                       // :75 rotate by -roty
    float* o = s_box + j * 8;
    o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
    o[3] = l[0] / 2.f; o[4] = l[1] / 2.f; o[5] = l[2] / 2.f;             // :76 bboxes_lwh / 2
    o[6] = cosf(a); o[7] = sinf(a);
  }
  __syncthreads();
  double part[A_N];
#pragma unroll
  for (int i = 0; i < A_N; ++i) part[i] = 0.0;
  // ---- vote loss (:62-84)
  for (int i = tid; i < n_seed; i += 256) {
    const float* sp = seeds_xyz + ((size_t)b * n_seed + i) * 3;
    bool inside = false;
    float best = INFINITY;
    int bj = 0;
    for (int j = 0; j < n_box; ++j) {
      const float* o = s_box + j * 8;
      const float dx = fabsf(sp[0] - o[0]), dy = fabsf(sp[1] - o[1]), dz = fabsf(sp[2] - o[2]);   // :62 tf.abs
      const float rx = o[6] * dx + o[7] * dz, ry = dy, rz = -o[7] * dx + o[6] * dz;               // :64-75
      inside = inside || (rx < o[3] && ry < o[4] && rz < o[5]);                                   // :76-78
      const float nrm = sqrtf(rx * rx + ry * ry + rz * rz);                                       // :80
      if (nrm < best) { best = nrm; bj = j; }                                                     // :81 argmin = first minimum
    }
    if (inside && n_box > 0) {
      const float* v = votes_xyz + ((size_t)b * n_seed + i) * 3;
      const float* o = s_box + bj * 8;
      part[A_VOTE] += (double)(fabsf(v[0] - o[0]) + fabsf(v[1] - o[1]) + fabsf(v[2] - o[2]));     // :86 L1 norm
    }
  }
  // ---- proposal assignment and the per-proposal losses (:141-228)
  for (int i = tid; i < n_prop; i += 256) {
    const float* pp = prop_xyz + ((size_t)b * n_prop + i) * 3;
    const float* po = prop_out + ((size_t)b * n_prop + i) * LPCH;
    float best = INFINITY;
    int bj = 0;
    for (int j = 0; j < n_box; ++j) {
      const float* o = s_box + j * 8;
      const float dx = pp[0] - o[0], dy = pp[1] - o[1], dz = pp[2] - o[2];
      const float d = sqrtf(dx * dx + dy * dy + dz * dz);                                          // :147
      if (d < best) { best = d; bj = j; }                                                          // :148-149
    }
    if (n_box == 0) continue;
    if (best > neg_thr) {                                                                          // :154,160-161
      part[A_OBJ_NEG] += cross_entropy(po, 2, 0);
      part[A_OBJ_OK] += in_top_1(po, 2, 0) ? 1.0 : 0.0;
      part[A_NNEG] += 1.0;
    }
    if (best < pos_thr) {                                                                          // :152
      const size_t g = (size_t)b * n_box + bj;
      const float* o = s_box + bj * 8;
      part[A_NPOS] += 1.0;
      part[A_OBJ_POS] += cross_entropy(po, 2, 1);                                                  // :158-159
      part[A_OBJ_OK] += in_top_1(po, 2, 1) ? 1.0 : 0.0;
      for (int a = 0; a < 3; ++a) part[A_CENTER] += huber1((double)po[2 + a] - (double)(o[a] - pp[a]));   // :169-172
      const int hl = head_lab[g], sl = size_lab[g];
      part[A_HCLS] += cross_entropy(po + 5, LNH, hl);                                              // :185-187
      part[A_HRES] += huber1((double)po[5 + LNH + hl] - (double)head_res[g]);                      // :189-193
      part[A_SCLS] += cross_entropy(po + 5 + 2 * LNH, LNS, sl);                                    // :196-198
      for (int a = 0; a < 3; ++a)
        part[A_SRES] += huber1((double)po[5 + 2 * LNH + LNS + sl * 3 + a] - (double)size_res[g * 3 + a]);   // :200-205
      part[A_SEM] += cross_entropy(po + LPCH - LNC, LNC, sem_lab[g]);                              // :210-214
      part[A_SEM_OK] += in_top_1(po + LPCH - LNC, LNC, sem_lab[g]) ? 1.0 : 0.0;
    }
  }
  // ---- dual centre term: every box takes its nearest proposal (:174-179)
  for (int j = tid; j < n_box; j += 256) {
    const float* o = s_box + j * 8;
    float best = INFINITY;
    int bi = 0;
    for (int i = 0; i < n_prop; ++i) {
      const float* pp = prop_xyz + ((size_t)b * n_prop + i) * 3;
      const float dx = pp[0] - o[0], dy = pp[1] - o[1], dz = pp[2] - o[2];
      const float d = sqrtf(dx * dx + dy * dy + dz * dz);
      if (d < best) { best = d; bi = i; }
    }
    if (n_prop > 0) {
      const float* pp = prop_xyz + ((size_t)b * n_prop + bi) * 3;
      const float* po = prop_out + ((size_t)b * n_prop + bi) * LPCH;
      for (int a = 0; a < 3; ++a) part[A_DUAL] += huber1((double)po[2 + a] - (double)(o[a] - pp[a]));
    }
  }
#pragma unroll
  for (int i = 0; i < A_N; ++i) {
    double v = part[i];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0 && v != 0.0) atomicAdd(&s_red[i], v);
  }
  __syncthreads();
  if (tid < A_N && s_red[tid] != 0.0) atomicAdd(&acc[tid], s_red[tid]);
}

// out: total, vote_reg, obj_cls, box, center, heading_cls, heading_residual, size_cls, size_residual, sem_cls,
//      obj_accuracy, sem_accuracy, n_positive, n_negative
__global__ void losses_finalize_kernel(int b, int n_seed, int n_box, const double* __restrict__ acc, double* __restrict__ out) {
  const double npos = acc[A_NPOS], nneg = acc[A_NNEG];
  const double vote = acc[A_VOTE] / ((double)b * n_seed);                                  // :86 mean over B*N
  const double obj = acc[A_OBJ_POS] / npos + acc[A_OBJ_NEG] / nneg;                        // :162-163
  const double center = acc[A_CENTER] / npos + acc[A_DUAL] / ((double)b * n_box);          // :172,179,182
  const double hcls = acc[A_HCLS] / npos, hres = acc[A_HRES] / npos;
  const double scls = acc[A_SCLS] / npos, sres = acc[A_SRES] / npos;
  const double sem = acc[A_SEM] / npos;
  const double box = center + 0.1 * hcls + hres + 0.1 * scls + sres;                       // :207
  out[0] = vote + 0.5 * obj + box + 0.1 * sem;                                             // :231
  out[1] = vote; out[2] = obj; out[3] = box; out[4] = center; out[5] = hcls; out[6] = hres; out[7] = scls; out[8] = sres;
  out[9] = sem;
  out[10] = acc[A_OBJ_OK] / (npos + nneg);                                                 // :164-166
  out[11] = acc[A_SEM_OK] / npos;                                                          // :216-217
  out[12] = npos; out[13] = nneg;
}

}  // namespace vnb

using namespace vnb;

extern "C" int vnb_votenet_losses(int b, int n_seed, int n_prop, int n_box, const float* seeds_xyz, const float* votes_xyz,
                                  const float* proposals_xyz, const float* proposals_output, const float* bboxes_xyz,
                                  const float* bboxes_lwh, const float* bboxes_roty, const int* semantic_labels,
                                  const int* heading_labels, const float* heading_residuals, const int* size_labels,
                                  const float* size_residuals, float positive_thres, float negative_thres, double* out14,
                                  void* workspace_128_bytes, void* stream) {
  VNB_REQUIRE(b > 0 && n_seed > 0 && n_prop > 0 && n_box >= 0, "votenet_losses: bad shape");
  VNB_REQUIRE((size_t)n_box * 32 <= 200 * 1024, "votenet_losses: at most 6400 ground-truth boxes per cloud");
  cudaStream_t st = as_stream(stream);
  double* acc = static_cast<double*>(workspace_128_bytes);
  VNB_CUDA(cudaMemsetAsync(acc, 0, A_N * sizeof(double), st));
  const size_t smem = (size_t)(n_box > 0 ? n_box : 1) * 32;
  if (smem > 48 * 1024) VNB_CUDA(cudaFuncSetAttribute(losses_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  losses_kernel<<<b, 256, smem, st>>>(n_seed, n_prop, n_box, seeds_xyz, votes_xyz, proposals_xyz, proposals_output, bboxes_xyz,
                                      bboxes_lwh, bboxes_roty, semantic_labels, heading_labels, heading_residuals, size_labels,
                                      size_residuals, positive_thres, negative_thres, acc);
  if (int rc = check_launch("votenet_losses")) return rc;
  losses_finalize_kernel<<<1, 1, 0, st>>>(b, n_seed, n_box, acc, out14);
  return check_launch("votenet_losses finalize");
}

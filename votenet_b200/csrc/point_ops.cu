// HBM-bound point ops: gather_point, query_ball_point, group_point, three_nn, three_interpolate, the FP-module
// front half, concat/split and box decode.  All integer/index outputs are bit-exact with the reference; the
// floating-point contraction of every reference expression is spelled with intrinsics (never left to -fmad).
#include "common.cuh"

#include <math.h>

#include <vector>

namespace vnb {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;
int g_bq_variant = 1;  // 0: brute-force scan, 1: cell grid + sparse-aware index bitmap (needs the workspace entry point)
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
char* err_buf() { return g_err; }
static std::vector<void (*)(int*)>& trap_setters() {
  static std::vector<void (*)(int*)> v;
  return v;
}
void register_trap_setter(void (*fn)(int*)) { trap_setters().push_back(fn); }
int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---------------------------------------------------------------------------------------------------------
// gather_point — reference gatherpointKernel, tf_sampling_g.cu:172-181.  One thread per output point.
__global__ void gather_point_kernel(int n, int m, const float* __restrict__ inp, const int* __restrict__ idx,
                                    float* __restrict__ out, int total) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int bi = t / m;
  int a = idx[t];
  const float* s = inp + ((size_t)bi * n + a) * 3;
  float* d = out + (size_t)t * 3;
  d[0] = s[0];
  d[1] = s[1];
  d[2] = s[2];
}

// ---------------------------------------------------------------------------------------------------------
// query_ball_point — reference query_ball_point_gpu, tf_grouping_g.cu:3-36.
// One WARP per query: lanes test 32 consecutive points per step, __ballot_sync + popc gives the
// order-preserving compaction ("first nsample in index order", :16-17), early exit once nsample hits are found.
// The predicate max(sqrtf(d2),1e-20f) < radius is evaluated as d2 <= d2_max, where d2_max is the largest float
// whose correctly-rounded sqrt is < radius (computed on the host; sqrt is monotone so no decision changes).
constexpr int BQ_WARPS = 8;
__global__ void __launch_bounds__(BQ_WARPS * 32) query_ball_kernel(int n, int m, float d2_max, int nsample,
                                                                    const float* __restrict__ xyz1,
                                                                    const float* __restrict__ xyz2,
                                                                    int* __restrict__ idx, int* __restrict__ pts_cnt) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * BQ_WARPS + (threadIdx.x >> 5);
  const int bi = blockIdx.y;
  if (j >= m) return;
  const float* p = xyz1 + (size_t)bi * n * 3;
  const float* q = xyz2 + ((size_t)bi * m + j) * 3;
  int* row = idx + ((size_t)bi * m + j) * nsample;
  const float qx = q[0], qy = q[1], qz = q[2];
  int cnt = 0;
  int first = -1;
  // 128 points per step: four independent loads / distance tests / ballots in flight, then the four masks are consumed
  // in index order (so compaction order and the early exit are those of the one-point-at-a-time scan)
  for (int k0 = 0; k0 < n && cnt < nsample; k0 += 128) {
    unsigned mask[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + 32 * u + lane;
      bool hit = false;
      if (k < n) {
        const float dx = qx - p[k * 3 + 0], dy = qy - p[k * 3 + 1], dz = qz - p[k * 3 + 2];
        hit = d2_ref_gpu(dx, dy, dz) <= d2_max;
      }
      mask[u] = __ballot_sync(0xffffffffu, hit);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (mask[u] && cnt < nsample) {
        if (first < 0) first = k0 + 32 * u + __ffs(mask[u]) - 1;
        const int pos = cnt + __popc(mask[u] & ((1u << lane) - 1u));
        if (((mask[u] >> lane) & 1u) && pos < nsample) row[pos] = k0 + 32 * u + lane;
        cnt += __popc(mask[u]);
      }
    }
  }
  if (cnt > nsample) cnt = nsample;
  if (cnt > 0)
    for (int l = cnt + lane; l < nsample; l += 32) row[l] = first;  // pad with the first hit (:26-29)
  if (lane == 0) pts_cnt[(size_t)bi * m + j] = cnt;
}

// ---------------------------------------------------------------------------------------------------------
// group_point — reference group_point_gpu, tf_grouping_g.cu:40-57.  One warp per gathered row, lanes over channels.
__global__ void group_point_kernel(int n, int c, int rows_per_batch, const float* __restrict__ points,
                                   const int* __restrict__ idx, float* __restrict__ out, long long total_rows) {
  long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= total_rows) return;
  int lane = threadIdx.x & 31;
  int bi = (int)(r / rows_per_batch);
  int ii = idx[r];
  const float* s = points + ((size_t)bi * n + ii) * c;
  float* d = out + (size_t)r * c;
  for (int l = lane; l < c; l += 32) d[l] = s[l];
}

// ---------------------------------------------------------------------------------------------------------
// three_nn — reference threenn_cpu, tf_interpolate.cpp:60-103.  d is the UN-FUSED float expression
// ((dx*dx + dy*dy) + dz*dz) (the reference is g++ -O2 without FMA), strict '<' insertion so the earlier k wins ties.
// (float)1e40 == +inf for unused slots (m < 3).
// One WARP per unknown point: lane l scans the known points l, l+32, ... keeping its own three best (same strict '<'
// insertion), then three rounds of warp arg-min on (distance, index) merge the 32 sorted triples — the reference's
// sequential scan keeps, among equal distances, the smaller index first, which is exactly the lexicographic minimum.
// Distances are >= 0 (or +inf for an empty slot), so their bit patterns order as unsigned integers.  16x shorter
// dependent chain than a thread per point (the thread-per-point form took 38 us for 1024 x 512 on a handful of SMs).
constexpr int NN_WARPS = 8;
__global__ void __launch_bounds__(NN_WARPS * 32) three_nn_kernel(int n, int m, const float* __restrict__ xyz1,
                                                                  const float* __restrict__ xyz2,
                                                                  float* __restrict__ dist, int* __restrict__ idx) {
  const int bi = blockIdx.y, lane = threadIdx.x & 31;
  const int j = blockIdx.x * NN_WARPS + (threadIdx.x >> 5);
  if (j >= n) return;  // whole warp
  const float* u = xyz1 + ((size_t)bi * n + j) * 3;
  const float x1 = u[0], y1 = u[1], z1 = u[2];
  const float* kn = xyz2 + (size_t)bi * m * 3;
  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;  // (float)1e40, tf_interpolate.cpp:66
  int i1 = 0, i2 = 0, i3 = 0;                         // :67
  for (int k = lane; k < m; k += 32) {
    const float dx = __fsub_rn(kn[k * 3 + 0], x1), dy = __fsub_rn(kn[k * 3 + 1], y1), dz = __fsub_rn(kn[k * 3 + 2], z1);
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));  // :73, un-fused
    if (d < b1) {
      b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
    } else if (d < b2) {
      b3 = b2; i3 = i2; b2 = d; i2 = k;
    } else if (d < b3) {
      b3 = d; i3 = k;
    }
  }
  float od[3];
  int oi[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const unsigned mybits = __float_as_uint(b1);
    unsigned wd;
    asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(wd) : "r"(mybits));
    int wi;
    const int cand = mybits == wd ? i1 : 0x7fffffff;
    asm volatile("redux.sync.min.s32 %0, %1, 0xffffffff;" : "=r"(wi) : "r"(cand));
    od[r] = __uint_as_float(wd);
    oi[r] = wi;
    if (mybits == wd && i1 == wi) {  // pop (several lanes only when the slot is empty everywhere: (+inf, 0))
      b1 = b2; i1 = i2; b2 = b3; i2 = i3; b3 = INFINITY; i3 = 0;
    }
  }
  if (lane == 0) {
    float* dd = dist + ((size_t)bi * n + j) * 3;
    int* di = idx + ((size_t)bi * n + j) * 3;
    dd[0] = od[0]; dd[1] = od[1]; dd[2] = od[2];
    di[0] = oi[0]; di[1] = oi[1]; di[2] = oi[2];
  }
}

// ---------------------------------------------------------------------------------------------------------
// three_interpolate — reference threeinterpolate_cpu, tf_interpolate.cpp:107-127.
// out = p[i1]*w1 + p[i2]*w2 + p[i3]*w3, un-fused, left to right (:119).  One warp per output row.
__global__ void three_interpolate_kernel(int m, int c, int n, const float* __restrict__ points,
                                         const int* __restrict__ idx, const float* __restrict__ weight,
                                         float* __restrict__ out, long long total_rows) {
  long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= total_rows) return;
  int lane = threadIdx.x & 31;
  int bi = (int)(r / n);
  const float w1 = weight[r * 3 + 0], w2 = weight[r * 3 + 1], w3 = weight[r * 3 + 2];
  const float* p = points + (size_t)bi * m * c;
  const float* p1 = p + (size_t)idx[r * 3 + 0] * c;
  const float* p2 = p + (size_t)idx[r * 3 + 1] * c;
  const float* p3 = p + (size_t)idx[r * 3 + 2] * c;
  float* o = out + (size_t)r * c;
  for (int l = lane; l < c; l += 32)
    o[l] = __fadd_rn(__fadd_rn(__fmul_rn(p1[l], w1), __fmul_rn(p2[l], w2)), __fmul_rn(p3[l], w3));
}

// ---------------------------------------------------------------------------------------------------------
// pointnet_fp_module front half, utils.py:279-286: d = max(d,1e-10); w = (1/d) / sum(1/d); interpolate; concat
// [interpolated (c2), skip (c1)].  One warp per unknown point.
__global__ void fp_interp_concat_kernel(int n, int m, int c1, int c2, const float* __restrict__ dist,
                                        const int* __restrict__ idx, const float* __restrict__ points1,
                                        const float* __restrict__ points2, float* __restrict__ out,
                                        long long total_rows) {
  long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= total_rows) return;
  int lane = threadIdx.x & 31;
  int bi = (int)(r / n);
  float d1 = fmaxf(dist[r * 3 + 0], 1e-10f), d2 = fmaxf(dist[r * 3 + 1], 1e-10f), d3 = fmaxf(dist[r * 3 + 2], 1e-10f);
  float r1 = __fdiv_rn(1.0f, d1), r2 = __fdiv_rn(1.0f, d2), r3 = __fdiv_rn(1.0f, d3);
  float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
  float w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm), w3 = __fdiv_rn(r3, norm);
  const float* p = points2 + (size_t)bi * m * c2;
  const float* p1 = p + (size_t)idx[r * 3 + 0] * c2;
  const float* p2 = p + (size_t)idx[r * 3 + 1] * c2;
  const float* p3 = p + (size_t)idx[r * 3 + 2] * c2;
  float* o = out + (size_t)r * (c1 + c2);
  for (int l = lane; l < c2; l += 32)
    o[l] = __fadd_rn(__fadd_rn(__fmul_rn(p1[l], w1), __fmul_rn(p2[l], w2)), __fmul_rn(p3[l], w3));
  if (points1 != nullptr) {
    const float* s = points1 + (size_t)r * c1;
    for (int l = lane; l < c1; l += 32) o[c2 + l] = s[l];
  }
}

__global__ void concat2_kernel(long long rows, int ca, int cb, const float* __restrict__ a,
                               const float* __restrict__ b, float* __restrict__ out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c = ca + cb;
  if (t >= rows * c) return;
  long long r = t / c;
  int l = (int)(t - r * c);
  out[t] = l < ca ? a[r * ca + l] : b[r * cb + (l - ca)];
}
__global__ void split2_kernel(long long rows, int ca, int cb, const float* __restrict__ in, float* __restrict__ a,
                              float* __restrict__ b) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c = ca + cb;
  if (t >= rows * c) return;
  long long r = t / c;
  int l = (int)(t - r * c);
  if (l < ca) a[r * ca + l] = in[t];
  else b[r * cb + (l - ca)] = in[t];
}

// Largest float t with sqrtf(t) < radius (host; IEEE sqrt is correctly rounded on both host and device).
float ball_d2_max(float radius) {
  float t = radius * radius;
  while (t > 0.0f && !(sqrtf(t) < radius)) t = nextafterf(t, 0.0f);
  for (;;) {
    float u = nextafterf(t, INFINITY);
    if (sqrtf(u) < radius) t = u;
    else break;
  }
  return t;
}

}  // namespace vnb

using namespace vnb;

extern "C" {

int vnb_abi_version(void) { return VNB_ABI_VERSION; }
const char* vnb_last_error(void) { return err_buf(); }
// debugging aid (not part of the drop-in boundary): host-mapped int[8] that a bounded wait fills in before it traps
int vnb_debug_trap_buffer(void* host_mapped_int8) {
  for (auto fn : vnb::trap_setters()) fn(static_cast<int*>(host_mapped_int8));
  return VNB_OK;
}
unsigned long long vnb_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int vnb_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out, void* stream) {
  VNB_REQUIRE(b >= 0 && n > 0 && m >= 0, "GatherPoint expects (batch_size,num_points,3) inp shape / (batch_size,num_result) idx shape");
  int total = b * m;
  if (total == 0) return VNB_OK;
  gather_point_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(n, m, inp, idx, out, total);
  return check_launch("gather_point");
}

int vnb_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                         int* idx, int* pts_cnt, void* stream) {
  VNB_REQUIRE(radius > 0, "QueryBallPoint expects positive radius");          // tf_grouping.cpp:71
  VNB_REQUIRE(nsample > 0, "QueryBallPoint expects positive nsample");        // tf_grouping.cpp:74
  VNB_REQUIRE(b >= 0 && n >= 0 && m >= 0, "QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.");
  if (b == 0 || m == 0) return VNB_OK;
  // max(sqrtf(d2),1e-20f) < radius: for radius <= 1e-20f nothing can hit (d2_max < 0 encodes that)
  float d2_max = (radius <= 1e-20f) ? -1.0f : ball_d2_max(radius);
  dim3 grid((m + BQ_WARPS - 1) / BQ_WARPS, b);
  query_ball_kernel<<<grid, BQ_WARPS * 32, 0, as_stream(stream)>>>(n, m, d2_max, nsample, xyz1, xyz2, idx, pts_cnt);
  return check_launch("query_ball_point");
}

int vnb_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out,
                    void* stream) {
  VNB_REQUIRE(b >= 0 && n > 0 && c > 0 && m >= 0 && nsample >= 0, "GroupPoint expects (batch_size, num_points, channel) points shape");
  long long rows = (long long)b * m * nsample;
  if (rows == 0) return VNB_OK;
  int wpb = 8;
  group_point_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, as_stream(stream)>>>(n, c, m * nsample, points,
                                                                                            idx, out, rows);
  return check_launch("group_point");
}

int vnb_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx, void* stream) {
  VNB_REQUIRE(b >= 0 && n >= 0 && m >= 0, "ThreeNN expects (b,n,3) xyz1 shape.");
  if (b == 0 || n == 0) return VNB_OK;
  dim3 grid((n + NN_WARPS - 1) / NN_WARPS, b);
  three_nn_kernel<<<grid, NN_WARPS * 32, 0, as_stream(stream)>>>(n, m, xyz1, xyz2, dist, idx);
  return check_launch("three_nn");
}

int vnb_three_interpolate(int b, int m, int c, int n, const float* points, const int* idx, const float* weight,
                          float* out, void* stream) {
  VNB_REQUIRE(b >= 0 && m > 0 && c > 0 && n >= 0, "ThreeInterpolate expects (b,m,c) points shape");
  long long rows = (long long)b * n;
  if (rows == 0) return VNB_OK;
  int wpb = 8;
  three_interpolate_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, as_stream(stream)>>>(m, c, n, points, idx,
                                                                                                  weight, out, rows);
  return check_launch("three_interpolate");
}

int vnb_fp_interpolate_concat(int b, int n, int m, int c1, int c2, const float* dist, const int* idx,
                              const float* points1, const float* points2, float* out, void* stream) {
  VNB_REQUIRE(b >= 0 && n >= 0 && m > 0 && c1 >= 0 && c2 > 0, "fp_interpolate_concat: bad shape");
  long long rows = (long long)b * n;
  if (rows == 0) return VNB_OK;
  int wpb = 8;
  fp_interp_concat_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, as_stream(stream)>>>(
      n, m, c1, c2, dist, idx, c1 > 0 ? points1 : nullptr, points2, out, rows);
  return check_launch("fp_interpolate_concat");
}

int vnb_concat2(int rows, int ca, int cb, const float* a, const float* b, float* out, void* stream) {
  VNB_REQUIRE(rows >= 0 && ca >= 0 && cb >= 0, "concat2: bad shape");
  long long total = (long long)rows * (ca + cb);
  if (total == 0) return VNB_OK;
  concat2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(rows, ca, cb, a, b, out);
  return check_launch("concat2");
}

int vnb_split2(int rows, int ca, int cb, const float* in, float* a, float* b, void* stream) {
  VNB_REQUIRE(rows >= 0 && ca >= 0 && cb >= 0, "split2: bad shape");
  long long total = (long long)rows * (ca + cb);
  if (total == 0) return VNB_OK;
  split2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(rows, ca, cb, in, a, b);
  return check_launch("split2");
}

}  // extern "C"

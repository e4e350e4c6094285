// fp32 SIMT implementations of the dense layers (precision = 0 of vnb_linear / vnb_sa_group_mlp_max).
// Exact-fp32 mode: used when results tighter than the fp16-operand tensor-core path are wanted, and as the on-device
// cross-check of the tcgen05 kernels.  Same fusion as the tensor-core path: the grouped (m,64,3+C) tensor, the
// concat, every activation and the max-pool input stay on chip (reference materialises all of them,
// utils.py:50-55,125-132).
#include "common.cuh"

namespace vnb {

// out[r, n] = act(sum_k in[r,k] W[k,n] + bias[n]) (+ res[r,n]).   64x64 tile, 16-wide k-steps, 4x4 per thread.
constexpr int LT = 64, LK = 16;
__global__ void __launch_bounds__(256) linear_simt_kernel(int rows, int cin, int cout, const float* __restrict__ in,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          const float* __restrict__ res, int act,
                                                          float* __restrict__ out_f32, __half* __restrict__ out_f16) {
  __shared__ float sA[LK][LT + 1];
  __shared__ float sB[LK][LT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int r0 = blockIdx.x * LT, n0 = blockIdx.y * LT;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < cin; k0 += LK) {
    for (int t = threadIdx.x; t < LT * LK; t += 256) {
      int rr = t / LK, kk = t % LK;
      int r = r0 + rr, k = k0 + kk;
      sA[kk][rr] = (r < rows && k < cin) ? in[(size_t)r * cin + k] : 0.f;
      int kk2 = t / LT, nn = t % LT;
      int k2 = k0 + kk2, n = n0 + nn;
      sB[kk2][nn] = (k2 < cin && n < cout) ? w[(size_t)k2 * cout + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; bb[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= cout) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (act == VNB_ACT_RELU) v = fmaxf(v, 0.f);
      if (res) v += res[(size_t)r * cout + n];
      if (out_f32) out_f32[(size_t)r * cout + n] = v;
      if (out_f16) out_f16[(size_t)r * cout + n] = __float2half_rn(v);
    }
  }
}

// One CTA (256 threads) per centroid: 64 grouped rows through 3 layers, max over the rows.
// Activations ping-pong between two shared buffers of 64 x CMAX floats.
template <int NS>
__global__ void __launch_bounds__(256) sa_simt_kernel(int n, int c, int m, const float* __restrict__ xyz,
                                                      const float* __restrict__ feat, const float* __restrict__ new_xyz,
                                                      const int* __restrict__ idx, int c1, int c2, int c3,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                                      const float* __restrict__ w3, const float* __restrict__ b3,
                                                      float* __restrict__ out, int cmax) {
  extern __shared__ float sm[];
  float* bufA = sm;                 // NS x cmax
  float* bufB = sm + NS * cmax;     // NS x cmax
  const int j = blockIdx.x, bi = blockIdx.y;
  const int cin = 3 + c;
  const float* cen = new_xyz + ((size_t)bi * m + j) * 3;
  const int* row = idx + ((size_t)bi * m + j) * NS;
  // gather + relative xyz + concat (utils.py:50-55)
  for (int t = threadIdx.x; t < NS * cin; t += blockDim.x) {
    int s = t / cin, l = t % cin;
    int pid = row[s];
    float v = (l < 3) ? xyz[((size_t)bi * n + pid) * 3 + l] - cen[l] : feat[((size_t)bi * n + pid) * c + (l - 3)];
    bufA[s * cmax + l] = v;
  }
  __syncthreads();
  auto layer = [&](const float* src, float* dst, int ci, int co, const float* w, const float* b) {
    for (int t = threadIdx.x; t < NS * co; t += blockDim.x) {
      int s = t / co, o = t % co;
      float acc = b[o];
      for (int k = 0; k < ci; ++k) acc = fmaf(src[s * cmax + k], w[(size_t)k * co + o], acc);
      dst[s * cmax + o] = fmaxf(acc, 0.f);
    }
    __syncthreads();
  };
  layer(bufA, bufB, cin, c1, w1, b1);
  layer(bufB, bufA, c1, c2, w2, b2);
  // last layer fused with the max-pool: thread o loops over the NS rows
  for (int o = threadIdx.x; o < c3; o += blockDim.x) {
    float mx = 0.f;  // post-ReLU values are >= 0
    for (int s = 0; s < NS; ++s) {
      float acc = b3[o];
      for (int k = 0; k < c2; ++k) acc = fmaf(bufA[s * cmax + k], w3[(size_t)k * c3 + o], acc);
      mx = fmaxf(mx, fmaxf(acc, 0.f));
    }
    out[((size_t)bi * m + j) * c3 + o] = mx;
  }
}

int linear_simt(int rows, int cin, int cout, const float* in, const float* w, const float* bias, const float* res,
                int act, float* out_f32, void* out_f16, cudaStream_t st) {
  dim3 grid((rows + LT - 1) / LT, (cout + LT - 1) / LT);
  linear_simt_kernel<<<grid, 256, 0, st>>>(rows, cin, cout, in, w, bias, res, act, out_f32,
                                           reinterpret_cast<__half*>(out_f16));
  return check_launch("linear (fp32 simt)");
}

int sa_simt(int b, int n, int c, int m, const float* xyz, const float* feat, const float* new_xyz, const int* idx,
            int c1, int c2, int c3, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
            const float* b3, float* out, cudaStream_t st) {
  int cmax = 3 + c;
  if (c1 > cmax) cmax = c1;
  if (c2 > cmax) cmax = c2;
  size_t smem = (size_t)2 * 64 * cmax * sizeof(float);
  if (smem > 220 * 1024) return set_err(VNB_ERR_INVALID, "sa_group_mlp_max(fp32): channel width %d too large", cmax);
  if (smem > 48 * 1024)
    VNB_CUDA(cudaFuncSetAttribute(sa_simt_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sa_simt_kernel<64><<<dim3(m, b), 256, smem, st>>>(n, c, m, xyz, feat, new_xyz, idx, c1, c2, c3, w1, b1, w2, b2, w3,
                                                    b3, out, cmax);
  return check_launch("sa_group_mlp_max (fp32 simt)");
}

}  // namespace vnb

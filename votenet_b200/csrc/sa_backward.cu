// Backward of the fused set-abstraction layer — SURVEY.md §8(f) rank 1, "(+ grads through fused MLP/max)".
//
// Forward (reference utils.py:49-55,120-132): X[r] = [xyz[idx[r]] - new_xyz, feat[idx[r]]] for the 64 grouped rows of a
// centroid, H1 = relu(X W1 + b1), H2 = relu(H1 W2 + b2), out = max_r relu(H2 W3 + b3).  The reference gets the backward
// from TensorFlow autodiff over materialised (B,m,64,C) tensors (plus GroupPointGrad, tf_grouping_g.cu:61-78).  The fused
// inference kernels keep nothing, so this kernel REMATERIALISES the forward per centroid in fp32 (one CTA walks centroids;
// X, H1, H2 live in shared memory), routes d(out) to each channel's arg-max row, and back-propagates:
//   dW3, db3, dH2 (only the arg-max rows are touched) -> dW2, db2, dH1 -> dW1, db1, dX -> scatter-add into d(feat),
//   d(xyz) and d(new_xyz) (the GroupPointGrad scatter fused in).
// Weight / bias gradients are accumulated with float reductions in global memory (like the reference's own GPU grad
// kernels, the result equals a sequential sum up to summation order).  fp32 SIMT: correctness first — the tensor-core
// forward is 100x faster than this; a tcgen05 backward is the obvious next step.
#include "common.cuh"

namespace vnb {

constexpr int SB_T = 256;    // threads
constexpr int SB_R = 64;     // rows per centroid (nsample)
constexpr int SB_RB = 8;     // rows per register block

// out[64][C] = act(in[64][K] W[K][C] + b): a thread owns (column, block of 8 rows); W is read once per k per block
__device__ void dense_rows(const float* __restrict__ in, int K, const float* __restrict__ W, const float* __restrict__ b, int C,
                           float* __restrict__ out, bool relu) {
  for (int o = threadIdx.x; o < C * (SB_R / SB_RB); o += SB_T) {
    const int c = o % C, r0 = (o / C) * SB_RB;
    float acc[SB_RB];
#pragma unroll
    for (int i = 0; i < SB_RB; ++i) acc[i] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(W + (size_t)k * C + c);
#pragma unroll
      for (int i = 0; i < SB_RB; ++i) acc[i] = fmaf(in[(r0 + i) * K + k], w, acc[i]);
    }
    const float bias = b[c];
#pragma unroll
    for (int i = 0; i < SB_RB; ++i) {
      const float v = acc[i] + bias;
      out[(r0 + i) * C + c] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

// dW[K][C] += in^T[K][64] d[64][C]  (global float reductions), db[C] += column sums of d
__device__ void weight_grad(const float* __restrict__ in, int K, const float* __restrict__ d, int C, float* __restrict__ dW,
                            float* __restrict__ db) {
  for (int o = threadIdx.x; o < K * C; o += SB_T) {
    const int c = o % C, k = o / C;
    float acc = 0.f;
#pragma unroll 8
    for (int r = 0; r < SB_R; ++r) acc = fmaf(in[r * K + k], d[r * C + c], acc);
    if (acc != 0.f) atomicAdd(dW + o, acc);
  }
  for (int c = threadIdx.x; c < C; c += SB_T) {
    float acc = 0.f;
    for (int r = 0; r < SB_R; ++r) acc += d[r * C + c];
    if (acc != 0.f) atomicAdd(db + c, acc);
  }
}

// din[64][K] = (d[64][C] W^T) (* relu'(h) when h != NULL); WT is W transposed, [C][K], so reads are coalesced over k
__device__ void data_grad(const float* __restrict__ d, int C, const float* __restrict__ WT, int K, const float* __restrict__ h,
                          float* __restrict__ din) {
  for (int o = threadIdx.x; o < K * (SB_R / SB_RB); o += SB_T) {
    const int k = o % K, r0 = (o / K) * SB_RB;
    float acc[SB_RB];
#pragma unroll
    for (int i = 0; i < SB_RB; ++i) acc[i] = 0.f;
    for (int c = 0; c < C; ++c) {
      const float w = __ldg(WT + (size_t)c * K + k);
#pragma unroll
      for (int i = 0; i < SB_RB; ++i) acc[i] = fmaf(d[(r0 + i) * C + c], w, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < SB_RB; ++i) {
      const int r = r0 + i;
      din[r * K + k] = (h == nullptr || h[r * K + k] > 0.f) ? acc[i] : 0.f;
    }
  }
}

struct SaBwdArgs {
  int n, c, m, total;                       // points per cloud, feature channels, centroids per cloud, b*m
  int c1, c2, c3;
  const float *xyz, *feat, *new_xyz;
  const int* idx;
  const float *w1, *b1, *w2, *b2, *w3, *b3; // (3+c,c1), (c1,c2), (c2,c3) row-major (TF [Cin,Cout])
  const float *w1t, *w2t;                   // transposes (c1,3+c), (c2,c1)
  const float* dout;                        // (b,m,c3)
  float *dfeat, *dxyz, *dnew_xyz;           // (b,n,c), (b,n,3), (b,m,3): pre-zeroed, accumulated
  float *dw1, *db1, *dw2, *db2, *dw3, *db3; // pre-zeroed, accumulated
};

__global__ void __launch_bounds__(SB_T) sa_backward_kernel(SaBwdArgs a) {
  extern __shared__ __align__(16) float sb[];
  const int K1 = 3 + a.c;
  float* X = sb;                       // [64][K1]
  float* H1 = X + SB_R * K1;           // [64][c1]
  float* H2 = H1 + SB_R * a.c1;        // [64][c2]
  float* D2 = H2 + SB_R * a.c2;        // [64][c2]   dH2
  float* D1 = D2 + SB_R * a.c2;        // [64][c1]   dH1; its head doubles as the layer-3 scratch [groups][c3] max / arg
  __shared__ int s_idx[SB_R];
  const int tid = threadIdx.x;
  for (int g = blockIdx.x; g < a.total; g += gridDim.x) {
    const int bi = g / a.m;
    __syncthreads();   // the previous centroid's buffers are free
    if (tid < SB_R) s_idx[tid] = a.idx[(size_t)g * SB_R + tid];
    __syncthreads();
    // ---- forward, rematerialised
    for (int o = tid; o < SB_R * K1; o += SB_T) {
      const int r = o / K1, k = o % K1;
      const size_t src = (size_t)bi * a.n + s_idx[r];
      X[o] = k < 3 ? a.xyz[src * 3 + k] - a.new_xyz[(size_t)g * 3 + k] : a.feat[src * a.c + (k - 3)];   // utils.py:50-55
    }
    __syncthreads();
    dense_rows(X, K1, a.w1, a.b1, a.c1, H1, true);
    __syncthreads();
    dense_rows(H1, a.c1, a.w2, a.b2, a.c2, H2, true);
    for (int o = tid; o < SB_R * a.c2; o += SB_T) D2[o] = 0.f;
    __syncthreads();
    // ---- layer 3 + max-pool: a thread owns a channel, walks the 64 rows; first maximum wins (torch.max / tf.reduce_max
    //      route the gradient to one arg-max; rows padded by the ball query are copies of row 0 and scatter to the same
    //      source point whichever copy is chosen)
    for (int c = tid; c < a.c3; c += SB_T) {
      float best = -INFINITY;
      int arg = 0;
      for (int r0 = 0; r0 < SB_R; r0 += SB_RB) {
        float acc[SB_RB];
#pragma unroll
        for (int i = 0; i < SB_RB; ++i) acc[i] = 0.f;
        for (int k = 0; k < a.c2; ++k) {
          const float w = __ldg(a.w3 + (size_t)k * a.c3 + c);
#pragma unroll
          for (int i = 0; i < SB_RB; ++i) acc[i] = fmaf(H2[(r0 + i) * a.c2 + k], w, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < SB_RB; ++i)
          if (acc[i] > best) { best = acc[i]; arg = r0 + i; }
      }
      const float gch = (best + a.b3[c] > 0.f) ? a.dout[(size_t)g * a.c3 + c] : 0.f;   // relu after the max
      if (gch != 0.f) {
        atomicAdd(a.db3 + c, gch);
        for (int k = 0; k < a.c2; ++k) {
          const float h = H2[arg * a.c2 + k];
          if (h != 0.f) atomicAdd(a.dw3 + (size_t)k * a.c3 + c, h * gch);
          if (h > 0.f) atomicAdd(D2 + arg * a.c2 + k, __ldg(a.w3 + (size_t)k * a.c3 + c) * gch);   // * relu'(H2)
        }
      }
    }
    __syncthreads();
    // ---- layer 2
    weight_grad(H1, a.c1, D2, a.c2, a.dw2, a.db2);
    data_grad(D2, a.c2, a.w2t, a.c1, H1, D1);
    __syncthreads();
    // ---- layer 1 and the inputs
    weight_grad(X, K1, D1, a.c1, a.dw1, a.db1);
    for (int o = tid; o < K1 * (SB_R / SB_RB); o += SB_T) {   // dX = D1 W1^T, scattered (GroupPointGrad fused in)
      const int k = o % K1, r0 = (o / K1) * SB_RB;
      float acc[SB_RB];
#pragma unroll
      for (int i = 0; i < SB_RB; ++i) acc[i] = 0.f;
      for (int cc = 0; cc < a.c1; ++cc) {
        const float w = __ldg(a.w1t + (size_t)cc * K1 + k);
#pragma unroll
        for (int i = 0; i < SB_RB; ++i) acc[i] = fmaf(D1[(r0 + i) * a.c1 + cc], w, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < SB_RB; ++i) {
        if (acc[i] == 0.f) continue;
        const size_t src = (size_t)bi * a.n + s_idx[r0 + i];
        if (k < 3) {
          if (a.dxyz != nullptr) atomicAdd(a.dxyz + src * 3 + k, acc[i]);
          if (a.dnew_xyz != nullptr) atomicAdd(a.dnew_xyz + (size_t)g * 3 + k, -acc[i]);
        } else if (a.dfeat != nullptr) {
          atomicAdd(a.dfeat + src * a.c + (k - 3), acc[i]);
        }
      }
    }
  }
}

}  // namespace vnb

using namespace vnb;

extern "C" int vnb_sa_group_mlp_max_backward(int b, int n, int c, int m, int nsample, const float* xyz, const float* feat,
                                             const float* new_xyz, const int* idx, int c1, int c2, int c3, const float* w1,
                                             const float* b1, const float* w2, const float* b2, const float* w3,
                                             const float* b3, const float* w1_t, const float* w2_t, const float* grad_out,
                                             float* grad_feat, float* grad_xyz, float* grad_new_xyz, float* grad_w1,
                                             float* grad_b1, float* grad_w2, float* grad_b2, float* grad_w3, float* grad_b3,
                                             void* stream) {
  VNB_REQUIRE(nsample == SB_R, "sa_group_mlp_max_backward: nsample must be 64 (got %d)", nsample);
  VNB_REQUIRE(b >= 0 && n > 0 && c >= 0 && m >= 0 && c1 > 0 && c2 > 0 && c3 > 0, "sa_group_mlp_max_backward: bad shape");
  if (b == 0 || m == 0) return VNB_OK;
  const size_t smem = (size_t)SB_R * ((3 + c) + 2 * c1 + 2 * c2) * sizeof(float);
  VNB_REQUIRE(smem <= 220 * 1024, "sa_group_mlp_max_backward: 64 x (3 + c + 2 c1 + 2 c2) floats exceed shared memory");
  cudaStream_t st = as_stream(stream);
  VNB_CUDA(cudaFuncSetAttribute(sa_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SaBwdArgs a;
  a.n = n; a.c = c; a.m = m; a.total = b * m; a.c1 = c1; a.c2 = c2; a.c3 = c3;
  a.xyz = xyz; a.feat = feat; a.new_xyz = new_xyz; a.idx = idx;
  a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.w3 = w3; a.b3 = b3; a.w1t = w1_t; a.w2t = w2_t; a.dout = grad_out;
  a.dfeat = grad_feat; a.dxyz = grad_xyz; a.dnew_xyz = grad_new_xyz;
  a.dw1 = grad_w1; a.db1 = grad_b1; a.dw2 = grad_w2; a.db2 = grad_b2; a.dw3 = grad_w3; a.db3 = grad_b3;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = b * m < sms * 2 ? b * m : sms * 2;
  sa_backward_kernel<<<grid, SB_T, smem, st>>>(a);
  return check_launch("sa_group_mlp_max_backward");
}

// Tensor-core (tcgen05 + TMEM) implementations of the dense layers: fp16 operands, fp32 accumulation in TMEM.
//
//  * pack_weight_kernel   W (cin,cout) f32  ->  fp16 [n_pad][k_pad] K-major image, 128-byte swizzled, in panels of 64
//                         k-columns: byte-for-byte what tcgen05.mma reads from shared memory, so a weight tile is
//                         staged by ONE 1-D bulk copy (cp.async.bulk, the TMA engine) per panel — no tensor map.
//  * linear_tc_kernel     (linear_tc.cu) out = act(in @ W + b) (+res).
//  * sa_tc_kernel         pointnet_sa_module's group -> 3-layer shared MLP -> max-pool in ONE kernel (utils.py:49-55,
//                         120-132): a tile is 2 centroids x 64 samples = 128 rows.  Layer-1 input rows are produced by
//                         the CTA's threads straight from the irregular gather (never materialised in HBM); layer
//                         outputs hop TMEM -> registers (bias+ReLU, fp16) -> shared memory as the next layer's operand;
//                         the last layer is issued TRANSPOSED (D^T = W3^T . H2^T, channels on TMEM lanes, samples on
//                         columns) so the 64-sample max-pool is a register-only reduction per thread.
//    Layer 1 of the wide layers (C_in = 3+128 / 3+256) is hoisted through the gather:
//       W1^T [rel_xyz, feat[idx]] + b1 = (feat W1f + b1)[idx] + W1x^T (xyz[idx] - centroid)
//    q = feat W1f + b1 is computed once per SOURCE point (n rows instead of m*64 grouped rows) by linear_tc_kernel with
//    fp16 output; the kernel gathers q rows (fp16, half the bytes of the fp32 feature rows the reference gathers) and
//    adds the rank-3 relative-xyz term in fp32.  Exact in real arithmetic; only float summation order changes.
#include "common.cuh"
#include "umma.cuh"

namespace vnb {

using namespace umma;

int linear_simt(int rows, int cin, int cout, const float* in, const float* w, const float* bias, const float* res,
                int act, float* out_f32, void* out_f16, cudaStream_t st);
int sa_simt(int b, int n, int c, int m, const float* xyz, const float* feat, const float* new_xyz, const int* idx,
            int c1, int c2, int c3, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
            const float* b3, float* out, cudaStream_t st);
int sa_ws2_dispatch(int b, int n, int m, const float* xyz, const float* new_xyz, const int* idx, const int* pts_cnt, int c1,
                    int c2, int c3,
                    const float* w1x, const float* b2, const float* b3, const void* w2_img, const void* w3_img,
                    const void* q, float* out, void* workspace, cudaStream_t st);  // sa_ws2.cu
int sa1_ws2_dispatch(int b, int n, int c, int m, const float* xyz, const float* feat, const float* new_xyz, const int* idx,
                     const int* pts_cnt,
                     int c1, int c2, int c3, const float* b1, const float* b2, const float* b3, const void* w1_img,
                     const void* w2_img, const void* w3_img, float* out, void* workspace, cudaStream_t st);  // sa1_ws2.cu
int linear_tc(int rows, int cin, int cout, const float* in, const void* w_img, const float* bias, const float* res,
              int act, float* out_f32, void* out_f16, cudaStream_t st);  // linear_tc.cu
int g_sa_sms = 0;    // tuning: SMs the fused SA kernels size their grid for (0 = all)
int g_sa_split = 1;  // tuning: CTAs (chunks of tiles) per SM
int g_sa_min_tpc = 24;  // tuning: fewest tiles a CTA of the fused SA kernels takes (fewer CTAs when tiles are scarce)
int g_sa_variant = 2;  // 2: MMA issuers park on their mbarriers (default), 3: they poll (experiments);

// ---------------------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(int cin, int cout, int k_pad, int n_pad, const float* __restrict__ w,
                                   __half* __restrict__ image) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pad * k_pad) return;
  int r = t / k_pad, k = t % k_pad;
  float v = (r < cout && k < cin) ? w[(size_t)k * cout + r] : 0.f;
  size_t off = (size_t)(k >> 6) * n_pad * 128 + sw128_offset((uint32_t)r, (uint32_t)k);
  *reinterpret_cast<__half*>(reinterpret_cast<char*>(image) + off) = __float2half_rn(v);
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace vnb

using namespace vnb;

namespace vnb {
size_t sa_rel_bytes(long long rows);              // sa_pack.cu
size_t sa_tile_table_bytes(int total_centroids);  // sa_pack.cu
}  // namespace vnb

extern "C" size_t vnb_sa_workspace_bytes(int b, int m, int nsample) {
  if (b <= 0 || m <= 0 || nsample <= 0) return 256;
  return sa_rel_bytes((long long)b * m * nsample) + sa_tile_table_bytes(b * m);  // rel table | tile header | tile table
}

extern "C" size_t vnb_weight_image_bytes(int cin, int cout) {
  if (cin <= 0 || cout <= 0) return 0;
  size_t k_pad = (size_t)round_up(cin, 16), n_pad = (size_t)round_up(cout, 16);
  return ((k_pad + 63) / 64) * n_pad * 128;
}

extern "C" int vnb_pack_weight_f16(int cin, int cout, const float* w, void* image, void* stream) {
  VNB_REQUIRE(cin > 0 && cout > 0, "pack_weight: bad shape");
  cudaStream_t st = as_stream(stream);
  VNB_CUDA(cudaMemsetAsync(image, 0, vnb_weight_image_bytes(cin, cout), st));
  int k_pad = round_up(cin, 16), n_pad = round_up(cout, 16);
  int total = k_pad * n_pad;
  pack_weight_kernel<<<(total + 255) / 256, 256, 0, st>>>(cin, cout, k_pad, n_pad, w, static_cast<__half*>(image));
  return check_launch("pack_weight");
}

extern "C" int vnb_linear(int rows, int cin, int cout, const float* in, const float* w_f32, const void* w_img,
                          const float* bias, const float* residual, int act, float* out_f32, void* out_f16,
                          int precision, void* stream) {
  VNB_REQUIRE(rows >= 0 && cin > 0 && cout > 0, "linear: bad shape");
  VNB_REQUIRE(out_f32 != nullptr || out_f16 != nullptr, "linear: no output buffer");
  VNB_REQUIRE(act == VNB_ACT_NONE || act == VNB_ACT_RELU, "linear: unknown activation");
  if (rows == 0) return VNB_OK;
  cudaStream_t st = as_stream(stream);
  if (precision == 0) {
    VNB_REQUIRE(w_f32 != nullptr, "linear(fp32): w_f32 missing");
    return linear_simt(rows, cin, cout, in, w_f32, bias, residual, act, out_f32, out_f16, st);
  }
  VNB_REQUIRE(precision == 1, "linear: precision must be 0 (fp32) or 1 (tensor cores)");
  VNB_REQUIRE(w_img != nullptr, "linear(tensor cores): packed weight image missing");
  return linear_tc(rows, cin, cout, in, w_img, bias, residual, act, out_f32, out_f16, st);
}

extern "C" int vnb_sa_group_mlp_max_counted(int b, int n, int c, int m, int nsample, const float* xyz, const float* feat,
                                            const float* new_xyz, const int* idx, const int* pts_cnt, int c1, int c2, int c3,
                                            const float* w1_f32, const float* b1, const float* w2_f32, const float* b2,
                                            const float* w3_f32, const float* b3, const void* w1_img, const void* w2_img,
                                            const void* w3_img, const void* q_f16, float* out, int precision,
                                            void* workspace, void* stream);

extern "C" int vnb_sa_group_mlp_max(int b, int n, int c, int m, int nsample, const float* xyz, const float* feat,
                                    const float* new_xyz, const int* idx, int c1, int c2, int c3, const float* w1_f32,
                                    const float* b1, const float* w2_f32, const float* b2, const float* w3_f32,
                                    const float* b3, const void* w1_img, const void* w2_img, const void* w3_img,
                                    const void* q_f16, float* out, int precision, void* workspace, void* stream) {
  return vnb_sa_group_mlp_max_counted(b, n, c, m, nsample, xyz, feat, new_xyz, idx, nullptr, c1, c2, c3, w1_f32, b1, w2_f32,
                                      b2, w3_f32, b3, w1_img, w2_img, w3_img, q_f16, out, precision, workspace, stream);
}

extern "C" int vnb_sa_group_mlp_max_counted(int b, int n, int c, int m, int nsample, const float* xyz, const float* feat,
                                            const float* new_xyz, const int* idx, const int* pts_cnt, int c1, int c2, int c3,
                                            const float* w1_f32, const float* b1, const float* w2_f32, const float* b2,
                                            const float* w3_f32, const float* b3, const void* w1_img, const void* w2_img,
                                            const void* w3_img, const void* q_f16, float* out, int precision,
                                            void* workspace, void* stream) {
  VNB_REQUIRE(nsample == 64, "sa_group_mlp_max: nsample must be 64 (got %d)", nsample);
  VNB_REQUIRE(b >= 0 && n > 0 && c >= 0 && m >= 0 && c1 > 0 && c2 > 0 && c3 > 0, "sa_group_mlp_max: bad shape");
  if (b == 0 || m == 0) return VNB_OK;
  cudaStream_t st = as_stream(stream);
  if (precision == 0) {
    VNB_REQUIRE(w1_f32 && w2_f32 && w3_f32, "sa_group_mlp_max(fp32): f32 weights missing");
    return sa_simt(b, n, c, m, xyz, feat, new_xyz, idx, c1, c2, c3, w1_f32, b1, w2_f32, b2, w3_f32, b3, out, st);
  }
  VNB_REQUIRE(precision == 1, "sa_group_mlp_max: precision must be 0 (fp32) or 1 (tensor cores)");
  VNB_REQUIRE((b * m) % 2 == 0, "sa_group_mlp_max(tensor cores): b*m must be even");
  VNB_REQUIRE(w2_img && w3_img, "sa_group_mlp_max(tensor cores): packed weight images missing");
  const bool hoist = c > 13;
  if (!hoist) {
    VNB_REQUIRE(w1_img != nullptr, "sa_group_mlp_max(tensor cores): w1_img missing");
    {  // narrow-input pipeline (sa1_ws2.cu)
      int rc = sa1_ws2_dispatch(b, n, c, m, xyz, feat, new_xyz, idx, pts_cnt, c1, c2, c3, b1, b2, b3, w1_img, w2_img, w3_img, out,
                                workspace, st);
      if (rc >= 0) return rc;
    }
  } else {
    VNB_REQUIRE(q_f16 != nullptr && w1_f32 != nullptr,
                "sa_group_mlp_max(tensor cores): hoisted layer 1 needs q_f16 and w1_f32 (rows 0..2)");
    {  // hoisted-layer-1 pipeline (sa_ws2.cu)
      int rc = sa_ws2_dispatch(b, n, m, xyz, new_xyz, idx, pts_cnt, c1, c2, c3, w1_f32, b2, b3, w2_img, w3_img, q_f16, out, workspace, st);
      if (rc >= 0) return rc;
    }
  }
  return set_err(VNB_ERR_INVALID,
                 "sa_group_mlp_max(tensor cores): no kernel instance for c=%d mlp=(%d,%d,%d)%s; use precision=0", c, c1,
                 c2, c3, workspace == nullptr ? " without a workspace (vnb_sa_workspace_bytes)" : "");
}

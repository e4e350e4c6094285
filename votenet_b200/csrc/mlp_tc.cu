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
int sa_ws_dispatch(int b, int n, int m, const float* xyz, const float* new_xyz, const int* idx, int c1, int c2, int c3,
                   const float* w1x, const float* b2, const float* b3, const void* w2_img, const void* w3_img,
                   const void* q, float* out, void* workspace, cudaStream_t st);
int sa1_ws_dispatch(int b, int n, int c, int m, const float* xyz, const float* feat, const float* new_xyz, const int* idx,
                    int c1, int c2, int c3, const float* b1, const float* b2, const float* b3, const void* w1_img,
                    const void* w2_img, const void* w3_img, float* out, void* workspace, cudaStream_t st);
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
int g_sa_variant = 2;  // 0: single-role kernel (sa_tc_kernel), 1: warp-specialised pipeline, first generation,
                       // 2: second generation (one MMA issuer per layer, one wave) where available

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

// ---------------------------------------------------------------------------------------------------------------------
// fused SA layer.  128 threads; persistent over tiles (tile = 2 centroids x 64 samples); weights stay resident.
//   HOIST = false : layer 1 on tensor cores from raw [rel_xyz, feat] rows (K padded to 16)          (sa1)
//   HOIST = true  : layer 1 = relu(q[idx] + W1x^T rel_xyz) in the producer                          (sa2-4, proposal)
template <int C1, int C2, int C3, bool HOIST>
struct SaCfg {
  static constexpr int P1 = C1 / 64, P2 = C2 / 64;               // panels of h1 / h2
  static constexpr int W1_BYTES = HOIST ? 0 : C1 * 128;          // [C1][16] in one panel
  static constexpr int W2_BYTES = P1 * C2 * 128;                 // [C2][C1]
  static constexpr int W3_BYTES = P2 * C3 * 128;                 // [C3][C2]
  static constexpr int A0_BYTES = HOIST ? 0 : 128 * 128;         // raw input rows, one panel
  static constexpr int H1_BYTES = P1 * 128 * 128;
  static constexpr int H2_BYTES = P2 * 128 * 128;
  static constexpr int OFF_W1 = 0;
  static constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
  static constexpr int OFF_W3 = OFF_W2 + W2_BYTES;
  static constexpr int OFF_A0 = OFF_W3 + W3_BYTES;
  static constexpr int OFF_H1 = OFF_A0 + A0_BYTES;
  static constexpr int OFF_H2 = OFF_H1 + H1_BYTES;
  static constexpr int OFF_F = OFF_H2 + H2_BYTES;                // floats: b1[C1] | b2[C2] | b3[C3] | wx[3][C1]
  static constexpr int NFLOAT = C1 + C2 + C3 + 3 * C1;
  static constexpr int OFF_BAR = OFF_F + NFLOAT * 4;
  static constexpr int SMEM = OFF_BAR + 64 + 1024;
  static constexpr int TM_D1 = 0;                                // TMEM columns
  static constexpr int TM_D2 = HOIST ? 0 : C1;
  static constexpr int TM_D3 = TM_D2 + C2;
  static constexpr int TM_USED = TM_D3 + C3;
  static constexpr int TM_COLS = TM_USED <= 32 ? 32 : TM_USED <= 64 ? 64 : TM_USED <= 128 ? 128 : TM_USED <= 256 ? 256 : 512;
  static_assert(TM_USED <= 512, "TMEM budget");
  static_assert(C1 % 64 == 0 && C2 % 64 == 0 && C3 % 128 == 0, "channel widths");
};

template <int C1, int C2, int C3, bool HOIST>
__global__ void __launch_bounds__(128) sa_tc_kernel(int n, int c, int m, int total_centroids, int tiles_per_cta,
                                                    const float* __restrict__ xyz, const float* __restrict__ feat,
                                                    const float* __restrict__ new_xyz, const int* __restrict__ idx,
                                                    const float* __restrict__ w1x /* (3,C1) f32: W1 rows 0..2 */,
                                                    const float* __restrict__ b1, const float* __restrict__ b2,
                                                    const float* __restrict__ b3, const char* __restrict__ w1_img,
                                                    const char* __restrict__ w2_img, const char* __restrict__ w3_img,
                                                    const __half* __restrict__ q, float* __restrict__ out) {
  using Cfg = SaCfg<C1, C2, C3, HOIST>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* sW1 = smem + Cfg::OFF_W1;
  uint8_t* sW2 = smem + Cfg::OFF_W2;
  uint8_t* sW3 = smem + Cfg::OFF_W3;
  uint8_t* sA0 = smem + Cfg::OFF_A0;
  uint8_t* sH1 = smem + Cfg::OFF_H1;
  uint8_t* sH2 = smem + Cfg::OFF_H2;
  float* sB1 = reinterpret_cast<float*>(smem + Cfg::OFF_F);
  float* sB2 = sB1 + C1;
  float* sB3 = sB2 + C2;
  float* sWx = sB3 + C3;  // [3][C1]
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_w + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    // resident weights: one bulk copy per image
    mbar_arrive_expect_tx(bar_w, (uint32_t)(Cfg::W1_BYTES + Cfg::W2_BYTES + Cfg::W3_BYTES));
    if (!HOIST) bulk_g2s(sW1, w1_img, Cfg::W1_BYTES, bar_w);
    bulk_g2s(sW2, w2_img, Cfg::W2_BYTES, bar_w);
    bulk_g2s(sW3, w3_img, Cfg::W3_BYTES, bar_w);
  }
  for (int i = tid; i < C1; i += 128) sB1[i] = b1 ? b1[i] : 0.f;
  for (int i = tid; i < C2; i += 128) sB2[i] = b2[i];
  for (int i = tid; i < C3; i += 128) sB3[i] = b3[i];
  if (HOIST)
    for (int i = tid; i < 3 * C1; i += 128) sWx[i] = w1x[i];
  if (!HOIST) {  // zero the raw-input panel once: only k < 16 is ever rewritten (padding columns must be finite)
    for (int i = tid; i < Cfg::A0_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sA0)[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc(tmem_ptr, Cfg::TM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  mbar_wait(bar_w, 0);

  uint32_t mma_phase = 0;
  const int ntiles = total_centroids / 2;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  const int tile_end = min(ntiles, ((int)blockIdx.x + 1) * tiles_per_cta);
  for (int tile = (int)blockIdx.x * tiles_per_cta; tile < tile_end; ++tile) {
    // ------------------------------------------------------------ producer: this thread's grouped row
    {
      const int g = tile * 2 + (tid >> 6);  // global centroid
      const int bi = g / m;
      const int pid = idx[(size_t)g * 64 + (tid & 63)];
      const float* pp = xyz + ((size_t)bi * n + pid) * 3;
      const float* cc = new_xyz + (size_t)g * 3;
      const float rx = pp[0] - cc[0], ry = pp[1] - cc[1], rz = pp[2] - cc[2];  // utils.py:51
      if (!HOIST) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        v[0] = rx; v[1] = ry; v[2] = rz;
        const float* f = feat + ((size_t)bi * n + pid) * c;
        for (int i = 0; i < c; ++i) v[3 + i] = f[i];  // c <= 13
        uint4 lo = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        uint4 hi = make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
        *reinterpret_cast<uint4*>(sA0 + sw128_offset((uint32_t)tid, 0)) = lo;
        *reinterpret_cast<uint4*>(sA0 + sw128_offset((uint32_t)tid, 8)) = hi;
      } else {
        const uint4* qr = reinterpret_cast<const uint4*>(q + ((size_t)bi * n + pid) * C1);
#pragma unroll 4
        for (int ch = 0; ch < C1 / 8; ++ch) {
          uint4 raw = qr[ch];
          const __half2* h = reinterpret_cast<const __half2*>(&raw);
          float o[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 f = __half22float2(h[i]);
            o[2 * i] = f.x; o[2 * i + 1] = f.y;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int ch_i = ch * 8 + i;
            float a = fmaf(sWx[ch_i], rx, o[i]);
            a = fmaf(sWx[C1 + ch_i], ry, a);
            a = fmaf(sWx[2 * C1 + ch_i], rz, a);
            o[i] = fmaxf(a, 0.f);
          }
          uint4 pk = make_uint4(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]), pack_h2(o[4], o[5]), pack_h2(o[6], o[7]));
          const uint32_t kk = (uint32_t)ch * 8;
          *reinterpret_cast<uint4*>(sH1 + (kk >> 6) * (128 * 128) + sw128_offset((uint32_t)tid, kk)) = pk;
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    // ------------------------------------------------------------ layer 1 on tensor cores (sa1 only)
    if (!HOIST) {
      if (tid == 0) {
        mma_f16_ss(tmem + Cfg::TM_D1, make_desc_sw128(smem_u32(sA0)), make_desc_sw128(smem_u32(sW1)),
                   make_idesc_f16_f32(128, C1), 0u);
        mma_commit(bar_mma);
      }
      mbar_wait(bar_mma, mma_phase); mma_phase ^= 1;
      tc_fence_after_sync();
      for (int cb = 0; cb < C1; cb += 32) {
        uint32_t v[32];
        tmem_ld_x32(tmem + lane_base + Cfg::TM_D1 + cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = fmaxf(__uint_as_float(v[ch * 8 + i]) + sB1[cb + ch * 8 + i], 0.f);
          uint4 pk = make_uint4(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]), pack_h2(o[4], o[5]), pack_h2(o[6], o[7]));
          const uint32_t kk = (uint32_t)(cb + ch * 8);
          *reinterpret_cast<uint4*>(sH1 + (kk >> 6) * (128 * 128) + sw128_offset((uint32_t)tid, kk)) = pk;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      tc_fence_after_sync();
    }

    // ------------------------------------------------------------ layer 2: D2[sample][c2] = H1 . W2^T
    if (tid == 0) {
      const uint32_t a0 = smem_u32(sH1), b0 = smem_u32(sW2);
      const uint32_t idesc = make_idesc_f16_f32(128, C2);
#pragma unroll
      for (int ks = 0; ks < C1 / 16; ++ks) {
        const uint32_t pan = (uint32_t)ks >> 2, kin = (uint32_t)ks & 3;
        mma_f16_ss(tmem + Cfg::TM_D2, make_desc_sw128(a0 + pan * (128 * 128) + kin * 32),
                   make_desc_sw128(b0 + pan * (C2 * 128) + kin * 32), idesc, ks > 0 ? 1u : 0u);
      }
      mma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase); mma_phase ^= 1;
    tc_fence_after_sync();
    for (int cb = 0; cb < C2; cb += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem + lane_base + Cfg::TM_D2 + cb, v);
      tmem_ld_wait();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaxf(__uint_as_float(v[ch * 8 + i]) + sB2[cb + ch * 8 + i], 0.f);
        uint4 pk = make_uint4(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]), pack_h2(o[4], o[5]), pack_h2(o[6], o[7]));
        const uint32_t kk = (uint32_t)(cb + ch * 8);
        *reinterpret_cast<uint4*>(sH2 + (kk >> 6) * (128 * 128) + sw128_offset((uint32_t)tid, kk)) = pk;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    // ------------------------------------------------------------ layer 3, transposed: D3[c3][sample] = W3^T . H2^T
    if (tid == 0) {
      const uint32_t a0 = smem_u32(sW3), b0 = smem_u32(sH2);
      const uint32_t idesc = make_idesc_f16_f32(128, 128);
#pragma unroll
      for (int hh = 0; hh < C3 / 128; ++hh) {
#pragma unroll
        for (int ks = 0; ks < C2 / 16; ++ks) {
          const uint32_t pan = (uint32_t)ks >> 2, kin = (uint32_t)ks & 3;
          mma_f16_ss(tmem + Cfg::TM_D3 + hh * 128,
                     make_desc_sw128(a0 + pan * (C3 * 128) + hh * (128 * 128) + kin * 32),
                     make_desc_sw128(b0 + pan * (128 * 128) + kin * 32), idesc, ks > 0 ? 1u : 0u);
        }
      }
      mma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase); mma_phase ^= 1;
    tc_fence_after_sync();
    // max-pool over the 64 samples of each centroid: pure register reduction (channel = TMEM lane = this thread)
#pragma unroll
    for (int hh = 0; hh < C3 / 128; ++hh) {
      const int ch = hh * 128 + tid;
      float mx[2];
#pragma unroll
      for (int gq = 0; gq < 2; ++gq) {
        float mval = -INFINITY;
#pragma unroll
        for (int cb = 0; cb < 64; cb += 32) {
          uint32_t v[32];
          tmem_ld_x32(tmem + lane_base + Cfg::TM_D3 + hh * 128 + gq * 64 + cb, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mval = fmaxf(mval, __uint_as_float(v[i]));
        }
        mx[gq] = mval;
      }
      // bias + ReLU commute with the max (both monotone)
      const float bb = sB3[ch];
      out[((size_t)tile * 2 + 0) * C3 + ch] = fmaxf(mx[0] + bb, 0.f);
      out[((size_t)tile * 2 + 1) * C3 + ch] = fmaxf(mx[1] + bb, 0.f);
    }
    tc_fence_before_sync();
    __syncthreads();  // TMEM and H1/H2 are reused by the next tile
    tc_fence_after_sync();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TM_COLS);
}

template <int C1, int C2, int C3, bool HOIST>
static int launch_sa_tc(int b, int n, int c, int m, const float* xyz, const float* feat, const float* new_xyz,
                        const int* idx, const float* w1x, const float* b1, const float* b2, const float* b3,
                        const void* w1_img, const void* w2_img, const void* w3_img, const void* q, float* out,
                        cudaStream_t st) {
  using Cfg = SaCfg<C1, C2, C3, HOIST>;
  auto kern = sa_tc_kernel<C1, C2, C3, HOIST>;
  VNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  int dev = 0, sms = 148;
  VNB_CUDA(cudaGetDevice(&dev));
  VNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ntiles = b * m / 2;
  // CTAs per SM limited by shared memory and by TMEM columns (512 per SM)
  int per_sm = (227 * 1024) / Cfg::SMEM;
  if (per_sm > 512 / Cfg::TM_COLS) per_sm = 512 / Cfg::TM_COLS;
  if (per_sm < 1) per_sm = 1;
  int tpc = ntiles / (2 * sms * per_sm);
  tpc = tpc < 2 ? 2 : (tpc > 16 ? 16 : tpc);
  const int grid = (ntiles + tpc - 1) / tpc;
  kern<<<grid, 128, Cfg::SMEM, st>>>(n, c, m, b * m, tpc, xyz, feat, new_xyz, idx, w1x, b1, b2, b3,
                                     static_cast<const char*>(w1_img), static_cast<const char*>(w2_img),
                                     static_cast<const char*>(w3_img), static_cast<const __half*>(q), out);
  return check_launch("sa_group_mlp_max (tcgen05)");
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
    if (g_sa_variant >= 2) {  // second-generation pipeline (sa1_ws2.cu)
      int rc = sa1_ws2_dispatch(b, n, c, m, xyz, feat, new_xyz, idx, pts_cnt, c1, c2, c3, b1, b2, b3, w1_img, w2_img, w3_img, out,
                                workspace, st);
      if (rc >= 0) return rc;
    }
    if (g_sa_variant >= 1) {  // warp-specialised, pipelined kernel (sa1_ws.cu)
      int rc = sa1_ws_dispatch(b, n, c, m, xyz, feat, new_xyz, idx, c1, c2, c3, b1, b2, b3, w1_img, w2_img, w3_img, out,
                               workspace, st);
      if (rc >= 0) return rc;
    }
    if (c1 == 64 && c2 == 64 && c3 == 128)
      return launch_sa_tc<64, 64, 128, false>(b, n, c, m, xyz, feat, new_xyz, idx, nullptr, b1, b2, b3, w1_img, w2_img,
                                              w3_img, nullptr, out, st);
  } else {
    VNB_REQUIRE(q_f16 != nullptr && w1_f32 != nullptr,
                "sa_group_mlp_max(tensor cores): hoisted layer 1 needs q_f16 and w1_f32 (rows 0..2)");
    if (g_sa_variant >= 2) {  // second-generation pipeline (sa_ws2.cu)
      int rc = sa_ws2_dispatch(b, n, m, xyz, new_xyz, idx, pts_cnt, c1, c2, c3, w1_f32, b2, b3, w2_img, w3_img, q_f16, out, workspace, st);
      if (rc >= 0) return rc;
    }
    if (g_sa_variant >= 1) {  // warp-specialised, pipelined kernel (sa_ws.cu)
      int rc = sa_ws_dispatch(b, n, m, xyz, new_xyz, idx, c1, c2, c3, w1_f32, b2, b3, w2_img, w3_img, q_f16, out, workspace, st);
      if (rc >= 0) return rc;
    }
    if (c1 == 128 && c2 == 128 && c3 == 256)
      return launch_sa_tc<128, 128, 256, true>(b, n, c, m, xyz, feat, new_xyz, idx, w1_f32, nullptr, b2, b3, nullptr,
                                               w2_img, w3_img, q_f16, out, st);
    if (c1 == 128 && c2 == 128 && c3 == 128)
      return launch_sa_tc<128, 128, 128, true>(b, n, c, m, xyz, feat, new_xyz, idx, w1_f32, nullptr, b2, b3, nullptr,
                                               w2_img, w3_img, q_f16, out, st);
  }
  return set_err(VNB_ERR_INVALID,
                 "sa_group_mlp_max(tensor cores): no kernel instance for c=%d mlp=(%d,%d,%d); use precision=0", c, c1,
                 c2, c3);
}

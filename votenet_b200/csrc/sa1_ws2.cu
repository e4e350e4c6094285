// Fused set-abstraction kernel for NARROW inputs (sa1: 3 + c <= 16 input channels), second generation:
// all three layers on the tensor cores, every stage double-buffered, every stage with its OWN issuing warp.
//
//   group -> [rel_xyz, feat] (K padded to 16) -> L1 -> L2 -> L3 -> max over the 64 samples        (utils.py:49-55,120-132)
//
//   warps 0-3    EPILOGUE1  D1 -> +b1, ReLU, fp16 -> H1[t%2]               (thread = tile row = TMEM lane)
//   warps 4-7    EPILOGUE2  D2 -> +b2, ReLU, fp16 -> H2[t%2]
//   warps 8-15   EPILOGUE3  D3 (channel per lane) -> max over a centroid's 64 samples, +b3, ReLU -> out;
//                           two warps per TMEM lane quadrant, one per centroid of the tile (a warp can pull only 64 B per
//                           cycle out of TMEM, measured: scripts/micro/tmem.cu)
//   warps 16-19  PRODUCER   one grouped row per thread: {rel_xyz, source row} from the helper kernel's table, c feature
//                           floats; table rows prefetched two tiles ahead and features one tile ahead so no global
//                           latency is exposed; fp16, swizzled 32-byte row of A0[t%2]
//   warp  20/21/22  MMA1/2/3   one thread each: M1(t): D1 = A0 . W1^T, M2(t): D2 = H1 . W2^T, M3(t): D3 = W3^T . H2^T
//                           (transposed: channels on TMEM lanes, so the 64-sample max-pool is a register reduction).
//                           Separate issuers: a layer never waits behind another layer's operands (the first generation's
//                           single in-order issuer kept only ~2 tiles in flight over 7 stages).
// TMEM: D1[2] (2 x 64) | D2[2] (2 x 64) | D3[2] (2 x 128) columns.  All hand-offs are mbarriers.  One CTA per SM, one wave:
// CTA i owns a contiguous chunk of tiles.
#include "common.cuh"
#include "umma.cuh"

namespace vnb {

using namespace umma;

namespace s1v2 {

constexpr int C1 = 64, C2 = 64, C3 = 128;
constexpr int W1_BYTES = C1 * 128;         // [C1][16] in one panel
constexpr int W2_BYTES = C2 * 128;         // [C2][C1=64]
constexpr int W3_BYTES = C3 * 128;         // [C3][C2=64]
constexpr int A0_BYTES = 128 * 128;        // per buffer
constexpr int H_BYTES = 128 * 128;         // H1 / H2 per buffer (64 columns = one panel)
constexpr int OFF_W1 = 0;
constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
constexpr int OFF_W3 = OFF_W2 + W2_BYTES;
constexpr int OFF_A0 = OFF_W3 + W3_BYTES;
constexpr int OFF_H1 = OFF_A0 + 2 * A0_BYTES;
constexpr int OFF_H2 = OFF_H1 + 2 * H_BYTES;
constexpr int OFF_F = OFF_H2 + 2 * H_BYTES;          // floats b1 | b2 | b3
constexpr int OFF_BAR = OFF_F + (C1 + C2 + C3) * 4;
constexpr int SMEM = OFF_BAR + 32 * 8 + 16 + 1024;
constexpr int TM_D1 = 0, TM_D2 = 2 * C1, TM_D3 = 2 * C1 + 2 * C2;
constexpr int TM_COLS = 512;
static_assert(TM_D3 + 2 * C3 <= 512, "TMEM budget");
constexpr int THREADS = 23 * 32;

__device__ __forceinline__ uint32_t pack_relu(float2 v) {  // fp16x2(max(v, 0)): rounding is monotone, so relu commutes
  __half2 h = __hmax2(__float22half2_rn(v), __float2half2_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(THREADS, 1) sa1_ws2_kernel(int c, int total_centroids, int tiles_per_cta,
                                                             const float4* __restrict__ rel,
                                                             const float* __restrict__ feat,
                                                             const float* __restrict__ b1, const float* __restrict__ b2,
                                                             const float* __restrict__ b3, const char* __restrict__ w1_img,
                                                             const char* __restrict__ w2_img,
                                                             const char* __restrict__ w3_img, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW1 = smem + OFF_W1;
  uint8_t* sW2 = smem + OFF_W2;
  uint8_t* sW3 = smem + OFF_W3;
  uint8_t* sA0 = smem + OFF_A0;
  uint8_t* sH1 = smem + OFF_H1;
  uint8_t* sH2 = smem + OFF_H2;
  float* sB1 = reinterpret_cast<float*>(smem + OFF_F);
  float* sB2 = sB1 + C1;
  float* sB3 = sB2 + C2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_w = bars;          // weights landed
  uint64_t* a0_full = bars + 1;    // [2] 128 producer arrivals
  uint64_t* m1_done = bars + 3;    // [2] commit
  uint64_t* h1_full = bars + 5;    // [2] 128 E1 arrivals
  uint64_t* d1_empty = bars + 7;   // [2] 128 E1 arrivals
  uint64_t* m2_done = bars + 9;    // [2] commit
  uint64_t* h2_full = bars + 11;   // [2] 128 E2 arrivals
  uint64_t* d2_empty = bars + 13;  // [2] 128 E2 arrivals
  uint64_t* m3_done = bars + 15;   // [2] commit
  uint64_t* d3_empty = bars + 17;  // [2] 256 E3 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a0_full[s], 128); mbar_init(&m1_done[s], 1); mbar_init(&h1_full[s], 128); mbar_init(&d1_empty[s], 128);
      mbar_init(&m2_done[s], 1); mbar_init(&h2_full[s], 128); mbar_init(&d2_empty[s], 128);
      mbar_init(&m3_done[s], 1); mbar_init(&d3_empty[s], 256);
    }
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_w, (uint32_t)(W1_BYTES + W2_BYTES + W3_BYTES));
    bulk_g2s(sW1, w1_img, W1_BYTES, bar_w);
    bulk_g2s(sW2, w2_img, W2_BYTES, bar_w);
    bulk_g2s(sW3, w3_img, W3_BYTES, bar_w);
  }
  for (int i = tid; i < C1; i += THREADS) sB1[i] = b1[i];
  for (int i = tid; i < C2; i += THREADS) sB2[i] = b2[i];
  for (int i = tid; i < C3; i += THREADS) sB3[i] = b3[i];
  // A0 padding columns (k >= 16 of the 64-column panel are never read; k in [3+c,16) must be finite): zero both buffers
  for (int i = tid; i < 2 * A0_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sA0)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(tmem_ptr, TM_COLS);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  const int ntiles = total_centroids / 2;
  const int first_tile = (int)blockIdx.x * tiles_per_cta;
  const int my_tiles = max(0, min(tiles_per_cta, ntiles - first_tile));
  // use u = t>>1 of a [2]-slotted barrier -> parity u&1
  auto par_of = [](int t) { return (uint32_t)((t >> 1) & 1); };

  if (warp >= 16 && warp < 20) {
    // ================================================================ PRODUCER: one grouped row per thread
    const int pt = tid - 512;
    const float4* rp = rel + (size_t)first_tile * 128 + pt;
    float4 r_cur = make_float4(0.f, 0.f, 0.f, 0.f), r_nxt = r_cur;   // table rows of tiles t, t+1
    if (my_tiles > 0) r_cur = __ldg(rp);
    if (my_tiles > 1) r_nxt = __ldg(rp + 128);
    float f_cur[13], f_nxt[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) f_cur[i] = f_nxt[i] = 0.f;
    if (my_tiles > 0) {
      const float* f = feat + (size_t)__float_as_int(r_cur.w) * c;
#pragma unroll
      for (int i = 0; i < 13; ++i)
        if (i < c) f_cur[i] = __ldg(f + i);
    }
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t & 1;
      // prefetch: features of tile t+1 (its table row arrived an iteration ago), table row of tile t+2
      float4 r_nn = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t + 1 < my_tiles) {
        const float* f = feat + (size_t)__float_as_int(r_nxt.w) * c;
#pragma unroll
        for (int i = 0; i < 13; ++i)
          if (i < c) f_nxt[i] = __ldg(f + i);
      }
      if (t + 2 < my_tiles) r_nn = __ldg(rp + (size_t)(t + 2) * 128);
      if (t >= 2) mbar_wait(&m1_done[s], par_of(t - 2));  // M1(t-2) finished reading A0[s]
      uint8_t* a0 = sA0 + s * A0_BYTES;
      *reinterpret_cast<uint4*>(a0 + sw128_offset((uint32_t)pt, 0)) =
          make_uint4(pack2(r_cur.x, r_cur.y), pack2(r_cur.z, f_cur[0]), pack2(f_cur[1], f_cur[2]), pack2(f_cur[3], f_cur[4]));
      *reinterpret_cast<uint4*>(a0 + sw128_offset((uint32_t)pt, 8)) =
          make_uint4(pack2(f_cur[5], f_cur[6]), pack2(f_cur[7], f_cur[8]), pack2(f_cur[9], f_cur[10]),
                     pack2(f_cur[11], f_cur[12]));
      fence_proxy_async_smem();
      mbar_arrive(&a0_full[s]);
      r_cur = r_nxt; r_nxt = r_nn;
#pragma unroll
      for (int i = 0; i < 13; ++i) f_cur[i] = f_nxt[i];
    }
  } else if (warp == 20) {
    // ================================================================ MMA1: D1[s] = A0[s] . W1^T  (single K = 16 step)
    if (lane == 0 && my_tiles > 0) {
      mbar_wait(bar_w, 0);
      const uint32_t id1 = make_idesc_f16_f32(128, C1);
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t & 1;
        mbar_wait(&a0_full[s], par_of(t));
        if (t >= 2) mbar_wait(&d1_empty[s], par_of(t - 2));
        tc_fence_after_sync();
        mma_f16_ss(tmem + TM_D1 + s * C1, make_desc_sw128(smem_u32(sA0 + s * A0_BYTES)), make_desc_sw128(smem_u32(sW1)),
                   id1, 0u);
        mma_commit(&m1_done[s]);
      }
    }
  } else if (warp == 21) {
    // ================================================================ MMA2: D2[s] = H1[s] . W2^T
    if (lane == 0 && my_tiles > 0) {
      mbar_wait(bar_w, 0);
      const uint32_t id2 = make_idesc_f16_f32(128, C2);
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t & 1;
        mbar_wait(&h1_full[s], par_of(t));
        if (t >= 2) mbar_wait(&d2_empty[s], par_of(t - 2));
        tc_fence_after_sync();
        const uint32_t a0 = smem_u32(sH1 + s * H_BYTES), b0 = smem_u32(sW2);
#pragma unroll
        for (int ks = 0; ks < C1 / 16; ++ks)
          mma_f16_ss(tmem + TM_D2 + s * C2, make_desc_sw128(a0 + ks * 32), make_desc_sw128(b0 + ks * 32), id2,
                     ks > 0 ? 1u : 0u);
        mma_commit(&m2_done[s]);
      }
    }
  } else if (warp == 22) {
    // ================================================================ MMA3: D3[s] = W3^T . H2[s]^T
    if (lane == 0 && my_tiles > 0) {
      mbar_wait(bar_w, 0);
      const uint32_t id3 = make_idesc_f16_f32(128, 128);
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t & 1;
        mbar_wait(&h2_full[s], par_of(t));
        if (t >= 2) mbar_wait(&d3_empty[s], par_of(t - 2));
        tc_fence_after_sync();
        const uint32_t a0 = smem_u32(sW3), b0 = smem_u32(sH2 + s * H_BYTES);
#pragma unroll
        for (int ks = 0; ks < C2 / 16; ++ks)
          mma_f16_ss(tmem + TM_D3 + s * C3, make_desc_sw128(a0 + ks * 32), make_desc_sw128(b0 + ks * 32), id3,
                     ks > 0 ? 1u : 0u);
        mma_commit(&m3_done[s]);
      }
    }
  } else if (warp < 8) {
    // ================================================================ EPILOGUE 1 (warps 0-3) / EPILOGUE 2 (warps 4-7)
    const bool e2 = warp >= 4;
    const int et = tid & 127;  // tile row == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint64_t* in_done = e2 ? m2_done : m1_done;      // accumulator ready
    uint64_t* out_free = e2 ? m3_done : m2_done;     // MMA that read our output buffer two tiles ago
    uint64_t* d_empty = e2 ? d2_empty : d1_empty;
    uint64_t* h_full = e2 ? h2_full : h1_full;
    const uint32_t tm = e2 ? TM_D2 : TM_D1;
    uint8_t* hbase = e2 ? sH2 : sH1;
    const float* bias = e2 ? sB2 : sB1;
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t & 1;
      mbar_wait(&in_done[s], par_of(t));
      tc_fence_after_sync();
      uint32_t v[2][32];
      tmem_ld_x32(tmem + lane_base + tm + s * 64, v[0]);       // both halves of the 64-column accumulator in flight
      tmem_ld_x32(tmem + lane_base + tm + s * 64 + 32, v[1]);
      if (t >= 2) mbar_wait(&out_free[s], par_of(t - 2));
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&d_empty[s]);                                 // accumulator drained into registers
      uint8_t* h = hbase + s * H_BYTES;
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 ba = *reinterpret_cast<const float4*>(bias + hb * 32 + ch * 8);
          const float4 bb = *reinterpret_cast<const float4*>(bias + hb * 32 + ch * 8 + 4);
          const uint32_t* vv = &v[hb][ch * 8];
          const uint4 pk = make_uint4(
              pack_relu(__fadd2_rn(make_float2(__uint_as_float(vv[0]), __uint_as_float(vv[1])), make_float2(ba.x, ba.y))),
              pack_relu(__fadd2_rn(make_float2(__uint_as_float(vv[2]), __uint_as_float(vv[3])), make_float2(ba.z, ba.w))),
              pack_relu(__fadd2_rn(make_float2(__uint_as_float(vv[4]), __uint_as_float(vv[5])), make_float2(bb.x, bb.y))),
              pack_relu(__fadd2_rn(make_float2(__uint_as_float(vv[6]), __uint_as_float(vv[7])), make_float2(bb.z, bb.w))));
          *reinterpret_cast<uint4*>(h + sw128_offset((uint32_t)et, (uint32_t)(hb * 32 + ch * 8))) = pk;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&h_full[s]);
    }
  } else if (warp < 16) {
    // ================================================================ EPILOGUE 3: D3 -> max-pool -> out
    // warp (8 + 4*g + q): TMEM lane quadrant q (channels 32q..32q+31), centroid g of the tile (columns 64g..64g+63)
    const int q = warp & 3, g = (warp - 8) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int ch = q * 32 + lane;
    const float bias3 = sB3[ch];
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = first_tile + t;
      const int s = t & 1;
      mbar_wait(&m3_done[s], par_of(t));
      tc_fence_after_sync();
      uint32_t v[2][32];
      tmem_ld_x32(tmem + lane_base + TM_D3 + s * C3 + g * 64, v[0]);
      tmem_ld_x32(tmem + lane_base + TM_D3 + s * C3 + g * 64 + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&d3_empty[s]);
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;  // four independent chains
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        m0 = fmaxf(fmaxf(m0, __uint_as_float(v[0][i])), __uint_as_float(v[1][i]));
        m1 = fmaxf(fmaxf(m1, __uint_as_float(v[0][i + 1])), __uint_as_float(v[1][i + 1]));
        m2 = fmaxf(fmaxf(m2, __uint_as_float(v[0][i + 2])), __uint_as_float(v[1][i + 2]));
        m3 = fmaxf(fmaxf(m3, __uint_as_float(v[0][i + 3])), __uint_as_float(v[1][i + 3]));
      }
      const float mval = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      out[((size_t)tile * 2 + g) * C3 + ch] = fmaxf(mval + bias3, 0.f);  // bias + ReLU commute with the max
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace s1v2

void launch_group_rel(int n, int m, long long rows, const float* xyz, const float* new_xyz, const int* idx, void* rel,
                      cudaStream_t st);  // sa_ws.cu

// returns -1 when no instance matches
int sa1_ws2_dispatch(int b, int n, int c, int m, const float* xyz, const float* feat, const float* new_xyz, const int* idx,
                     int c1, int c2, int c3, const float* b1, const float* b2, const float* b3, const void* w1_img,
                     const void* w2_img, const void* w3_img, float* out, void* workspace, cudaStream_t st) {
  if (workspace == nullptr || !(c1 == 64 && c2 == 64 && c3 == 128) || c > 13) return -1;
  auto kern = s1v2::sa1_ws2_kernel;
  VNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, s1v2::SMEM));
  int dev = 0, sms = 148;
  VNB_CUDA(cudaGetDevice(&dev));
  VNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long rows = (long long)b * m * 64;
  launch_group_rel(n, m, rows, xyz, new_xyz, idx, workspace, st);
  if (int rc = check_launch("sa_group_mlp_max: grouped relative coordinates")) return rc;
  const int ntiles = b * m / 2;
  const int tpc = (ntiles + sms - 1) / sms;          // one wave: one CTA per SM, contiguous chunks
  const int grid = (ntiles + tpc - 1) / tpc;
  kern<<<grid, s1v2::THREADS, s1v2::SMEM, st>>>(c, b * m, tpc, static_cast<const float4*>(workspace), feat, b1, b2, b3,
                                                static_cast<const char*>(w1_img), static_cast<const char*>(w2_img),
                                                static_cast<const char*>(w3_img), out);
  return check_launch("sa_group_mlp_max (tcgen05, warp-specialised v2, narrow input)");
}

}  // namespace vnb

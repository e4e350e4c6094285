// Warp-specialised, software-pipelined fused set-abstraction kernel for NARROW inputs (sa1: 3 + c <= 16 input channels):
// all three layers on the tensor cores, every stage double-buffered.
//
//   group -> [rel_xyz, feat] (K padded to 16) -> L1 -> L2 -> L3 -> max over the 64 samples        (utils.py:49-55,120-132)
//
//   warps 12-15 PRODUCER   one grouped row per thread: {rel_xyz, source row} from the helper kernel's table (prefetched
//                          one tile ahead), c feature floats, fp16, swizzled 32-byte row of A0[t%2]        -> a0_full
//   warp  16    MMA        M1(t): D1 = A0 . W1^T    M2(t): D2 = H1 . W2^T    M3(t): D3 = W3^T . H2^T (transposed)
//                          issued as M1(k), M2(k-1), M3(k-2) per step so the pipe always has independent work
//   warps 0-3   EPILOGUE1  D1 -> +b1, ReLU, fp16 -> H1[t%2]
//   warps 4-7   EPILOGUE2  D2 -> +b2, ReLU, fp16 -> H2[t%2]
//   warps 8-11  EPILOGUE3  D3 (channel per lane) -> max over each centroid's 64 samples, +b3, ReLU -> out
// TMEM: D1[2] (2 x C1) | D2[2] (2 x C2) | D3[2] (2 x 128) columns.  All hand-offs are mbarriers.
#include "common.cuh"
#include "umma.cuh"

namespace vnb {

using namespace umma;

__device__ __forceinline__ uint32_t s1_pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int C1, int C2, int C3>
struct S1Cfg {
  static_assert(C1 == 64 && C2 == 64 && C3 == 128, "instance for mlp (64,64,128)");
  static constexpr int W1_BYTES = C1 * 128;         // [C1][16] in one panel
  static constexpr int W2_BYTES = C2 * 128;         // [C2][C1=64]
  static constexpr int W3_BYTES = C3 * 128;         // [C3][C2=64]
  static constexpr int A0_BYTES = 128 * 128;        // per buffer
  static constexpr int H_BYTES = 128 * 128;         // H1 / H2 per buffer (64 columns = one panel)
  static constexpr int OFF_W1 = 0;
  static constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
  static constexpr int OFF_W3 = OFF_W2 + W2_BYTES;
  static constexpr int OFF_A0 = OFF_W3 + W3_BYTES;
  static constexpr int OFF_H1 = OFF_A0 + 2 * A0_BYTES;
  static constexpr int OFF_H2 = OFF_H1 + 2 * H_BYTES;
  static constexpr int OFF_F = OFF_H2 + 2 * H_BYTES;          // floats b1 | b2 | b3
  static constexpr int OFF_BAR = OFF_F + (C1 + C2 + C3) * 4;
  static constexpr int SMEM = OFF_BAR + 32 * 8 + 16 + 1024;
  static constexpr int TM_D1 = 0, TM_D2 = 2 * C1, TM_D3 = 2 * C1 + 2 * C2;
  static constexpr int TM_COLS = 512;
  static_assert(TM_D3 + 2 * C3 <= 512, "TMEM budget");
};

constexpr int S1_THREADS = 17 * 32;

template <int C1, int C2, int C3>
__global__ void __launch_bounds__(S1_THREADS, 1) sa1_ws_kernel(int c, int total_centroids, int tiles_per_cta,
                                                               const float4* __restrict__ rel,
                                                               const float* __restrict__ feat,
                                                               const float* __restrict__ b1, const float* __restrict__ b2,
                                                               const float* __restrict__ b3, const char* __restrict__ w1_img,
                                                               const char* __restrict__ w2_img,
                                                               const char* __restrict__ w3_img, float* __restrict__ out) {
  using Cfg = S1Cfg<C1, C2, C3>;
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_w = bars;          // weights landed
  uint64_t* a0_full = bars + 1;    // [2] 128 producer arrivals
  uint64_t* m1_done = bars + 3;    // [2] commit
  uint64_t* h1_full = bars + 5;    // [2] 128 E1 arrivals
  uint64_t* d1_empty = bars + 7;   // [2] 128 E1 arrivals
  uint64_t* m2_done = bars + 9;    // [2] commit
  uint64_t* h2_full = bars + 11;   // [2] 128 E2 arrivals
  uint64_t* d2_empty = bars + 13;  // [2] 128 E2 arrivals
  uint64_t* m3_done = bars + 15;   // [2] commit
  uint64_t* d3_empty = bars + 17;  // [2] 128 E3 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a0_full[s], 128); mbar_init(&m1_done[s], 1); mbar_init(&h1_full[s], 128); mbar_init(&d1_empty[s], 128);
      mbar_init(&m2_done[s], 1); mbar_init(&h2_full[s], 128); mbar_init(&d2_empty[s], 128);
      mbar_init(&m3_done[s], 1); mbar_init(&d3_empty[s], 128);
    }
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_w, (uint32_t)(Cfg::W1_BYTES + Cfg::W2_BYTES + Cfg::W3_BYTES));
    bulk_g2s(sW1, w1_img, Cfg::W1_BYTES, bar_w);
    bulk_g2s(sW2, w2_img, Cfg::W2_BYTES, bar_w);
    bulk_g2s(sW3, w3_img, Cfg::W3_BYTES, bar_w);
  }
  for (int i = tid; i < C1; i += S1_THREADS) sB1[i] = b1[i];
  for (int i = tid; i < C2; i += S1_THREADS) sB2[i] = b2[i];
  for (int i = tid; i < C3; i += S1_THREADS) sB3[i] = b3[i];
  // A0 padding columns (k >= 16 of the 64-column panel are never read; k in [3+c,16) must be finite): zero both buffers
  for (int i = tid; i < 2 * Cfg::A0_BYTES / 16; i += S1_THREADS) reinterpret_cast<uint4*>(sA0)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(tmem_ptr, Cfg::TM_COLS);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  const int ntiles = total_centroids / 2;
  const int first_tile = (int)blockIdx.x * tiles_per_cta;
  const int my_tiles = min(tiles_per_cta, ntiles - first_tile);
  // use index u of a [2]-slotted barrier for tile t is t>>1 -> parity (t>>1)&1
  auto par_of = [](int t) { return (uint32_t)((t >> 1) & 1); };

  if (warp >= 12 && warp < 16) {
    // ================================================================ PRODUCER: one grouped row per thread
    const int pt = tid - 384;
    float4 rl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (my_tiles > 0) rl = __ldg(rel + (size_t)first_tile * 128 + pt);
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t & 1;
      const float4 cur = rl;
      if (t + 1 < my_tiles) rl = __ldg(rel + (size_t)(first_tile + t + 1) * 128 + pt);
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
      v[0] = cur.x; v[1] = cur.y; v[2] = cur.z;
      const float* f = feat + (size_t)__float_as_int(cur.w) * c;
#pragma unroll
      for (int i = 0; i < 13; ++i)
        if (i < c) v[3 + i] = __ldg(f + i);
      if (t >= 2) mbar_wait(&m1_done[s], par_of(t - 2));  // M1(t-2) finished reading A0[s]
      uint8_t* a0 = sA0 + s * Cfg::A0_BYTES;
      *reinterpret_cast<uint4*>(a0 + sw128_offset((uint32_t)pt, 0)) =
          make_uint4(s1_pack(v[0], v[1]), s1_pack(v[2], v[3]), s1_pack(v[4], v[5]), s1_pack(v[6], v[7]));
      *reinterpret_cast<uint4*>(a0 + sw128_offset((uint32_t)pt, 8)) =
          make_uint4(s1_pack(v[8], v[9]), s1_pack(v[10], v[11]), s1_pack(v[12], v[13]), s1_pack(v[14], v[15]));
      fence_proxy_async_smem();
      mbar_arrive(&a0_full[s]);
    }
  } else if (warp == 16) {
    // ================================================================ MMA issuer
    if (lane == 0 && my_tiles > 0) {
      mbar_wait(bar_w, 0);
      const uint32_t id1 = make_idesc_f16_f32(128, C1), id2 = make_idesc_f16_f32(128, C2), id3 = make_idesc_f16_f32(128, 128);
      for (int k = 0; k < my_tiles + 2; ++k) {
        if (k - 2 >= 0 && k - 2 < my_tiles) {  // M3(t): D3[s] = W3^T . H2[s]^T
          const int t = k - 2, s = t & 1;
          mbar_wait(&h2_full[s], par_of(t));
          if (t >= 2) mbar_wait(&d3_empty[s], par_of(t - 2));
          tc_fence_after_sync();
          const uint32_t a0 = smem_u32(sW3), b0 = smem_u32(sH2 + s * Cfg::H_BYTES);
#pragma unroll
          for (int hh = 0; hh < C3 / 128; ++hh)
#pragma unroll
            for (int ks = 0; ks < C2 / 16; ++ks)
              mma_f16_ss(tmem + Cfg::TM_D3 + s * C3 + hh * 128, make_desc_sw128(a0 + hh * (128 * 128) + ks * 32),
                         make_desc_sw128(b0 + ks * 32), id3, ks > 0 ? 1u : 0u);
          mma_commit(&m3_done[s]);
        }
        if (k - 1 >= 0 && k - 1 < my_tiles) {  // M2(t): D2[s] = H1[s] . W2^T
          const int t = k - 1, s = t & 1;
          mbar_wait(&h1_full[s], par_of(t));
          if (t >= 2) mbar_wait(&d2_empty[s], par_of(t - 2));
          tc_fence_after_sync();
          const uint32_t a0 = smem_u32(sH1 + s * Cfg::H_BYTES), b0 = smem_u32(sW2);
#pragma unroll
          for (int ks = 0; ks < C1 / 16; ++ks)
            mma_f16_ss(tmem + Cfg::TM_D2 + s * C2, make_desc_sw128(a0 + ks * 32), make_desc_sw128(b0 + ks * 32), id2,
                       ks > 0 ? 1u : 0u);
          mma_commit(&m2_done[s]);
        }
        if (k < my_tiles) {  // M1(t): D1[s] = A0[s] . W1^T  (single K = 16 step)
          const int t = k, s = t & 1;
          mbar_wait(&a0_full[s], par_of(t));
          if (t >= 2) mbar_wait(&d1_empty[s], par_of(t - 2));
          tc_fence_after_sync();
          mma_f16_ss(tmem + Cfg::TM_D1 + s * C1, make_desc_sw128(smem_u32(sA0 + s * Cfg::A0_BYTES)),
                     make_desc_sw128(smem_u32(sW1)), id1, 0u);
          mma_commit(&m1_done[s]);
        }
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
    const uint32_t tm = e2 ? Cfg::TM_D2 : Cfg::TM_D1;
    uint8_t* hbase = e2 ? sH2 : sH1;
    const float* bias = e2 ? sB2 : sB1;
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t & 1;
      mbar_wait(&in_done[s], par_of(t));
      if (t >= 2) mbar_wait(&out_free[s], par_of(t - 2));
      tc_fence_after_sync();
      uint8_t* h = hbase + s * Cfg::H_BYTES;
#pragma unroll 1
      for (int cb = 0; cb < 64; cb += 32) {
        uint32_t v[32];
        tmem_ld_x32(tmem + lane_base + tm + s * 64 + cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 ba = *reinterpret_cast<const float4*>(bias + cb + ch * 8);
          const float4 bb = *reinterpret_cast<const float4*>(bias + cb + ch * 8 + 4);
          const uint4 pk = make_uint4(
              s1_pack(fmaxf(__uint_as_float(v[ch * 8 + 0]) + ba.x, 0.f), fmaxf(__uint_as_float(v[ch * 8 + 1]) + ba.y, 0.f)),
              s1_pack(fmaxf(__uint_as_float(v[ch * 8 + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(v[ch * 8 + 3]) + ba.w, 0.f)),
              s1_pack(fmaxf(__uint_as_float(v[ch * 8 + 4]) + bb.x, 0.f), fmaxf(__uint_as_float(v[ch * 8 + 5]) + bb.y, 0.f)),
              s1_pack(fmaxf(__uint_as_float(v[ch * 8 + 6]) + bb.z, 0.f), fmaxf(__uint_as_float(v[ch * 8 + 7]) + bb.w, 0.f)));
          *reinterpret_cast<uint4*>(h + sw128_offset((uint32_t)et, (uint32_t)(cb + ch * 8))) = pk;
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&d_empty[s]);
      fence_proxy_async_smem();
      mbar_arrive(&h_full[s]);
    }
  } else if (warp < 12) {
    // ================================================================ EPILOGUE 3: D3 -> max-pool -> out (thread = channel)
    const int et = tid - 256;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = first_tile + t;
      const int s = t & 1;
      mbar_wait(&m3_done[s], par_of(t));
      tc_fence_after_sync();
#pragma unroll 1
      for (int part = 0; part < (C3 / 128) * 2; ++part) {
        const int hh = part >> 1, gq = part & 1;
        float mval = -INFINITY;
#pragma unroll 1
        for (int cb = 0; cb < 64; cb += 32) {
          uint32_t v[32];
          tmem_ld_x32(tmem + lane_base + Cfg::TM_D3 + s * C3 + hh * 128 + gq * 64 + cb, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mval = fmaxf(mval, __uint_as_float(v[i]));
        }
        const int ch = hh * 128 + et;
        out[((size_t)tile * 2 + gq) * C3 + ch] = fmaxf(mval + sB3[ch], 0.f);  // bias + ReLU commute with the max
      }
      tc_fence_before_sync();
      mbar_arrive(&d3_empty[s]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TM_COLS);
}

void launch_group_rel(int n, int m, long long rows, const float* xyz, const float* new_xyz, const int* idx,
                      const int* pts_cnt, void* rel, cudaStream_t st);  // sa_ws.cu

// returns -1 when no instance matches
int sa1_ws_dispatch(int b, int n, int c, int m, const float* xyz, const float* feat, const float* new_xyz, const int* idx,
                    int c1, int c2, int c3, const float* b1, const float* b2, const float* b3, const void* w1_img,
                    const void* w2_img, const void* w3_img, float* out, void* workspace, cudaStream_t st) {
  if (workspace == nullptr || !(c1 == 64 && c2 == 64 && c3 == 128) || c > 13) return -1;
  using Cfg = S1Cfg<64, 64, 128>;
  auto kern = sa1_ws_kernel<64, 64, 128>;
  VNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  int dev = 0, sms = 148;
  VNB_CUDA(cudaGetDevice(&dev));
  VNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long rows = (long long)b * m * 64;
  launch_group_rel(n, m, rows, xyz, new_xyz, idx, nullptr, workspace, st);
  if (int rc = check_launch("sa_group_mlp_max: grouped relative coordinates")) return rc;
  const int ntiles = b * m / 2;
  int tpc = ntiles / (2 * sms);
  tpc = tpc < 2 ? 2 : (tpc > 32 ? 32 : tpc);
  const int grid = (ntiles + tpc - 1) / tpc;
  kern<<<grid, S1_THREADS, Cfg::SMEM, st>>>(c, b * m, tpc, static_cast<const float4*>(workspace), feat, b1, b2, b3,
                                            static_cast<const char*>(w1_img), static_cast<const char*>(w2_img),
                                            static_cast<const char*>(w3_img), out);
  return check_launch("sa_group_mlp_max (tcgen05, warp-specialised, narrow input)");
}

}  // namespace vnb

// Warp-specialised, software-pipelined version of the fused set-abstraction kernel (hoisted layer 1; sa2..sa4, proposal).
//
// Same arithmetic as sa_tc_kernel (mlp_tc.cu) — group -> relu(q[idx] + W1x^T rel_xyz) -> L2 -> L3 -> max-pool
// (reference utils.py:49-55,120-132) — but the four stages of a 128-row tile run on different warps and overlap with
// the tensor core across tiles:
//
//   warps 8-15  PRODUCER   gather q rows (16 threads per row, 16-byte coalesced chunks, 8 independent loads in flight per
//                          thread), add the rank-3 relative-xyz term, ReLU, fp16, write the swizzled A operand H1[t%2]
//                          -> h1_full[t%2].  The relative coordinates + q-row index of every grouped row come from a
//                          tiny helper kernel (group_rel_kernel) so the gather addresses carry no dependent-load chain.
//   warp  16    MMA        one elected thread issues  M2(t): D2[t%2] = H1[t%2] . W2^T            -> m2_done[t%2] (commit)
//                                                      M3(t): D3 = W3^T . H2^T  (transposed)      -> m3_done     (commit)
//   warps 0-3   EPILOGUE2  D2[t%2] (TMEM) -> +bias, ReLU, fp16 -> H2 (B operand of M3)            -> d2_empty[t%2], h2_full
//   warps 4-7   EPILOGUE3  D3 (TMEM, channel per lane) -> max over each centroid's 64 samples, +bias, ReLU -> out; d3_empty
//
// Issue order on the tensor pipe is M2(0), M2(1), M3(0), M2(2), M3(1), ... so the pipe works on tile t+1's layer 2 while
// the epilogue warps turn tile t's D2 into H2.  All hand-offs are mbarriers; tcgen05.commit signals MMA completion.
// Weights (W2, W3 images) are staged once per CTA by bulk copies and stay resident; one CTA per SM, each owning a contiguous chunk of tiles.
#include "common.cuh"
#include "umma.cuh"

namespace vnb {

using namespace umma;

__device__ __forceinline__ uint32_t pack_h2f(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int C1, int C2, int C3>
struct WsCfg {
  static constexpr int P1 = C1 / 64, P2 = C2 / 64;
  static constexpr int W2_BYTES = P1 * C2 * 128;
  static constexpr int W3_BYTES = P2 * C3 * 128;
  static constexpr int H1_BYTES = P1 * 128 * 128;   // per buffer
  static constexpr int H2_BYTES = P2 * 128 * 128;
  static constexpr int OFF_W2 = 0;
  static constexpr int OFF_W3 = OFF_W2 + W2_BYTES;
  static constexpr int OFF_H1 = OFF_W3 + W3_BYTES;          // two buffers
  static constexpr int OFF_H2 = OFF_H1 + 2 * H1_BYTES;
  static constexpr int OFF_REL = OFF_H2 + H2_BYTES;          // float4[2][128]: rel xyz + q row index of each tile row
  static constexpr int OFF_F = OFF_REL + 2 * 128 * 16;           // floats: b2[C2] | b3[C3]
  static constexpr int OFF_BAR = OFF_F + (C2 + C3) * 4;
  static constexpr int SMEM = OFF_BAR + 16 * 8 + 16 + 1024;
  static constexpr int TM_D2 = 0;                            // two buffers of C2 columns
  static constexpr int TM_D3 = 2 * C2;
  static constexpr int TM_USED = TM_D3 + C3;
  static constexpr int TM_COLS = TM_USED <= 256 ? 256 : 512;
  static_assert(TM_USED <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

constexpr int WS_THREADS = 17 * 32;
constexpr int WS_PRODUCERS = 256;

// relative coordinates (utils.py:51) and flat source-row index of every grouped row: rel[(g*64+s)] = {xyz[idx]-c, row}
// pts_cnt (optional): rows beyond the centroid's slot (16 / 32 / 64 rows, sa_pack.cu) are never read and are skipped.
__global__ void group_rel_kernel(int n, int m, long long total_rows, const float* __restrict__ xyz,
                                 const float* __restrict__ new_xyz, const int* __restrict__ idx,
                                 const int* __restrict__ pts_cnt, float4* __restrict__ rel) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_rows) return;
  const int g = (int)(t >> 6);
  if (pts_cnt != nullptr) {
    const int c = pts_cnt[g];
    const int slot = (c <= 0 || c > 32) ? 64 : (c > 16 ? 32 : 16);
    if ((int)(t & 63) >= slot) return;
  }
  const int bi = g / m;
  const int pid = idx[t];
  const float* pp = xyz + ((size_t)bi * n + pid) * 3;
  const float* cc = new_xyz + (size_t)g * 3;
  rel[t] = make_float4(pp[0] - cc[0], pp[1] - cc[1], pp[2] - cc[2], __int_as_float(bi * n + pid));
}

template <int C1, int C2, int C3>
__global__ void __launch_bounds__(WS_THREADS, 1) sa_ws_kernel(int n, int m, int total_centroids, int tiles_per_cta,
                                                              const float4* __restrict__ rel,
                                                              const float* __restrict__ w1x /* (3,C1) */,
                                                              const float* __restrict__ b2, const float* __restrict__ b3,
                                                              const char* __restrict__ w2_img,
                                                              const char* __restrict__ w3_img,
                                                              const __half* __restrict__ q, float* __restrict__ out) {
  using Cfg = WsCfg<C1, C2, C3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* sW2 = smem + Cfg::OFF_W2;
  uint8_t* sW3 = smem + Cfg::OFF_W3;
  uint8_t* sH1 = smem + Cfg::OFF_H1;
  uint8_t* sH2 = smem + Cfg::OFF_H2;
  float4* sRel = reinterpret_cast<float4*>(smem + Cfg::OFF_REL);
  float* sB2 = reinterpret_cast<float*>(smem + Cfg::OFF_F);
  float* sB3 = sB2 + C2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_w = bars + 0;
  uint64_t* h1_full = bars + 1;    // [2] 256 producer arrivals
  uint64_t* m2_done = bars + 3;    // [2] tcgen05.commit
  uint64_t* d2_empty = bars + 5;   // [2] 128 epilogue-2 arrivals
  uint64_t* h2_full = bars + 7;    //     128 epilogue-2 arrivals
  uint64_t* m3_done = bars + 8;    //     tcgen05.commit
  uint64_t* d3_empty = bars + 9;   //     128 epilogue-3 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&h1_full[s], WS_PRODUCERS); mbar_init(&m2_done[s], 1); mbar_init(&d2_empty[s], 128); }
    mbar_init(h2_full, 128); mbar_init(m3_done, 1); mbar_init(d3_empty, 128);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_w, (uint32_t)(Cfg::W2_BYTES + Cfg::W3_BYTES));
    bulk_g2s(sW2, w2_img, Cfg::W2_BYTES, bar_w);
    bulk_g2s(sW3, w3_img, Cfg::W3_BYTES, bar_w);
  }
  for (int i = tid; i < C2; i += WS_THREADS) sB2[i] = b2[i];
  for (int i = tid; i < C3; i += WS_THREADS) sB3[i] = b3[i];
  if (warp == 0) tmem_alloc(tmem_ptr, Cfg::TM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  const int ntiles = total_centroids / 2;
  // a CTA owns a contiguous chunk of tiles; more CTAs than SMs, so the hardware scheduler balances the chunks over
  // whatever SMs are free (other kernels of overlapping steps may hold some)
  const int first_tile = (int)blockIdx.x * tiles_per_cta;
  const int my_tiles = min(tiles_per_cta, ntiles - first_tile);

  if (warp >= 8 && warp < 16) {
    // ================================================================ PRODUCER (256 threads)
    const int pt = tid - 256;          // 0..255
    const int chunk = pt & 15;         // 8 channels [8*chunk, 8*chunk+8)
    const int rsub = pt >> 4;          // rows rsub, rsub+16, ...
    float wx[8], wy[8], wz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      wx[i] = w1x[chunk * 8 + i]; wy[i] = w1x[C1 + chunk * 8 + i]; wz[i] = w1x[2 * C1 + chunk * 8 + i];
    }
    float4 relreg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pt < 128 && my_tiles > 0) relreg = __ldg(rel + (size_t)first_tile * 128 + pt);
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = first_tile + t;
      const int s = t & 1;
      if (pt < 128) {
        sRel[s * 128 + pt] = relreg;                                                    // this tile
        if (t + 1 < my_tiles) relreg = __ldg(rel + (size_t)(tile + 1) * 128 + pt);      // prefetch the next
      }
      named_bar_sync(1, WS_PRODUCERS);  // sRel[s] visible; its previous readers (tile t-2) passed the barrier of tile t-1
      if (t >= 2) mbar_wait(&m2_done[s], (uint32_t)(((t >> 1) - 1) & 1));  // M2(t-2) finished reading H1[s]
      uint8_t* h1 = sH1 + s * Cfg::H1_BYTES;
      const float4* srel = sRel + s * 128;
      uint4 raw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)  // all gathers first: 8 independent 16-byte loads in flight per thread
        raw[i] = __ldg(reinterpret_cast<const uint4*>(q + (size_t)__float_as_int(srel[rsub + 16 * i].w) * C1) + chunk);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rsub + 16 * i;
        const float4 rl = srel[r];
        const __half2* hh = reinterpret_cast<const __half2*>(&raw[i]);
        float o[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __half22float2(hh[j]);
          o[2 * j] = f.x; o[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = fmaf(wx[j], rl.x, o[j]);
          a = fmaf(wy[j], rl.y, a);
          a = fmaf(wz[j], rl.z, a);
          o[j] = fmaxf(a, 0.f);
        }
        const uint4 pk = make_uint4(pack_h2f(o[0], o[1]), pack_h2f(o[2], o[3]), pack_h2f(o[4], o[5]), pack_h2f(o[6], o[7]));
        const uint32_t kk = (uint32_t)chunk * 8;
        *reinterpret_cast<uint4*>(h1 + (kk >> 6) * (128 * 128) + sw128_offset((uint32_t)r, kk)) = pk;
      }
      fence_proxy_async_smem();
      mbar_arrive(&h1_full[s]);
    }
  } else if (warp == 16) {
    // ================================================================ MMA issuer (one thread)
    if (lane == 0 && my_tiles > 0) {
      mbar_wait(bar_w, 0);
      const uint32_t idesc2 = make_idesc_f16_f32(128, C2);
      const uint32_t idesc3 = make_idesc_f16_f32(128, 128);
      auto issue_m2 = [&](int t) {
        const int s = t & 1;
        mbar_wait(&h1_full[s], (uint32_t)((t >> 1) & 1));
        if (t >= 2) mbar_wait(&d2_empty[s], (uint32_t)(((t >> 1) - 1) & 1));  // E2(t-2) drained D2[s]
        tc_fence_after_sync();
        const uint32_t a0 = smem_u32(sH1 + s * Cfg::H1_BYTES), b0 = smem_u32(sW2);
#pragma unroll
        for (int ks = 0; ks < C1 / 16; ++ks) {
          const uint32_t pan = (uint32_t)ks >> 2, kin = (uint32_t)ks & 3;
          mma_f16_ss(tmem + Cfg::TM_D2 + s * C2, make_desc_sw128(a0 + pan * (128 * 128) + kin * 32),
                     make_desc_sw128(b0 + pan * (C2 * 128) + kin * 32), idesc2, ks > 0 ? 1u : 0u);
        }
        mma_commit(&m2_done[s]);
      };
      issue_m2(0);
      for (int t = 0; t < my_tiles; ++t) {
        if (t + 1 < my_tiles) issue_m2(t + 1);
        mbar_wait(h2_full, (uint32_t)(t & 1));
        if (t >= 1) mbar_wait(d3_empty, (uint32_t)((t - 1) & 1));  // E3(t-1) drained D3
        tc_fence_after_sync();
        const uint32_t a0 = smem_u32(sW3), b0 = smem_u32(sH2);
#pragma unroll
        for (int hh = 0; hh < C3 / 128; ++hh) {
#pragma unroll
          for (int ks = 0; ks < C2 / 16; ++ks) {
            const uint32_t pan = (uint32_t)ks >> 2, kin = (uint32_t)ks & 3;
            mma_f16_ss(tmem + Cfg::TM_D3 + hh * 128, make_desc_sw128(a0 + pan * (C3 * 128) + hh * (128 * 128) + kin * 32),
                       make_desc_sw128(b0 + pan * (128 * 128) + kin * 32), idesc3, ks > 0 ? 1u : 0u);
          }
        }
        mma_commit(m3_done);
      }
    }
  } else if (warp < 4) {
    // ================================================================ EPILOGUE 2: D2 -> H2   (thread = tile row)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t & 1;
      mbar_wait(&m2_done[s], (uint32_t)((t >> 1) & 1));
      if (t >= 1) mbar_wait(m3_done, (uint32_t)((t - 1) & 1));  // M3(t-1) finished reading H2
      tc_fence_after_sync();
#pragma unroll
      for (int cb = 0; cb < C2; cb += 32) {
        uint32_t v[32];
        tmem_ld_x32(tmem + lane_base + Cfg::TM_D2 + s * C2 + cb, v);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 ba = *reinterpret_cast<const float4*>(sB2 + cb + ch * 8);
          const float4 bb = *reinterpret_cast<const float4*>(sB2 + cb + ch * 8 + 4);
          const uint4 pk = make_uint4(
              pack_h2f(fmaxf(__uint_as_float(v[ch * 8 + 0]) + ba.x, 0.f), fmaxf(__uint_as_float(v[ch * 8 + 1]) + ba.y, 0.f)),
              pack_h2f(fmaxf(__uint_as_float(v[ch * 8 + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(v[ch * 8 + 3]) + ba.w, 0.f)),
              pack_h2f(fmaxf(__uint_as_float(v[ch * 8 + 4]) + bb.x, 0.f), fmaxf(__uint_as_float(v[ch * 8 + 5]) + bb.y, 0.f)),
              pack_h2f(fmaxf(__uint_as_float(v[ch * 8 + 6]) + bb.z, 0.f), fmaxf(__uint_as_float(v[ch * 8 + 7]) + bb.w, 0.f)));
          const uint32_t kk = (uint32_t)(cb + ch * 8);
          *reinterpret_cast<uint4*>(sH2 + (kk >> 6) * (128 * 128) + sw128_offset((uint32_t)tid, kk)) = pk;
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&d2_empty[s]);
      fence_proxy_async_smem();
      mbar_arrive(h2_full);
    }
  } else {
    // ================================================================ EPILOGUE 3: D3 -> max-pool -> out   (thread = channel)
    const int et = tid - 128;  // 0..127 == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = first_tile + t;
      mbar_wait(m3_done, (uint32_t)(t & 1));
      tc_fence_after_sync();
      // one (channel half, centroid) at a time, not unrolled: keeps a single 32-register TMEM load live
#pragma unroll 1
      for (int part = 0; part < (C3 / 128) * 2; ++part) {
        const int hh = part >> 1, gq = part & 1;
        float mval = -INFINITY;
#pragma unroll 1
        for (int cb = 0; cb < 64; cb += 32) {
          uint32_t v[32];
          tmem_ld_x32(tmem + lane_base + Cfg::TM_D3 + hh * 128 + gq * 64 + cb, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mval = fmaxf(mval, __uint_as_float(v[i]));
        }
        const int ch = hh * 128 + et;
        // bias + ReLU commute with the max (both monotone)
        out[((size_t)tile * 2 + gq) * C3 + ch] = fmaxf(mval + sB3[ch], 0.f);
      }
      tc_fence_before_sync();
      mbar_arrive(d3_empty);  // D3 is free for M3(t+1)
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TM_COLS);
}

void launch_group_rel(int n, int m, long long rows, const float* xyz, const float* new_xyz, const int* idx,
                      const int* pts_cnt, void* rel, cudaStream_t st) {
  group_rel_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(n, m, rows, xyz, new_xyz, idx, pts_cnt,
                                                                   static_cast<float4*>(rel));
}

template <int C1, int C2, int C3>
static int launch_ws(int b, int n, int m, const float* xyz, const float* new_xyz, const int* idx, const float* w1x,
                     const float* b2, const float* b3, const void* w2_img, const void* w3_img, const void* q, float* out,
                     void* workspace, cudaStream_t st) {
  using Cfg = WsCfg<C1, C2, C3>;
  auto kern = sa_ws_kernel<C1, C2, C3>;
  VNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  int dev = 0, sms = 148;
  VNB_CUDA(cudaGetDevice(&dev));
  VNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ntiles = b * m / 2;
  float4* rel = static_cast<float4*>(workspace);
  const long long rows = (long long)b * m * 64;
  group_rel_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(n, m, rows, xyz, new_xyz, idx, nullptr, rel);
  if (int rc = check_launch("sa_group_mlp_max: grouped relative coordinates")) return rc;
  int tpc = ntiles / (2 * sms);
  tpc = tpc < 2 ? 2 : (tpc > 16 ? 16 : tpc);
  const int grid = (ntiles + tpc - 1) / tpc;
  kern<<<grid, WS_THREADS, Cfg::SMEM, st>>>(n, m, b * m, tpc, rel, w1x, b2, b3, static_cast<const char*>(w2_img),
                                            static_cast<const char*>(w3_img), static_cast<const __half*>(q), out);
  return check_launch("sa_group_mlp_max (tcgen05, warp-specialised)");
}

// returns -1 when no instance matches
int sa_ws_dispatch(int b, int n, int m, const float* xyz, const float* new_xyz, const int* idx, int c1, int c2, int c3,
                   const float* w1x, const float* b2, const float* b3, const void* w2_img, const void* w3_img,
                   const void* q, float* out, void* workspace, cudaStream_t st) {
  if (workspace == nullptr) return -1;
  if (c1 == 128 && c2 == 128 && c3 == 256)
    return launch_ws<128, 128, 256>(b, n, m, xyz, new_xyz, idx, w1x, b2, b3, w2_img, w3_img, q, out, workspace, st);
  if (c1 == 128 && c2 == 128 && c3 == 128)
    return launch_ws<128, 128, 128>(b, n, m, xyz, new_xyz, idx, w1x, b2, b3, w2_img, w3_img, q, out, workspace, st);
  return -1;
}

}  // namespace vnb

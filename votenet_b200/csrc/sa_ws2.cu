// Fused set-abstraction kernel with hoisted layer 1 (sa2..sa4, proposal), second generation.
//
// Same arithmetic as sa_ws_kernel (sa_ws.cu) — group -> relu(q[idx] + W1x^T rel_xyz) -> L2 -> L3 -> max-pool
// (reference utils.py:49-55,120-132) — with the serialisations of the first generation removed:
//   * layer 3 is issued K-CHUNK-MAJOR (k-steps of chunk 0 for both channel halves, then chunk 1, ...) and H2 is handed
//     over in four 32-column chunks, each with its own full/free mbarrier pair: the tensor pipe starts M3(t) when the
//     first quarter of H2(t) exists and EPILOGUE2(t+1) refills a chunk as soon as M3(t) has consumed it, so E2 and M3
//     overlap although H2 is single-buffered (shared memory is full).  (Two 64-column chunks — NCHUNK = C2 / 64, the code
//     is parametrised — make EPILOGUE2 17 % shorter but lengthen M3's tail behind it: tile period 3200 -> 3600 cycles.);
//   * layers 2 and 3 have separate issuing warps: M2(t+1) never waits behind M3(t)'s operands or vice versa;
//   * EPILOGUE3 uses two warps per TMEM lane quadrant (a warp pulls at most 64 B/cycle out of TMEM) and drains D3 in
//     ~350 cycles, hidden behind M2 on the tensor pipe;
//   * all TMEM loads of a stage are in flight before the first wait; packed fp32x2 math (FFMA2 / FADD2) and fp16x2 ReLU;
//   * one wave: one CTA per SM, contiguous chunks of tiles (one prologue, one pipeline fill per SM).
// Measured (vnb_debug_sa_trace + ncu, profiles/): the kernel is SHARED-MEMORY-BANDWIDTH bound, not tensor- or ALU-bound —
// both operands of every tcgen05.mma stream from shared memory (192 KB per tile = 1536 cycles at 128 B/cycle, exactly the
// tensor time) while the producer / epilogue stores and the broadcast loads add ~1500 wavefronts per tile on the same
// banks; period ~2900 cycles per tile.  Variants that ADD shared-memory traffic or fixed latencies lost: cp.async gathers
// transformed in place (+512 wavefronts), two alternating EPILOGUE2 groups, one polling lane per warp.  The next step is
// structural: layer-2's A operand (H1) in TMEM via tcgen05.st (TS-mode MMA), which removes a third of the traffic.
//
//   warps 0-3    EPILOGUE2  D2[t%2] -> +b2, ReLU, fp16 -> H2 chunk by chunk
//   warps 4-11   EPILOGUE3  D3 (channel per lane) -> max over a centroid's 64 samples, +b3, ReLU -> out
//   warps 12-19  PRODUCER   gather q rows (16 threads per row, 8 independent 16-byte loads in flight per thread, issued
//                           BEFORE the wait for the H1 buffer), + rank-3 relative-xyz term, ReLU, fp16 -> H1[t%2]
//   warp  20     MMA2       D2[t%2] = H1[t%2] . W2^T
//   warp  21     MMA3       D3 = W3^T . H2^T   (transposed; chunk-major)
#include "common.cuh"
#include "umma.cuh"

namespace vnb {

using namespace umma;

namespace s2v2 {

template <int C1, int C2, int C3>
struct Cfg {
  static_assert(C1 == 128 && C2 == 128 && (C3 == 128 || C3 == 256), "instances: (128,128,256) and (128,128,128)");
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
  static constexpr int OFF_F = OFF_REL + 2 * 128 * 16;       // floats: b2[C2] | b3[C3]
  static constexpr int OFF_BAR = OFF_F + (C2 + C3) * 4;
  static constexpr int SMEM = OFF_BAR + 32 * 8 + 16 + 1024;
  static constexpr int TM_D2 = 0;                            // two buffers of C2 columns
  static constexpr int TM_D3 = 2 * C2;
  static constexpr int TM_COLS = 512;
  static_assert(TM_D3 + C3 <= 512, "TMEM budget");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static constexpr int NSUB = C2 / 32;                       // EPILOGUE2 works in 32-column steps (one tcgen05.ld each)
  static constexpr int NCHUNK = C2 / 32;                     // H2 hand-over granularity: 32 columns = 2 k-steps
};

constexpr int THREADS = 22 * 32;
constexpr int PRODUCERS = 256;

// fp16x2{v.x, v.y} = relu(round(.)) in ONE instruction (F2FP.RELU.F16.F32.PACK_AB); rounding is monotone, so relu commutes
__device__ __forceinline__ uint32_t pack_relu(float2 v) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v.y), "f"(v.x));
  return d;
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int C1, int C2, int C3>
__global__ void __launch_bounds__(THREADS, 1) sa_ws2_kernel(const int* __restrict__ hdr,
                                                            const int* __restrict__ tile_cid,
                                                            const float4* __restrict__ rel,
                                                            const float* __restrict__ w1x /* (3,C1) */,
                                                            const float* __restrict__ b2, const float* __restrict__ b3,
                                                            const char* __restrict__ w2_img,
                                                            const char* __restrict__ w3_img,
                                                            const __half* __restrict__ q, float* __restrict__ out,
                                                            int min_tpc, long long* __restrict__ trace) {
  using K = Cfg<C1, C2, C3>;
  // debugging aid (vnb_debug_sa_trace): CTA 0 stamps clock64() at the start (after its input wait) and the end of every
  // stage of its first 64 tiles: trace[(role * 64 + t) * 2 + {0,1}], roles 0 P, 1 M2, 2 E2, 3 M3, 4 E3
  const bool tr = trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0;
#define S2_STAMP(role, t, ph) \
  if (tr && (t) < 64) trace[((role) * 64 + (t)) * 2 + (ph)] = clock64();
  constexpr int NCH = K::NCHUNK, NSUB = K::NSUB, SPC = NSUB / NCH;   // hand-over chunks, 32-column steps, steps per chunk
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* sW2 = smem + K::OFF_W2;
  uint8_t* sW3 = smem + K::OFF_W3;
  uint8_t* sH1 = smem + K::OFF_H1;
  uint8_t* sH2 = smem + K::OFF_H2;
  float4* sRel = reinterpret_cast<float4*>(smem + K::OFF_REL);
  float* sB2 = reinterpret_cast<float*>(smem + K::OFF_F);
  float* sB3 = sB2 + C2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
  uint64_t* bar_w = bars + 0;
  uint64_t* h1_full = bars + 1;     // [2]   256 producer arrivals
  uint64_t* m2_done = bars + 3;     // [2]   tcgen05.commit
  uint64_t* d2_empty = bars + 5;    // [2]   128 epilogue-2 arrivals
  uint64_t* h2c_full = bars + 7;    // [NCH] 128 epilogue-2 arrivals: chunk c of H2(t) written
  uint64_t* m3c_done = bars + 11;   // [NCH] tcgen05.commit: M3(t) has consumed chunk c (the last one == D3(t) complete)
  uint64_t* d3_empty = bars + 15;   //       256 epilogue-3 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 32);
  static_assert(NCH <= 4 && NSUB % NCH == 0, "barrier table laid out for at most 4 chunks");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // tile table written by sa_pack_tiles_kernel (sa_pack.cu): tiles of 2 x 64-row, then 4 x 32-row, then 8 x 16-row slots
  const int ntiles = __ldg(hdr), t64 = __ldg(hdr + 1), t32 = __ldg(hdr + 2);
  // CTAs that share the tiles: a CTA has a fixed cost (prologue, weight copy, pipeline fill and drain ~ 10 us), so with few
  // tiles (the packed deeper levels) fewer CTAs with >= min_tpc tiles each occupy far fewer SM-microseconds — what counts
  // when other forwards' kernels are waiting for SMs; the surplus CTAs exit at once
  const int active = max(1, min((int)gridDim.x, (ntiles + min_tpc - 1) / min_tpc));
  const int tiles_per_cta = (ntiles + active - 1) / active;
  const int first_tile = (int)blockIdx.x * tiles_per_cta;
  const int my_tiles = max(0, min(tiles_per_cta, ntiles - first_tile));
  auto shift_of = [&](int tile) { return tile < t64 ? 6 : (tile < t64 + t32 ? 5 : 4); };  // log2(rows per slot)
  if (my_tiles == 0) return;  // whole CTA, before any barrier / TMEM / bulk copy exists

  if (tid == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&h1_full[s], PRODUCERS); mbar_init(&m2_done[s], 1); mbar_init(&d2_empty[s], 128); }
    for (int c = 0; c < NCH; ++c) { mbar_init(&h2c_full[c], 128); mbar_init(&m3c_done[c], 1); }
    mbar_init(d3_empty, 256);
    fence_barrier_init();
    if (my_tiles > 0) {  // a CTA without tiles must not leave a bulk copy in flight when it exits
      mbar_arrive_expect_tx(bar_w, (uint32_t)(K::W2_BYTES + K::W3_BYTES));
      bulk_g2s(sW2, w2_img, K::W2_BYTES, bar_w);
      bulk_g2s(sW3, w3_img, K::W3_BYTES, bar_w);
    }
  }
  for (int i = tid; i < C2; i += THREADS) sB2[i] = b2[i];
  for (int i = tid; i < C3; i += THREADS) sB3[i] = b3[i];
  if (warp == 0) tmem_alloc(tmem_ptr, K::TM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  if (warp >= 12 && warp < 20) {
    // ================================================================ PRODUCER (256 threads)
    const int pt = tid - 384;          // 0..255
    const int chunk = pt & 15;         // 8 channels [8*chunk, 8*chunk+8)
    const int rsub = pt >> 4;          // rows rsub, rsub+16, ...
    float2 wx[4], wy[4], wz[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      wx[i] = make_float2(w1x[chunk * 8 + 2 * i], w1x[chunk * 8 + 2 * i + 1]);
      wy[i] = make_float2(w1x[C1 + chunk * 8 + 2 * i], w1x[C1 + chunk * 8 + 2 * i + 1]);
      wz[i] = make_float2(w1x[2 * C1 + chunk * 8 + 2 * i], w1x[2 * C1 + chunk * 8 + 2 * i + 1]);
    }
    // tile row pt = sample (pt mod slot) of the centroid in slot (pt / slot); empty slots read row 0 (discarded later)
    auto cid_of = [&](int tile) { return __ldg(tile_cid + (size_t)tile * 8 + (pt >> shift_of(tile))); };
    auto src_row = [&](int tile, int cid) {
      return cid < 0 ? (size_t)0 : (size_t)cid * 64 + (size_t)(pt & ((1 << shift_of(tile)) - 1));
    };
    float4 relreg = make_float4(0.f, 0.f, 0.f, 0.f);
    int cid_nxt = -1;  // centroid of this row in tile t+1 (requested one tile before its table row)
    if (pt < 128 && my_tiles > 0) relreg = __ldg(rel + src_row(first_tile, cid_of(first_tile)));
    if (pt < 128 && my_tiles > 1) cid_nxt = cid_of(first_tile + 1);
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = first_tile + t;
      const int s = t & 1;
      if (pt < 128) {
        sRel[s * 128 + pt] = relreg;                                                    // this tile
        if (t + 1 < my_tiles) relreg = __ldg(rel + src_row(tile + 1, cid_nxt));         // prefetch the next
        if (t + 2 < my_tiles) cid_nxt = cid_of(tile + 2);
      }
      named_bar(1, PRODUCERS);  // sRel[s] visible; its previous readers (tile t-2) passed the barrier of tile t-1
      const float4* srel = sRel + s * 128;
      uint4 raw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)  // all gathers first: 8 independent 16-byte loads in flight per thread
        raw[i] = __ldg(reinterpret_cast<const uint4*>(q + (size_t)__float_as_int(srel[rsub + 16 * i].w) * C1) + chunk);
      if (t >= 2) mbar_wait(&m2_done[s], (uint32_t)(((t >> 1) - 1) & 1));  // M2(t-2) finished reading H1[s]
      if (warp == 12) { S2_STAMP(0, t, 0) }
      uint8_t* h1 = sH1 + s * K::H1_BYTES;
      float4 rl_next = srel[rsub];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rsub + 16 * i;
        const float4 rl = rl_next;
        if (i + 1 < 8) rl_next = srel[r + 16];  // before this row's store: shared loads are kept behind earlier shared stores
        const float2 rx = make_float2(rl.x, rl.x), ry = make_float2(rl.y, rl.y), rz = make_float2(rl.z, rl.z);
        const __half2* hh = reinterpret_cast<const __half2*>(&raw[i]);
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 a = __ffma2_rn(wx[j], rx, __half22float2(hh[j]));
          a = __ffma2_rn(wy[j], ry, a);
          a = __ffma2_rn(wz[j], rz, a);
          pk[j] = pack_relu(a);
        }
        const uint32_t kk = (uint32_t)chunk * 8;
        *reinterpret_cast<uint4*>(h1 + (kk >> 6) * (128 * 128) + sw128_offset((uint32_t)r, kk)) =
            make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&h1_full[s]);
      if (warp == 12) { S2_STAMP(0, t, 1) }
      if (pt < 128 && t + 1 < my_tiles) {
        // relreg holds row pt of the NEXT tile by now: pull its q row (C1 halves = 2 lines) into L1, so the next
        // iteration's gathers are L1 hits instead of L2 round trips (no registers are held across the tile for it)
        const char* rowp = reinterpret_cast<const char*>(q + (size_t)__float_as_int(relreg.w) * C1);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rowp));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rowp + 128));
      }
    }
  } else if (warp == 20) {
    // ================================================================ MMA2: D2[s] = H1[s] . W2^T
    if (lane == 0 && my_tiles > 0) {
      mbar_wait(bar_w, 0);
      const uint32_t idesc2 = make_idesc_f16_f32(128, C2);
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t & 1;
        mbar_wait(&h1_full[s], (uint32_t)((t >> 1) & 1));
        if (t >= 2) mbar_wait(&d2_empty[s], (uint32_t)(((t >> 1) - 1) & 1));  // E2(t-2) drained D2[s]
        tc_fence_after_sync();
        S2_STAMP(1, t, 0)
        const uint32_t a0 = smem_u32(sH1 + s * K::H1_BYTES), b0 = smem_u32(sW2);
#pragma unroll
        for (int ks = 0; ks < C1 / 16; ++ks) {
          const uint32_t pan = (uint32_t)ks >> 2, kin = (uint32_t)ks & 3;
          mma_f16_ss(tmem + K::TM_D2 + s * C2, make_desc_sw128(a0 + pan * (128 * 128) + kin * 32),
                     make_desc_sw128(b0 + pan * (C2 * 128) + kin * 32), idesc2, ks > 0 ? 1u : 0u);
        }
        mma_commit(&m2_done[s]);
        S2_STAMP(1, t, 1)
      }
    }
  } else if (warp == 21) {
    // ================================================================ MMA3: D3 = W3^T . H2^T, chunk-major
    if (lane == 0 && my_tiles > 0) {
      mbar_wait(bar_w, 0);
      const uint32_t idesc3 = make_idesc_f16_f32(128, 128);
      const uint32_t a0 = smem_u32(sW3), b0 = smem_u32(sH2);
      for (int t = 0; t < my_tiles; ++t) {
        const uint32_t par = (uint32_t)(t & 1);
        if (t >= 1) mbar_wait(d3_empty, (uint32_t)((t - 1) & 1));  // E3(t-1) drained D3
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          mbar_wait(&h2c_full[c], par);
          tc_fence_after_sync();
          if (c == 0) { S2_STAMP(3, t, 0) }
#pragma unroll
          for (int k2 = 0; k2 < 2 * SPC; ++k2) {
            const uint32_t ks = (uint32_t)(2 * SPC * c + k2), pan = ks >> 2, kin = ks & 3;
#pragma unroll
            for (int hh = 0; hh < C3 / 128; ++hh)
              mma_f16_ss(tmem + K::TM_D3 + hh * 128, make_desc_sw128(a0 + pan * (C3 * 128) + hh * (128 * 128) + kin * 32),
                         make_desc_sw128(b0 + pan * (128 * 128) + kin * 32), idesc3, ks > 0 ? 1u : 0u);
          }
          mma_commit(&m3c_done[c]);
          if (c == NCH - 1) { S2_STAMP(3, t, 1) }
        }
      }
    }
  } else if (warp < 4) {
    // ================================================================ EPILOGUE 2: D2 -> H2   (thread = tile row)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t & 1;
      mbar_wait(&m2_done[s], (uint32_t)((t >> 1) & 1));
      tc_fence_after_sync();
      if (warp == 0) { S2_STAMP(2, t, 0) }
      uint32_t v[2][32];
      tmem_ld_x32(tmem + lane_base + K::TM_D2 + s * C2, v[0]);
#pragma unroll
      for (int c = 0; c < NSUB; ++c) {
        tmem_ld_wait();  // 32-column step c is in registers
        if (c + 1 < NSUB) {
          tmem_ld_x32(tmem + lane_base + K::TM_D2 + s * C2 + (c + 1) * 32, v[(c + 1) & 1]);  // in flight during the math
        } else {
          tc_fence_before_sync();
          mbar_arrive(&d2_empty[s]);  // the whole accumulator has been read
        }
        if (t >= 1 && c % SPC == 0) mbar_wait(&m3c_done[c / SPC], (uint32_t)((t - 1) & 1));  // M3(t-1) has consumed this chunk of H2
        const uint32_t* vc = v[c & 1];
        // bias of group ch+1 is loaded BEFORE group ch is stored (shared loads are kept behind earlier shared stores)
        float4 ba = *reinterpret_cast<const float4*>(sB2 + c * 32), bb = *reinterpret_cast<const float4*>(sB2 + c * 32 + 4);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float4 na = ba, nb = bb;
          if (ch + 1 < 4) {
            na = *reinterpret_cast<const float4*>(sB2 + c * 32 + (ch + 1) * 8);
            nb = *reinterpret_cast<const float4*>(sB2 + c * 32 + (ch + 1) * 8 + 4);
          }
          const uint32_t* vv = vc + ch * 8;
          const uint4 pk = make_uint4(
              pack_relu(__fadd2_rn(make_float2(__uint_as_float(vv[0]), __uint_as_float(vv[1])), make_float2(ba.x, ba.y))),
              pack_relu(__fadd2_rn(make_float2(__uint_as_float(vv[2]), __uint_as_float(vv[3])), make_float2(ba.z, ba.w))),
              pack_relu(__fadd2_rn(make_float2(__uint_as_float(vv[4]), __uint_as_float(vv[5])), make_float2(bb.x, bb.y))),
              pack_relu(__fadd2_rn(make_float2(__uint_as_float(vv[6]), __uint_as_float(vv[7])), make_float2(bb.z, bb.w))));
          const uint32_t kk = (uint32_t)(c * 32 + ch * 8);
          *reinterpret_cast<uint4*>(sH2 + (kk >> 6) * (128 * 128) + sw128_offset((uint32_t)tid, kk)) = pk;
          ba = na; bb = nb;
        }
        if (c % SPC == SPC - 1) {
          fence_proxy_async_smem();
          mbar_arrive(&h2c_full[c / SPC]);
        }
      }
      if (warp == 0) { S2_STAMP(2, t, 1) }
    }
  } else if (warp < 12) {
    // ================================================================ EPILOGUE 3: D3 -> max-pool -> out
    // warp (4 + 4*g + qd): TMEM lane quadrant qd (channels hh*128 + 32qd..+31), centroid g of the tile (columns 64g..64g+63)
    const int qd = warp & 3, g = (warp - 4) >> 2;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    float bias3[C3 / 128];
#pragma unroll
    for (int hh = 0; hh < C3 / 128; ++hh) bias3[hh] = sB3[hh * 128 + qd * 32 + lane];
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = first_tile + t;
      const int shift = shift_of(tile);
      // centroids whose samples sit in this warp's 64 columns: one (64-row slots), two (32) or four (16)
      int cid[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int slot = shift == 6 ? g : (shift == 5 ? 2 * g + (k >> 1) : 4 * g + k);
        cid[k] = __ldg(tile_cid + (size_t)tile * 8 + slot);
      }
      mbar_wait(&m3c_done[NCH - 1], (uint32_t)(t & 1));  // the last chunk's commit == D3(t) complete
      tc_fence_after_sync();
      if (warp == 4) { S2_STAMP(4, t, 0) }
#pragma unroll
      for (int hh = 0; hh < C3 / 128; ++hh) {
        uint32_t v[2][32];
        tmem_ld_x32(tmem + lane_base + K::TM_D3 + hh * 128 + g * 64, v[0]);
        tmem_ld_x32(tmem + lane_base + K::TM_D3 + hh * 128 + g * 64 + 32, v[1]);
        tmem_ld_wait();
        if (hh == C3 / 128 - 1) {
          tc_fence_before_sync();
          mbar_arrive(d3_empty);  // D3 has been read: free for M3(t+1)
        }
        float mb[4];  // maxima of the four 16-column blocks (independent chains)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t* vv = &v[k >> 1][(k & 1) * 16];
          float a0 = __uint_as_float(vv[0]), a1 = __uint_as_float(vv[1]), a2 = __uint_as_float(vv[2]), a3 = __uint_as_float(vv[3]);
#pragma unroll
          for (int i = 4; i < 16; i += 4) {
            a0 = fmaxf(a0, __uint_as_float(vv[i])); a1 = fmaxf(a1, __uint_as_float(vv[i + 1]));
            a2 = fmaxf(a2, __uint_as_float(vv[i + 2])); a3 = fmaxf(a3, __uint_as_float(vv[i + 3]));
          }
          mb[k] = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        }
        if (shift == 6) {
          mb[0] = fmaxf(fmaxf(mb[0], mb[1]), fmaxf(mb[2], mb[3]));
        } else if (shift == 5) {
          mb[0] = fmaxf(mb[0], mb[1]);
          mb[2] = fmaxf(mb[2], mb[3]);
        }
        // bias + ReLU commute with the max (both monotone)
        const int ch = hh * 128 + qd * 32 + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool emit = shift == 6 ? k == 0 : (shift == 5 ? (k & 1) == 0 : true);
          if (emit && cid[k] >= 0) out[(size_t)cid[k] * C3 + ch] = fmaxf(mb[k] + bias3[hh], 0.f);
        }
      }
      if (warp == 4) { S2_STAMP(4, t, 1) }
    }
  }
#undef S2_STAMP
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, K::TM_COLS);
}

}  // namespace s2v2

extern int g_sa_sms, g_sa_split, g_sa_min_tpc;  // mlp_tc.cu
extern long long* g_sa_trace;                   // sa1_ws2.cu (vnb_debug_sa_trace)
size_t sa_rel_bytes(long long rows);                                                                  // sa_pack.cu
int launch_sa_pack(int total_centroids, const int* pts_cnt, int* hdr, int* tile_cid, cudaStream_t st);  // sa_pack.cu
void launch_group_rel(int n, int m, long long rows, const float* xyz, const float* new_xyz, const int* idx,
                      const int* pts_cnt, void* rel, cudaStream_t st);  // sa_ws.cu

template <int C1, int C2, int C3>
static int s2v2_launch(int b, int n, int m, const float* xyz, const float* new_xyz, const int* idx, const int* pts_cnt,
                       const float* w1x,
                       const float* b2, const float* b3, const void* w2_img, const void* w3_img, const void* q, float* out,
                       void* workspace, cudaStream_t st) {
  using K = s2v2::Cfg<C1, C2, C3>;
  auto kern = s2v2::sa_ws2_kernel<C1, C2, C3>;
  VNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  int dev = 0, sms = 148;
  VNB_CUDA(cudaGetDevice(&dev));
  VNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long rows = (long long)b * m * 64;
  launch_group_rel(n, m, rows, xyz, new_xyz, idx, pts_cnt, workspace, st);
  if (int rc = check_launch("sa_group_mlp_max: grouped relative coordinates")) return rc;
  int* hdr = reinterpret_cast<int*>(static_cast<char*>(workspace) + sa_rel_bytes(rows));
  int* tile_cid = hdr + 64;
  if (int rc = launch_sa_pack(b * m, pts_cnt, hdr, tile_cid, st)) return rc;
  if (g_sa_sms > 0 && g_sa_sms < sms) sms = g_sa_sms;  // leave room for concurrently running FPS CTAs
  sms *= g_sa_split;
  const int cap = b * m / 2 + 3;
  const int grid = sms < cap ? sms : cap;             // one wave: contiguous chunks of the (device-side) tile count
  kern<<<grid, s2v2::THREADS, K::SMEM, st>>>(hdr, tile_cid, static_cast<const float4*>(workspace), w1x, b2, b3,
                                             static_cast<const char*>(w2_img), static_cast<const char*>(w3_img),
                                             static_cast<const __half*>(q), out, g_sa_min_tpc, g_sa_trace);
  return check_launch("sa_group_mlp_max (tcgen05, warp-specialised v2)");
}

// returns -1 when no instance matches
int sa_ws2_dispatch(int b, int n, int m, const float* xyz, const float* new_xyz, const int* idx, const int* pts_cnt,
                    int c1, int c2, int c3,
                    const float* w1x, const float* b2, const float* b3, const void* w2_img, const void* w3_img,
                    const void* q, float* out, void* workspace, cudaStream_t st) {
  if (workspace == nullptr) return -1;
  if (c1 == 128 && c2 == 128 && c3 == 256)
    return s2v2_launch<128, 128, 256>(b, n, m, xyz, new_xyz, idx, pts_cnt, w1x, b2, b3, w2_img, w3_img, q, out, workspace, st);
  if (c1 == 128 && c2 == 128 && c3 == 128)
    return s2v2_launch<128, 128, 128>(b, n, m, xyz, new_xyz, idx, pts_cnt, w1x, b2, b3, w2_img, w3_img, q, out, workspace, st);
  return -1;
}

}  // namespace vnb

// Fused set-abstraction kernel for NARROW inputs (sa1: 3 + c <= 16 input channels), second generation:
// all three layers on the tensor cores, every stage double-buffered, every stage with its OWN issuing warp, and TWO
// warp groups per epilogue stage that alternate tiles.
//
//   group -> [rel_xyz, feat] (K padded to 16) -> L1 -> L2 -> L3 -> max over the 64 samples        (utils.py:49-55,120-132)
//
//   warps 0-3 / 4-7      EPILOGUE1 (even / odd tiles)  D1[s] -> ReLU, fp16 -> H1[t%4]        (thread = tile row = TMEM lane)
//   warps 8-11 / 12-15   EPILOGUE2 (even / odd tiles)  D2[s] -> ReLU, fp16 -> H2[t%4]        (biases b1, b2 ride on the MMAs)
//   warps 16-19 / 20-23  EPILOGUE3 (even / odd tiles)  D3[s] (channel per lane) -> max over each centroid's 64 samples,
//                        +b3, ReLU -> out
//   warps 24-27          PRODUCER   one grouped row per thread: {rel_xyz, source row} from the helper kernel's table, c
//                        feature floats; table rows prefetched two tiles ahead and features one tile ahead so no global
//                        latency is exposed; fp16, swizzled 32-byte row of A0[s]
//   warps 28 / 29 / 30   MMA1/2/3   one thread each: M1(t): D1 = A0 . W1^T, M2(t): D2 = H1 . W2^T, M3(t): D3 = W3^T . H2^T
//                        (transposed: channels on TMEM lanes, so the 64-sample max-pool is a register reduction).
// Why this shape (stage timeline measured with vnb_debug_sa_trace, profiles/):
//   * one in-order MMA issuer kept only ~2 tiles in flight over 7 stages -> one issuer per layer;
//   * an epilogue pass over one tile is a chain of fixed latencies in ONE warp per TMEM lane quadrant (mbarrier try_wait
//     ~90 cycles even when complete, tcgen05.ld ~100, fence + arrive ~100, ...): ~1100 cycles however little math is
//     left -> two warp groups per stage, each owning one of the two buffers (s = tile parity), halve the period;
//   * a warp pulls at most 64 B/cycle out of TMEM (scripts/micro/tmem.cu): the max-pool needs 8 warps;
//   * shared-memory pointers must stay in the shared address space (umma.cuh: smem_align_1024) and loads must be issued
//     before earlier stores in program order (the compiler never hoists LDS over STS).
// TMEM: D1[2] (2 x 64) | D2[2] (2 x 64) | D3[2] (2 x 128) columns; shared memory: A0[2], H1[4], H2[4].  All hand-offs are
// mbarriers (MMA issuers poll, everything else parks).  One CTA per SM, one wave:
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
constexpr int NH = 4;                      // H1 / H2 buffers (shared memory is plentiful here; TMEM allows only 2 x D)
constexpr int OFF_H2 = OFF_H1 + NH * H_BYTES;
constexpr int OFF_ONES = OFF_H2 + NH * H_BYTES;      // A panel [128 rows][k 0..15]: column 0 = 1.0, rest 0  (bias k-step of M2)
constexpr int OFF_B2P = OFF_ONES + 128 * 128;        // B panel [C2 rows][k 0..15]: column 0 = b2[n], rest 0
constexpr int OFF_F = OFF_B2P + C2 * 128;            // floats b3
constexpr int OFF_BAR = OFF_F + C3 * 4;
constexpr int SMEM = OFF_BAR + 40 * 8 + 16 + 1024;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
constexpr int TM_D1 = 0, TM_D2 = 2 * C1, TM_D3 = 2 * C1 + 2 * C2;
constexpr int TM_COLS = 512;
static_assert(TM_D3 + 2 * C3 <= 512, "TMEM budget");
constexpr int THREADS = 31 * 32;
constexpr int TRACE_T = 64;

// fp16x2{lo, hi} = relu(round(.)) in ONE instruction (F2FP.RELU.F16.F32.PACK_AB); rounding is monotone, so relu commutes
__device__ __forceinline__ uint32_t pack_relu2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// BULK (c <= 4): the helper kernel has written the finished fp16 operand rows (group_a0_kernel, sa_pack.cu); a tile's A0 is
// staged by one bulk copy per slot into a ring of NB un-swizzled 2 KB buffers (K columns 0..7), K columns 8..15 — zeros and
// the two constant ones that carry b1 — are ONE constant block every buffer's descriptor points at (its LBO).  No register
// ever holds prefetched data: the register-gathering producer was pinned to one L2 round trip per tile (1235 cycles —
// ptxas multiplexes the look-ahead loads of several tiles onto six scoreboards, so waiting for the oldest waits for all).
constexpr int NB = 4;                       // A0 ring buffers (BULK)
constexpr int A0B_BYTES = 128 * 16;         // one buffer: 128 rows x 8 halves
constexpr int OFF_A0K1 = NB * A0B_BYTES;    // constant K-half block, relative to OFF_A0
constexpr int OFF_CID = OFF_A0K1 + A0B_BYTES;            // this CTA's slice of the tile table, relative to OFF_A0
constexpr int CIDCAP = (2 * A0_BYTES - OFF_CID) / 32;    // tiles cached (the rest is read from global memory)

template <bool BULK>
__global__ void __launch_bounds__(THREADS, 1) sa1_ws2_kernel(int c, const int* __restrict__ hdr,
                                                             const int* __restrict__ tile_cid,
                                                             const float4* __restrict__ rel,
                                                             const float* __restrict__ feat,
                                                             const float* __restrict__ b1, const float* __restrict__ b2,
                                                             const float* __restrict__ b3, const char* __restrict__ w1_img,
                                                             const char* __restrict__ w2_img,
                                                             const char* __restrict__ w3_img, float* __restrict__ out,
                                                             long long* __restrict__ trace, int mma_spin, int min_tpc) {
  // debugging aid: CTA 0 stamps clock64() at the start (after its input wait) and the end of every stage of its first
  // TRACE_T tiles: trace[(role * TRACE_T + t) * 2 + {0,1}], roles P, M1, M2, M3, E1, E2, E3(g=0), E3(g=1)
  const bool tr = trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0;
#define S1_STAMP(role, t, ph) \
  if (tr && (t) < TRACE_T) trace[((role) * TRACE_T + (t)) * 2 + (ph)] = clock64();
#define MMA_WAIT(bar, par) \
  do { if (mma_spin) mbar_wait_spin(bar, par); else mbar_wait(bar, par); } while (0)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* sW1 = smem + OFF_W1;
  uint8_t* sW2 = smem + OFF_W2;
  uint8_t* sW3 = smem + OFF_W3;
  uint8_t* sA0 = smem + OFF_A0;
  uint8_t* sH1 = smem + OFF_H1;
  uint8_t* sH2 = smem + OFF_H2;
  uint8_t* sOnes = smem + OFF_ONES;
  uint8_t* sB2P = smem + OFF_B2P;
  float* sB3 = reinterpret_cast<float*>(smem + OFF_F);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_w = bars;          // weights landed
  uint64_t* a0_full = bars + 1;    // [2]  128 producer arrivals                        (slot t&1, parity (t>>1)&1)
  uint64_t* m1_done = bars + 3;    // [2]  commit
  uint64_t* d1_empty = bars + 5;   // [2]  128 E1 arrivals
  uint64_t* d2_empty = bars + 7;   // [2]  128 E2 arrivals
  uint64_t* d3_empty = bars + 9;   // [2]  128 E3 arrivals
  uint64_t* h1_full = bars + 11;   // [NH] 128 E1 arrivals                              (slot t&3, parity (t>>2)&1)
  uint64_t* m2_done = bars + 15;   // [NH] commit: D2[t&1] ready, H1[t&3] free again
  uint64_t* h2_full = bars + 19;   // [NH] 128 E2 arrivals
  uint64_t* m3_done = bars + 23;   // [NH] commit: D3[t&1] ready, H2[t&3] free again
  uint64_t* a0r_full = bars + 27;  // [NB] BULK: expect_tx arrive of the producer + the bytes of the tile's bulk copies
  uint64_t* a0r_free = bars + 31;  // [NB] BULK: commit: M1 has read the ring buffer
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 36);
  int* sCid = reinterpret_cast<int*>(sA0 + OFF_CID);  // BULK: tile table rows of this CTA's first CIDCAP tiles

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
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a0_full[s], 128); mbar_init(&m1_done[s], 1); mbar_init(&d1_empty[s], 128); mbar_init(&d2_empty[s], 128);
      mbar_init(&d3_empty[s], 128);
    }
    for (int k = 0; k < NH; ++k) {
      mbar_init(&h1_full[k], 128); mbar_init(&m2_done[k], 1); mbar_init(&h2_full[k], 128); mbar_init(&m3_done[k], 1);
    }
    for (int k = 0; k < NB; ++k) { mbar_init(&a0r_full[k], 1); mbar_init(&a0r_free[k], 1); }
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_w, (uint32_t)(W1_BYTES + W2_BYTES + W3_BYTES));
    bulk_g2s(sW1, w1_img, W1_BYTES, bar_w);
    bulk_g2s(sW2, w2_img, W2_BYTES, bar_w);
    bulk_g2s(sW3, w3_img, W3_BYTES, bar_w);
  }
  for (int i = tid; i < C3; i += THREADS) sB3[i] = b3[i];
  // A0 padding columns (k >= 16 of the 64-column panel are never read; k in [3+c,14) must be finite): zero both buffers;
  // zero the two bias panels
  for (int i = tid; i < 2 * A0_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sA0)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (128 * 128 + C2 * 128) / 16; i += THREADS) reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0, 0, 0, 0);
  if (BULK) {
    __syncthreads();  // the zero fill above covers the region the two tables below live in
    // K columns 8..15 of every tile row: zeros, and k = 14, 15 = the constant 1 that multiplies b1 (hi, lo)
    if (tid < 128) reinterpret_cast<uint4*>(sA0 + OFF_A0K1)[tid] = make_uint4(0u, 0u, 0u, pack2(1.f, 1.f));
    const int ncache = min(my_tiles, CIDCAP) * 8;
    for (int i = tid; i < ncache; i += THREADS) sCid[i] = __ldg(tile_cid + (size_t)first_tile * 8 + i);
  }
  if (warp == 0) tmem_alloc(tmem_ptr, TM_COLS);
  // No thread may touch an mbarrier before thread 0 has initialised it: the previous CTA on this SM (possibly another
  // kernel) leaves arbitrary bytes there, and re-initialising a barrier other threads are parked on is undefined — seen
  // as sporadic launch failures when different fused kernels alternate on an SM (scripts/gpu_stress.py, mix "sa_all").
  __syncthreads();
  mbar_wait(bar_w, 0);  // weight images have landed (every thread observes the barrier: the patch below follows the copy)
  __syncthreads();
  // Biases of layers 1 and 2 ride on the tensor cores (an epilogue pass is ALU-pipe bound, profiles/micro_thr.txt):
  //   layer 1: A0 columns 14, 15 are the constant 1 (written by the producer), rows k = 14, 15 of the W1 image are b1;
  //   layer 2: one extra k-step multiplies the constant panel sOnes (columns 0, 1 = 1) with sB2P (columns 0, 1 = b2).
  // fp16 biases, fp32 accumulation.  Layer 3's bias is added after the max-pool (one add per output).
  // Each bias is split into fp16 hi + lo parts on two columns (~22 significant bits).
  if (tid < 128)
    *reinterpret_cast<__half2*>(sOnes + sw128_offset((uint32_t)tid, 0)) = __floats2half2_rn(1.f, 1.f);
  if (tid < C2) {
    const float bv = b2[tid];
    const __half hi = __float2half_rn(bv);
    *reinterpret_cast<__half2*>(sB2P + sw128_offset((uint32_t)tid, 0)) = __halves2half2(hi, __float2half_rn(bv - __half2float(hi)));
  }
  if (tid < C1) {
    const float bv = b1[tid];
    const __half hi = __float2half_rn(bv);
    *reinterpret_cast<__half2*>(sW1 + sw128_offset((uint32_t)tid, 14)) = __halves2half2(hi, __float2half_rn(bv - __half2float(hi)));
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  // use number t>>1 of a [2]-slotted barrier -> parity (t>>1)&1; use number t>>2 of an [NH]-slotted one -> (t>>2)&1
  auto par_of = [](int t) { return (uint32_t)((t >> 1) & 1); };
  auto par4 = [](int t) { return (uint32_t)((t >> 2) & 1); };

  if (BULK && warp >= 24 && warp < 28) {
    // ================================================================ PRODUCER (BULK): one warp, one bulk copy per slot
    // lane g < (slots of the tile) copies the rows of the centroid in slot g: (1 << shift) rows x 16 bytes, contiguous in
    // the helper's table; lane 0 announces the tile's byte count on the ring buffer's barrier first.
    if (warp == 24) {
      const uint4* table = reinterpret_cast<const uint4*>(rel);
      for (int t = 0; t < my_tiles; ++t) {
        const int slot = t & (NB - 1);
        const int shift = shift_of(first_tile + t);
        const int nslot = 128 >> shift;
        int cid = -1;
        if (lane < nslot) cid = t < CIDCAP ? sCid[t * 8 + lane] : __ldg(tile_cid + (size_t)(first_tile + t) * 8 + lane);
        const unsigned mask = __ballot_sync(0xffffffffu, cid >= 0);
        if (t >= NB) mbar_wait(&a0r_free[slot], (uint32_t)(((t / NB) - 1) & 1));  // M1(t - NB) has read the buffer
        S1_STAMP(0, t, 0)
        const uint32_t bytes = 16u << shift;
        if (lane == 0) mbar_arrive_expect_tx(&a0r_full[slot], (uint32_t)__popc(mask) * bytes);
        __syncwarp();
        if (cid >= 0)
          bulk_g2s(sA0 + slot * A0B_BYTES + lane * bytes, table + (size_t)cid * 64, bytes, &a0r_full[slot]);
        S1_STAMP(0, t, 1)
      }
    }
  } else if (warp >= 24 && warp < 28) {
    // ================================================================ PRODUCER: one grouped row per thread
    const int pt = tid - 768;
    // tile row pt = sample (pt mod slot) of the centroid in slot (pt / slot); empty slots read row 0 (discarded later)
    auto cid_of = [&](int tile) { return __ldg(tile_cid + (size_t)tile * 8 + (pt >> shift_of(tile))); };
    auto src_row = [&](int tile, int cid) {
      return cid < 0 ? (size_t)0 : (size_t)cid * 64 + (size_t)(pt & ((1 << shift_of(tile)) - 1));
    };
    {  // (c <= 4 takes the BULK instance)
    float4 r_cur = make_float4(0.f, 0.f, 0.f, 0.f), r_nxt = r_cur;   // table rows of tiles t, t+1
    if (my_tiles > 0) r_cur = __ldg(rel + src_row(first_tile, cid_of(first_tile)));
    if (my_tiles > 1) r_nxt = __ldg(rel + src_row(first_tile + 1, cid_of(first_tile + 1)));
    float f_cur[11], f_nxt[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) f_cur[i] = f_nxt[i] = 0.f;
    if (my_tiles > 0) {
      const float* f = feat + (size_t)__float_as_int(r_cur.w) * c;
#pragma unroll
      for (int i = 0; i < 11; ++i)
        if (i < c) f_cur[i] = __ldg(f + i);
    }
    for (int t = 0; t < my_tiles; ++t) {
      const int s = t & 1;
      // prefetch: features of tile t+1 (its table row arrived an iteration ago), table row of tile t+2
      float4 r_nn = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t + 1 < my_tiles) {
        const float* f = feat + (size_t)__float_as_int(r_nxt.w) * c;
#pragma unroll
        for (int i = 0; i < 11; ++i)
          if (i < c) f_nxt[i] = __ldg(f + i);
      }
      if (t + 2 < my_tiles) r_nn = __ldg(rel + src_row(first_tile + t + 2, cid_of(first_tile + t + 2)));
      if (t >= 2) mbar_wait(&m1_done[s], par_of(t - 2));  // M1(t-2) finished reading A0[s]
      if (warp == 24) { S1_STAMP(0, t, 0) }
      uint8_t* a0 = sA0 + s * A0_BYTES;
      *reinterpret_cast<uint4*>(a0 + sw128_offset((uint32_t)pt, 0)) =
          make_uint4(pack2(r_cur.x, r_cur.y), pack2(r_cur.z, f_cur[0]), pack2(f_cur[1], f_cur[2]), pack2(f_cur[3], f_cur[4]));
      *reinterpret_cast<uint4*>(a0 + sw128_offset((uint32_t)pt, 8)) =
          make_uint4(pack2(f_cur[5], f_cur[6]), pack2(f_cur[7], f_cur[8]), pack2(f_cur[9], f_cur[10]),
                     pack2(1.f, 1.f));  // k = 14, 15: the constant 1 that multiplies b1 (hi, lo); needs c <= 11
      fence_proxy_async_smem();
      mbar_arrive(&a0_full[s]);
      if (warp == 24) { S1_STAMP(0, t, 1) }
      r_cur = r_nxt; r_nxt = r_nn;
#pragma unroll
      for (int i = 0; i < 11; ++i) f_cur[i] = f_nxt[i];
    }
    }
  } else if (warp == 28) {
    // ================================================================ MMA1: D1[s] = A0[s] . W1^T  (single K = 16 step)
    // (all 32 lanes run the loop; one elected lane issues — see umma.cuh: elect_one)
    if (my_tiles > 0) {
      const uint32_t id1 = make_idesc_f16_f32(128, C1);
      const uint64_t bd = make_desc_sw128(smem_u32(sW1));
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t & 1;
        const int slot = t & (NB - 1);
        if (BULK) MMA_WAIT(&a0r_full[slot], (uint32_t)((t / NB) & 1)); else MMA_WAIT(&a0_full[s], par_of(t));
        if (t >= 2) MMA_WAIT(&d1_empty[s], par_of(t - 2));
        tc_fence_after_sync();
        S1_STAMP(1, t, 0)
        if (elect_one()) {
          // BULK: un-swizzled A (8-row groups 128 B apart; the K columns 8..15 are the shared constant block)
          const uint64_t ad = BULK ? make_desc_nosw(smem_u32(sA0 + slot * A0B_BYTES), (uint32_t)(OFF_A0K1 - slot * A0B_BYTES), 128u)
                                   : make_desc_sw128(smem_u32(sA0 + s * A0_BYTES));
          mma_f16_ss(tmem + TM_D1 + s * C1, ad, bd, id1, 0u);
          mma_commit(&m1_done[s]);
          if (BULK) mma_commit(&a0r_free[slot]);
        }
        __syncwarp();
        S1_STAMP(1, t, 1)
      }
    }
  } else if (warp == 29) {
    // ================================================================ MMA2: D2[s] = H1[t%4] . W2^T  (+ b2)
    if (my_tiles > 0) {
      const uint32_t id2 = make_idesc_f16_f32(128, C2);
      const uint64_t bd = make_desc_sw128(smem_u32(sW2));
      const uint64_t ones = make_desc_sw128(smem_u32(sOnes)), bias = make_desc_sw128(smem_u32(sB2P));
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t & 1;
        MMA_WAIT(&h1_full[t & 3], par4(t));
        if (t >= 2) MMA_WAIT(&d2_empty[s], par_of(t - 2));
        tc_fence_after_sync();
        S1_STAMP(2, t, 0)
        if (elect_one()) {
          const uint64_t ad = make_desc_sw128(smem_u32(sH1 + (t & 3) * H_BYTES));
#pragma unroll
          for (int ks = 0; ks < C1 / 16; ++ks)  // +32 bytes per k-step == +2 in the descriptor's address field
            mma_f16_ss(tmem + TM_D2 + s * C2, ad + (uint64_t)(2 * ks), bd + (uint64_t)(2 * ks), id2, ks > 0 ? 1u : 0u);
          mma_f16_ss(tmem + TM_D2 + s * C2, ones, bias, id2, 1u);  // + b2
          mma_commit(&m2_done[t & 3]);
        }
        __syncwarp();
        S1_STAMP(2, t, 1)
      }
    }
  } else if (warp == 30) {
    // ================================================================ MMA3: D3[s] = W3^T . H2[t%4]^T
    if (my_tiles > 0) {
      const uint32_t id3 = make_idesc_f16_f32(128, 128);
      const uint64_t ad = make_desc_sw128(smem_u32(sW3));
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t & 1;
        MMA_WAIT(&h2_full[t & 3], par4(t));
        if (t >= 2) MMA_WAIT(&d3_empty[s], par_of(t - 2));
        tc_fence_after_sync();
        S1_STAMP(3, t, 0)
        if (elect_one()) {
          const uint64_t bd = make_desc_sw128(smem_u32(sH2 + (t & 3) * H_BYTES));
#pragma unroll
          for (int ks = 0; ks < C2 / 16; ++ks)
            mma_f16_ss(tmem + TM_D3 + s * C3, ad + (uint64_t)(2 * ks), bd + (uint64_t)(2 * ks), id3, ks > 0 ? 1u : 0u);
          mma_commit(&m3_done[t & 3]);
        }
        __syncwarp();
        S1_STAMP(3, t, 1)
      }
    }
  } else if (warp < 16) {
    // ================================================================ EPILOGUE 1 (warps 0-7) / EPILOGUE 2 (warps 8-15)
    // warp group (warp >> 2) & 1 owns the tiles of that parity, i.e. buffer s of every stage
    const bool e2 = warp >= 8;
    const int s = (warp >> 2) & 1;
    const int et = tid & 127;  // tile row == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint64_t* d_empty = (e2 ? d2_empty : d1_empty) + s;
    const uint32_t tacc = tmem + lane_base + (e2 ? TM_D2 : TM_D1) + s * 64;
    uint8_t* hbase = sH1 + (e2 ? NH * H_BYTES : 0);   // sH2 follows sH1; pointers stay provably shared-memory
    const bool stamp = (warp & 3) == 0;
    const int role = e2 ? (s ? 9 : 5) : (s ? 8 : 4);  // trace roles: E1 even 4 / odd 8, E2 even 5 / odd 9
    for (int t = s; t < my_tiles; t += 2) {
      // accumulator ready: E1 <- m1_done[s], E2 <- m2_done[t&3]
      if (e2) mbar_wait(&m2_done[t & 3], par4(t)); else mbar_wait(&m1_done[s], par_of(t));
      tc_fence_after_sync();
      if (stamp) { S1_STAMP(role, t, 0) }
      uint8_t* h = hbase + (t & 3) * H_BYTES;
      // the MMA that read our output buffer NH tiles ago: E1 <- m2_done, E2 <- m3_done
      uint64_t* out_free = (e2 ? m3_done : m2_done) + (t & 3);
      uint64_t* h_full = (e2 ? h2_full : h1_full) + (t & 3);
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {  // four 16-column quarters (ptxas hoists the next load; x32 loads overflow 64 registers)
        uint32_t v[16];
        tmem_ld_x16(tacc + qq * 16, v);
        if (qq == 0 && t >= NH) mbar_wait(out_free, par4(t - NH));
        tmem_ld_wait();
        if (qq == 3) {
          tc_fence_before_sync();
          mbar_arrive(d_empty);  // accumulator drained into registers
        }
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {  // bias already in the accumulator: relu + fp16 pack is one instruction per pair
          const uint4 pk = make_uint4(pack_relu2(__uint_as_float(v[gi * 8 + 0]), __uint_as_float(v[gi * 8 + 1])),
                                      pack_relu2(__uint_as_float(v[gi * 8 + 2]), __uint_as_float(v[gi * 8 + 3])),
                                      pack_relu2(__uint_as_float(v[gi * 8 + 4]), __uint_as_float(v[gi * 8 + 5])),
                                      pack_relu2(__uint_as_float(v[gi * 8 + 6]), __uint_as_float(v[gi * 8 + 7])));
          *reinterpret_cast<uint4*>(h + sw128_offset((uint32_t)et, (uint32_t)(qq * 16 + gi * 8))) = pk;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(h_full);
      if (stamp) { S1_STAMP(role, t, 1) }
    }
  } else if (warp < 24) {
    // ================================================================ EPILOGUE 3: D3 -> max-pool -> out
    // warp (16 + 4*s + q): tiles of parity s, TMEM lane quadrant q (channels 32q..32q+31), both centroids of the tile.
    // (Splitting every tile between the two groups by column half instead halves the accumulator hold time but makes
    // every warp pay its fixed per-tile cost on every tile: tile period 1100 instead of 905 cycles, measured.)
    const int q = warp & 3, s = (warp >> 2) & 1;
    const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + TM_D3 + s * C3;
    const int ch = q * 32 + lane;
    const float bias3 = sB3[ch];
    for (int t = s; t < my_tiles; t += 2) {
      const int tile = first_tile + t;
      const int shift = shift_of(tile);
      // the tile's eight slot entries, requested BEFORE the wait for the accumulator so the (L2) latency hides behind it
      // (BULK: from the CTA's shared-memory copy of the table — when this stage paces the kernel the wait returns at once)
      int4 ca, cb;
      if (BULK && t < CIDCAP) {
        ca = reinterpret_cast<const int4*>(sCid + t * 8)[0];
        cb = reinterpret_cast<const int4*>(sCid + t * 8)[1];
      } else {
        ca = __ldg(reinterpret_cast<const int4*>(tile_cid + (size_t)tile * 8));
        cb = __ldg(reinterpret_cast<const int4*>(tile_cid + (size_t)tile * 8) + 1);
      }
      mbar_wait(&m3_done[t & 3], par4(t));
      tc_fence_after_sync();
      if (q == 0) { S1_STAMP(6 + s, t, 0) }
      // eight 16-column blocks; the load of block k+1 is in flight while block k is reduced (two 16-register buffers —
      // left to itself ptxas stopped hoisting the next load once the per-slot reduction was added, and eight exposed
      // TMEM round trips made this stage pace the whole kernel)
      float mbv[8];
      {
        uint32_t v[2][16];
        tmem_ld_x16(tacc, v[0]);
#pragma unroll
        for (int blk = 0; blk < 8; ++blk) {
          tmem_ld_wait();
          if (blk + 1 < 8) {
            tmem_ld_x16(tacc + (blk + 1) * 16, v[(blk + 1) & 1]);
          } else {
            tc_fence_before_sync();
            mbar_arrive(&d3_empty[s]);  // D3[s] has been read
          }
          const uint32_t* w = v[blk & 1];
          float a0 = fmaxf(__uint_as_float(w[0]), __uint_as_float(w[4])), a1 = fmaxf(__uint_as_float(w[1]), __uint_as_float(w[5]));
          float a2 = fmaxf(__uint_as_float(w[2]), __uint_as_float(w[6])), a3 = fmaxf(__uint_as_float(w[3]), __uint_as_float(w[7]));
          a0 = fmaxf(fmaxf(a0, __uint_as_float(w[8])), __uint_as_float(w[12]));
          a1 = fmaxf(fmaxf(a1, __uint_as_float(w[9])), __uint_as_float(w[13]));
          a2 = fmaxf(fmaxf(a2, __uint_as_float(w[10])), __uint_as_float(w[14]));
          a3 = fmaxf(fmaxf(a3, __uint_as_float(w[11])), __uint_as_float(w[15]));
          mbv[blk] = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        }
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        // centroids whose samples sit in this half of the tile: one (64-row slots), two (32) or four (16)
        int cid[4];
        if (shift == 6) {
          cid[0] = g == 0 ? ca.x : ca.y; cid[1] = cid[2] = cid[3] = -1;
        } else if (shift == 5) {
          cid[0] = g == 0 ? ca.x : ca.z; cid[1] = g == 0 ? ca.y : ca.w; cid[2] = cid[3] = -1;
        } else {
          cid[0] = g == 0 ? ca.x : cb.x; cid[1] = g == 0 ? ca.y : cb.y; cid[2] = g == 0 ? ca.z : cb.z; cid[3] = g == 0 ? ca.w : cb.w;
        }
        const float* mb = &mbv[4 * g];
        // bias + ReLU commute with the max
        if (shift == 6) {
          if (cid[0] >= 0) out[(size_t)cid[0] * C3 + ch] = fmaxf(fmaxf(fmaxf(mb[0], mb[1]), fmaxf(mb[2], mb[3])) + bias3, 0.f);
        } else if (shift == 5) {
          if (cid[0] >= 0) out[(size_t)cid[0] * C3 + ch] = fmaxf(fmaxf(mb[0], mb[1]) + bias3, 0.f);
          if (cid[1] >= 0) out[(size_t)cid[1] * C3 + ch] = fmaxf(fmaxf(mb[2], mb[3]) + bias3, 0.f);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (cid[k] >= 0) out[(size_t)cid[k] * C3 + ch] = fmaxf(mb[k] + bias3, 0.f);
        }
      }
      if (q == 0) { S1_STAMP(6 + s, t, 1) }
    }
  }
#undef S1_STAMP
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace s1v2

extern int g_sa_variant;  // mlp_tc.cu
long long* g_sa_trace = nullptr;  // debugging: device buffer of 12 x 64 x 2 int64 (vnb_debug_sa_trace)

extern int g_sa_sms, g_sa_split, g_sa_min_tpc;  // mlp_tc.cu
size_t sa_rel_bytes(long long rows);                                                                  // sa_pack.cu
int launch_sa_pack(int total_centroids, const int* pts_cnt, int* hdr, int* tile_cid, cudaStream_t st);  // sa_pack.cu
void launch_group_rel(int n, int m, long long rows, const float* xyz, const float* new_xyz, const int* idx,
                      const int* pts_cnt, void* rel, cudaStream_t st);  // sa_pack.cu
void launch_group_a0(int n, int m, int c, long long rows, const float* xyz, const float* feat, const float* new_xyz,
                     const int* idx, const int* pts_cnt, void* a0, cudaStream_t st);  // sa_pack.cu

// returns -1 when no instance matches
int sa1_ws2_dispatch(int b, int n, int c, int m, const float* xyz, const float* feat, const float* new_xyz, const int* idx,
                     const int* pts_cnt, int c1, int c2, int c3, const float* b1, const float* b2, const float* b3, const void* w1_img,
                     const void* w2_img, const void* w3_img, float* out, void* workspace, cudaStream_t st) {
  if (workspace == nullptr || !(c1 == 64 && c2 == 64 && c3 == 128) || c > 11) return -1;  // k = 14, 15 carry the bias
  const bool bulk = c <= 4;   // the helper writes finished fp16 operand rows; the kernel stages tiles by bulk copies
  auto kern = bulk ? s1v2::sa1_ws2_kernel<true> : s1v2::sa1_ws2_kernel<false>;
  VNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, s1v2::SMEM));
  int dev = 0, sms = 148;
  VNB_CUDA(cudaGetDevice(&dev));
  VNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long rows = (long long)b * m * 64;
  if (bulk) launch_group_a0(n, m, c, rows, xyz, feat, new_xyz, idx, pts_cnt, workspace, st);
  else launch_group_rel(n, m, rows, xyz, new_xyz, idx, pts_cnt, workspace, st);
  if (int rc = check_launch("sa_group_mlp_max: grouped relative coordinates")) return rc;
  int* hdr = reinterpret_cast<int*>(static_cast<char*>(workspace) + sa_rel_bytes(rows));
  int* tile_cid = hdr + 64;
  if (int rc = launch_sa_pack(b * m, pts_cnt, hdr, tile_cid, st)) return rc;
  if (g_sa_sms > 0 && g_sa_sms < sms) sms = g_sa_sms;  // leave room for concurrently running FPS CTAs
  sms *= g_sa_split;
  const int cap = b * m / 2 + 3;
  const int grid = sms < cap ? sms : cap;             // one wave: contiguous chunks of the (device-side) tile count
  kern<<<grid, s1v2::THREADS, s1v2::SMEM, st>>>(c, hdr, tile_cid, static_cast<const float4*>(workspace), feat, b1, b2, b3,
                                                static_cast<const char*>(w1_img), static_cast<const char*>(w2_img),
                                                static_cast<const char*>(w3_img), out, g_sa_trace, g_sa_variant == 3 ? 1 : 0, g_sa_min_tpc);  // MMA issuers park (default) or poll (sa_variant 3)
  return check_launch("sa_group_mlp_max (tcgen05, warp-specialised v2, narrow input)");
}

}  // namespace vnb

// Fused feature-propagation module (+ optionally the voting module) on the tensor cores: ONE kernel per FP level.
//
//   pointnet_fp_module (reference utils.py:266-294):  w = (1/max(d,1e-10)) / sum, interpolated = sum_i w_i points2[idx_i],
//       x = concat[interpolated (256), points1 (256)] -> 1x1 conv + BN + ReLU (512 -> 256) -> (256 -> 256)
//   voting module (reference model.py:53-61), fused behind the LAST fp level:
//       seeds = concat[seeds_xyz (3), seeds_feat (256)] -> FC+BN+ReLU (259 -> 256) -> (256 -> 256) -> FC (256 -> 259)
//       votes = seeds + offset; votes_xyz = votes[:, :3], votes_feat = votes[:, 3:]
//
// The unfused path is 2 + 7 launches (interpolate/concat, 2 + 2 + 3 linears, concat, split) that stream ~300 MB of fp32
// activations through HBM per forward and hold an SM per 128 x 128 tile for ~10 us of latency each — 89 us of a 470 us
// step with twelve forwards in flight (scripts/gpu_stress.py).  Every row is independent, so here a CTA owns 128 rows
// and walks them through ALL layers; activations never leave the SM:
//   * layer 0 streams its A operand: the 256 loader threads interpolate / copy one 64-column chunk at a time straight
//     into a 2-slot ring of 128-byte-swizzled K-major panels (the concat tensor is never materialised);
//   * every layer's output goes TMEM -> registers (bias, ReLU, fp16) -> a resident panel buffer that is the next layer's
//     A operand (two 64 KB buffers, ping-pong; fp16 is what the unfused path feeds its tensor cores with as well);
//   * weights stream from L2 through a 5-slot ring of 16 KB sub-chunks (64 k-columns x 128 output columns of the
//     pre-swizzled image, one cp.async.bulk each), issued by a dedicated warp that runs ahead across layer boundaries;
//   * one elected thread issues tcgen05.mma (M 128, N 128 / 16, K 16), accumulators in TMEM columns 0..271;
//   * the 3 xyz input columns of the first vote layer are a rank-3 fp32 update in the epilogue (K stays 256, exact in
//     real arithmetic), the last vote layer is issued with its 259 output columns permuted [features | xyz] so that the
//     epilogue adds the residual and writes votes_feat / votes_xyz directly (no concat, no split).
// Warps 0-7: loaders, then epilogues (TMEM lane quadrant = warp & 3, column half = warp >> 2); warp 8: weight ring;
// warp 9: MMA issue.  The two single-role warps POLL their mbarriers (a parked try_wait wakes ~300 cycles after the
// arrive, umma.cuh, and every one of the ~50 weight sub-chunks of a tile would pay that on the critical path).
#include "common.cuh"
#include "umma.cuh"

namespace vnb {

using namespace umma;

namespace fpc {

constexpr int THREADS = 10 * 32;
constexpr int WORKERS = 256;            // loader / epilogue threads
constexpr int CW = 256;                 // width of every hidden layer, of points1 and of points2
constexpr int PANEL = 128 * 128;        // bytes: 128 rows x 64 fp16 columns, SW128 K-major
constexpr int NSTAGE = 5;               // weight ring slots (16 KB each)
constexpr int MAX_LAYERS = 5;
constexpr int OFF_X = 0;                // 4 panels
constexpr int OFF_Y = 4 * PANEL;        // 4 panels; panels 0, 1 double as the layer-0 A ring (Y is first written by layer 1)
constexpr int OFF_W = 8 * PANEL;
constexpr int OFF_BAR = OFF_W + NSTAGE * PANEL;
constexpr int BIAS_LD = 272;            // floats per layer
constexpr int OFF_BIAS = OFF_BAR + 512; // float [MAX_LAYERS][BIAS_LD] biases | float [3][CW] xyz rows of the first vote layer
constexpr int SMEM = OFF_BIAS + (MAX_LAYERS * BIAS_LD + 3 * CW) * 4 + 1024;
static_assert(SMEM <= 227 * 1024, "shared memory budget");
constexpr int TM_COLS = 512;
constexpr int TM_RES = 272;             // TMEM columns TM_RES..511: the vote residual (fp32 output of the fp module), parked
constexpr int RES_TM_COLS = TM_COLS - TM_RES;   // 240 of its 256 columns fit behind the 272 accumulator columns

struct Params {
  int rows_total, n, m;                  // rows_total = b * n unknown points, m known points per cloud
  const float* dist;                     // (b, n, 3)
  const int* idx;                        // (b, n, 3)
  const float* points1;                  // (b, n, 256) skip features
  const float* points2;                  // (b, m, 256) known features
  int n_layers;                          // 2 (fp module) or 5 (fp module + vote module)
  const char* w_img[MAX_LAYERS];         // packed fp16 images; layer 2's image holds rows 3.. of its weight (K = 256);
                                         // layer 4's image has its output columns permuted [3..258, 0..2]
  const float* bias[MAX_LAYERS];         // layer 4's bias permuted like its columns
  int k_pad[MAX_LAYERS], n_pad[MAX_LAYERS], n_out[MAX_LAYERS];
  const float* seeds_xyz;                // (rows, 3)                       (vote only)
  const float* w_vote_xyz;               // (3, 256) rows 0..2 of layer 2's weight, fp32   (vote only)
  float* fp_out;                         // (rows, 256) output of the fp module (= seed features), fp32
  float* votes_xyz;                      // (rows, 3)                       (vote only)
  float* votes_feat;                     // (rows, 256)                     (vote only)
  long long* trace;                      // debugging (vnb_debug_sa_trace): CTA 0 stamps clock64() per stage, see FP_STAMP
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_relu_h2(float lo, float hi) {  // fp16x2(relu(.)) in one F2FP.RELU
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__global__ void __launch_bounds__(THREADS, 1) fp_chain_kernel(const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* sX = smem + OFF_X;
  uint8_t* sY = smem + OFF_Y;
  uint8_t* sW = smem + OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;                 // [NSTAGE] weight sub-chunk landed (tx)
  uint64_t* w_empty = bars + NSTAGE;       // [NSTAGE] commit: the MMAs reading the slot have completed
  uint64_t* a_full = bars + 2 * NSTAGE;    // [2] 256 loader arrivals: layer-0 A chunk written
  uint64_t* a_empty = bars + 2 * NSTAGE + 2;  // [2] commit: the MMAs reading the A slot have completed
  uint64_t* acc_full = bars + 2 * NSTAGE + 4;  // commit: all MMAs of the current layer completed (one phase per layer)
  uint64_t* act_full = bars + 2 * NSTAGE + 5;  // 256 arrivals: the epilogue has written the next layer's A operand
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 8);
  float* sBias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* sWxyz = sBias + MAX_LAYERS * BIAS_LD;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = (int)blockIdx.x * 128;
  const int L = P.n_layers;
  // stage trace of CTA 0: slot 0 kernel start; 1 prologue done; 8 + 2c / 9 + 2c loader chunk c loads issued / stored;
  // 32 + 4l .. : layer l first MMA issued, last MMA issued, accumulator seen by the epilogue, epilogue done
  const bool tr = P.trace != nullptr && blockIdx.x == 0 && lane == 0;
#define FP_STAMP(slot) if (tr) P.trace[slot] = clock64();
  if (warp == 0) { FP_STAMP(0) }
  // Interpolation indices / weights of this thread's four rows (loader role: 8 threads per row, 4 rows per thread),
  // requested before the prologue: the first chunk of layer 0 waits on a chain of two global round trips (idx -> points2),
  // the first of them now overlaps the barrier / bias / TMEM set-up.
  const int c8 = tid & 7, rsub = tid >> 3;
  int gi[4][3];
  float gw[4][3];
  bool ok[4];
  const float* p2base[4];
  if (warp < 8) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int gr = row0 + rsub + 32 * p;
      ok[p] = gr < P.rows_total;
      const int r = ok[p] ? gr : 0;
      const float d1 = fmaxf(__ldg(P.dist + (size_t)r * 3), 1e-10f), d2 = fmaxf(__ldg(P.dist + (size_t)r * 3 + 1), 1e-10f),
                  d3 = fmaxf(__ldg(P.dist + (size_t)r * 3 + 2), 1e-10f);  // utils.py:279
      const float r1 = __fdiv_rn(1.0f, d1), r2 = __fdiv_rn(1.0f, d2), r3 = __fdiv_rn(1.0f, d3);
      const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);  // utils.py:280-282
      gw[p][0] = __fdiv_rn(r1, norm); gw[p][1] = __fdiv_rn(r2, norm); gw[p][2] = __fdiv_rn(r3, norm);
#pragma unroll
      for (int i = 0; i < 3; ++i) gi[p][i] = __ldg(P.idx + (size_t)r * 3 + i);
      p2base[p] = P.points2 + (size_t)(r / P.n) * P.m * CW;
    }
  }

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], WORKERS); mbar_init(&a_empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(act_full, WORKERS);
    fence_barrier_init();
  }
  // biases (and the rank-3 xyz weights) are read by every epilogue thread for every column: stage them once
  for (int l = 0; l < L; ++l)
    for (int i = tid; i < P.n_out[l]; i += THREADS) sBias[l * BIAS_LD + i] = __ldg(P.bias[l] + i);
  if (L > 2)
    for (int i = tid; i < 3 * CW; i += THREADS) sWxyz[i] = __ldg(P.w_vote_xyz + i);
  if (warp == 0) tmem_alloc(tmem_ptr, TM_COLS);
  tc_fence_before_sync();
  __syncthreads();  // barriers initialised before anyone touches them
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  if (warp == 0) { FP_STAMP(1) }

  if (warp == 8) {
    // ================================================================ weight ring: all sub-chunks of all layers, in order
    if (lane == 0) {
      int g = 0;
      for (int l = 0; l < L; ++l) {
        const int nch = (P.k_pad[l] + 63) >> 6, npad = P.n_pad[l];
        const int npc = (npad + 127) >> 7;
        for (int c = 0; c < nch; ++c)
          for (int j = 0; j < npc; ++j, ++g) {
            const int s = g % NSTAGE;
            if (g >= NSTAGE) mbar_wait_spin(&w_empty[s], (uint32_t)((g / NSTAGE - 1) & 1));
            const int rws = min(128, npad - 128 * j);
            mbar_arrive_expect_tx(&w_full[s], (uint32_t)(rws * 128));
            bulk_g2s(sW + s * PANEL, P.w_img[l] + ((size_t)c * npad + 128 * j) * 128, (uint32_t)(rws * 128), &w_full[s]);
          }
      }
    }
  } else if (warp == 9) {
    // ================================================================ MMA issue (all lanes loop, one elected lane issues)
    int g = 0;
    for (int l = 0; l < L; ++l) {
      const int nch = (P.k_pad[l] + 63) >> 6, npad = P.n_pad[l];
      const int npc = (npad + 127) >> 7;
      // layer l >= 1 reads the resident buffer the epilogue of layer l-1 wrote: X after even layers, Y after odd ones
      const uint32_t abase = smem_u32((l & 1) ? sX : sY);
      if (l >= 1) {
        mbar_wait_spin(act_full, (uint32_t)((l - 1) & 1));
        tc_fence_after_sync();
      }
      for (int c = 0; c < nch; ++c) {
        const int kc = min(64, P.k_pad[l] - 64 * c);
        uint32_t a0;
        if (l == 0) {
          mbar_wait_spin(&a_full[c & 1], (uint32_t)((c >> 1) & 1));
          tc_fence_after_sync();
          a0 = smem_u32(sY) + (uint32_t)(c & 1) * PANEL;
        } else {
          a0 = abase + (uint32_t)c * PANEL;
        }
        for (int j = 0; j < npc; ++j, ++g) {
          const int s = g % NSTAGE;
          mbar_wait_spin(&w_full[s], (uint32_t)((g / NSTAGE) & 1));
          tc_fence_after_sync();
          const int rws = min(128, npad - 128 * j);
          const uint32_t idesc = make_idesc_f16_f32(128, (uint32_t)rws);
          const uint32_t b0 = smem_u32(sW) + (uint32_t)s * PANEL;
          if (c == 0 && j == 0) { FP_STAMP(32 + 4 * l) }
          if (elect_one()) {
            for (int ks = 0; ks < kc / 16; ++ks)
              mma_f16_ss(tmem + (uint32_t)(128 * j), make_desc_sw128(a0 + (uint32_t)ks * 32),
                         make_desc_sw128(b0 + (uint32_t)ks * 32), idesc, (c > 0 || ks > 0) ? 1u : 0u);
            mma_commit(&w_empty[s]);
            if (l == 0 && j == npc - 1) mma_commit(&a_empty[c & 1]);
            if (c == nch - 1 && j == npc - 1) mma_commit(acc_full);
          }
          __syncwarp();
          if (c == nch - 1 && j == npc - 1) { FP_STAMP(33 + 4 * l) }
        }
      }
    }
  } else {
    // ================================================================ workers: layer-0 loader, then the epilogues
    // ---- layer 0: [interpolated (256) | points1 (256)] in 64-column chunks; 8 threads per row, 4 rows per thread
    {
      for (int c = 0; c < 8; ++c) {
        const int s = c & 1;
        uint4 pk[4];
        if (c < 4) {  // interpolated columns 64c + 8 c8 .. + 8   (tf_interpolate.cpp:107-127)
          const int col = 64 * c + 8 * c8;
          f32x8 v[4][3];  // all 12 independent 32-byte loads of the chunk in flight before the first use
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int i = 0; i < 3; ++i) v[p][i] = ldg_f32x8(p2base[p] + (size_t)gi[p][i] * CW + col);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              o[e] = __fadd_rn(__fadd_rn(__fmul_rn(v[p][0].v[e], gw[p][0]), __fmul_rn(v[p][1].v[e], gw[p][1])),
                               __fmul_rn(v[p][2].v[e], gw[p][2]));
            pk[p] = ok[p] ? make_uint4(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]), pack_h2(o[4], o[5]), pack_h2(o[6], o[7]))
                          : make_uint4(0, 0, 0, 0);
          }
        } else {  // skip features, columns 64 (c-4) + 8 c8 .. + 8 of points1
          const int col = 64 * (c - 4) + 8 * c8;
          f32x8 v[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int gr = ok[p] ? row0 + rsub + 32 * p : 0;
            v[p] = ldg_f32x8(P.points1 + (size_t)gr * CW + col);
          }
#pragma unroll
          for (int p = 0; p < 4; ++p)
            pk[p] = ok[p] ? make_uint4(pack_h2(v[p].v[0], v[p].v[1]), pack_h2(v[p].v[2], v[p].v[3]),
                                       pack_h2(v[p].v[4], v[p].v[5]), pack_h2(v[p].v[6], v[p].v[7]))
                          : make_uint4(0, 0, 0, 0);
        }
        if (warp == 0) { FP_STAMP(8 + 2 * c) }
        if (c >= 2) mbar_wait(&a_empty[s], (uint32_t)(((c >> 1) - 1) & 1));  // the MMAs of chunk c-2 have read the slot
        uint8_t* dst = sY + s * PANEL;
#pragma unroll
        for (int p = 0; p < 4; ++p)
          *reinterpret_cast<uint4*>(dst + sw128_offset((uint32_t)(rsub + 32 * p), (uint32_t)(8 * c8))) = pk[p];
        fence_proxy_async_smem();
        mbar_arrive(&a_full[s]);
        if (warp == 0) { FP_STAMP(9 + 2 * c) }
      }
    }
    // ---- epilogues: thread = (row = TMEM lane 32 q + lane, column half hf)
    const int q = warp & 3, hf = warp >> 2;
    const int row = q * 32 + lane;
    const int grow = row0 + row;
    const bool live = grow < P.rows_total;
    const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16);
    float res_tail[16];   // vote residual columns 240..255 (held by the hf == 1 threads from layer 1 to the last layer)
#pragma unroll
    for (int i = 0; i < 16; ++i) res_tail[i] = 0.f;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    if (L > 2 && live) {
      sx = __ldg(P.seeds_xyz + (size_t)grow * 3); sy = __ldg(P.seeds_xyz + (size_t)grow * 3 + 1);
      sz = __ldg(P.seeds_xyz + (size_t)grow * 3 + 2);
    }
    for (int l = 0; l < L; ++l) {
      mbar_wait(acc_full, (uint32_t)(l & 1));
      tc_fence_after_sync();
      if (warp == 0) { FP_STAMP(34 + 4 * l) }
      const bool last = l == L - 1;
      const bool vote_out = last && L > 2;          // [features | xyz] + residual -> votes
      const bool fp_store = l == 1;                 // output of the fp module, fp32 (seed features)
      const bool relu = !vote_out;                  // every layer but the last vote layer is conv/FC + BN + ReLU
      const bool xyz_term = l == 2;                 // first vote layer: + seeds_xyz . W[0:3]
      uint8_t* dst = (l & 1) ? sY : sX;             // next layer's A operand
      const float* bias = sBias + l * BIAS_LD;
      uint32_t v[2][32];
      tmem_ld_x32(tacc + (uint32_t)(hf * 128), v[0]);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {  // 4 x 32 columns of this thread's half; the next TMEM load is in flight during the math
        const int col0 = hf * 128 + cb * 32;
        // residual of the vote output = the fp module's output, parked in TMEM (columns TM_RES..511) and in res_tail by
        // this very thread in layer 1's epilogue: no trip through global memory
        uint32_t rsd[32];
        if (vote_out) {
          if (col0 + 32 <= RES_TM_COLS) {
            tmem_ld_x32(tacc + (uint32_t)(TM_RES + col0), rsd);
          } else {
            uint32_t r16[16];
            tmem_ld_x16(tacc + (uint32_t)(TM_RES + col0), r16);
#pragma unroll
            for (int i = 0; i < 16; ++i) { rsd[i] = r16[i]; rsd[16 + i] = __float_as_uint(res_tail[i]); }
          }
        }
        tmem_ld_wait();
        if (cb + 1 < 4) tmem_ld_x32(tacc + (uint32_t)(col0 + 32), v[(cb + 1) & 1]);
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + col0 + i);
          x[i] = __uint_as_float(v[cb & 1][i]) + b4.x; x[i + 1] = __uint_as_float(v[cb & 1][i + 1]) + b4.y;
          x[i + 2] = __uint_as_float(v[cb & 1][i + 2]) + b4.z; x[i + 3] = __uint_as_float(v[cb & 1][i + 3]) + b4.w;
        }
        if (xyz_term) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 wx = *reinterpret_cast<const float4*>(sWxyz + col0 + i);
            const float4 wy = *reinterpret_cast<const float4*>(sWxyz + CW + col0 + i);
            const float4 wz = *reinterpret_cast<const float4*>(sWxyz + 2 * CW + col0 + i);
            x[i] = __fmaf_rn(sz, wz.x, __fmaf_rn(sy, wy.x, __fmaf_rn(sx, wx.x, x[i])));
            x[i + 1] = __fmaf_rn(sz, wz.y, __fmaf_rn(sy, wy.y, __fmaf_rn(sx, wx.y, x[i + 1])));
            x[i + 2] = __fmaf_rn(sz, wz.z, __fmaf_rn(sy, wy.z, __fmaf_rn(sx, wx.z, x[i + 2])));
            x[i + 3] = __fmaf_rn(sz, wz.w, __fmaf_rn(sy, wy.w, __fmaf_rn(sx, wx.w, x[i + 3])));
          }
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = fmaxf(x[i], 0.f);
        }
        if (vote_out) {
          if (live) {
            float* o = P.votes_feat + (size_t)grow * CW + col0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              st_f32x8(o + 8 * i, x[8 * i] + __uint_as_float(rsd[8 * i]), x[8 * i + 1] + __uint_as_float(rsd[8 * i + 1]),
                       x[8 * i + 2] + __uint_as_float(rsd[8 * i + 2]), x[8 * i + 3] + __uint_as_float(rsd[8 * i + 3]),
                       x[8 * i + 4] + __uint_as_float(rsd[8 * i + 4]), x[8 * i + 5] + __uint_as_float(rsd[8 * i + 5]),
                       x[8 * i + 6] + __uint_as_float(rsd[8 * i + 6]), x[8 * i + 7] + __uint_as_float(rsd[8 * i + 7]));
          }
        } else {
          if (fp_store && L > 2) {   // park the vote residual on chip (the last 16 columns do not fit: registers)
            uint32_t xr[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) xr[i] = __float_as_uint(x[i]);
            if (col0 + 32 <= RES_TM_COLS) {
              tmem_st_x32(tacc + (uint32_t)(TM_RES + col0), xr);
            } else {
              uint32_t r16[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) { r16[i] = xr[i]; res_tail[i] = x[16 + i]; }
              tmem_st_x16(tacc + (uint32_t)(TM_RES + col0), r16);
            }
          }
          if (fp_store && live && P.fp_out != nullptr) {
            float* o = P.fp_out + (size_t)grow * CW + col0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              st_f32x8(o + 8 * i, x[8 * i], x[8 * i + 1], x[8 * i + 2], x[8 * i + 3], x[8 * i + 4], x[8 * i + 5], x[8 * i + 6],
                       x[8 * i + 7]);
          }
          if (!last) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 pk = make_uint4(pack_h2(x[8 * i], x[8 * i + 1]), pack_h2(x[8 * i + 2], x[8 * i + 3]),
                                          pack_h2(x[8 * i + 4], x[8 * i + 5]), pack_h2(x[8 * i + 6], x[8 * i + 7]));
              const uint32_t kk = (uint32_t)(col0 + 8 * i);
              *reinterpret_cast<uint4*>(dst + (kk >> 6) * PANEL + sw128_offset((uint32_t)row, kk)) = pk;
            }
          }
        }
      }
      if (vote_out && hf == 1) {  // xyz offsets: permuted columns 256..258
        uint32_t v16[16];
        tmem_ld_x16(tacc + 256u, v16);
        tmem_ld_wait();
        if (live) {
          P.votes_xyz[(size_t)grow * 3 + 0] = sx + (__uint_as_float(v16[0]) + bias[256]);
          P.votes_xyz[(size_t)grow * 3 + 1] = sy + (__uint_as_float(v16[1]) + bias[257]);
          P.votes_xyz[(size_t)grow * 3 + 2] = sz + (__uint_as_float(v16[2]) + bias[258]);
        }
      }
      if (fp_store && L > 2) tmem_st_wait();
      if (!last) {
        fence_proxy_async_smem();   // the next layer's MMAs read `dst` through the async proxy
        tc_fence_before_sync();     // ... and overwrite the accumulator these loads have drained
        mbar_arrive(act_full);
      }
      if (warp == 0) { FP_STAMP(35 + 4 * l) }
    }
  }
#undef FP_STAMP
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace fpc

extern long long* g_sa_trace;  // sa1_ws2.cu (vnb_debug_sa_trace)

int fp_chain_launch(const fpc::Params& p0, cudaStream_t st) {
  fpc::Params p = p0;
  p.trace = g_sa_trace;
  VNB_CUDA(cudaFuncSetAttribute(fpc::fp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fpc::SMEM));
  const int grid = (p.rows_total + 127) / 128;
  fpc::fp_chain_kernel<<<grid, fpc::THREADS, fpc::SMEM, st>>>(p);
  return check_launch("fp_module_fused (tcgen05)");
}

}  // namespace vnb

using namespace vnb;

extern "C" int vnb_fp_module_fused(int b, int n, int m, int c1, int c2, const float* dist, const int* idx,
                                   const float* points1, const float* points2, int n_fp_layers,
                                   const void* const* fp_w_img, const float* const* fp_bias, const int* fp_cout,
                                   float* fp_out, int n_vote_layers, const void* const* vote_w_img,
                                   const float* const* vote_bias, const int* vote_cout, const float* vote_w0_xyz_f32,
                                   const float* seeds_xyz, float* votes_xyz, float* votes_feat, void* stream) {
  VNB_REQUIRE(b >= 0 && n > 0 && m > 0, "fp_module_fused: bad shape");
  VNB_REQUIRE(c1 == fpc::CW && c2 == fpc::CW, "fp_module_fused: points1 / points2 must have 256 channels (got %d / %d)", c1, c2);
  VNB_REQUIRE(n_fp_layers == 2 && fp_cout[0] == fpc::CW && fp_cout[1] == fpc::CW, "fp_module_fused: fp mlp must be [256, 256]");
  VNB_REQUIRE(n_vote_layers == 0 || (n_vote_layers == 3 && vote_cout[0] == fpc::CW && vote_cout[1] == fpc::CW &&
                                     vote_cout[2] == fpc::CW + 3),
              "fp_module_fused: vote units must be [256, 256, 259]");
  VNB_REQUIRE(dist && idx && points1 && points2, "fp_module_fused: null buffer");
  VNB_REQUIRE(fp_out != nullptr || n_vote_layers > 0, "fp_module_fused: fp_out may be NULL only with the voting module fused behind");
  if (n_vote_layers) VNB_REQUIRE(vote_w0_xyz_f32 && seeds_xyz && votes_xyz && votes_feat, "fp_module_fused: null vote buffer");
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };   // 256-bit global accesses
  VNB_REQUIRE(al32(points1) && al32(points2) && al32(fp_out) && al32(votes_feat),
              "fp_module_fused: points1 / points2 / fp_out / votes_feat must be 32-byte aligned");
  if (b == 0) return VNB_OK;
  fpc::Params p = {};
  p.rows_total = b * n; p.n = n; p.m = m;
  p.dist = dist; p.idx = idx; p.points1 = points1; p.points2 = points2;
  p.n_layers = 2 + n_vote_layers;
  p.w_img[0] = static_cast<const char*>(fp_w_img[0]); p.bias[0] = fp_bias[0]; p.k_pad[0] = c1 + c2; p.n_pad[0] = fpc::CW; p.n_out[0] = fpc::CW;
  p.w_img[1] = static_cast<const char*>(fp_w_img[1]); p.bias[1] = fp_bias[1]; p.k_pad[1] = fpc::CW; p.n_pad[1] = fpc::CW; p.n_out[1] = fpc::CW;
  for (int i = 0; i < n_vote_layers; ++i) {
    p.w_img[2 + i] = static_cast<const char*>(vote_w_img[i]);
    p.bias[2 + i] = vote_bias[i];
    p.k_pad[2 + i] = fpc::CW;
    p.n_pad[2 + i] = i == 2 ? round_up(fpc::CW + 3, 16) : fpc::CW;
    p.n_out[2 + i] = vote_cout[i];
  }
  p.seeds_xyz = seeds_xyz; p.w_vote_xyz = vote_w0_xyz_f32;
  p.fp_out = fp_out; p.votes_xyz = votes_xyz; p.votes_feat = votes_feat;
  return fp_chain_launch(p, as_stream(stream));
}

// out = act(in @ W + b) (+ residual) on the tensor cores: fp16 operands, fp32 accumulation in TMEM.
// (FullyConnected / 1x1 Conv2D + folded BN + ReLU of the reference: utils.py:291-292, model.py:53-57, utils.py:149-155;
//  also the hoisted layer-1 pre-GEMM of the fused SA kernels, see mlp_tc.cu.)
//
// One CTA (256 threads) per 128-row x <=128-column output tile.  K is streamed in 64-column chunks (one SW128 panel)
// through a 2-stage ring: the activation chunk is read from global memory with 8 threads per row (256 contiguous
// bytes), converted fp32 -> fp16 in registers and written as the 128-byte-swizzled K-major A operand; the weight chunk
// arrives as ONE bulk copy of the pre-swizzled image panel (cp.async.bulk, completes a transaction barrier).  One
// thread issues tcgen05.mma; tcgen05.commit frees the stage.  Epilogue: TMEM -> registers (row per lane),
// bias/ReLU/residual, 16-byte vector stores; the eight warps split the columns in two halves.
// A CTA is a chain of latencies (global load -> convert -> barrier -> MMA -> commit), so the kernel is sized for
// CO-RESIDENCY, not for a big tile: 66 KB of shared memory and 128 TMEM columns per CTA let three CTAs (of this or of
// another step's launch) share an SM and hide each other's stalls — with several forwards in flight the GPU is
// throughput-bound on the sum of all kernels' SM-time (scripts/gpu_stress.py).
#include "common.cuh"
#include "umma.cuh"

namespace vnb {

using namespace umma;

constexpr int LT_THREADS = 256;
constexpr int LT_KC = 64;              // K chunk = one 64-column SW128 panel
constexpr int LT_A_STAGE = 128 * 128;  // 128 rows x 64 k x 2 B
constexpr int LT_B_STAGE = 128 * 128;  // up to 128 n-rows x 64 k x 2 B
constexpr int LT_SMEM = 2 * (LT_A_STAGE + LT_B_STAGE) + 512 /*bias*/ + 128 /*barriers*/ + 1024 /*align slack*/;

__device__ __forceinline__ uint32_t lt_pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(LT_THREADS, 3) linear_tc_kernel(int rows, int cin, int cout, int k_pad, int n_pad,
                                                                  const float* __restrict__ in,
                                                                  const char* __restrict__ w_img,
                                                                  const float* __restrict__ bias,
                                                                  const float* __restrict__ res, int act,
                                                                  float* __restrict__ out_f32,
                                                                  __half* __restrict__ out_f16) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* sA[2] = {smem, smem + LT_A_STAGE};
  uint8_t* sB[2] = {smem + 2 * LT_A_STAGE, smem + 2 * LT_A_STAGE + LT_B_STAGE};
  float* sBias = reinterpret_cast<float*>(smem + 2 * (LT_A_STAGE + LT_B_STAGE));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (LT_A_STAGE + LT_B_STAGE) + 512);
  uint64_t* full_b = bars;       // [2] weights landed
  uint64_t* empty = bars + 2;    // [2] MMAs reading the stage have completed
  uint64_t* done = bars + 4;     // all MMAs of the tile completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * 128;
  const int nt0 = blockIdx.y * 128;      // first image row (= output column) of this n-tile
  const int nt = min(128, n_pad - nt0);  // multiple of 16
  const int nchunks = (k_pad + LT_KC - 1) / LT_KC;

  if (tid == 0) {
    mbar_init(&full_b[0], 1); mbar_init(&full_b[1], 1);
    mbar_init(&empty[0], 1);  mbar_init(&empty[1], 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (tid < 128) sBias[tid] = (bias != nullptr && nt0 + tid < cout) ? bias[nt0 + tid] : 0.f;
  if (warp == 0) tmem_alloc(tmem_ptr, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = *tmem_ptr;
  const uint32_t idesc = make_idesc_f16_f32(128, (uint32_t)nt);

  const int c8 = tid & 7;     // 16-byte (8 x fp16) chunk of the 64-wide K chunk
  const int rsub = tid >> 3;  // rows rsub, rsub+32, rsub+64, rsub+96
  const bool vec_in = (cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  const bool vec8_in = (cin % 8 == 0) && ((reinterpret_cast<uintptr_t>(in) & 31) == 0);   // one 256-bit load per 8 floats

  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    const int kc0 = c * LT_KC;
    const int kc = min(LT_KC, k_pad - kc0);  // multiple of 16
    if (c >= 2) mbar_wait(&empty[s], (uint32_t)(((c >> 1) - 1) & 1));  // stage free again
    if (tid == 0) {
      mbar_arrive_expect_tx(&full_b[s], (uint32_t)(nt * 128));
      bulk_g2s(sB[s], w_img + (size_t)(kc0 >> 6) * n_pad * 128 + (size_t)nt0 * 128, (uint32_t)(nt * 128), &full_b[s]);
    }
    if (c8 * 8 < kc) {
      const int kb = kc0 + c8 * 8;
      float v[4][8];
#pragma unroll
      for (int p = 0; p < 4; ++p) {  // all loads first (4 independent 32-byte reads per thread)
        const int gr = row0 + rsub + 32 * p;
        const float* src = in + (size_t)gr * cin + kb;
        if (gr < rows && vec8_in && kb + 8 <= cin) {
          const f32x8 a = ldg_f32x8(src);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[p][i] = a.v[i];
        } else if (gr < rows && vec_in && kb + 8 <= cin) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src));
          const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
          v[p][0] = a.x; v[p][1] = a.y; v[p][2] = a.z; v[p][3] = a.w;
          v[p][4] = b.x; v[p][5] = b.y; v[p][6] = b.z; v[p][7] = b.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[p][i] = (gr < rows && kb + i < cin) ? __ldg(src + i) : 0.f;
        }
      }
      const uint32_t kk = (uint32_t)c8 * 8;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const uint4 pk = make_uint4(lt_pack_h2(v[p][0], v[p][1]), lt_pack_h2(v[p][2], v[p][3]),
                                    lt_pack_h2(v[p][4], v[p][5]), lt_pack_h2(v[p][6], v[p][7]));
        *reinterpret_cast<uint4*>(sA[s] + sw128_offset((uint32_t)(rsub + 32 * p), kk)) = pk;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&full_b[s], (uint32_t)((c >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t a0 = smem_u32(sA[s]), b0 = smem_u32(sB[s]);
      for (int ks = 0; ks < kc / 16; ++ks)
        mma_f16_ss(tmem_d, make_desc_sw128(a0 + (uint32_t)ks * 32), make_desc_sw128(b0 + (uint32_t)ks * 32), idesc,
                   (c > 0 || ks > 0) ? 1u : 0u);
      mma_commit(&empty[s]);
      if (c == nchunks - 1) mma_commit(done);
    }
  }
  mbar_wait(done, 0);
  tc_fence_after_sync();
  // epilogue: warp w reads TMEM lanes 32*(w&3).. (tile rows); warps 0-3 take the first half of the columns, 4-7 the rest
  {
    const int q = warp & 3, half = warp >> 2;
    const int orow = row0 + q * 32 + lane;
    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
    const int nblk = nt / 16, b0 = half * ((nblk + 1) / 2), b1 = half ? nblk : (nblk + 1) / 2;
    const bool ok32 = (cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(out_f32) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(res) & 15) == 0);
    const bool ok16 = (cout % 8 == 0) && ((reinterpret_cast<uintptr_t>(out_f16) & 15) == 0);
    const bool use_vec = ((out_f32 == nullptr && res == nullptr) || ok32) && (out_f16 == nullptr || ok16);
    const bool wide32 = (cout % 8 == 0) && ((reinterpret_cast<uintptr_t>(out_f32) & 31) == 0);   // 256-bit stores
    const bool wide16 = (cout % 16 == 0) && ((reinterpret_cast<uintptr_t>(out_f16) & 31) == 0);
    for (int blk = b0; blk < b1; ++blk) {
      const int cc = blk * 16;
      uint32_t v[16];
      tmem_ld_x16(taddr + (uint32_t)cc, v);
      tmem_ld_wait();
      if (orow >= rows) continue;
      float x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        x[i] = __uint_as_float(v[i]) + sBias[cc + i];
        if (act == VNB_ACT_RELU) x[i] = fmaxf(x[i], 0.f);
      }
      const int n0 = nt0 + cc;
      const size_t o = (size_t)orow * cout + n0;
      if (n0 + 16 <= cout && use_vec) {
        if (res != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(res + o) + i);
            x[4 * i] += r4.x; x[4 * i + 1] += r4.y; x[4 * i + 2] += r4.z; x[4 * i + 3] += r4.w;
          }
        }
        if (out_f32 != nullptr) {
          if (wide32) {   // whole 32-byte sectors per lane (a lane owns a row: every access touches 32 lines anyway)
            st_f32x8(out_f32 + o, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
            st_f32x8(out_f32 + o + 8, x[8], x[9], x[10], x[11], x[12], x[13], x[14], x[15]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(out_f32 + o)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
          }
        }
        if (out_f16 != nullptr) {
          if (wide16) {
            st_b32x8(out_f16 + o, lt_pack_h2(x[0], x[1]), lt_pack_h2(x[2], x[3]), lt_pack_h2(x[4], x[5]), lt_pack_h2(x[6], x[7]),
                     lt_pack_h2(x[8], x[9]), lt_pack_h2(x[10], x[11]), lt_pack_h2(x[12], x[13]), lt_pack_h2(x[14], x[15]));
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i)
              reinterpret_cast<uint4*>(out_f16 + o)[i] =
                  make_uint4(lt_pack_h2(x[8 * i], x[8 * i + 1]), lt_pack_h2(x[8 * i + 2], x[8 * i + 3]),
                             lt_pack_h2(x[8 * i + 4], x[8 * i + 5]), lt_pack_h2(x[8 * i + 6], x[8 * i + 7]));
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (n0 + i < cout) {
            float y = x[i];
            if (res != nullptr) y += res[o + i];
            if (out_f32 != nullptr) out_f32[o + i] = y;
            if (out_f16 != nullptr) out_f16[o + i] = __float2half_rn(y);
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
}

int linear_tc(int rows, int cin, int cout, const float* in, const void* w_img, const float* bias, const float* res,
              int act, float* out_f32, void* out_f16, cudaStream_t st) {
  const int k_pad = round_up(cin, 16), n_pad = round_up(cout, 16);
  VNB_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
  dim3 grid((rows + 127) / 128, (n_pad + 127) / 128);
  linear_tc_kernel<<<grid, LT_THREADS, LT_SMEM, st>>>(rows, cin, cout, k_pad, n_pad, in, static_cast<const char*>(w_img),
                                                     bias, res, act, out_f32, static_cast<__half*>(out_f16));
  return check_launch("linear (tcgen05)");
}

}  // namespace vnb

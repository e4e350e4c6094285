/* votenet_b200 — C ABI of the B200 (sm_100a) VoteNet / PointNet++ inference hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / TensorFlow types.  Each entry point is the
 * C-ABI twin of one reference launcher or CPU loop (cited per function; paths relative to the reference repo
 * qq456cvb/VoteNet).  Conventions, all mirroring the reference's native side:
 *   - every pointer is a DEVICE pointer to dense, row-major, contiguous data; inputs are const;
 *   - the CALLER allocates every output and workspace buffer; the library never allocates or frees (the one exception
 *     is the explicit vnb_peer_alloc / vnb_peer_free pair of the multi-GPU exchange);
 *   - launches are asynchronous on the given `stream` (a cudaStream_t passed as void*); no hidden host sync;
 *     every entry point is re-entrant and CUDA-graph capturable; the only process-global state is the experiment knobs of
 *     vnb_set_tuning (documented there as not synchronised) and the debugging hooks;
 *   - return 0 on success; VNB_ERR_INVALID for a shape / attribute violation (where the reference op raises
 *     errors::InvalidArgument), VNB_ERR_CUDA for a CUDA launch error (which the reference never checks).
 *     vnb_last_error() returns a thread-local message for the last non-zero return.
 * There is no CPU fallback: without a CUDA device every compute entry point returns VNB_ERR_CUDA.
 */
#ifndef VOTENET_B200_H_
#define VOTENET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VNB_OK 0
#define VNB_ERR_INVALID 1
#define VNB_ERR_CUDA 2

#define VNB_ABI_VERSION 1

int vnb_abi_version(void);
const char* vnb_last_error(void);
/* number of kernel launches this library has issued in this process (all threads); bench.py reports it */
unsigned long long vnb_launch_count(void);
/* Tuning knobs for experiments (defaults are what bench.py uses) — PROCESS-GLOBAL and not synchronised: set them before
 * launching from several threads, never concurrently with calls (the entry points themselves keep no other state and are
 * re-entrant).  "fps_variant" 1 = bucket-pruned sampler for 2048 < n <= 20480 (default) / 0 = register-cluster sampler only,
 * "fps_cluster" CTAs per cluster of the latter (0 = auto), "ball_query_variant" 0 = exhaustive scan / 1 = grid + bitmap,
 * "bq_grid_min_n", "sa_variant" 2 = MMA issuers park (default) / 3 = poll, "sa_sms", "sa_split", "sa_min_tpc" (fewest
 * tiles per CTA of the fused SA kernels), "nms_cluster" CTAs per cloud of the NMS kernel (1, 2, 4, 8), "fps_ablate"
 * (timing experiments only: wrong results).  Unknown keys return VNB_ERR_INVALID. */
int vnb_set_tuning(const char* key, int value);
/* Debugging aid: when given a device buffer of 16 x 10 int64, the next bucket-pruned FPS launches accumulate, per warp of
 * cloud 0, dependency-anchored cycle counts of the round's phases into it (scripts/gpu_fps_phases.py); NULL = off. */
int vnb_debug_fps_profile(void* device_buffer_16x10_i64);
/* Debugging aid: when given a device buffer of 12 x 64 x 2 int64, CTA 0 of the next fused tensor-core launches stamps
 * clock64() at the start / end of every pipeline stage of its first 64 tiles (sa1: roles P, M1, M2, M3, E1, E2, E3a, E3b;
 * hoisted SA kernels: P, M2, E2, M3, E3; fused FP kernel: loader chunks and per-layer marks — scripts/gpu_trace_*.py
 * decode them); NULL = off. */
int vnb_debug_sa_trace(void* device_buffer_8x64x2_i64);
/* debugging aid: host-mapped int[8]; a bounded device-side wait that gives up stores {source line, blockDim.x,
 * blockIdx.x, threadIdx.x, gridDim.x} there before it traps */
int vnb_debug_trap_buffer(void* host_mapped_int8);

/* ------------------------------------------------------------------ tf_ops/sampling ------------------- */

/* farthestpointsamplingLauncher(b,n,m,inp,temp,out)            tf_ops/sampling/tf_sampling_g.cu:203-205
 * (op FarthestPointSample, tf_sampling.cpp:95-123; python farthest_point_sample(npoint, inp), tf_sampling.py:48-56)
 * xyz (b,n,3) f32 -> out_idx (b,m) i32.  Bit-identical index sequence to the reference kernel (tie rule included).
 * The reference's (32,n) `temp` scratch is not needed: the running min-distances live in registers.
 * Limits: 1 <= n <= 32768; m >= 1 (m <= 0 is VNB_ERR_INVALID like the op's npoint>0 check, tf_sampling.cpp:99). */
int vnb_farthest_point_sample(int b, int n, int m, const float* xyz, int* out_idx, void* stream);

/* Same result as vnb_farthest_point_sample, for inputs that are themselves FPS-ordered (the nested levels sa2..sa4 and
 * the proposal module run FPS on the previous level's FPS output, utils.py:43-45, model.py:89-93): the output is then
 * the identity prefix 0..m-1 unless exact float ties reorder it.  This entry point PROVES that per cloud with a fully
 * parallel check and runs the sequential sampler only for clouds where the proof fails; it is correct (bit-identical
 * to vnb_farthest_point_sample) for ANY input, just not faster when the input is not FPS-ordered.
 * workspace: vnb_fps_nested_workspace_bytes(b, m) bytes. */
size_t vnb_fps_nested_workspace_bytes(int b, int m);
int vnb_farthest_point_sample_nested(int b, int n, int m, const float* xyz, int* out_idx, void* workspace, void* stream);

/* FPS with provenance tracking.  vnb_farthest_point_sample_ties is vnb_farthest_point_sample that also reports, per
 * cloud, the first round whose arg-max was NOT unique (two points shared the maximal running distance exactly and the
 * reference's tie rule decided); 0x7fffffff if every round had a unique winner, 0 when the kernel in use cannot tell.
 * Only rounds < track_rounds are reported (<= 0: all m rounds) — a nested level of m' picks needs m' of them.  The
 * tracking costs ~8 % of a round (making it conditional on the round index cost more than it saved: measured).
 * vnb_farthest_point_sample_nested_hint is vnb_farthest_point_sample_nested for an input that is the GATHER, in order,
 * of the first n picks of such a parent call (new_xyz of the previous level, utils.py:42-45): if the parent's first
 * tie round is >= m, every round j < m had a unique arg-max p_j over the parent's superset, p_j is in the subset, the
 * running distances are the same floats, hence the subset's FPS picks position j — the identity prefix holds with no
 * check at all.  Clouds whose hint is smaller go to the sequential sampler.  The
 * caller vouches for the provenance (wrong provenance = wrong indices); results are otherwise bit-identical. */
int vnb_farthest_point_sample_ties(int b, int n, int m, const float* xyz, int* out_idx, int* first_tie_round,
                                   int track_rounds, void* stream);
/* vnb_farthest_point_sample_nested that also reports what it proved: proven_rounds[cloud] = m when the identity prefix
 * 0..m-1 was PROVEN for the cloud (every round j < m of FPS over these n points picks position j, under the reference's
 * index-based tie rule), else 0.  That array is a valid hint (parent_first_tie_round) for
 * vnb_farthest_point_sample_nested_hint on any PREFIX xyz[:, :n'] of the same points, n' <= n, for m' <= m picks: a
 * competitor inside the prefix is inside the proven set too, it has the same running distance and the same index
 * there, so position j wins the prefix's round j as well.  One proof at the first nested level (2048 -> 1024) thus
 * covers the deeper levels and the proposal module's sampling (utils.py:42-45, model.py:89-93), which all sample
 * prefixes of it — without tie tracking in the big sa1 sampler. */
int vnb_farthest_point_sample_nested_proof(int b, int n, int m, const float* xyz, int* out_idx, void* workspace,
                                           int* proven_rounds, void* stream);
int vnb_farthest_point_sample_nested_hint(int b, int n, int m, const float* xyz, int* out_idx, void* workspace,
                                          const int* parent_first_tie_round, void* stream);

/* gatherpointLauncher(b,n,m,inp,idx,out)                        tf_sampling_g.cu:206-208
 * inp (b,n,3), idx (b,m) -> out (b,m,3). */
int vnb_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out, void* stream);

/* ------------------------------------------------------------------ tf_ops/grouping ------------------- */

/* queryBallPointLauncher(b,n,m,radius,nsample,xyz1,xyz2,idx,pts_cnt)   tf_grouping_g.cu:125-128
 * xyz1 (b,n,3) searched set, xyz2 (b,m,3) queries -> idx (b,m,nsample) i32, pts_cnt (b,m) i32.
 * First `nsample` hits in ascending index order, padded with the first hit; rows of empty balls are left untouched
 * exactly like the reference (tf_grouping_g.cu:14-34).  radius > 0, nsample > 0 (tf_grouping.cpp:71,74). */
int vnb_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                         int* idx, int* pts_cnt, void* stream);

/* Same outputs, bit for bit, computed through a uniform grid + per-query index bitmap instead of the exhaustive scan
 * (see csrc/ball_query_grid.cu).  workspace: vnb_query_ball_point_workspace_bytes(b,n) bytes; with workspace == NULL
 * (or n < 1024, tuning knob bq_grid_min_n, or n > 32768) this is vnb_query_ball_point. */
size_t vnb_query_ball_point_workspace_bytes(int b, int n);
int vnb_query_ball_point_ws(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                            int* idx, int* pts_cnt, void* workspace, void* stream);

/* vnb_query_ball_point_ws in two halves.  _prepare bins the searched set xyz1 into the workspace (it needs neither the
 * queries nor nsample, so it can run while the sampler that produces xyz2 is still working; a no-op whenever the _ws call
 * would take the exhaustive scan); _prepared answers the queries from a workspace prepared with the same b, n, radius
 * and xyz1.  _ws == _prepare followed by _prepared on the same stream. */
int vnb_query_ball_point_prepare(int b, int n, float radius, const float* xyz1, void* workspace, void* stream);
int vnb_query_ball_point_prepared(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                                  int* idx, int* pts_cnt, void* workspace, void* stream);

/* groupPointLauncher(b,n,c,m,nsample,points,idx,out)            tf_grouping_g.cu:133-136
 * points (b,n,c), idx (b,m,nsample) -> out (b,m,nsample,c). */
int vnb_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out,
                    void* stream);

/* ------------------------------------------------------------------ tf_ops/3d_interpolation ----------- */

/* threenn_cpu(b,n,m,xyz1,xyz2,dist,idx)                         tf_interpolate.cpp:60-103
 * xyz1 (b,n,3) unknown, xyz2 (b,m,3) known -> dist (b,n,3) SQUARED f32, idx (b,n,3) i32 (ascending, strict <). */
int vnb_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx, void* stream);

/* threeinterpolate_cpu(b,m,c,n,points,idx,weight,out)           tf_interpolate.cpp:107-127
 * points (b,m,c), idx (b,n,3), weight (b,n,3) -> out (b,n,c). */
int vnb_three_interpolate(int b, int m, int c, int n, const float* points, const int* idx, const float* weight,
                          float* out, void* stream);

/* ------------------------------------------------------------------ tf_ops/3d_nms --------------------- */

/* NonMaxSuppression3DOp::Compute                                 tf_nms3d.cpp:283-305 (+ :202-273)
 * bbox (b,k,8,3), scores (b,k), objectiveness (b,k,2), 0 <= iou_threshold <= 1.   k <= 1024 (one CTA per cloud holds
 * the cloud's suppression bitmask in shared memory); any b.  NaN scores rank last (the order stays total).
 * Outputs (all caller-allocated):
 *   keep      (b,k)   u8   1 where the box survives NMS (the per-cloud keep mask)
 *   out_idx   (b*k,2) i32  rows (batch, box) of the survivors in the reference's GLOBAL descending-score order
 *                          (exact score ties: ascending (batch, box) — the reference's order among exactly equal
 *                          scores is a libstdc++ heap artefact)
 *   out_count (1)     i32  number of valid rows in out_idx
 * workspace: vnb_nms3d_workspace_bytes(b,k) bytes of device memory. */
size_t vnb_nms3d_workspace_bytes(int b, int k);
/* After a vnb_nms3d / vnb_decode_nms3d call the u32 at this byte offset of its workspace holds the number of clipped
 * pairs whose IoU fell within 1e-5 of iou_threshold — the only pairs on which a last-bit difference from the reference's
 * x86 arithmetic could change the keep mask (0 on every parity-test input). */
size_t vnb_nms3d_near_threshold_offset(int b, int k);
int vnb_nms3d(int b, int k, const float* bbox, const float* scores, const float* objectiveness, float iou_threshold,
              uint8_t* keep, int* out_idx, int* out_count, void* workspace, void* stream);

/* Box decode (model.py:100-129) + NonMaxSuppression3DOp (tf_nms3d.cpp:202-305) as ONE kernel, one CTA per cloud —
 * what the forward path calls.  Inputs as vnb_decode_boxes, outputs as vnb_decode_boxes + vnb_nms3d, plus
 *   out_key (b*k) u32   the order-preserving score key of every row of out_idx (descending): together with out_idx
 *                       and out_count this is a rank's sorted detection list, the input of vnb_merge_detections.
 * and, where the pointers are not NULL, the output gathers of model.py:135-137 over this batch, rows in out_idx order:
 *   bboxes_pred (b*k,8,3) f32 = bboxes[out_idx], class_scores_pred (b*k,10) f32, batch_idx (b*k) i32 (first out_count rows).
 * workspace: vnb_nms3d_workspace_bytes(b,k). */
int vnb_decode_nms3d(int b, int k, const float* proposals_xyz, const float* proposals_output,
                     const float* class_mean_size, float iou_threshold, float* bboxes, float* scores,
                     float* objectness, float* class_scores, uint8_t* keep, int* out_idx, uint32_t* out_key,
                     int* out_count, float* bboxes_pred, float* class_scores_pred, int* batch_idx, void* workspace,
                     void* stream);

/* Global ordering of NMS survivors across an all-gathered batch and the output gathers of
 * model.py:135-137.  `gathered` holds `world` (<= 64) records of `rank_stride` bytes; record r carries, at the given
 * byte offsets, its rank's sorted detection list — rows (local_batch, box) i32 at off_idx, their score keys u32 at
 * off_key, the row count i32 at off_count (all three as written by vnb_decode_nms3d) — and the decoded boxes
 * (b,k,8,3) f32 at off_bboxes / class scores (b,k,10) f32 at off_class_scores.  A k-way merge by rank (own position +
 * binary searches in the other lists) writes
 *   out_idx (world*b*k,2) i32  rows (global_batch, box), global_batch = rank*b + local, descending score, exact ties by
 *                              ascending (global_batch, box) — the reference's (Nnms,2) output for the whole sharded
 *                              batch (tf_nms3d.cpp:240-272);  out_count (1) i32 = Nnms
 * and, where the pointers are not NULL, the three gathers of model.py:135-137:
 *   bboxes_pred (world*b*k,8,3) f32, class_scores_pred (world*b*k,10) f32, batch_idx (world*b*k) i32  (first Nnms rows). */
int vnb_merge_detections(int world, int b, int k, const void* gathered, size_t rank_stride, size_t off_idx,
                         size_t off_key, size_t off_count, size_t off_bboxes, size_t off_class_scores, int* out_idx,
                         int* out_count, float* bboxes_pred, float* class_scores_pred, int* batch_idx, void* stream);

/* One-sided all-gather of the detection records over NVLink peer memory (csrc/peer.cu) — what the multi-GPU path uses
 * instead of a collective rendezvous.  Every rank owns an inbox of `world` records per slot and `world` int32 sequence
 * flags per slot; `peer_inbox[p]` / `peer_flags[p]` are device pointers to the slot's buffers in rank p's memory
 * (CUDA-IPC mappings for p != rank; HOST arrays of `world` pointers, read during the call).
 *   vnb_peer_push_record: stores this rank's record (nbytes, 16-byte aligned and sized) into row `rank` of every
 *     peer's inbox, then raises flag[rank] = seq there (system-scope release).  One launch, `world` CTAs.
 *   vnb_peer_wait: one warp waits until all `world` local flags of the slot have reached seq (acquire); enqueue it
 *     in front of vnb_merge_detections on the same stream.  A peer that never arrives traps after a few seconds.
 * seq must increase with every reuse of a slot; a slot may be reused once the merge that consumed it is complete on
 * every rank (the Python side keeps two inbox slots per in-flight forward, which makes that hold by stream order). */
int vnb_peer_push_record(int world, int rank, const void* record, size_t nbytes, void* const* peer_inbox,
                         int* const* peer_flags, int seq, void* stream);
int vnb_peer_wait(int world, const int* flags, int seq, void* stream);
/* Lets kernels of the CURRENT device load / store memory of `peer_device` (cudaDeviceEnablePeerAccess; a no-op when it is
 * already enabled or peer_device is the current device).  Call once per peer before the first vnb_peer_push_record. */
int vnb_peer_enable_access(int peer_device);
/* The inbox memory.  vnb_peer_alloc: a dedicated, zeroed device allocation on the current device + its 64-byte CUDA IPC
 * handle (exchange it with the peers by any host channel); vnb_peer_open: map a peer's allocation into the CURRENT
 * device's context (peer-to-peer over NVLink; the exporting process must keep the allocation alive); vnb_peer_close /
 * vnb_peer_free undo them.  The only entry points of the library that allocate device memory. */
int vnb_peer_alloc(size_t nbytes, void** dev_ptr, unsigned char* ipc_handle_64);
int vnb_peer_open(const unsigned char* ipc_handle_64, void** dev_ptr);
int vnb_peer_close(void* dev_ptr);
int vnb_peer_free(void* dev_ptr);

/* ------------------------------------------------------------------ fused layers (boundary B) --------- */
/* The dense arithmetic the reference delegates to TensorFlow/Tensorpack (Conv2D 1x1 + BN(EMA) + ReLU + reduce_max,
 * utils.py:120-158,286-292; FullyConnected, model.py:53-61).  BatchNorm is folded into W,b by the host. */

/* Packed weight image for the tensor-core kernels: fp16, [n_pad rows][k_pad cols] K-major, 128-byte-swizzled panels
 * of 64 k-columns (the exact shared-memory image tcgen05.mma consumes).  n_pad = round_up(cout,16),
 * k_pad = round_up(cin,16).  vnb_pack_weight_f16 converts W (cin,cout) f32 row-major (TF [Cin,Cout]) on device. */
size_t vnb_weight_image_bytes(int cin, int cout);
int vnb_pack_weight_f16(int cin, int cout, const float* w, void* image, void* stream);

#define VNB_ACT_NONE 0
#define VNB_ACT_RELU 1

/* out[r, :] = act(in[r, :] @ W + bias) (+ residual[r, :])       rows x cin -> rows x cout
 * precision 0: fp32 SIMT (W = (cin,cout) f32 in `w_f32`);  precision 1: fp16 tensor cores, fp32 accumulate
 * (W = packed image in `w_img`).  Either out_f32 (rows,cout) or out_f16 (rows,cout) (or both) may be given. */
int vnb_linear(int rows, int cin, int cout, const float* in, const float* w_f32, const void* w_img, const float* bias,
               const float* residual, int act, float* out_f32, void* out_f16, int precision, void* stream);

/* pointnet_sa_module grouping + shared MLP (3 layers) + max-pool, fused        utils.py:49-55,120-132
 *   new_points[b,j,:] = max_s relu(L3(relu(L2(relu(L1([xyz[idx]-new_xyz, feat[idx]]))))))
 * xyz (b,n,3), feat (b,n,c) f32, new_xyz (b,m,3), idx (b,m,nsample) with nsample == 64 -> out (b,m,c3) f32.
 * Layer i has folded weights w{i} (cin_i,cout_i) f32 and bias b{i}; cin_1 = 3 + c (relative xyz first).
 * precision 0: fp32 SIMT, weights read from w*_f32.
 * precision 1: tensor cores.  Needs w2_img / w3_img (packed) and
 *    - c <= 13 : layer 1 also on tensor cores from w1_img (K padded to 16);
 *    - c  > 13 : layer 1 is hoisted through the gather: q = feat @ W1[3:,:] + b1 is computed ONCE per source point
 *                by the caller (vnb_linear, fp16 output (b*n, c1)) and passed as `q_f16`; the kernel adds the
 *                rank-3 relative-xyz term W1[0:3,:] (w1_f32 rows 0..2) per grouped row in fp32. */
int vnb_sa_group_mlp_max(int b, int n, int c, int m, int nsample, const float* xyz, const float* feat,
                         const float* new_xyz, const int* idx, int c1, int c2, int c3, const float* w1_f32,
                         const float* b1, const float* w2_f32, const float* b2, const float* w3_f32, const float* b3,
                         const void* w1_img, const void* w2_img, const void* w3_img, const void* q_f16, float* out,
                         int precision, void* workspace, void* stream);
/* workspace (optional, may be NULL): vnb_sa_workspace_bytes(b,m,nsample) bytes; the pipelined tensor-core kernel of the
 * hoisted variant needs it (grouped relative coordinates); without it the single-role kernel runs. */
size_t vnb_sa_workspace_bytes(int b, int m, int nsample);
/* Same, with the ball query's pts_cnt (b,m) i32 (NULL = unknown): query_ball_point pads a row of cnt < nsample hits with
 * copies of its first hit (tf_grouping_g.cu:26-29), so only cnt rows of a group are distinct and the padded rows cannot
 * change the max-pool.  The tensor-core kernels then run each centroid in a slot of 16, 32 or 64 rows (the smallest that
 * holds its distinct rows) and pack 8 / 4 / 2 centroids into a 128-row MMA tile (csrc/sa_pack.cu) — bit-identical
 * outputs, 26-56 % of the tiles on SUN-RGB-D-shaped clouds.  pts_cnt must be the count returned for `idx`. */
int vnb_sa_group_mlp_max_counted(int b, int n, int c, int m, int nsample, const float* xyz, const float* feat,
                                 const float* new_xyz, const int* idx, const int* pts_cnt, int c1, int c2, int c3,
                                 const float* w1_f32, const float* b1, const float* w2_f32, const float* b2,
                                 const float* w3_f32, const float* b3, const void* w1_img, const void* w2_img,
                                 const void* w3_img, const void* q_f16, float* out, int precision, void* workspace,
                                 void* stream);

/* pointnet_fp_module front half: inverse-distance weights + three_interpolate + concat     utils.py:279-286
 * dist (b,n,3), idx (b,n,3) from vnb_three_nn; points2 (b,m,c2) known features; points1 (b,n,c1) skip features
 * -> out (b,n,c2+c1) = [interpolated, points1]. */
int vnb_fp_interpolate_concat(int b, int n, int m, int c1, int c2, const float* dist, const int* idx,
                              const float* points1, const float* points2, float* out, void* stream);

/* pointnet_fp_module (utils.py:266-294) as ONE tensor-core kernel, optionally followed by the voting module
 * (model.py:53-61) in the same kernel: a CTA walks 128 rows through every layer, activations stay on the SM
 * (csrc/fp_chain.cu).  dist / idx from vnb_three_nn; points1 (b,n,256) skip features; points2 (b,m,256) known features.
 *   fp layers (n_fp_layers == 2, widths [256,256]): packed images (vnb_pack_weight_f16) of the BN-folded weights
 *     (512,256) and (256,256), biases; fp_out (b*n,256) f32 receives the module's output (the seed features).
 *   vote layers (n_vote_layers == 0 or 3, widths [256,256,259]): vote_w_img[0] = image of rows 3.. of the first FC weight
 *     (K = 256; its xyz rows 0..2 are passed as vote_w0_xyz_f32 (3,256) f32 and applied as an fp32 rank-3 update);
 *     vote_w_img[2] / vote_bias[2] = last FC with its 259 output columns permuted [3..258, 0..2];
 *     seeds_xyz (b*n,3) -> votes_xyz (b*n,3) = seeds_xyz + offset[:, :3], votes_feat (b*n,256) = fp_out + offset[:, 3:].
 *     The residual never leaves the SM (it is parked in tensor memory), so with the voting module fused fp_out may be NULL
 *     when the caller does not need the seed features.
 * points1, points2, fp_out and votes_feat must be 32-byte aligned (256-bit global accesses).
 * The pointer arrays are HOST arrays of device pointers (read during the call).  Tensor-core path only (fp16
 * operands, fp32 accumulation); other widths return VNB_ERR_INVALID (use the unfused entry points). */
int vnb_fp_module_fused(int b, int n, int m, int c1, int c2, const float* dist, const int* idx, const float* points1,
                        const float* points2, int n_fp_layers, const void* const* fp_w_img,
                        const float* const* fp_bias, const int* fp_cout, float* fp_out, int n_vote_layers,
                        const void* const* vote_w_img, const float* const* vote_bias, const int* vote_cout,
                        const float* vote_w0_xyz_f32, const float* seeds_xyz, float* votes_xyz, float* votes_feat,
                        void* stream);

/* ------------------------------------------------------------------ evaluator -- */

/* iou_3d(bbox1, bbox2)                                            evaluator.py:26-39
 * boxes_a, boxes_b (n,8,3) f32 corner boxes (corners 0..3 = top face, y of corner 0 > y of corner 4) -> out (n) f64:
 * BEV intersection area of the two top-face quadrilaterals (exact convex clipping in double precision; the reference
 * calls shapely) x height overlap / union of volumes. */
int vnb_iou3d_pairs(int n, const float* boxes_a, const float* boxes_b, double* out, void* stream);

/* eval_det_cls(pred, gt, ovthresh)                                evaluator.py:77-151 (+ voc_ap :42-74)
 * One class.  Detections: det_boxes (nd,8,3), det_scores (nd), det_img (nd) i32 image index in [0, nimg);
 * ground truth: gt_boxes (ng,8,3) grouped by image, gt_offsets (nimg+1) i32 (image i owns rows gt_offsets[i] ..
 * gt_offsets[i+1]-1).  Outputs, all f64 and in descending-confidence order (equal confidences in input order):
 * rec (nd), prec (nd), ap (1).  npos = ng.  workspace: vnb_eval_det_cls_workspace_bytes(nd, ng). */
size_t vnb_eval_det_cls_workspace_bytes(int nd, int ng);
int vnb_eval_det_cls(int nd, int ng, int nimg, const float* det_boxes, const float* det_scores, const int* det_img,
                     const float* gt_boxes, const int* gt_offsets, double ovthresh, double* rec, double* prec,
                     double* ap, void* workspace, void* stream);

/* ------------------------------------------------------------------ input stage */

/* Point subsampling + upright-depth -> upright-camera axis flip + training augmentation, fused.
 * dataset.py:185-190 (subsample, flip_axis_to_camera sunutils.py:70-77), :219-231,302-308 (x / z flips, rotation about
 * y, scale).  raw_upright_depth (b,n_raw,3) f32; choice (b,n) i32 indices into the raw cloud (NULL: the first n points);
 * per cloud and all optional (NULL = skip): flip_x, flip_z (b) u8, roty_angle (b) f64 radians, scale (b) f64 — the
 * host's random draws.  -> xyz (b,n,3) f32 and, if not NULL, height (b,n) f32 = floor_y - y (the height-above-floor
 * input feature BASELINE.json names).  Arithmetic in double, one rounding to float. */
int vnb_prepare_input(int b, int n_raw, int n, const float* raw_upright_depth, const int* choice,
                      const unsigned char* flip_x, const unsigned char* flip_z, const double* roty_angle,
                      const double* scale, float* xyz, float* height, double floor_y, void* stream);

/* ------------------------------------------------------------------ losses ---- */

/* Label assignment + training losses (forward values)              model.py:62-84, 141-231
 * seeds_xyz, votes_xyz (b,n_seed,3); proposals_xyz (b,n_prop,3); proposals_output (b,n_prop,79); dense ground truth as
 * the reference feeds it: bboxes_xyz, bboxes_lwh (b,n_box,3), bboxes_roty (b,n_box), semantic / heading / size labels
 * (b,n_box) i32, heading_residuals (b,n_box), size_residuals (b,n_box,3); thresholds config.py:4-5 (0.3 / 0.6).
 * -> out14 (14) f64: total_cost, vote_reg_loss, obj_cls_loss, box_loss, center_loss (incl. the dual term),
 *    heading_cls_loss, heading_residual_loss, size_cls_loss, size_residual_loss, sem_cls_loss, obj_accuracy,
 *    sem_accuracy, number of positive proposals, number of negative proposals.  Means over empty sets are NaN (as
 *    tf.reduce_mean).  workspace: 128 bytes. */
int vnb_votenet_losses(int b, int n_seed, int n_prop, int n_box, const float* seeds_xyz, const float* votes_xyz,
                       const float* proposals_xyz, const float* proposals_output, const float* bboxes_xyz,
                       const float* bboxes_lwh, const float* bboxes_roty, const int* semantic_labels,
                       const int* heading_labels, const float* heading_residuals, const int* size_labels,
                       const float* size_residuals, float positive_thres, float negative_thres, double* out14,
                       void* workspace_128_bytes, void* stream);

/* ------------------------------------------------------------------ backward ops ---- */

/* scatteraddpointLauncher(b,n,m,out_g,idx,inp_g)              tf_ops/sampling/tf_sampling_g.cu:183-192,209-211
 * (op GatherPointGrad, tf_sampling.cpp:150-178; registered as the gradient of GatherPoint, tf_sampling.py:43-47)
 * out_g (b,m,3), idx (b,m) -> inp_g (b,n,3) = scatter-add; inp_g is zeroed first (tf_sampling.cpp:174). */
int vnb_gather_point_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g, void* stream);

/* groupPointGradLauncher(b,n,c,m,nsample,grad_out,idx,grad_points)   tf_ops/grouping/tf_grouping_g.cu:61-78,137-141
 * (op GroupPointGrad, tf_grouping.cpp:173-208; gradient of GroupPoint, tf_grouping.py:42-46)
 * grad_out (b,m,nsample,c), idx (b,m,nsample) -> grad_points (b,n,c), zeroed first (tf_grouping.cpp:203). */
int vnb_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                         float* grad_points, void* stream);

/* threeinterpolate_grad_cpu(b,n,c,m,grad_out,idx,weight,grad_points)  tf_ops/3d_interpolation/tf_interpolate.cpp:131-153
 * (op ThreeInterpolateGrad, :225-262, CPU-only in the reference; gradient of ThreeInterpolate, tf_interpolate.py:29-34)
 * grad_out (b,n,c), idx (b,n,3), weight (b,n,3) -> grad_points (b,m,c), zeroed first (:255).
 * All three accumulate with float atomics (as the reference's GPU kernels do): equal to a sequential sum up to the
 * rounding of a different summation order. */
int vnb_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int* idx, const float* weight,
                               float* grad_points, void* stream);

/* Backward of the fused set-abstraction layer (vnb_sa_group_mlp_max): the gradients TensorFlow autodiff produces for
 * utils.py:49-55,120-132 (group -> 3 x (1x1 conv + ReLU) -> reduce_max), including the GroupPointGrad scatter
 * (tf_grouping_g.cu:61-78), in one kernel that rematerialises the forward per centroid in fp32 (csrc/sa_backward.cu).
 * w1 (3+c,c1), w2 (c1,c2), w3 (c2,c3) BN-folded, row-major; w1_t (c1,3+c), w2_t (c2,c1) their transposes;
 * grad_out (b,m,c3).  Every grad_* output must be ZEROED by the caller and is accumulated into (float reductions):
 * grad_feat (b,n,c), grad_xyz (b,n,3), grad_new_xyz (b,m,3) — any of the three may be NULL — and grad_w1..3 / grad_b1..3 shaped like the weights. */
int vnb_sa_group_mlp_max_backward(int b, int n, int c, int m, int nsample, const float* xyz, const float* feat,
                                  const float* new_xyz, const int* idx, int c1, int c2, int c3, const float* w1,
                                  const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                                  const float* w1_t, const float* w2_t, const float* grad_out, float* grad_feat,
                                  float* grad_xyz, float* grad_new_xyz, float* grad_w1, float* grad_b1, float* grad_w2,
                                  float* grad_b2, float* grad_w3, float* grad_b3, void* stream);

/* row-wise concat / split helpers: out (rows, ca+cb) = [a (rows,ca), b (rows,cb)]  and the inverse */
int vnb_concat2(int rows, int ca, int cb, const float* a, const float* b, float* out, void* stream);
int vnb_split2(int rows, int ca, int cb, const float* in, float* a, float* b, void* stream);

/* box decode                                                    model.py:100-129 (+ dataset.py:36-49 mean sizes)
 * proposals_xyz (b,k,3), proposals_output (b,k,79), class_mean_size (10,3)
 * -> bboxes (b,k,8,3), scores (b,k) = max class logit (model.py:133), objectness (b,k,2), class_scores (b,k,10). */
int vnb_decode_boxes(int b, int k, const float* proposals_xyz, const float* proposals_output,
                     const float* class_mean_size, float* bboxes, float* scores, float* objectness,
                     float* class_scores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VOTENET_B200_H_ */

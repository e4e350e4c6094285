"""VoteNet inference tower on B200 — the graph of /root/reference/model.py:34-61,85-137 wired over the B200 ops.

`VoteNetB200.forward` is the readable, allocation-per-call form used by the parity tests (every intermediate is
returned); `votenet_b200.engine.Engine` is the pre-allocated, multi-stream, CUDA-graph form used by bench.py.
"""
import numpy as np
import torch

from . import tf_nms3d
from ._lib import check, dptr, lib, stream_ptr
from .config import NC, PROPOSAL_CHANNELS, VoteNetConfig
from .synth import CLASS_MEAN_SIZE
from . import tf_interpolate, utils
from .utils import (PRECISION_TENSOR, WeightStore, fp_fusable, fp_module_fused, linear, pointnet_fp_module,
                    pointnet_sa_module, vote_layers_fused)


def decode_boxes(proposals_xyz, proposals_output, class_mean_size):
    """model.py:100-129 + the score / objectness slices NMS consumes (model.py:133)."""
    b, k, ch = proposals_output.shape
    assert ch == PROPOSAL_CHANNELS
    dev = proposals_output.device
    bboxes = torch.empty((b, k, 8, 3), dtype=torch.float32, device=dev)
    scores = torch.empty((b, k), dtype=torch.float32, device=dev)
    objectness = torch.empty((b, k, 2), dtype=torch.float32, device=dev)
    class_scores = torch.empty((b, k, NC), dtype=torch.float32, device=dev)
    check(lib.vnb_decode_boxes(b, k, dptr(proposals_xyz, torch.float32), dptr(proposals_output, torch.float32),
                               dptr(class_mean_size, torch.float32), dptr(bboxes), dptr(scores), dptr(objectness),
                               dptr(class_scores), stream_ptr()))
    return bboxes, scores, objectness, class_scores


def decode_nms3d(proposals_xyz, proposals_output, class_mean_size, iou_threshold):
    """model.py:100-137 in one kernel (csrc/nms3d.cu): box decode, NMS3D and the output gathers.  -> dict with
    dec_bboxes (B,K,8,3), dec_scores (B,K), dec_objectness (B,K,2), dec_class_scores (B,K,10), nms_keep (B,K) u8,
    nms_idx (B*K,2) rows (batch, box) in global descending-score order, nms_key (B*K,) order-preserving score keys,
    nms_count (1,), and bboxes_pred (B*K,8,3) / class_scores_pred (B*K,10) / batch_idx (B*K,) (first nms_count rows)."""
    b, k, _ = proposals_xyz.shape
    dev = proposals_xyz.device
    f32, i32 = torch.float32, torch.int32
    E = lambda shape, dt=f32: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
    o = dict(dec_bboxes=E((b, k, 8, 3)), dec_scores=E((b, k)), dec_objectness=E((b, k, 2)), dec_class_scores=E((b, k, NC)),
             nms_keep=E((b, k), torch.uint8), nms_idx=torch.zeros((max(b * k, 1), 2), dtype=i32, device=dev),
             nms_key=torch.zeros((max(b * k, 1),), dtype=i32, device=dev), nms_count=torch.zeros((1,), dtype=i32, device=dev),
             bboxes_pred=E((max(b * k, 1), 8, 3)), class_scores_pred=E((max(b * k, 1), NC)),
             batch_idx=E((max(b * k, 1),), i32))
    ws = torch.empty((lib.vnb_nms3d_workspace_bytes(b, k),), dtype=torch.uint8, device=dev)
    check(lib.vnb_decode_nms3d(b, k, dptr(proposals_xyz, f32), dptr(proposals_output, f32), dptr(class_mean_size, f32),
                               float(iou_threshold), dptr(o["dec_bboxes"]), dptr(o["dec_scores"]), dptr(o["dec_objectness"]),
                               dptr(o["dec_class_scores"]), dptr(o["nms_keep"]), dptr(o["nms_idx"]), dptr(o["nms_key"]),
                               dptr(o["nms_count"]), dptr(o["bboxes_pred"]), dptr(o["class_scores_pred"]),
                               dptr(o["batch_idx"]), dptr(ws), stream_ptr()))
    return o


class VoteNetB200:
    def __init__(self, cfg: VoteNetConfig, weights, device="cuda", precision=PRECISION_TENSOR):
        self.cfg = cfg
        self.device = torch.device(device)
        self.store = WeightStore(weights, device=self.device, eps=cfg.bn_eps, precision=precision)
        self.class_mean_size = torch.as_tensor(CLASS_MEAN_SIZE, device=self.device).contiguous()

    @torch.no_grad()
    def forward(self, xyz, feats, run_nms=True):
        """xyz (B,N,3), feats (B,N,C) CUDA f32 -> dict of every intermediate + detections."""
        cfg, W = self.cfg, self.store
        prec = W.precision
        out = {}
        l_xyz, l_pts = [xyz], [feats]
        for li, sa in enumerate(cfg.sa):                                                      # model.py:39-46
            nx, npts, idx = pointnet_sa_module(l_xyz[-1], l_pts[-1], sa.npoint, sa.radius, sa.nsample, list(sa.mlp),
                                               None, False, f"sa{li + 1}", weights=W)
            l_xyz.append(nx); l_pts.append(npts)
            out[f"sa{li + 1}_xyz"], out[f"sa{li + 1}_points"], out[f"sa{li + 1}_idx"] = nx, npts, idx
        l3_points = pointnet_fp_module(l_xyz[3], l_xyz[4], l_pts[3], l_pts[4], list(cfg.fp_mlp), "fp1", weights=W)   # :48
        seeds_xyz = l_xyz[2]                                                                  # :50
        b, ns, _ = seeds_xyz.shape
        cf = cfg.seed_feat_dim
        votes_xyz = torch.empty((b, ns, 3), dtype=torch.float32, device=xyz.device)
        votes_points = torch.empty((b, ns, cf), dtype=torch.float32, device=xyz.device)
        if (utils.FUSE_FP and fp_fusable(l_pts[2], l3_points, cfg.fp_mlp, prec) and tuple(cfg.vote_units) == (256, 256, 259)):
            # fp2 (:49) and the voting module (:53-61) in ONE kernel (csrc/fp_chain.cu)
            dist, idx3 = tf_interpolate.three_nn(l_xyz[2], l_xyz[3])
            seeds_points = torch.empty((b, ns, cf), dtype=torch.float32, device=xyz.device)
            vl, x0 = vote_layers_fused(W, [f"voting{i}" for i in range(3)])
            fp_module_fused(dist, idx3, l_pts[2].contiguous(), l3_points.contiguous(),
                            [W.layer(f"fp2/conv_{i}") for i in range(2)], seeds_points,
                            vote=(vl, x0, seeds_xyz.contiguous(), votes_xyz, votes_points))
            out["votes"] = torch.cat([votes_xyz, votes_points], 2)
        else:
            seeds_points = pointnet_fp_module(l_xyz[2], l_xyz[3], l_pts[2], l3_points, list(cfg.fp_mlp), "fp2", weights=W)  # :49
            seeds = torch.empty((b * ns, 3 + cf), dtype=torch.float32, device=xyz.device)
            check(lib.vnb_concat2(b * ns, 3, cf, dptr(seeds_xyz), dptr(seeds_points), dptr(seeds), stream_ptr()))  # :53
            h = seeds
            nv = len(cfg.vote_units)
            for i in range(nv):                                                               # :55-56, residual :60
                h = linear(h, W.layer(f"voting{i}"), act=i < nv - 1, precision=prec, residual=seeds if i == nv - 1 else None)
            out["votes"] = h.reshape(b, ns, 3 + cf)
            check(lib.vnb_split2(b * ns, 3, cf, dptr(h), dptr(votes_xyz), dptr(votes_points), stream_ptr()))  # :61,:85
        out["fp1_points"], out["fp2_points"] = l3_points, seeds_points
        p = cfg.proposal
        prop_xyz, prop_out, pidx = pointnet_sa_module(votes_xyz, votes_points, p.npoint, p.radius, p.nsample,
                                                      list(p.mlp), list(p.mlp2), False, "proposal",
                                                      sample_xyz=seeds_xyz, weights=W)         # :89-93
        out["proposals_xyz"], out["proposals_output"], out["proposal_idx"] = prop_xyz, prop_out, pidx
        if not run_nms:
            bboxes, scores, objectness, class_scores = decode_boxes(prop_xyz, prop_out, self.class_mean_size)
            out.update(dec_bboxes=bboxes, dec_scores=scores, dec_objectness=objectness, dec_class_scores=class_scores)
            return out
        # decode (:100-129) + NMS3D (:133) + the three output gathers (:135-137) in one launch
        out.update(decode_nms3d(prop_xyz, prop_out, self.class_mean_size, cfg.nms_iou))
        return out

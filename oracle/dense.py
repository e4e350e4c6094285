"""TEST INFRASTRUCTURE — torch-CPU fp32 restatement of the reference's TensorFlow/Tensorpack glue.  NOT product code.

Restates, call for call:
  * sample_and_group / pointnet_sa_module   /root/reference/utils.py:25-61, 93-158   (max-pool, NHWC path)
  * pointnet_fp_module                      /root/reference/utils.py:266-294
  * voting module                           /root/reference/model.py:53-61
  * proposal module + decode + NMS          /root/reference/model.py:89-137 (+ dataset.py:36-49 mean sizes)

The dense arithmetic itself (Conv2D 1x1, FullyConnected, BatchNorm(EMA), ReLU, reduce_max) lives in TensorFlow 1.x /
Tensorpack, which are un-vendored, un-pinned dependencies absent from /root/reference and from this image, and the
reference has no tests pinning it => **parity unpinned** for this part.  Semantics restated from
the call sites with Tensorpack defaults: conv/FC have a bias; BNReLU = BatchNorm(eps=1e-5, EMA statistics at
inference) then ReLU; kernels are [Cin, Cout].
Index ops are served by the C oracle (oracle.ops).
"""
import numpy as np
import torch

from . import ops

_F = torch.float32


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=_F) if not torch.is_tensor(a) else a.to(_F)


def dense_layer(x, weights, name, eps=1e-5):
    """x (..., Cin) -> (..., Cout): affine (+ BatchNorm on EMA stats + ReLU when the layer has BN)."""
    y = x @ weights[f"{name}/W"] + weights[f"{name}/b"]
    if f"{name}/bn/gamma" in weights:
        inv = torch.rsqrt(weights[f"{name}/bn/variance/EMA"] + eps) * weights[f"{name}/bn/gamma"]
        y = (y - weights[f"{name}/bn/mean/EMA"]) * inv + weights[f"{name}/bn/beta"]
        y = torch.relu(y)
    return y


def sample_and_group(npoint, radius, nsample, xyz, points, sample_xyz=None):
    """utils.py:25-61 (knn=False, use_xyz=True)."""
    xyz_np = xyz.numpy()
    fps_idx = ops.farthest_point_sample(npoint, sample_xyz.numpy() if sample_xyz is not None else xyz_np)
    new_xyz = ops.gather_point(xyz_np, fps_idx)                       # :42-45
    idx, pts_cnt = ops.query_ball_point(radius, nsample, xyz_np, new_xyz)  # :49
    grouped_xyz = ops.group_point(xyz_np, idx)                         # :50
    grouped_xyz = grouped_xyz - new_xyz[:, :, None, :]                 # :51
    if points is not None:
        grouped_points = ops.group_point(points.numpy(), idx)          # :53
        new_points = np.concatenate([grouped_xyz, grouped_points], -1)  # :55  relative xyz FIRST
    else:
        new_points = grouped_xyz
    return _t(new_xyz), _t(new_points), idx, pts_cnt, fps_idx


def pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2, scope, weights, sample_xyz=None, eps=1e-5):
    """utils.py:93-158 -> (new_xyz (B,m,3), new_points (B,m,C), idx (B,m,ns), extras)."""
    new_xyz, new_points, idx, pts_cnt, fps_idx = sample_and_group(npoint, radius, nsample, xyz, points, sample_xyz)
    outs = []
    for b in range(new_points.shape[0]):  # per cloud to bound memory
        h = new_points[b]
        for i in range(len(mlp)):
            h = dense_layer(h, weights, f"{scope}/conv{i}", eps)       # :125-127
        h = h.max(dim=1).values                                        # :132
        if mlp2 is not None:
            for i in range(len(mlp2)):
                h = dense_layer(h, weights, f"{scope}/conv_post_{i}", eps)  # :149-155 (last layer has no BN/ReLU)
        outs.append(h)
    return new_xyz, torch.stack(outs, 0), idx, dict(pts_cnt=pts_cnt, fps_idx=fps_idx)


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, scope, weights, eps=1e-5):
    """utils.py:266-294."""
    dist, idx = ops.three_nn(xyz1.numpy(), xyz2.numpy())              # :278
    dist = torch.clamp_min(_t(dist), 1e-10)                            # :279
    norm = (1.0 / dist).sum(dim=2, keepdim=True)                       # :280
    weight = (1.0 / dist) / norm                                       # :282
    interpolated = _t(ops.three_interpolate(points2.numpy(), idx, weight.numpy()))  # :283
    new_points1 = torch.cat([interpolated, points1], dim=2) if points1 is not None else interpolated  # :286
    for i in range(len(mlp)):
        new_points1 = dense_layer(new_points1, weights, f"{scope}/conv_{i}", eps)   # :290-292
    return new_points1, dict(dist=dist, idx=idx, weight=weight, interpolated=interpolated)


def decode_boxes(proposals_xyz, proposals_output, class_mean_size, NH=12, NS=10, NC=10):
    """model.py:100-129 -> dict(bboxes (B,K,8,3), size, heading, center, scores (B,K), objectness (B,K,2))."""
    po = proposals_output
    cms = _t(class_mean_size)
    B, K = po.shape[:2]
    size_cls = torch.argmax(po[..., 5 + 2 * NH: 5 + 2 * NH + NS], dim=-1)                    # :115
    onehot = torch.nn.functional.one_hot(size_cls, NS).to(_F)                                 # :116
    size_res_all = po[..., 5 + 2 * NH + NS: 5 + 2 * NH + 4 * NS].reshape(B, K, NS, 3)
    size_residual = (onehot[..., None] * size_res_all).sum(dim=2)                             # :117-118
    size_pred = cms[size_cls] * torch.clamp_min(1 + size_residual, 1e-6)                      # :119
    center = proposals_xyz + po[..., 2:5]                                                     # :121
    heading_cls = torch.argmax(po[..., 5:5 + NH], dim=-1)                                     # :122
    h_onehot = torch.nn.functional.one_hot(heading_cls, NH).to(_F)
    heading_res = (h_onehot * po[..., 5 + NH:5 + 2 * NH]).sum(dim=2)                          # :124-125
    pi32 = torch.tensor(np.pi, dtype=_F)
    heading = torch.remainder((heading_cls.to(_F) * 2 + heading_res) * pi32 / float(NH), 2 * pi32)  # :126
    c, s = torch.cos(heading), torch.sin(heading)                                             # :103-104
    z, o = torch.zeros_like(c), torch.ones_like(c)
    rot = torch.stack([c, z, s, z, o, z, -s, z, c], -1).reshape(B, K, 3, 3)                   # :107
    l, w, h = size_pred[..., 0], size_pred[..., 1], size_pred[..., 2]                         # :108 lwh(xzy)
    corners = torch.stack([l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2,
                           h / 2, h / 2, h / 2, h / 2, -h / 2, -h / 2, -h / 2, -h / 2,
                           w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2], -1).reshape(B, K, 3, 8)
    bboxes = torch.einsum("ijkl,ijlm->ijmk", rot, corners) + center[:, :, None, :]            # :112
    return dict(bboxes=bboxes, size=size_pred, heading=heading, center=center,
                size_cls=size_cls, heading_cls=heading_cls,
                scores=po[..., -NC:].max(dim=-1).values, objectness=po[..., :2], class_scores=po[..., -NC:])


def votenet_forward(xyz, feats, weights, cfg, class_mean_size, run_nms=True):
    """Inference tower, model.py:34-61,85-137.  xyz (B,N,3), feats (B,N,C) -> dict of every intermediate."""
    xyz, feats = _t(xyz), _t(feats)
    eps = cfg.bn_eps
    out = {}
    l_xyz, l_pts = [xyz], [feats]
    for li, sa in enumerate(cfg.sa):                                                          # :39-46
        nx, npts, idx, ex = pointnet_sa_module(l_xyz[-1], l_pts[-1], sa.npoint, sa.radius, sa.nsample, sa.mlp, None,
                                               f"sa{li + 1}", weights, eps=eps)
        l_xyz.append(nx); l_pts.append(npts)
        out[f"sa{li + 1}_xyz"], out[f"sa{li + 1}_points"], out[f"sa{li + 1}_idx"] = nx, npts, idx
        out[f"sa{li + 1}_fps"], out[f"sa{li + 1}_cnt"] = ex["fps_idx"], ex["pts_cnt"]
    l3_points, ex = pointnet_fp_module(l_xyz[3], l_xyz[4], l_pts[3], l_pts[4], cfg.fp_mlp, "fp1", weights, eps)  # :48
    out["fp1_points"], out["fp1_nn_idx"] = l3_points, ex["idx"]
    seeds_points, ex = pointnet_fp_module(l_xyz[2], l_xyz[3], l_pts[2], l3_points, cfg.fp_mlp, "fp2", weights, eps)  # :49
    out["fp2_points"], out["fp2_nn_idx"] = seeds_points, ex["idx"]
    seeds_xyz = l_xyz[2]                                                                      # :50
    seeds = torch.cat([seeds_xyz, seeds_points], 2)                                           # :53
    off = seeds.reshape(-1, seeds.shape[-1])
    for i in range(len(cfg.vote_units)):                                                      # :55-56
        off = dense_layer(off, weights, f"voting{i}", eps)
    votes = seeds + off.reshape(seeds.shape)                                                  # :60
    out["votes"] = votes
    votes_xyz, votes_points = votes[:, :, :3].contiguous(), votes[:, :, 3:].contiguous()      # :61,:85
    p = cfg.proposal
    prop_xyz, prop_out, pidx, ex = pointnet_sa_module(votes_xyz, votes_points, p.npoint, p.radius, p.nsample, p.mlp,
                                                      p.mlp2, "proposal", weights, sample_xyz=seeds_xyz, eps=eps)  # :89-93
    out["proposals_xyz"], out["proposals_output"], out["proposal_idx"] = prop_xyz, prop_out, pidx
    out["proposal_fps"] = ex["fps_idx"]
    dec = decode_boxes(prop_xyz, prop_out, class_mean_size)
    out.update({f"dec_{k}": v for k, v in dec.items()})
    if run_nms:
        nms_idx, keep = ops.NMS3D(dec["bboxes"].numpy(), dec["scores"].numpy(), dec["objectness"].numpy(),
                                  cfg.nms_iou, return_keep=True)                              # :133
        out["nms_idx"], out["nms_keep"] = nms_idx, keep
    return out

"""Trainable form of the set-abstraction and feature-propagation modules — SURVEY.md §8(f) rank 1.

The reference trains through TensorFlow autodiff: the dense layers (Conv2D 1x1 / BN / ReLU / reduce_max, utils.py:120-132,
286-292) get TensorFlow's own gradients, and only the custom index ops carry registered gradients — GatherPointGrad
(tf_sampling.py:43-47), GroupPointGrad (tf_grouping.py:42-46), ThreeInterpolateGrad (tf_interpolate.py:29-34).  This module
is the same split: the index ops run the B200 kernels forward AND backward (csrc/grad_ops.cu, registered on
tf_sampling.gather_point / tf_grouping.group_point / tf_interpolate.three_interpolate as torch.autograd.Functions), the
dense layers are differentiable torch ops (cuBLAS), wired exactly like utils.py:25-61,93-158,266-294.  Sampling and
neighbour search (FPS, ball query, three_nn) are index outputs and carry no gradient, as in the reference
(tf_sampling.py:57 NoGradient, tf_grouping.py:21).

The fused tensor-core kernels (csrc/sa_ws2.cu, sa1_ws2.cu, fp_chain.cu) remain the inference path; `test_train.py` checks
that this trainable graph computes the same forward (fp32) and that its gradients match a float64 CPU autograd
restatement built on the oracle's indices, plus the reference's own gradient-error tests end to end."""
import torch

from . import tf_grouping, tf_interpolate, tf_sampling


def sample_and_group(npoint, radius, nsample, xyz, points, sample_xyz=None):
    """utils.py:25-61 (knn=False, use_xyz=True): -> new_xyz (B,m,3), new_points (B,m,ns,3+C), idx, grouped_xyz.
    Differentiable w.r.t. `points` and `xyz` through group_point / gather_point."""
    fps_in = sample_xyz if sample_xyz is not None else xyz
    with torch.no_grad():
        fps = tf_sampling.farthest_point_sample_nested if fps_in.shape[1] <= 4096 else tf_sampling.farthest_point_sample
        fps_idx = fps(npoint, fps_in.detach())
    new_xyz = tf_sampling.gather_point(xyz, fps_idx)                                         # :42-45
    with torch.no_grad():
        idx, _ = tf_grouping.query_ball_point(radius, nsample, xyz.detach(), new_xyz.detach())   # :49
    grouped_xyz = tf_grouping.group_point(xyz, idx) - new_xyz.unsqueeze(2)                   # :50-51
    if points is not None:
        new_points = torch.cat([grouped_xyz, tf_grouping.group_point(points, idx)], -1)      # :53-55
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def _dense(x, W, b, act):
    y = x @ W + b
    return torch.relu(y) if act else y


def pointnet_sa_module_trainable(xyz, points, npoint, radius, nsample, layers, mlp2_layers=None, sample_xyz=None):
    """utils.py:93-158 (max pooling, use_xyz).  `layers` / `mlp2_layers`: lists of (W (Cin,Cout), b (Cout,)) tensors — the
    BN-folded affine maps, possibly requiring grad.  -> (new_xyz, new_points (B,m,C), idx)."""
    new_xyz, h, idx, _ = sample_and_group(npoint, radius, nsample, xyz, points, sample_xyz)
    for W, b in layers:                                   # :120-127 1x1 conv + BN + ReLU
        h = _dense(h, W, b, True)
    h = h.max(dim=2).values                               # :132 reduce_max over nsample
    if mlp2_layers:                                       # :149-155 (last layer linear)
        for i, (W, b) in enumerate(mlp2_layers):
            h = _dense(h, W, b, i < len(mlp2_layers) - 1)
    return new_xyz, h, idx


def pointnet_fp_module_trainable(xyz1, xyz2, points1, points2, layers):
    """utils.py:266-294: three_nn -> inverse-distance weights -> three_interpolate -> concat skip -> dense layers."""
    with torch.no_grad():
        dist, idx = tf_interpolate.three_nn(xyz1.detach(), xyz2.detach())                    # :278
        dist = torch.clamp(dist, min=1e-10)                                                  # :279
        norm = (1.0 / dist).sum(dim=2, keepdim=True)                                         # :280-281
        weight = ((1.0 / dist) / norm).contiguous()                                          # :282
    h = tf_interpolate.three_interpolate(points2, idx, weight)                               # :283
    if points1 is not None:
        h = torch.cat([h, points1], dim=2)                                                   # :286
    for W, b in layers:                                                                      # :291-292
        h = _dense(h, W, b, True)
    return h


# ---------------------------------------------------------------------------------------------------------------------
# The set-abstraction layer as ONE differentiable op: fused forward kernel + fused backward kernel (csrc/sa_backward.cu)
class _FusedSAFn(torch.autograd.Function):
    """out (B,m,C3) = max_r relu(relu(relu([xyz[idx]-new_xyz, feat[idx]] W1+b1) W2+b2) W3+b3) with hand-written forward AND
    backward kernels: forward = the fp32 fused kernel (vnb_sa_group_mlp_max, precision 0); backward rematerialises the
    forward per centroid and returns gradients for feat, xyz, new_xyz and the six parameters in one launch."""

    @staticmethod
    def forward(ctx, xyz, feat, new_xyz, idx, W1, b1, W2, b2, W3, b3):
        from ._lib import check, dptr, lib, stream_ptr

        b, n, _ = xyz.shape
        c, m, ns = feat.shape[2], idx.shape[1], idx.shape[2]
        args = [t.detach().contiguous() for t in (xyz, feat, new_xyz, W1, b1, W2, b2, W3, b3)]
        xyz_, feat_, new_xyz_, W1_, b1_, W2_, b2_, W3_, b3_ = args
        out = torch.empty((b, m, W3_.shape[1]), dtype=torch.float32, device=xyz.device)
        check(lib.vnb_sa_group_mlp_max(b, n, c, m, ns, dptr(xyz_), dptr(feat_), dptr(new_xyz_), dptr(idx), W1_.shape[1],
                                       W2_.shape[1], W3_.shape[1], dptr(W1_), dptr(b1_), dptr(W2_), dptr(b2_), dptr(W3_), dptr(b3_),
                                       None, None, None, None, dptr(out), 0, None, stream_ptr()))
        ctx.save_for_backward(xyz_, feat_, new_xyz_, idx, W1_, b1_, W2_, b2_, W3_, b3_)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from ._lib import check, dptr, lib, stream_ptr

        xyz, feat, new_xyz, idx, W1, b1, W2, b2, W3, b3 = ctx.saved_tensors
        b, n, _ = xyz.shape
        c, m, ns = feat.shape[2], idx.shape[1], idx.shape[2]
        Z = torch.zeros_like
        g_feat, g_xyz, g_new = Z(feat), Z(xyz), Z(new_xyz)
        gW1, gb1, gW2, gb2, gW3, gb3 = Z(W1), Z(b1), Z(W2), Z(b2), Z(W3), Z(b3)
        # named, so that they outlive the launch (a temporary freed at once would be handed to the next allocation)
        W1t, W2t, gout = W1.t().contiguous(), W2.t().contiguous(), grad_out.contiguous()
        check(lib.vnb_sa_group_mlp_max_backward(b, n, c, m, ns, dptr(xyz), dptr(feat), dptr(new_xyz), dptr(idx), W1.shape[1],
                                                W2.shape[1], W3.shape[1], dptr(W1), dptr(b1), dptr(W2), dptr(b2), dptr(W3),
                                                dptr(b3), dptr(W1t), dptr(W2t),
                                                dptr(gout), dptr(g_feat), dptr(g_xyz), dptr(g_new), dptr(gW1),
                                                dptr(gb1), dptr(gW2), dptr(gb2), dptr(gW3), dptr(gb3), stream_ptr()))
        return g_xyz, g_feat, g_new, None, gW1, gb1, gW2, gb2, gW3, gb3


def pointnet_sa_module_fused_trainable(xyz, points, npoint, radius, nsample, layers, sample_xyz=None):
    """pointnet_sa_module (utils.py:93-158, three-layer mlp, max pooling) as fused forward + fused backward kernels.
    `layers` = [(W1,b1),(W2,b2),(W3,b3)].  -> (new_xyz, new_points (B,m,C3), idx); differentiable w.r.t. points, xyz and
    the parameters."""
    fps_in = sample_xyz if sample_xyz is not None else xyz
    with torch.no_grad():
        fps = tf_sampling.farthest_point_sample_nested if fps_in.shape[1] <= 4096 else tf_sampling.farthest_point_sample
        fps_idx = fps(npoint, fps_in.detach())
    new_xyz = tf_sampling.gather_point(xyz, fps_idx)
    with torch.no_grad():
        idx, _ = tf_grouping.query_ball_point(radius, nsample, xyz.detach(), new_xyz.detach())
    (W1, b1), (W2, b2), (W3, b3) = layers
    out = _FusedSAFn.apply(xyz, points, new_xyz, idx, W1, b1, W2, b2, W3, b3)
    return new_xyz, out, idx

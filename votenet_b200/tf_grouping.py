"""Drop-in for the reference's tf_ops/grouping/tf_grouping.py over CUDA tensors.  select_top_k / knn_point are off
the VoteNet path (knn=False, utils.py:46-49) and not provided."""
import torch

from ._lib import check, dptr, lib, stream_ptr


def query_ball_point(radius, nsample, xyz1, xyz2):
    """xyz1 (B,N,3) searched set, xyz2 (B,M,3) queries -> (idx (B,M,nsample) i32, pts_cnt (B,M) i32).
    Reference: tf_grouping.py:8-20 (QueryBallPoint, tf_grouping.cpp:67-106).  As in the reference, rows of empty
    balls are never written; the output buffer is zero-initialised here so they read as 0."""
    if xyz1.dim() != 3 or xyz1.shape[2] != 3:
        raise ValueError("QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.")  # tf_grouping.cpp:79
    if xyz2.dim() != 3 or xyz2.shape[2] != 3:
        raise ValueError("QueryBallPoint expects (batch_size, npoint, 3) xyz2 shape.")    # tf_grouping.cpp:84
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = torch.zeros((b, m, int(nsample)), dtype=torch.int32, device=xyz1.device)
    cnt = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
    # large searched sets go through the grid + index-bitmap kernel (bit-identical outputs); it needs scratch memory
    ws = None
    if n >= 1024:
        ws = torch.empty((lib.vnb_query_ball_point_workspace_bytes(b, n),), dtype=torch.uint8, device=xyz1.device)
    check(lib.vnb_query_ball_point_ws(b, n, m, float(radius), int(nsample), dptr(xyz1, torch.float32, "xyz1"),
                                      dptr(xyz2, torch.float32, "xyz2"), dptr(idx), dptr(cnt), dptr(ws), stream_ptr()))
    return idx, cnt


class _GroupPointFn(torch.autograd.Function):
    """GroupPoint with its registered gradient (@tf.RegisterGradient('GroupPoint'), tf_grouping.py:42-46)."""

    @staticmethod
    def forward(ctx, points, idx):
        ctx.save_for_backward(points, idx)
        return _group_point_fwd(points, idx)

    @staticmethod
    def backward(ctx, grad_out):
        points, idx = ctx.saved_tensors
        return group_point_grad(points, idx, grad_out.contiguous()), None


def group_point(points, idx):
    """points (B,N,C) f32, idx (B,M,S) i32 -> (B,M,S,C) f32.   Reference: tf_grouping.py:33-41 (tf_grouping.cpp:143-171).
    Differentiable w.r.t. points (GroupPointGrad) when points requires grad."""
    if torch.is_grad_enabled() and points.requires_grad:
        return _GroupPointFn.apply(points, idx)
    return _group_point_fwd(points, idx)


def _group_point_fwd(points, idx):
    if points.dim() != 3:
        raise ValueError("GroupPoint expects (batch_size, num_points, channel) points shape")  # tf_grouping.cpp:149
    if idx.dim() != 3 or idx.shape[0] != points.shape[0]:
        raise ValueError("GroupPoint expects (batch_size, npoints, nsample) idx shape")        # tf_grouping.cpp:155
    b, n, c = points.shape
    _, m, s = idx.shape
    out = torch.empty((b, m, s, c), dtype=torch.float32, device=points.device)
    check(lib.vnb_group_point(b, n, c, m, s, dptr(points, torch.float32, "points"), dptr(idx, torch.int32, "idx"),
                              dptr(out), stream_ptr()))
    return out


def group_point_grad(points, idx, grad_out):
    """Gradient of group_point w.r.t. points: (B,N,C) <- scatter-add of grad_out (B,M,S,C) at idx (B,M,S).
    Reference: GroupPointGrad, tf_grouping.py:42-46 / tf_grouping.cpp:173-208 / tf_grouping_g.cu:61-78."""
    if points.dim() != 3:
        raise ValueError("GroupPointGrad expects (batch_size, num_points, channel) points shape")   # tf_grouping.cpp:179
    if idx.dim() != 3 or idx.shape[0] != points.shape[0]:
        raise ValueError("GroupPointGrad expects (batch_size, npoints, nsample) idx shape")         # tf_grouping.cpp:185
    b, n, c = points.shape
    _, m, s = idx.shape
    if grad_out.dim() != 4 or tuple(grad_out.shape) != (b, m, s, c):
        raise ValueError("GroupPointGrad expects (batch_size, npoints, nsample, channel) grad_out shape")  # :190
    grad_points = torch.empty((b, n, c), dtype=torch.float32, device=points.device)
    check(lib.vnb_group_point_grad(b, n, c, m, s, dptr(grad_out, torch.float32, "grad_out"), dptr(idx, torch.int32, "idx"),
                                   dptr(grad_points), stream_ptr()))
    return grad_points

"""Drop-in for the reference's tf_ops/3d_interpolation/tf_interpolate.py over CUDA tensors (the reference runs these
two ops on ONE CPU thread with a device<->host bounce, tf_interpolate.cpp:187,222)."""
import torch

from ._lib import check, dptr, lib, stream_ptr


def three_nn(xyz1, xyz2):
    """xyz1 (b,n,3) unknown, xyz2 (b,m,3) known -> (dist (b,n,3) SQUARED f32, idx (b,n,3) i32).
    Reference: tf_interpolate.py:8-17 (ThreeNN, tf_interpolate.cpp:60-103,157-187)."""
    if xyz1.dim() != 3 or xyz1.shape[2] != 3:
        raise ValueError("ThreeNN expects (b,n,3) xyz1 shape.")  # tf_interpolate.cpp:163
    if xyz2.dim() != 3 or xyz2.shape[2] != 3:
        raise ValueError("ThreeNN expects (b,m,3) xyz2 shape.")  # tf_interpolate.cpp:168
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
    idx = torch.empty((b, n, 3), dtype=torch.int32, device=xyz1.device)
    check(lib.vnb_three_nn(b, n, m, dptr(xyz1, torch.float32, "xyz1"), dptr(xyz2, torch.float32, "xyz2"), dptr(dist),
                           dptr(idx), stream_ptr()))
    return dist, idx


class _ThreeInterpolateFn(torch.autograd.Function):
    """ThreeInterpolate with its registered gradient (@tf.RegisterGradient('ThreeInterpolate'), tf_interpolate.py:29-34:
    gradient w.r.t. points only, None for idx and weight)."""

    @staticmethod
    def forward(ctx, points, idx, weight):
        ctx.save_for_backward(points, idx, weight)
        return _three_interpolate_fwd(points, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        points, idx, weight = ctx.saved_tensors
        return three_interpolate_grad(points, idx, weight, grad_out.contiguous()), None, None


def three_interpolate(points, idx, weight):
    """points (b,m,c), idx (b,n,3) i32, weight (b,n,3) -> (b,n,c).
    Reference: tf_interpolate.py:19-28 (ThreeInterpolate, tf_interpolate.cpp:107-127,191-222).
    Differentiable w.r.t. points (ThreeInterpolateGrad) when points requires grad."""
    if torch.is_grad_enabled() and points.requires_grad:
        return _ThreeInterpolateFn.apply(points, idx, weight)
    return _three_interpolate_fwd(points, idx, weight)


def _three_interpolate_fwd(points, idx, weight):
    if points.dim() != 3:
        raise ValueError("ThreeInterpolate expects (b,m,c) points shape")  # tf_interpolate.cpp:197
    b, m, c = points.shape
    if idx.dim() != 3 or idx.shape[0] != b or idx.shape[2] != 3:
        raise ValueError("ThreeInterpolate expects (b,n,3) idx shape")     # tf_interpolate.cpp:203
    n = idx.shape[1]
    if tuple(weight.shape) != (b, n, 3):
        raise ValueError("ThreeInterpolate expects (b,n,3) weight shape")  # tf_interpolate.cpp:206
    out = torch.empty((b, n, c), dtype=torch.float32, device=points.device)
    check(lib.vnb_three_interpolate(b, m, c, n, dptr(points, torch.float32, "points"), dptr(idx, torch.int32, "idx"),
                                    dptr(weight, torch.float32, "weight"), dptr(out), stream_ptr()))
    return out


def three_interpolate_grad(points, idx, weight, grad_out):
    """Gradient of three_interpolate w.r.t. points: (b,m,c) <- sum over the 3 neighbours of grad_out (b,n,c) * weight.
    Reference: ThreeInterpolateGrad, tf_interpolate.py:29-34 / tf_interpolate.cpp:131-153,225-262 (one CPU thread)."""
    if points.dim() != 3:
        raise ValueError("ThreeInterpolateGrad expects (b,m,c) points shape")      # tf_interpolate.cpp:231
    b, m, c = points.shape
    if idx.dim() != 3 or idx.shape[0] != b or idx.shape[2] != 3:
        raise ValueError("ThreeInterpolateGrad expects (b,n,3) idx shape")         # tf_interpolate.cpp:236
    n = idx.shape[1]
    if tuple(weight.shape) != (b, n, 3):
        raise ValueError("ThreeInterpolateGrad expects (b,n,3) weight shape")      # tf_interpolate.cpp:240
    if tuple(grad_out.shape) != (b, n, c):
        raise ValueError("ThreeInterpolateGrad expects (b,n,c) grad_out shape")    # tf_interpolate.cpp:243
    grad_points = torch.empty((b, m, c), dtype=torch.float32, device=points.device)
    check(lib.vnb_three_interpolate_grad(b, n, c, m, dptr(grad_out, torch.float32, "grad_out"), dptr(idx, torch.int32, "idx"),
                                         dptr(weight, torch.float32, "weight"), dptr(grad_points), stream_ptr()))
    return grad_points

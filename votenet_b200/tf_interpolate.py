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


def three_interpolate(points, idx, weight):
    """points (b,m,c), idx (b,n,3) i32, weight (b,n,3) -> (b,n,c).
    Reference: tf_interpolate.py:19-28 (ThreeInterpolate, tf_interpolate.cpp:107-127,191-222)."""
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

"""GPU parity tests (`-m gpu`) of the backward ops through the Python drop-ins / C ABI:
GatherPointGrad, GroupPointGrad (against the oracle AND the reference's own GPU kernels compiled unmodified) and
ThreeInterpolateGrad (against the oracle == the reference's CPU op).  Accumulation uses float atomics, as in the
reference's GPU kernels, so the comparison is to the rounding of a reordered sum (1e-5 relative), not bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import ops as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def T(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def P(t):
    return C.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def ref_gpu(cuda):
    return C.CDLL(O.REF_GPU_PATH) if os.path.exists(O.REF_GPU_PATH) else None


@pytest.mark.parametrize("b,n,m", [(8, 20000, 2048), (2, 100, 300), (1, 7, 1)])
def test_gather_point_grad(cuda, ref_gpu, b, n, m):
    from votenet_b200.tf_sampling import gather_point_grad

    rng = np.random.default_rng(n + m)
    idx = rng.integers(0, min(n, 97), (b, m)).astype(np.int32)   # heavy duplication: many adds per destination
    g = rng.standard_normal((b, m, 3)).astype(np.float32)
    want = O.gather_point_grad(n, idx, g)
    inp = torch.zeros((b, n, 3), device=cuda)
    got = gather_point_grad(inp, T(idx, cuda), T(g, cuda))
    torch.cuda.synchronize()
    assert rel_err(got.cpu().numpy(), want) < TOL
    if ref_gpu is not None:
        ref = torch.full((b, n, 3), 7.0, device=cuda)
        assert ref_gpu.ref_gpu_gather_point_grad(b, n, m, P(T(g, cuda)), P(T(idx, cuda)), P(ref)) == 0
        assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < TOL


@pytest.mark.parametrize("b,n,m,ns,c", [(8, 2048, 1024, 64, 128), (2, 500, 40, 64, 259), (2, 64, 8, 3, 5), (1, 1024, 256, 64, 256)])
def test_group_point_grad(cuda, ref_gpu, b, n, m, ns, c):
    from votenet_b200.tf_grouping import group_point_grad

    rng = np.random.default_rng(n + m + c)
    idx = rng.integers(0, n, (b, m, ns)).astype(np.int32)
    idx[:, :, ns // 2:] = idx[:, :, :1]                      # ball-query padding: the first hit repeated
    g = rng.standard_normal((b, m, ns, c)).astype(np.float32)
    want = O.group_point_grad(n, idx, g)
    pts = torch.zeros((b, n, c), device=cuda)
    got = group_point_grad(pts, T(idx, cuda), T(g, cuda))
    torch.cuda.synchronize()
    assert rel_err(got.cpu().numpy(), want) < TOL
    if ref_gpu is not None and b * m * ns * c < 5e7:
        ref = torch.full((b, n, c), 7.0, device=cuda)
        assert ref_gpu.ref_gpu_group_point_grad(b, n, c, m, ns, P(T(g, cuda)), P(T(idx, cuda)), P(ref)) == 0
        assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < TOL


@pytest.mark.parametrize("b,n,m,c", [(8, 1024, 512, 256), (8, 512, 256, 256), (2, 333, 17, 7), (1, 5, 3, 4)])
def test_three_interpolate_grad(cuda, b, n, m, c):
    from votenet_b200.tf_interpolate import three_interpolate, three_interpolate_grad

    rng = np.random.default_rng(n + m + c)
    idx = rng.integers(0, m, (b, n, 3)).astype(np.int32)
    w = rng.random((b, n, 3), dtype=np.float32)
    w /= w.sum(-1, keepdims=True)
    g = rng.standard_normal((b, n, c)).astype(np.float32)
    want = O.three_interpolate_grad(m, idx, w, g)
    pts = torch.as_tensor(rng.standard_normal((b, m, c)).astype(np.float32), device=cuda)
    got = three_interpolate_grad(pts, T(idx, cuda), T(w, cuda), T(g, cuda))
    torch.cuda.synchronize()
    assert rel_err(got.cpu().numpy(), want) < TOL
    # adjoint identity against the forward op on the device
    fwd = three_interpolate(pts, T(idx, cuda), T(w, cuda))
    lhs = float((fwd.double() * T(g, cuda).double()).sum())
    rhs = float((got.double() * pts.double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))


def test_grad_ops_reject_bad_shapes(cuda):
    from votenet_b200.tf_grouping import group_point_grad
    from votenet_b200.tf_interpolate import three_interpolate_grad
    from votenet_b200.tf_sampling import gather_point_grad

    z = torch.zeros
    with pytest.raises(ValueError):
        gather_point_grad(z((2, 10, 3), device=cuda), z((2, 4), dtype=torch.int32, device=cuda), z((2, 5, 3), device=cuda))
    with pytest.raises(ValueError):
        group_point_grad(z((2, 10, 4), device=cuda), z((2, 3, 2), dtype=torch.int32, device=cuda), z((2, 3, 2, 5), device=cuda))
    with pytest.raises(ValueError):
        three_interpolate_grad(z((2, 6, 4), device=cuda), z((2, 5, 3), dtype=torch.int32, device=cuda),
                               z((2, 5, 3), device=cuda), z((2, 5, 3), device=cuda))


def _numeric_grad(fn, x, eps=1e-2):
    """Central differences of sum(fn(x) * g) w.r.t. x, g fixed random — the ops are linear in x, so this is exact up to
    float rounding (the reference's tests use tf.test.compute_gradient_error the same way)."""
    g = torch.randn_like(fn(x))
    num = torch.zeros_like(x)
    flat, nflat = x.view(-1), num.view(-1)
    for i in range(flat.numel()):
        old = flat[i].item()
        flat[i] = old + eps; hi = float((fn(x).double() * g.double()).sum())
        flat[i] = old - eps; lo = float((fn(x).double() * g.double()).sum())
        flat[i] = old
        nflat[i] = (hi - lo) / (2 * eps)
    return g, num


def test_group_point_gradient_error_like_reference(cuda):
    """The reference's own test (tf_ops/grouping/tf_grouping_op_test.py:9-25): points (1,128,16), xyz1 (1,128,3),
    xyz2 (1,8,3), radius 0.3, nsample 32; gradient error of group_point(points, idx) w.r.t. points < 1e-4."""
    from votenet_b200.tf_grouping import group_point, query_ball_point

    rng = np.random.default_rng(0)
    points = T(rng.random((1, 128, 16), dtype=np.float32), cuda)
    xyz1, xyz2 = T(rng.random((1, 128, 3), dtype=np.float32), cuda), T(rng.random((1, 8, 3), dtype=np.float32), cuda)
    idx, _ = query_ball_point(0.3, 32, xyz1, xyz2)
    g, num = _numeric_grad(lambda p: group_point(p, idx), points.clone())
    p = points.clone().requires_grad_(True)
    out = group_point(p, idx)
    assert tuple(out.shape) == (1, 8, 32, 16)
    out.backward(g)
    assert float((p.grad - num).abs().max()) < 1e-4 * max(1.0, float(num.abs().max()))


def test_three_interpolate_gradient_error_like_reference(cuda):
    """The reference's own test (tf_ops/3d_interpolation/tf_interpolate_op_test.py:9-21): points (1,8,16), xyz1 (1,128,3),
    xyz2 (1,8,3), weight = 1/3; gradient error of three_interpolate w.r.t. points < 1e-4."""
    from votenet_b200.tf_interpolate import three_interpolate, three_nn

    rng = np.random.default_rng(1)
    points = T(rng.random((1, 8, 16), dtype=np.float32), cuda)
    xyz1, xyz2 = T(rng.random((1, 128, 3), dtype=np.float32), cuda), T(rng.random((1, 8, 3), dtype=np.float32), cuda)
    dist, idx = three_nn(xyz1, xyz2)
    weight = torch.ones_like(dist) / 3.0
    g, num = _numeric_grad(lambda p: three_interpolate(p, idx, weight), points.clone())
    p = points.clone().requires_grad_(True)
    out = three_interpolate(p, idx, weight)
    assert tuple(out.shape) == (1, 128, 16)
    out.backward(g)
    assert float((p.grad - num).abs().max()) < 1e-4 * max(1.0, float(num.abs().max()))


def test_gather_point_autograd(cuda):
    from votenet_b200.tf_sampling import farthest_point_sample, gather_point

    rng = np.random.default_rng(2)
    xyz = T(rng.random((2, 200, 3), dtype=np.float32), cuda)
    idx = farthest_point_sample(32, xyz)
    x = xyz.clone().requires_grad_(True)
    out = gather_point(x, idx)
    g = torch.randn_like(out)
    out.backward(g)
    want = torch.zeros_like(xyz)
    for b in range(2):
        want[b].index_add_(0, idx[b].long(), g[b])
    assert torch.allclose(x.grad, want, atol=1e-6)

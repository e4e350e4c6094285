"""Trainable SA / FP modules: custom index-op gradients (B200 kernels) + torch dense layers.
Checks (`-m gpu`): same forward as the fused inference kernels in fp32 mode; gradients equal to a float64 CPU autograd
restatement on the oracle's indices; and the reference's own gradient-error test shape (tf_grouping_op_test.py:9-25) taken
end to end through a set-abstraction layer."""
import numpy as np
import pytest
import torch

from oracle import ops as O

pytestmark = pytest.mark.gpu


def _params(rng, dims, device, dtype=torch.float32, grad=True):
    out = []
    for cin, cout in zip(dims[:-1], dims[1:]):
        W = torch.as_tensor(rng.standard_normal((cin, cout)) * np.sqrt(2.0 / cin), dtype=dtype, device=device).requires_grad_(grad)
        b = torch.as_tensor(rng.standard_normal(cout) * 0.05, dtype=dtype, device=device).requires_grad_(grad)
        out.append((W, b))
    return out


def _cpu_sa(xyz, points, npoint, radius, nsample, layers):
    """float64 torch-CPU restatement of utils.py:25-61,120-132 on the ORACLE's indices."""
    x32 = xyz.detach().numpy().astype(np.float32)
    fps = O.farthest_point_sample(npoint, x32)
    new_xyz32 = O.gather_point(x32, fps)
    idx, _ = O.query_ball_point(radius, nsample, x32, new_xyz32)
    bi = torch.arange(xyz.shape[0])[:, None]
    new_xyz = xyz[bi, torch.as_tensor(fps.astype(np.int64))]
    tidx = torch.as_tensor(idx.astype(np.int64))
    g_xyz = xyz[bi[:, :, None], tidx] - new_xyz[:, :, None, :]
    h = torch.cat([g_xyz, points[bi[:, :, None], tidx]], -1)
    for W, b in layers:
        h = torch.relu(h @ W + b)
    return new_xyz, h.max(dim=2).values, idx


def test_trainable_sa_forward_and_gradients(cuda):
    from votenet_b200.train import pointnet_sa_module_trainable
    from votenet_b200.utils import WeightStore, pointnet_sa_module

    rng = np.random.default_rng(0)
    b, n, c, m, r, ns = 2, 1500, 16, 128, 0.25, 64
    xyz_np = rng.random((b, n, 3)).astype(np.float32)
    pts_np = rng.standard_normal((b, n, c)).astype(np.float32)
    dims = [3 + c, 64, 64, 128]
    layers = _params(np.random.default_rng(1), dims, cuda)
    xyz = torch.as_tensor(xyz_np, device=cuda)
    pts = torch.as_tensor(pts_np, device=cuda).requires_grad_(True)
    new_xyz, out, idx = pointnet_sa_module_trainable(xyz, pts, m, r, ns, layers)
    # (a) same forward as the fused inference path in exact-fp32 mode
    w = {}
    for i, (W, bb) in enumerate(layers):
        w[f"s/conv{i}/W"] = W.detach().cpu(); w[f"s/conv{i}/b"] = bb.detach().cpu()
    store = WeightStore(w, device=cuda, precision=0)
    nx2, out2, idx2 = pointnet_sa_module(xyz, pts.detach(), m, r, ns, list(dims[1:]), None, False, "s", weights=store)
    assert torch.equal(idx, idx2) and torch.equal(new_xyz, nx2)
    assert (out.detach() - out2).abs().max().item() < 2e-5 * out2.abs().max().item()
    # (b) gradients vs float64 CPU autograd on the oracle's indices
    g = torch.as_tensor(rng.standard_normal(out.shape).astype(np.float32), device=cuda)
    out.backward(g)
    xyz_c = torch.as_tensor(xyz_np, dtype=torch.float64)
    pts_c = torch.as_tensor(pts_np, dtype=torch.float64).requires_grad_(True)
    layers_c = [(W.detach().cpu().double().requires_grad_(True), bb.detach().cpu().double().requires_grad_(True)) for W, bb in layers]
    _, out_c, idx_c = _cpu_sa(xyz_c, pts_c, m, r, ns, layers_c)
    assert np.array_equal(idx.cpu().numpy(), idx_c)
    out_c.backward(g.cpu().double())
    scale = pts_c.grad.abs().max().item()
    assert (pts.grad.cpu().double() - pts_c.grad).abs().max().item() < 1e-4 * scale
    for (W, bb), (Wc, bc) in zip(layers, layers_c):
        assert (W.grad.cpu().double() - Wc.grad).abs().max().item() < 1e-4 * Wc.grad.abs().max().item()
        assert (bb.grad.cpu().double() - bc.grad).abs().max().item() < 1e-4 * max(bc.grad.abs().max().item(), 1e-6)


def test_gradient_error_like_the_reference_test(cuda):
    """tf_grouping_op_test.py:9-25: points (1,128,16), xyz1 (1,128,3), 8 centroids, radius 0.3, nsample 32,
    compute_gradient_error(points -> grouped_points) < 1e-4.  Same sizes through sample_and_group (FPS picks the 8
    centroids): the map points -> new_points is linear, so central differences are exact up to rounding."""
    from votenet_b200.train import sample_and_group

    rng = np.random.default_rng(3)
    xyz = torch.as_tensor(rng.random((1, 128, 3)).astype(np.float32), device=cuda)
    pts0 = rng.random((1, 128, 16)).astype(np.float32)
    proj = torch.as_tensor(rng.standard_normal((1, 8, 32, 19)).astype(np.float32), device=cuda)

    def f(p):
        return (sample_and_group(8, 0.3, 32, xyz, p)[1].double() * proj.double()).sum()

    pts = torch.as_tensor(pts0, device=cuda).requires_grad_(True)
    f(pts).backward()
    analytic = pts.grad.cpu().numpy().astype(np.float64)
    eps = 0.25
    worst = 0.0
    for _ in range(48):
        i, j = int(rng.integers(0, 128)), int(rng.integers(0, 16))
        d = np.zeros_like(pts0); d[0, i, j] = eps
        num = (f(torch.as_tensor(pts0 + d, device=cuda)).item() - f(torch.as_tensor(pts0 - d, device=cuda)).item()) / (2 * eps)
        worst = max(worst, abs(num - analytic[0, i, j]))
    assert worst < 1e-4, worst
    assert np.abs(analytic).max() > 0.1


def test_trainable_fp_gradients(cuda):
    from votenet_b200.train import pointnet_fp_module_trainable

    rng = np.random.default_rng(5)
    b, n, m, c1, c2 = 2, 300, 90, 24, 40
    xyz1 = torch.as_tensor(rng.random((b, n, 3)).astype(np.float32), device=cuda)
    xyz2 = torch.as_tensor(rng.random((b, m, 3)).astype(np.float32), device=cuda)
    p1 = torch.as_tensor(rng.standard_normal((b, n, c1)).astype(np.float32), device=cuda).requires_grad_(True)
    p2 = torch.as_tensor(rng.standard_normal((b, m, c2)).astype(np.float32), device=cuda).requires_grad_(True)
    layers = _params(np.random.default_rng(6), [c1 + c2, 48, 32], cuda)
    out = pointnet_fp_module_trainable(xyz1, xyz2, p1, p2, layers)
    g = torch.as_tensor(rng.standard_normal(out.shape).astype(np.float32), device=cuda)
    out.backward(g)
    # float64 CPU restatement on the oracle's three_nn
    d, idx = O.three_nn(xyz1.cpu().numpy(), xyz2.cpu().numpy())
    d = np.maximum(d.astype(np.float32), 1e-10)
    wgt = (1.0 / d) / (1.0 / d).sum(2, keepdims=True)
    p1c = p1.detach().cpu().double().requires_grad_(True)
    p2c = p2.detach().cpu().double().requires_grad_(True)
    bi = torch.arange(b)[:, None, None]
    h = (p2c[bi, torch.as_tensor(idx.astype(np.int64))] * torch.as_tensor(wgt, dtype=torch.float64)[..., None]).sum(2)
    h = torch.cat([h, p1c], 2)
    for W, bb in layers:
        h = torch.relu(h @ W.detach().cpu().double() + bb.detach().cpu().double())
    h.backward(g.cpu().double())
    assert (out.detach().cpu().double() - h.detach()).abs().max().item() < 1e-4
    assert (p2.grad.cpu().double() - p2c.grad).abs().max().item() < 1e-4 * p2c.grad.abs().max().item()
    assert (p1.grad.cpu().double() - p1c.grad).abs().max().item() < 1e-4 * p1c.grad.abs().max().item()


@pytest.mark.parametrize("b,n,c,m,dims,r", [(2, 1500, 16, 128, (64, 64, 128), 0.25), (1, 900, 128, 64, (128, 128, 256), 0.3),
                                            (2, 700, 1, 40, (64, 64, 128), 0.2)])
def test_fused_sa_backward_kernel(cuda, b, n, c, m, dims, r):
    """vnb_sa_group_mlp_max_backward (one rematerialising kernel) against the composed autograd path (validated above
    against float64 CPU autograd): gradients w.r.t. features, xyz and all six parameters."""
    from votenet_b200.train import pointnet_sa_module_fused_trainable, pointnet_sa_module_trainable

    rng = np.random.default_rng(c)
    xyz_np = rng.random((b, n, 3)).astype(np.float32)
    pts_np = rng.standard_normal((b, n, c)).astype(np.float32)
    g_np = None
    grads = []
    for fn in (pointnet_sa_module_trainable, pointnet_sa_module_fused_trainable):
        layers = _params(np.random.default_rng(1), [3 + c] + list(dims), cuda)
        xyz = torch.as_tensor(xyz_np, device=cuda).requires_grad_(True)
        pts = torch.as_tensor(pts_np, device=cuda).requires_grad_(True)
        _, out, idx = fn(xyz, pts, m, r, 64, layers)
        if g_np is None:
            g_np = rng.standard_normal(tuple(out.shape)).astype(np.float32)
        out.backward(torch.as_tensor(g_np, device=cuda))
        grads.append((out.detach(), [pts.grad, xyz.grad] + [t.grad for Wb in layers for t in Wb]))
    (o_ref, g_ref), (o_got, g_got) = grads
    assert (o_ref - o_got).abs().max().item() < 2e-5 * o_ref.abs().max().item()
    names = ["d points", "d xyz", "dW1", "db1", "dW2", "db2", "dW3", "db3"]
    for nme, a, bb in zip(names, g_ref, g_got):
        scale = max(a.abs().max().item(), 1e-6)
        assert (a - bb).abs().max().item() < 2e-4 * scale, (nme, (a - bb).abs().max().item(), scale)

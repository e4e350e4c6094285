"""GPU parity tests (`-m gpu`) of the dense layers: fp32 SIMT kernels to ~1e-5 and the tcgen05 tensor-core kernels
(fp16 operands, fp32 accumulate) to the 1e-3 relative tolerance BASELINE.json states, against torch-CPU fp32 / the
dense oracle on identical inputs."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import dense as D
from oracle import ops as O

pytestmark = pytest.mark.gpu

TOL_FP32 = 2e-5
TOL_TC = 1e-3  # BASELINE.json: "within 1e-3 relative on the float MLP/interpolation outputs"


def T(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


@pytest.mark.parametrize("rows,cin,cout,act,res", [(1000, 512, 256, True, False), (8192, 259, 256, True, False),
                                                   (1024, 256, 259, False, True), (256, 128, 79, False, False),
                                                   (4096, 128, 128, False, False), (5, 6, 64, True, False),
                                                   (130, 16, 16, False, False), (129, 272, 272, True, True)])
@pytest.mark.parametrize("precision", [0, 1])
def test_linear(cuda, rows, cin, cout, act, res, precision):
    from votenet_b200.utils import Layer, linear

    g = torch.Generator().manual_seed(rows + cin + cout)
    x = torch.randn(rows, cin, generator=g)
    W = torch.randn(cin, cout, generator=g) * (2.0 / cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    r = torch.randn(rows, cout, generator=g) if res else None
    ref = x.double() @ W.double() + b.double()
    if act:
        ref = torch.relu(ref)
    if res:
        ref = ref + r.double()
    layer = Layer(W, b, cuda)
    got = linear(x.to(cuda), layer, act, precision, residual=r.to(cuda) if res else None)
    torch.cuda.synchronize()
    assert got.shape == (rows, cout)
    e = rel_err(got.cpu().numpy(), ref.numpy())
    assert e < (TOL_FP32 if precision == 0 else TOL_TC), f"rel err {e:.3e}"
    if precision == 1:  # fp16 output path (used for the hoisted layer-1 pre-GEMM)
        got16 = linear(x.to(cuda), layer, act, 1, residual=r.to(cuda) if res else None, out_f16=True)
        assert got16.dtype == torch.float16
        assert rel_err(got16.float().cpu().numpy(), ref.numpy()) < 2e-3


def _sa_case(b, n, m, c, mlp, r, seed):
    rng = np.random.default_rng(seed)
    xyz = rng.random((b, n, 3), dtype=np.float32)
    feat = rng.standard_normal((b, n, c)).astype(np.float32)
    fps = O.farthest_point_sample(m, xyz)
    new_xyz = O.gather_point(xyz, fps)
    idx, _ = O.query_ball_point(r, 64, xyz, new_xyz)
    g = torch.Generator().manual_seed(seed)
    w = {}
    cin = 3 + c
    for i, co in enumerate(mlp):
        w[f"s/conv{i}/W"] = torch.randn(cin, co, generator=g) * (2.0 / cin) ** 0.5
        w[f"s/conv{i}/b"] = torch.randn(co, generator=g) * 0.05
        w[f"s/conv{i}/bn/gamma"] = torch.rand(co, generator=g) * 0.4 + 0.8
        w[f"s/conv{i}/bn/beta"] = torch.randn(co, generator=g) * 0.05
        w[f"s/conv{i}/bn/mean/EMA"] = torch.randn(co, generator=g) * 0.05
        w[f"s/conv{i}/bn/variance/EMA"] = torch.rand(co, generator=g) * 0.4 + 0.8
        cin = co
    # oracle: grouped tensor -> conv/BN/ReLU x3 -> max   (utils.py:50-55,120-132)
    grouped = np.concatenate([O.group_point(xyz, idx) - new_xyz[:, :, None, :], O.group_point(feat, idx)], -1)
    h = torch.as_tensor(grouped)
    for i in range(3):
        h = D.dense_layer(h, w, f"s/conv{i}")
    ref = h.max(dim=2).values
    return xyz, feat, new_xyz, idx, w, ref


@pytest.mark.parametrize("b,n,m,c,mlp,r", [(2, 3000, 256, 1, (64, 64, 128), 0.15), (1, 2000, 128, 3, (64, 64, 128), 0.2),
                                           (2, 1024, 128, 128, (128, 128, 256), 0.25), (1, 512, 64, 256, (128, 128, 256), 0.4),
                                           (2, 1024, 64, 256, (128, 128, 128), 0.3), (1, 1500, 96, 7, (64, 64, 128), 0.25)])
@pytest.mark.parametrize("precision", [0, 1])
def test_sa_group_mlp_max(cuda, b, n, m, c, mlp, r, precision):
    from votenet_b200.utils import WeightStore, sa_group_mlp_max

    xyz, feat, new_xyz, idx, w, ref = _sa_case(b, n, m, c, mlp, r, seed=b * 1000 + n + c)
    store = WeightStore(w, device=cuda, precision=precision)
    layers = [store.layer(f"s/conv{i}") for i in range(3)]
    got = sa_group_mlp_max(T(xyz, cuda), T(feat, cuda), T(new_xyz, cuda), T(idx, cuda), layers, precision, store, "s")
    torch.cuda.synchronize()
    e = rel_err(got.cpu().numpy(), ref.numpy())
    assert e < (TOL_FP32 if precision == 0 else TOL_TC), f"rel err {e:.3e}"


def test_fp_module_and_helpers(cuda):
    from votenet_b200._lib import check, dptr, lib, stream_ptr
    from votenet_b200.utils import WeightStore, pointnet_fp_module

    rng = np.random.default_rng(77)
    b, n, m, c1, c2 = 2, 512, 256, 256, 256
    xyz1 = rng.random((b, n, 3), dtype=np.float32); xyz2 = rng.random((b, m, 3), dtype=np.float32)
    p1 = rng.standard_normal((b, n, c1)).astype(np.float32); p2 = rng.standard_normal((b, m, c2)).astype(np.float32)
    g = torch.Generator().manual_seed(5)
    w = {}
    cin = c1 + c2
    for i, co in enumerate((256, 256)):
        w[f"fp/conv_{i}/W"] = torch.randn(cin, co, generator=g) * (2.0 / cin) ** 0.5
        w[f"fp/conv_{i}/b"] = torch.randn(co, generator=g) * 0.05
        w[f"fp/conv_{i}/bn/gamma"] = torch.rand(co, generator=g) * 0.4 + 0.8
        w[f"fp/conv_{i}/bn/beta"] = torch.randn(co, generator=g) * 0.05
        w[f"fp/conv_{i}/bn/mean/EMA"] = torch.randn(co, generator=g) * 0.05
        w[f"fp/conv_{i}/bn/variance/EMA"] = torch.rand(co, generator=g) * 0.4 + 0.8
        cin = co
    ref, ex = D.pointnet_fp_module(torch.as_tensor(xyz1), torch.as_tensor(xyz2), torch.as_tensor(p1), torch.as_tensor(p2),
                                   (256, 256), "fp", w)
    for precision, tol in ((0, TOL_FP32), (1, TOL_TC)):
        store = WeightStore(w, device=cuda, precision=precision)
        got = pointnet_fp_module(T(xyz1, cuda), T(xyz2, cuda), T(p1, cuda), T(p2, cuda), [256, 256], "fp", weights=store)
        torch.cuda.synchronize()
        assert rel_err(got.cpu().numpy(), ref.numpy()) < tol
    # interpolate+concat front half alone: float-exact up to the 1/d weights (1e-6)
    cat = torch.empty((b * n, c1 + c2), device=cuda)
    td, ti, tp1, tp2 = T(ex["dist"].numpy(), cuda), T(ex["idx"], cuda), T(p1, cuda), T(p2, cuda)  # keep alive
    check(lib.vnb_fp_interpolate_concat(b, n, m, c1, c2, dptr(td), dptr(ti), dptr(tp1), dptr(tp2), dptr(cat), stream_ptr()))
    torch.cuda.synchronize()
    want = np.concatenate([ex["interpolated"].numpy(), p1], -1).reshape(b * n, -1)
    assert np.abs(cat.cpu().numpy() - want).max() < 1e-5
    # concat2 / split2 round trip
    a = torch.randn(100, 3, device=cuda); bb = torch.randn(100, 256, device=cuda)
    o = torch.empty(100, 259, device=cuda)
    check(lib.vnb_concat2(100, 3, 256, dptr(a), dptr(bb), dptr(o), stream_ptr()))
    assert torch.equal(o, torch.cat([a, bb], 1))
    a2 = torch.empty_like(a); b2 = torch.empty_like(bb)
    check(lib.vnb_split2(100, 3, 256, dptr(o), dptr(a2), dptr(b2), stream_ptr()))
    assert torch.equal(a2, a) and torch.equal(b2, bb)


@pytest.mark.parametrize("b,n,m", [(2, 512, 256), (3, 200, 77), (8, 1024, 512)])
def test_fp_vote_fused(cuda, b, n, m):
    """vnb_fp_module_fused (fp module + voting module in one tensor-core kernel) against the dense oracle
    (utils.py:266-294 + model.py:53-61), and against the unfused kernels; ragged row counts included."""
    from votenet_b200 import utils as U
    from votenet_b200.tf_interpolate import three_nn

    rng = np.random.default_rng(1234 + n)
    xyz1 = rng.random((b, n, 3), dtype=np.float32); xyz2 = rng.random((b, m, 3), dtype=np.float32)
    p1 = rng.standard_normal((b, n, 256)).astype(np.float32); p2 = rng.standard_normal((b, m, 256)).astype(np.float32)
    g = torch.Generator().manual_seed(9)
    w = {}
    for name, cin, cout, bn in (("fp2/conv_0", 512, 256, True), ("fp2/conv_1", 256, 256, True), ("voting0", 259, 256, True),
                                ("voting1", 256, 256, True), ("voting2", 256, 259, False)):
        w[f"{name}/W"] = torch.randn(cin, cout, generator=g) * (2.0 / cin) ** 0.5
        w[f"{name}/b"] = torch.randn(cout, generator=g) * 0.05
        if bn:
            w[f"{name}/bn/gamma"] = torch.rand(cout, generator=g) * 0.4 + 0.8
            w[f"{name}/bn/beta"] = torch.randn(cout, generator=g) * 0.05
            w[f"{name}/bn/mean/EMA"] = torch.randn(cout, generator=g) * 0.05
            w[f"{name}/bn/variance/EMA"] = torch.rand(cout, generator=g) * 0.4 + 0.8
    seeds_feat, _ = D.pointnet_fp_module(torch.as_tensor(xyz1), torch.as_tensor(xyz2), torch.as_tensor(p1),
                                         torch.as_tensor(p2), (256, 256), "fp2", w)
    seeds = torch.cat([torch.as_tensor(xyz1), seeds_feat], 2)
    off = seeds.reshape(-1, 259)
    for i in range(3):
        off = D.dense_layer(off, w, f"voting{i}")
    votes = (seeds + off.reshape(seeds.shape)).numpy()

    store = U.WeightStore(w, device=cuda, precision=1)
    txyz1, txyz2, tp1, tp2 = T(xyz1, cuda), T(xyz2, cuda), T(p1, cuda), T(p2, cuda)
    dist, idx = three_nn(txyz1, txyz2)
    fp_out = torch.empty((b, n, 256), device=cuda)
    vx = torch.empty((b, n, 3), device=cuda); vf = torch.empty((b, n, 256), device=cuda)
    vl, x0 = U.vote_layers_fused(store, ["voting0", "voting1", "voting2"])
    U.fp_module_fused(dist, idx, tp1, tp2, [store.layer("fp2/conv_0"), store.layer("fp2/conv_1")], fp_out,
                      vote=(vl, x0, txyz1, vx, vf))
    torch.cuda.synchronize()
    assert rel_err(fp_out.cpu().numpy(), seeds_feat.numpy()) < TOL_TC
    assert rel_err(vf.cpu().numpy(), votes[..., 3:]) < TOL_TC
    assert np.abs(vx.cpu().numpy() - votes[..., :3]).max() < TOL_TC * max(1.0, np.abs(votes[..., :3]).max())
    # the vote residual lives in tensor memory: without a seed-feature buffer (the engine's call) the votes are the same bits
    vx2 = torch.empty_like(vx); vf2 = torch.empty_like(vf)
    U.fp_module_fused(dist, idx, tp1, tp2, [store.layer("fp2/conv_0"), store.layer("fp2/conv_1")], None,
                      vote=(vl, x0, txyz1, vx2, vf2))
    torch.cuda.synchronize()
    assert torch.equal(vf2, vf) and torch.equal(vx2, vx)
    # fp-only form of the same kernel == what pointnet_fp_module dispatches to; and the unfused kernels agree to 1e-3
    a = U.pointnet_fp_module(txyz1, txyz2, tp1, tp2, [256, 256], "fp2", weights=store)
    U.FUSE_FP = False
    try:
        c = U.pointnet_fp_module(txyz1, txyz2, tp1, tp2, [256, 256], "fp2", weights=store)
    finally:
        U.FUSE_FP = True
    torch.cuda.synchronize()
    assert torch.equal(a, fp_out)
    assert rel_err(a.cpu().numpy(), c.cpu().numpy()) < TOL_TC


def test_decode_boxes(cuda):
    from votenet_b200 import synth
    from votenet_b200.model import decode_boxes

    rng = np.random.default_rng(8)
    b, k = 3, 256
    pxyz = (rng.random((b, k, 3), dtype=np.float32) * 6 - 3).astype(np.float32)
    pout = rng.standard_normal((b, k, 79)).astype(np.float32)
    ref = D.decode_boxes(torch.as_tensor(pxyz), torch.as_tensor(pout), synth.CLASS_MEAN_SIZE)
    bboxes, scores, obj, cls = decode_boxes(T(pxyz, cuda), T(pout, cuda), T(synth.CLASS_MEAN_SIZE, cuda))
    torch.cuda.synchronize()
    assert np.abs(bboxes.cpu().numpy() - ref["bboxes"].numpy()).max() < 2e-5
    assert np.array_equal(scores.cpu().numpy(), ref["scores"].numpy())
    assert np.array_equal(obj.cpu().numpy(), ref["objectness"].numpy())
    assert np.array_equal(cls.cpu().numpy(), ref["class_scores"].numpy())


def test_sa_kernels_many_tiles_and_issuer_modes(cuda):
    """The pipelined tensor-core kernels (sa_ws2.cu / sa1_ws2.cu) with many tiles per CTA (rings of barriers wrap many
    times), CTAs with ragged tile counts, and both MMA-issuer wait modes (parked / polling): same bits either way, and
    within tolerance of the oracle and of the exact-fp32 SIMT twin."""
    from votenet_b200._lib import check, lib
    from votenet_b200.utils import WeightStore, sa_group_mlp_max

    for (b, n, m, c, mlp, r) in [(8, 2048, 1024, 128, (128, 128, 256), 0.2), (3, 1024, 254, 256, (128, 128, 128), 0.3),
                                 (4, 6000, 2048, 1, (64, 64, 128), 0.15), (1, 3000, 130, 3, (64, 64, 128), 0.2)]:
        xyz, feat, new_xyz, idx, w, ref = _sa_case(b, n, m, c, mlp, r, seed=77 + m)
        store = WeightStore(w, device=cuda, precision=1)
        layers = [store.layer(f"s/conv{i}") for i in range(3)]
        args = (T(xyz, cuda), T(feat, cuda), T(new_xyz, cuda), T(idx, cuda), layers, 1, store, "s")
        outs = []
        try:
            for v in (2, 3):
                check(lib.vnb_set_tuning(b"sa_variant", v))
                outs.append(sa_group_mlp_max(*args))
                torch.cuda.synchronize()
        finally:
            check(lib.vnb_set_tuning(b"sa_variant", 2))
        assert torch.equal(outs[0], outs[1])
        store32 = WeightStore(w, device=cuda, precision=0)
        exact = sa_group_mlp_max(T(xyz, cuda), T(feat, cuda), T(new_xyz, cuda), T(idx, cuda),
                                 [store32.layer(f"s/conv{i}") for i in range(3)], 0, store32, "s")
        assert rel_err(exact.cpu().numpy(), ref.numpy()) < 2e-5
        assert rel_err(outs[0].cpu().numpy(), exact.cpu().numpy()) < TOL_TC
        assert rel_err(outs[0].cpu().numpy(), ref.numpy()) < TOL_TC


@pytest.mark.parametrize("b,n,m,c,mlp", [(8, 2048, 1024, 128, (128, 128, 256)), (3, 1024, 254, 256, (128, 128, 128)),
                                         (4, 6000, 2048, 1, (64, 64, 128)), (1, 3000, 130, 3, (64, 64, 128)),
                                         (2, 600, 6, 128, (128, 128, 256))])
def test_sa_tile_packing_is_bit_identical(cuda, b, n, m, c, mlp):
    """vnb_sa_group_mlp_max_counted: with the ball query's pts_cnt the fused kernels run centroids in 16 / 32 / 64-row
    slots (csrc/sa_pack.cu) and skip the padded duplicate rows — outputs must equal the unpacked run bit for bit, for
    every mix of slot classes (engineered counts: 0, 1, 16, 17, 32, 33, 64; ragged last tiles of every class) and also
    match the oracle."""
    from votenet_b200.utils import WeightStore, sa_group_mlp_max

    rng = np.random.default_rng(b * 7 + m)
    xyz = rng.random((b, n, 3), dtype=np.float32)
    feat = rng.standard_normal((b, n, c)).astype(np.float32)
    new_xyz = O.gather_point(xyz, O.farthest_point_sample(m, xyz))
    # engineered groups with the reference's padding rule: cnt distinct neighbours, then the first one repeated
    cnt = rng.choice(np.asarray([1, 5, 16, 17, 30, 32, 33, 50, 64], np.int32), size=(b, m))
    cnt[0, :3] = (16, 32, 64)
    full, full_cnt = O.query_ball_point(0.45, 64, xyz, new_xyz)   # big ball: (almost) every group has 64 real hits
    cnt = np.minimum(cnt, full_cnt).astype(np.int32)
    idx = full.copy()
    for i in range(b):
        for j in range(m):
            idx[i, j, int(cnt[i, j]):] = full[i, j, 0]              # keep the first cnt hits, pad with the first one
    g = torch.Generator().manual_seed(3)
    w = {}
    cin = 3 + c
    for i, co in enumerate(mlp):
        w[f"s/conv{i}/W"] = torch.randn(cin, co, generator=g) * (2.0 / cin) ** 0.5
        w[f"s/conv{i}/b"] = torch.randn(co, generator=g) * 0.05
        w[f"s/conv{i}/bn/gamma"] = torch.rand(co, generator=g) * 0.4 + 0.8   # BN keys make the oracle apply BN + ReLU
        w[f"s/conv{i}/bn/beta"] = torch.randn(co, generator=g) * 0.05
        w[f"s/conv{i}/bn/mean/EMA"] = torch.randn(co, generator=g) * 0.05
        w[f"s/conv{i}/bn/variance/EMA"] = torch.rand(co, generator=g) * 0.4 + 0.8
        cin = co
    store = WeightStore(w, device=cuda, precision=1)
    layers = [store.layer(f"s/conv{i}") for i in range(3)]
    args = (T(xyz, cuda), T(feat, cuda), T(new_xyz, cuda), T(idx, cuda), layers, 1, store, "s")
    plain = sa_group_mlp_max(*args)
    packed = sa_group_mlp_max(*args, pts_cnt=T(cnt, cuda))
    cnt0 = cnt.copy(); cnt0[:, ::5] = 0            # "empty ball" entries keep the full 64-row slot
    packed0 = sa_group_mlp_max(*args, pts_cnt=T(cnt0, cuda))
    torch.cuda.synchronize()
    assert torch.equal(plain, packed)
    assert torch.equal(plain, packed0)
    grouped = np.concatenate([O.group_point(xyz, idx) - new_xyz[:, :, None, :], O.group_point(feat, idx)], -1)
    h = torch.as_tensor(grouped)
    for i in range(3):
        h = D.dense_layer(h, w, f"s/conv{i}")
    assert rel_err(packed.cpu().numpy(), h.max(dim=2).values.numpy()) < TOL_TC

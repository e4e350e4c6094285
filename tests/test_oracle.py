"""CPU tests of the oracle (`-m "not gpu"`): the C restatement is pinned against (a) golden vectors produced by the
REAL reference sources (tests/golden/ref_*.npz), (b) the live reference library oracle/_ref when it is built, (c) the
only known answer the reference ships (the tf_nms3d.py demo), and (d) the algorithmic properties SURVEY.md lists."""
import numpy as np
import pytest

from conftest import random_boxes
from helpers import golden, make_golden
from oracle import ops

needs_ref = pytest.mark.skipif(not ops.ref.available, reason="oracle/_ref not built (needs /root/reference)")


# ---------------------------------------------------------------- golden vectors from the real reference
def test_golden_interpolate_config1():
    g = golden("ref_interpolate_config1")
    xyz1, xyz2, pts = make_golden.interp_inputs()
    dist, idx = ops.three_nn(xyz1, xyz2)
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(dist, g["dist"])
    out = ops.three_interpolate(pts, idx, make_golden.fp_weights(dist))
    assert np.array_equal(out[:, ::32], g["out_rows"])
    assert abs(out.astype(np.float64).sum() - float(g["out_sum"])) < 1e-9


@pytest.mark.parametrize("name", ["ref_nms_random", "ref_nms_degenerate", "ref_nms_thr0"])
def test_golden_nms(name):
    g = golden(name)
    boxes, scores, obj = make_golden.nms_inputs(int(g["seed"]), int(g["b"]), int(g["k"]), bool(g["degenerate"]))
    sel = ops.NMS3D(boxes, scores, obj, float(g["thr"]))
    assert np.array_equal(sel, g["selected"])
    inter = np.array([ops.intersection2d(boxes[0, i], boxes[0, j]) for i in range(16) for j in range(16)], np.float32)
    assert np.array_equal(inter, g["inter2d_16x16"])


def test_golden_fps_ballquery_self_consistency():
    g = golden("oracle_fps_ballquery")
    x = make_golden.fps_inputs()
    f = ops.farthest_point_sample(64, x)
    assert np.array_equal(f, g["fps"])
    bi, bc = ops.query_ball_point(0.12, 16, x, ops.gather_point(x, f))
    assert np.array_equal(bi, g["ball_idx"]) and np.array_equal(bc, g["ball_cnt"])


# ---------------------------------------------------------------- the reference's only known answer
def _demo_boxes():
    def roty(t):
        c, s = np.cos(t), np.sin(t)
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])

    def bbox(l, w, h, ang=None):  # tf_nms3d.py:21-28
        x = [l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2]
        y = [h / 2, h / 2, h / 2, h / 2, -h / 2, -h / 2, -h / 2, -h / 2]
        z = [w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2]
        c = np.vstack([x, y, z])
        if ang:
            c = roty(ang) @ c
        return c.T

    return np.array([[bbox(1, 1, 1), bbox(0.8, 0.8, 0.8, np.pi / 4 * 3)]]).astype("float32")


def test_nms_demo_known_answer():
    """tf_nms3d.py:30-46: unit cube vs 0.8-cube yawed 135 deg, thr 0.5 -> IoU 0.491408676 < thr -> [[0,1],[0,0]]."""
    bb = _demo_boxes()
    scores = np.array([[0.5, 0.6]], "float32")
    obj = np.array([[[0.3, 0.7], [0.4, 0.6]]], "float32")
    assert ops.NMS3D(bb, scores, obj, 0.5).tolist() == [[0, 1], [0, 0]]
    assert abs(ops.intersection2d(bb[0, 0], bb[0, 1]) - 0.622741759) < 1e-6
    assert abs(ops.iou3d(bb[0, 0], bb[0, 1]) - 0.491408676) < 1e-6
    assert ops.NMS3D(bb, scores, obj, 0.49).tolist() == [[0, 1]]
    assert ops.intersection2d(bb[0, 0], bb[0, 0]) == 1.0
    far = bb[0, 0] + np.float32(5.0)
    assert ops.intersection2d(bb[0, 0], far) == 0.0
    with pytest.raises(ValueError):
        ops.NMS3D(bb, scores, obj, 1.5)  # tf_nms3d.cpp:300


# ---------------------------------------------------------------- live comparison with the compiled reference
@needs_ref
@pytest.mark.parametrize("seed", range(4))
def test_three_nn_interpolate_vs_reference(seed):
    rng = np.random.default_rng(100 + seed)
    b, n, m, c = 2, 300 + 37 * seed, 64 + seed, 33
    xyz1 = rng.random((b, n, 3), dtype=np.float32)
    xyz2 = rng.random((b, m, 3), dtype=np.float32)
    if seed == 3:  # exact distance ties: known points on a lattice, duplicated
        xyz2 = np.round(xyz2 * 4) / 4
        xyz1 = np.round(xyz1 * 4) / 4
    d0, i0 = ops.ref.three_nn(xyz1, xyz2)
    d1, i1 = ops.three_nn(xyz1, xyz2)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    pts = rng.standard_normal((b, m, c)).astype(np.float32)
    w = make_golden.fp_weights(d0)
    assert np.array_equal(ops.ref.three_interpolate(pts, i0, w), ops.three_interpolate(pts, i1, w))


@needs_ref
def test_three_nn_fewer_than_three_known_points():
    rng = np.random.default_rng(5)
    xyz1 = rng.random((1, 10, 3), dtype=np.float32)
    for m in (1, 2):
        xyz2 = rng.random((1, m, 3), dtype=np.float32)
        d0, i0 = ops.ref.three_nn(xyz1, xyz2)
        d1, i1 = ops.three_nn(xyz1, xyz2)
        assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
        assert np.isinf(d1[..., m:]).all() and (i1[..., m:] == 0).all()  # (float)1e40 = inf, idx 0 (A.3)


@needs_ref
@pytest.mark.parametrize("seed,deg", [(0, False), (1, False), (2, True), (3, True)])
def test_intersection_vs_reference(seed, deg):
    rng = np.random.default_rng(200 + seed)
    boxes = random_boxes(rng, 1, 40, spread=1.0, degenerate=deg)[0]
    for i in range(40):
        for j in range(40):
            assert ops.intersection2d(boxes[i], boxes[j]) == ops.ref.intersection2d(boxes[i], boxes[j]), (i, j)


@needs_ref
@pytest.mark.parametrize("seed,deg,thr", [(0, False, 0.25), (1, False, 0.1), (2, True, 0.25), (3, False, 0.0), (4, False, 1.0)])
def test_nms_vs_reference(seed, deg, thr):
    rng = np.random.default_rng(300 + seed)
    b, k = 3, 80
    boxes = random_boxes(rng, b, k, spread=1.5, degenerate=deg)
    scores = rng.standard_normal((b, k)).astype(np.float32)
    obj = rng.standard_normal((b, k, 2)).astype(np.float32)
    assert np.array_equal(ops.NMS3D(boxes, scores, obj, thr), ops.ref.NMS3D(boxes, scores, obj, thr))


@needs_ref
def test_nms_exact_score_ties_follow_libstdcxx_heap():
    """The reference's order among EXACTLY equal scores is whatever std::priority_queue does; the oracle restates
    libstdc++'s push_heap/pop_heap so that even this matches."""
    rng = np.random.default_rng(42)
    b, k = 2, 64
    boxes = random_boxes(rng, b, k, spread=1.0)
    scores = rng.integers(0, 4, (b, k)).astype(np.float32)  # many exact ties
    obj = rng.standard_normal((b, k, 2)).astype(np.float32)
    assert np.array_equal(ops.NMS3D(boxes, scores, obj, 0.25), ops.ref.NMS3D(boxes, scores, obj, 0.25))


# ---------------------------------------------------------------- algorithmic properties
def test_fps_starts_at_zero_and_nested_prefix():
    """A.1 + fact 8: first index is 0; FPS of an FPS-ordered set is its own prefix (no exact ties in random data)."""
    rng = np.random.default_rng(3)
    x = rng.random((2, 5000, 3), dtype=np.float32)
    f1 = ops.farthest_point_sample(512, x)
    assert (f1[:, 0] == 0).all()
    assert all(len(set(r.tolist())) == 512 for r in f1)
    l1 = ops.gather_point(x, f1)
    f2 = ops.farthest_point_sample(128, l1)
    assert np.array_equal(f2, np.tile(np.arange(128, dtype=np.int32), (2, 1)))


def test_fps_tie_rule():
    """Equal maxima: lowest (k mod 512) wins, then lowest k (A.1).  Points 600 and 100 are equidistant duplicates."""
    x = np.zeros((1, 700, 3), np.float32)
    x[0, 600] = (1, 0, 0)
    x[0, 100] = (1, 0, 0)   # slot 100;  600 mod 512 = 88 -> 600 wins although 100 < 600
    assert ops.farthest_point_sample(2, x)[0].tolist() == [0, 600]
    x[0, 600] = 0
    x[0, 612] = (1, 0, 0)   # slot 100 again: same slot -> lower k (100) wins
    assert ops.farthest_point_sample(2, x)[0].tolist() == [0, 100]


def test_ball_query_padding_rule():
    """A.2: hits in ascending index order, padded with the FIRST hit; count capped at nsample."""
    rng = np.random.default_rng(9)
    x = rng.random((1, 400, 3), dtype=np.float32)
    q = x[:, :5]
    idx, cnt = ops.query_ball_point(0.2, 8, x, q)
    d = np.linalg.norm(x[0][None] - q[0][:, None], axis=-1)
    for j in range(5):
        hits = np.nonzero(d[j] < 0.2 - 1e-6)[0]
        c = min(len(hits), 8)
        assert cnt[0, j] == c or abs(cnt[0, j] - c) <= 1  # boundary ulps
        row = idx[0, j]
        assert (np.diff(row[:cnt[0, j]]) > 0).all()
        assert (row[cnt[0, j]:] == row[0]).all()
    # empty ball: row untouched, count 0
    far = np.full((1, 1, 3), 10.0, np.float32)
    idx, cnt = ops.query_ball_point(0.1, 4, x, far, fill=-7)
    assert cnt[0, 0] == 0 and (idx == -7).all()


def test_group_and_gather_are_exact_copies():
    rng = np.random.default_rng(4)
    pts = rng.standard_normal((2, 50, 7)).astype(np.float32)
    idx = rng.integers(0, 50, (2, 6, 5)).astype(np.int32)
    g = ops.group_point(pts, idx)
    for b in range(2):
        assert np.array_equal(g[b], pts[b][idx[b]])


def test_golden_interpolate_grad_config1():
    """three_interpolate_grad of the oracle == the fixture produced by the reference's own ThreeInterpolateGradOp."""
    xyz1, xyz2, _ = make_golden.interp_inputs()
    dist, idx = ops.three_nn(xyz1, xyz2)
    w = make_golden.fp_weights(dist)
    g = np.random.default_rng(2).standard_normal((1, 1024, 256)).astype(np.float32)
    gp = ops.three_interpolate_grad(256, idx, w, g)
    fx = golden("ref_interpolate_grad_config1")
    assert np.array_equal(gp[:, ::16], fx["grad_rows"])
    assert abs(float(gp.astype(np.float64).sum()) - float(fx["grad_sum"])) < 1e-9


# ------------------------------------------------------------------------------------------------ backward ops
@pytest.mark.parametrize("seed", [0, 1])
def test_grad_ops_oracle_vs_reference_and_numpy(seed):
    """The C restatement of the three scatter-add gradients: three_interpolate_grad is compared with the reference's own
    ThreeInterpolateGradOp::Compute (compiled unmodified) bit for bit; all three with a float64 numpy scatter-add, and
    with the adjoint identity <g, op(x)> == <op_grad(g), x> of the forward ops."""
    rng = np.random.default_rng(seed)
    b, n, m, c, ns = 2, 300, 64, 20, 8
    # three_interpolate_grad
    idx = rng.integers(0, m, (b, n, 3)).astype(np.int32)
    w = rng.random((b, n, 3), dtype=np.float32)
    g = rng.standard_normal((b, n, c)).astype(np.float32)
    got = ops.three_interpolate_grad(m, idx, w, g)
    if ops.ref.available:
        assert np.array_equal(got, ops.ref.three_interpolate_grad(m, idx, w, g))
    want = np.zeros((b, m, c))
    for i in range(b):
        for k in range(3):
            np.add.at(want[i], idx[i, :, k], g[i].astype(np.float64) * w[i, :, k:k + 1])
    assert np.abs(got - want).max() < 1e-4
    pts = rng.standard_normal((b, m, c)).astype(np.float32)
    lhs = float((ops.three_interpolate(pts, idx, w).astype(np.float64) * g).sum())
    assert abs(lhs - float((got.astype(np.float64) * pts).sum())) < 1e-3 * max(1.0, abs(lhs))
    # group_point_grad
    gidx = rng.integers(0, n, (b, m, ns)).astype(np.int32)
    gg = rng.standard_normal((b, m, ns, c)).astype(np.float32)
    got = ops.group_point_grad(n, gidx, gg)
    want = np.zeros((b, n, c))
    for i in range(b):
        np.add.at(want[i], gidx[i].reshape(-1), gg[i].reshape(-1, c).astype(np.float64))
    assert np.abs(got - want).max() < 1e-4
    x = rng.standard_normal((b, n, c)).astype(np.float32)
    lhs = float((ops.group_point(x, gidx).astype(np.float64) * gg).sum())
    assert abs(lhs - float((got.astype(np.float64) * x).sum())) < 1e-3 * max(1.0, abs(lhs))
    # gather_point_grad (duplicates in idx accumulate)
    pidx = rng.integers(0, 50, (b, m)).astype(np.int32)
    pg = rng.standard_normal((b, m, 3)).astype(np.float32)
    got = ops.gather_point_grad(n, pidx, pg)
    want = np.zeros((b, n, 3))
    for i in range(b):
        np.add.at(want[i], pidx[i], pg[i].astype(np.float64))
    assert np.abs(got - want).max() < 1e-5
    assert (got[:, 50:] == 0).all()

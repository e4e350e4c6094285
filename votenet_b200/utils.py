"""PointNet++ layer library — drop-in for the reference's utils.py (pointnet_sa_module / pointnet_fp_module,
/root/reference/utils.py:93-158,266-294) over CUDA tensors, with the fusion the reference cannot express:
group_point(xyz) - centroid, group_point(features), concat, the 1x1-conv stack, BN, ReLU and reduce_max are ONE kernel
(`vnb_sa_group_mlp_max`); three_interpolate + concat are one kernel feeding the tensor-core layers.

`scope` names the weight set exactly as in the reference (tf.variable_scope(scope), utils.py:114,277); weights come from
a `WeightStore` (BatchNorm folded, fp16 tensor-core images packed once).
"""
import torch

from . import tf_grouping, tf_interpolate, tf_sampling
from ._lib import check, dptr, lib, stream_ptr
from .weights import fold_bn

PRECISION_FP32 = 0     # fp32 SIMT kernels
PRECISION_TENSOR = 1   # fp16 operands on tcgen05 tensor cores, fp32 accumulate
FUSE_FP = True         # pointnet_fp_module: one fused kernel when the shapes allow it (tests flip this to cross-check)
HOIST_MIN_C = 14       # feature widths above 13 use the hoisted layer-1 formulation (see csrc/mlp_tc.cu)


class Layer:
    """One dense layer on the device: folded fp32 weight (cin,cout), bias, lazily packed tensor-core image."""

    def __init__(self, W, b, device):
        self.W = W.to(device=device, dtype=torch.float32).contiguous()
        self.b = b.to(device=device, dtype=torch.float32).contiguous()
        self.cin, self.cout = self.W.shape
        self._img = None

    @property
    def img(self):
        if self._img is None:
            nbytes = lib.vnb_weight_image_bytes(self.cin, self.cout)
            img = torch.empty((nbytes,), dtype=torch.uint8, device=self.W.device)
            check(lib.vnb_pack_weight_f16(self.cin, self.cout, dptr(self.W), dptr(img), stream_ptr()))
            self._img = img
        return self._img

    def rows(self, lo, hi):
        """Sub-layer using input rows [lo,hi) of W (no bias) — used to split W1 into its xyz / feature parts."""
        return Layer(self.W[lo:hi].contiguous(), torch.zeros_like(self.b), self.W.device)


class WeightStore:
    """Device-side weight set keyed by the reference's layer names ('sa1/conv0', 'fp1/conv_1', 'voting2', ...)."""

    def __init__(self, weights, device="cuda", eps=1e-5, precision=PRECISION_TENSOR):
        self.device = torch.device(device)
        self.precision = precision
        self.eps = eps
        self._raw = weights
        self._layers = {}

    def layer(self, name):
        if name not in self._layers:
            W, b = fold_bn(self._raw, name, self.eps)
            self._layers[name] = Layer(W, b, self.device)
        return self._layers[name]

    def derived(self, key, make):
        if key not in self._layers:
            self._layers[key] = make()
        return self._layers[key]


def linear(x, layer, act, precision, residual=None, out_f16=False):
    """rows x cin -> rows x cout through vnb_linear.  Returns fp32 (or fp16 when out_f16)."""
    rows, cin = x.shape
    assert cin == layer.cin
    if out_f16:
        out = torch.empty((rows, layer.cout), dtype=torch.float16, device=x.device)
        o32, o16 = None, out
    else:
        out = torch.empty((rows, layer.cout), dtype=torch.float32, device=x.device)
        o32, o16 = out, None
    check(lib.vnb_linear(rows, cin, layer.cout, dptr(x, torch.float32, "in"), dptr(layer.W),
                         dptr(layer.img) if precision == PRECISION_TENSOR else None, dptr(layer.b),
                         dptr(residual, torch.float32, "residual"), 1 if act else 0, dptr(o32), dptr(o16),
                         int(precision), stream_ptr()))
    return out


def sa_group_mlp_max(xyz, points, new_xyz, idx, layers, precision, store=None, scope=None, pts_cnt=None):
    """Fused grouping + 3-layer shared MLP + max-pool (utils.py:49-55,120-132).  pts_cnt (the ball query's second output)
    lets the tensor-core kernels skip the padded duplicate rows (same result, fewer tiles)."""
    b, n, _ = xyz.shape
    c = points.shape[2]
    m, ns = idx.shape[1], idx.shape[2]
    l1, l2, l3 = layers
    out = torch.empty((b, m, l3.cout), dtype=torch.float32, device=xyz.device)
    q = None
    w1_img = w2_img = w3_img = None
    ws = None
    if precision == PRECISION_TENSOR:
        ws = torch.empty((lib.vnb_sa_workspace_bytes(b, m, ns),), dtype=torch.uint8, device=xyz.device)
        w2_img, w3_img = l2.img, l3.img
        if c >= HOIST_MIN_C:
            # layer 1 hoisted through the gather: q = feat @ W1[3:] + b1, once per source point, fp16
            lf = store.derived(scope + "/conv0:feat", lambda: Layer(l1.W[3:].contiguous(), l1.b, l1.W.device))
            q = linear(points.reshape(b * n, c), lf, act=False, precision=precision, out_f16=True)
        else:
            w1_img = l1.img
    check(lib.vnb_sa_group_mlp_max_counted(b, n, c, m, ns, dptr(xyz, torch.float32, "xyz"),
                                           dptr(points, torch.float32, "points"), dptr(new_xyz, torch.float32, "new_xyz"),
                                           dptr(idx, torch.int32, "idx"), dptr(pts_cnt, torch.int32, "pts_cnt"), l1.cout,
                                           l2.cout, l3.cout, dptr(l1.W), dptr(l1.b), dptr(l2.W), dptr(l2.b), dptr(l3.W),
                                           dptr(l3.b), dptr(w1_img), dptr(w2_img), dptr(w3_img), dptr(q), dptr(out),
                                           int(precision), dptr(ws), stream_ptr()))
    return out


def pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2, group_all, scope, bn=True, pooling="max",
                       knn=False, use_xyz=True, use_nchw=False, sample_xyz=None, *, weights):
    """PointNet Set Abstraction module — signature of /root/reference/utils.py:93-94 plus the `weights` store.
    Returns (new_xyz (B,npoint,3), new_points (B,npoint,mlp[-1] or mlp2[-1]), idx (B,npoint,nsample))."""
    if group_all or knn or pooling != "max" or not use_xyz or points is None or not bn:
        raise NotImplementedError("only the configuration VoteNet uses is on the B200 hot path: group_all=False, "
                                  "knn=False, pooling='max', use_xyz=True, bn=True, points given (SURVEY.md §2.1)")
    if len(mlp) != 3:
        raise NotImplementedError("pointnet_sa_module: the fused kernel implements the 3-layer shared MLP VoteNet uses")
    prec = weights.precision
    # utils.py:42-45.  (the _nested entry point is bit-identical; it is only faster when the input is FPS-ordered,
    # which is the case for every VoteNet level but the first)
    fps_in = sample_xyz if sample_xyz is not None else xyz
    fps = tf_sampling.farthest_point_sample_nested if fps_in.shape[1] <= 4096 else tf_sampling.farthest_point_sample
    fps_idx = fps(npoint, fps_in)
    new_xyz = tf_sampling.gather_point(xyz, fps_idx)
    idx, pts_cnt = tf_grouping.query_ball_point(radius, nsample, xyz, new_xyz)                           # utils.py:49
    layers = [weights.layer(f"{scope}/conv{i}") for i in range(3)]
    for l, co in zip(layers, mlp):
        assert l.cout == co, f"{scope}: weight set does not match mlp={mlp}"
    new_points = sa_group_mlp_max(xyz, points, new_xyz, idx, layers, prec, weights, scope, pts_cnt)      # :50-55,120-132
    if mlp2 is not None:                                                                                  # :149-155
        b, m, c = new_points.shape
        h = new_points.reshape(b * m, c)
        for i in range(len(mlp2)):
            h = linear(h, weights.layer(f"{scope}/conv_post_{i}"), act=i < len(mlp2) - 1, precision=prec)
        new_points = h.reshape(b, m, -1)
    return new_xyz, new_points, idx


def _ptr_array(tensors):
    import ctypes as C
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _int_array(vals):
    import ctypes as C
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def vote_layers_fused(weights, names):
    """Weight set of the voting FC stack (model.py:53-57) in the form the fused kernel wants: first layer split into its
    xyz rows (fp32, rank-3 update) and feature rows (tensor-core image); last layer with its output columns permuted
    [features | xyz] so that the kernel writes votes_feat / votes_xyz directly."""
    l0, l1, l2 = (weights.layer(nm) for nm in names)
    dev = l0.W.device
    f0 = weights.derived(names[0] + ":feat", lambda: Layer(l0.W[3:].contiguous(), l0.b, dev))
    x0 = weights.derived(names[0] + ":xyz", lambda: Layer(l0.W[:3].contiguous(), torch.zeros_like(l0.b), dev))
    p2 = weights.derived(names[2] + ":perm", lambda: Layer(torch.cat([l2.W[:, 3:], l2.W[:, :3]], 1).contiguous(),
                                                           torch.cat([l2.b[3:], l2.b[:3]]).contiguous(), dev))
    return [f0, l1, p2], x0


def fp_module_fused(dist, idx, points1, points2, fp_layers, fp_out, vote=None, stream=None):
    """vnb_fp_module_fused: three_interpolate + concat + the fp 1x1-conv stack (+ the voting FC stack) in one kernel.
    vote = (layers, xyz_layer, seeds_xyz, votes_xyz, votes_feat) or None."""
    b, n, _ = idx.shape
    m, c2 = points2.shape[1], points2.shape[2]
    c1 = points1.shape[2]
    sp = stream_ptr(stream)
    if vote is None:
        check(lib.vnb_fp_module_fused(b, n, m, c1, c2, dptr(dist), dptr(idx), dptr(points1), dptr(points2), len(fp_layers),
                                      _ptr_array([l.img for l in fp_layers]), _ptr_array([l.b for l in fp_layers]),
                                      _int_array([l.cout for l in fp_layers]), dptr(fp_out), 0, None, None, None, None,
                                      None, None, None, sp))
    else:
        vl, x0, seeds_xyz, votes_xyz, votes_feat = vote
        check(lib.vnb_fp_module_fused(b, n, m, c1, c2, dptr(dist), dptr(idx), dptr(points1), dptr(points2), len(fp_layers),
                                      _ptr_array([l.img for l in fp_layers]), _ptr_array([l.b for l in fp_layers]),
                                      _int_array([l.cout for l in fp_layers]), dptr(fp_out), len(vl),
                                      _ptr_array([l.img for l in vl]), _ptr_array([l.b for l in vl]),
                                      _int_array([vl[0].cout, vl[1].cout, vl[2].cout]), dptr(x0.W), dptr(seeds_xyz),
                                      dptr(votes_xyz), dptr(votes_feat), sp))


def fp_fusable(points1, points2, mlp, precision):
    return (precision == PRECISION_TENSOR and points1 is not None and points1.shape[-1] == 256 and points2.shape[-1] == 256
            and tuple(mlp) == (256, 256))


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, scope, bn=True, *, weights):
    """PointNet Feature Propagation module — signature of /root/reference/utils.py:266 plus the `weights` store."""
    if not bn:
        raise NotImplementedError("bn=False is not on the VoteNet path")
    prec = weights.precision
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    c2 = points2.shape[2]
    c1 = points1.shape[2] if points1 is not None else 0
    dist, idx = tf_interpolate.three_nn(xyz1, xyz2)                                            # utils.py:278
    if fp_fusable(points1, points2, mlp, prec) and FUSE_FP:                                    # :279-292 in one kernel
        out = torch.empty((b, n, mlp[-1]), dtype=torch.float32, device=xyz1.device)
        fp_module_fused(dist, idx, points1.contiguous(), points2.contiguous(),
                        [weights.layer(f"{scope}/conv_{i}") for i in range(len(mlp))], out)
        return out
    cat = torch.empty((b * n, c2 + c1), dtype=torch.float32, device=xyz1.device)
    check(lib.vnb_fp_interpolate_concat(b, n, m, c1, c2, dptr(dist), dptr(idx), dptr(points1, torch.float32, "points1"),
                                        dptr(points2, torch.float32, "points2"), dptr(cat), stream_ptr()))  # :279-286
    h = cat
    for i in range(len(mlp)):                                                                  # :290-292
        h = linear(h, weights.layer(f"{scope}/conv_{i}"), act=True, precision=prec)
    return h.reshape(b, n, -1)

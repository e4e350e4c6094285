"""TEST INFRASTRUCTURE — numpy front-end of the CPU oracle.  NOT product code.

Function names, argument order and output dtypes/shapes mirror the reference's Python wrappers
(/root/reference/tf_ops/sampling/tf_sampling.py:29-56, grouping/tf_grouping.py:8-41,
3d_interpolation/tf_interpolate.py:8-28, 3d_nms/tf_nms3d.py:11-12) so parity tests read like reference calls.

``ops.<fn>``      -> our C restatement (oracle/oracle.c)
``ops.ref.<fn>``  -> the reference's own CPU sources compiled unmodified (oracle/_ref/libvotenet_ref_cpu.so);
                     ``ops.ref.available`` tells whether that library was built.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_REF_CPU_PATH = os.path.join(_HERE, "_ref", "libvotenet_ref_cpu.so")
REF_GPU_PATH = os.path.join(_HERE, "_ref", "libvotenet_ref_gpu.so")


def build(force=False):
    """Compile the oracle (and, if /root/reference is mounted, the reference itself) via oracle/Makefile."""
    if force or not os.path.exists(_LIB_PATH) or os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", _HERE] + (["-B"] if force else []), check=True)


def _load():
    if not os.path.exists(_LIB_PATH):
        build()
    return C.CDLL(_LIB_PATH)


_lib = _load()
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


for _name in ("vno_intersection2d", "vno_area2d", "vno_area3d", "vno_iou3d"):
    getattr(_lib, _name).restype = C.c_float


def farthest_point_sample(npoint, inp):
    """(b,n,3) f32 -> (b,npoint) i32.  Restates tf_sampling_g.cu:105-170."""
    inp, pi = _f(inp)
    b, n, _ = inp.shape
    out = np.zeros((b, npoint), np.int32)
    _lib.vno_fps(b, n, int(npoint), pi, out.ctypes.data_as(_ip))
    return out


def gather_point(inp, idx):
    inp, pi = _f(inp)
    idx, pidx = _i(idx)
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = np.zeros((b, m, 3), np.float32)
    _lib.vno_gather_point(b, n, m, pi, pidx, out.ctypes.data_as(_fp))
    return out


def query_ball_point(radius, nsample, xyz1, xyz2, fill=0):
    """-> (idx (b,m,nsample) i32, pts_cnt (b,m) i32).  Rows of empty balls keep `fill` (reference leaves them
    uninitialised, tf_grouping_g.cu:14-34)."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.full((b, m, nsample), fill, np.int32)
    cnt = np.zeros((b, m), np.int32)
    _lib.vno_query_ball_point(b, n, m, C.c_float(radius), int(nsample), p1, p2, idx.ctypes.data_as(_ip),
                              cnt.ctypes.data_as(_ip))
    return idx, cnt


def group_point(points, idx):
    points, pp = _f(points)
    idx, pidx = _i(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.zeros((b, m, ns, c), np.float32)
    _lib.vno_group_point(b, n, c, m, ns, pp, pidx, out.ctypes.data_as(_fp))
    return out


def three_nn(xyz1, xyz2):
    """xyz1 (b,n,3) unknown, xyz2 (b,m,3) known -> (dist (b,n,3) squared f32, idx (b,n,3) i32)."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.zeros((b, n, 3), np.float32)
    idx = np.zeros((b, n, 3), np.int32)
    _lib.vno_three_nn(b, n, m, p1, p2, dist.ctypes.data_as(_fp), idx.ctypes.data_as(_ip))
    return dist, idx


def three_interpolate(points, idx, weight):
    points, pp = _f(points)
    idx, pidx = _i(idx)
    weight, pw = _f(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.zeros((b, n, c), np.float32)
    _lib.vno_three_interpolate(b, m, c, n, pp, pidx, pw, out.ctypes.data_as(_fp))
    return out


def gather_point_grad(n, idx, out_g):
    """out_g (b,m,3), idx (b,m) -> inp_g (b,n,3).  Restates tf_sampling_g.cu:183-192."""
    out_g, pg = _f(out_g)
    idx, pidx = _i(idx)
    b, m = idx.shape
    inp_g = np.zeros((b, n, 3), np.float32)
    _lib.vno_gather_point_grad(b, int(n), m, pg, pidx, inp_g.ctypes.data_as(_fp))
    return inp_g


def group_point_grad(n, idx, grad_out):
    """grad_out (b,m,ns,c), idx (b,m,ns) -> grad_points (b,n,c).  Restates tf_grouping_g.cu:61-78."""
    grad_out, pg = _f(grad_out)
    idx, pidx = _i(idx)
    b, m, ns = idx.shape
    c = grad_out.shape[3]
    gp = np.zeros((b, n, c), np.float32)
    _lib.vno_group_point_grad(b, int(n), c, m, ns, pg, pidx, gp.ctypes.data_as(_fp))
    return gp


def three_interpolate_grad(m, idx, weight, grad_out):
    """grad_out (b,n,c), idx (b,n,3), weight (b,n,3) -> grad_points (b,m,c).  Restates tf_interpolate.cpp:131-153."""
    grad_out, pg = _f(grad_out)
    idx, pidx = _i(idx)
    weight, pw = _f(weight)
    b, n, c = grad_out.shape
    gp = np.zeros((b, m, c), np.float32)
    _lib.vno_three_interpolate_grad(b, n, c, int(m), pg, pidx, pw, gp.ctypes.data_as(_fp))
    return gp


def intersection2d(box1, box2):
    _, p1 = _f(box1)
    _, p2 = _f(box2)
    a1, p1 = _f(box1)
    a2, p2 = _f(box2)
    return float(_lib.vno_intersection2d(p1, p2))


def iou3d(box_i, box_j):
    a1, p1 = _f(box_i)
    a2, p2 = _f(box_j)
    return float(_lib.vno_iou3d(p1, p2))


def NMS3D(bboxes, scores, objectiveness, iou_threshold, return_keep=False):
    """-> (Nnms,2) i32 rows (batch, box) in the reference's global pop order (tf_nms3d.cpp:202-273).
    Raises ValueError where the reference op raises InvalidArgument (:287-300)."""
    bboxes, pb = _f(bboxes)
    scores, ps = _f(scores)
    objectiveness, po = _f(objectiveness)
    if bboxes.ndim != 4 or bboxes.shape[2:] != (8, 3):
        raise ValueError("3D NMS expects (batch_size, nbbox, 8, 3) bbox shape.")
    b, k = bboxes.shape[:2]
    if scores.shape != (b, k):
        raise ValueError("3D NMS expects (batch_size, nbbox) scores shape.")
    if objectiveness.shape != (b, k, 2):
        raise ValueError("3D NMS expects (batch_size, nbbox, 2) objectiveness shape.")
    out = np.zeros((max(b * k, 1), 2), np.int32)
    keep = np.zeros((b, k), np.uint8)
    cnt = _lib.vno_nms3d(b, k, pb, ps, po, C.c_float(float(iou_threshold)), out.ctypes.data_as(_ip),
                         keep.ctypes.data_as(C.POINTER(C.c_ubyte)))
    if cnt < 0:
        raise ValueError("iou_threshold must be in [0, 1]")
    res = out[:cnt].copy()
    return (res, keep.astype(bool)) if return_keep else res


class _Ref:
    """The reference's own CPU op sources, compiled unmodified (oracle/Makefile -> oracle/_ref/)."""

    def __init__(self):
        self.lib = C.CDLL(_REF_CPU_PATH) if os.path.exists(_REF_CPU_PATH) else None
        if self.lib is not None:
            for nm in ("ref_intersection2d", "ref_area2d", "ref_area3d"):
                getattr(self.lib, nm).restype = C.c_float

    @property
    def available(self):
        return self.lib is not None

    def three_nn(self, xyz1, xyz2):
        xyz1, p1 = _f(xyz1)
        xyz2, p2 = _f(xyz2)
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dist = np.zeros((b, n, 3), np.float32)
        idx = np.zeros((b, n, 3), np.int32)
        rc = self.lib.ref_three_nn(b, n, m, p1, p2, dist.ctypes.data_as(_fp), idx.ctypes.data_as(_ip))
        assert rc == 0
        return dist, idx

    def three_interpolate(self, points, idx, weight):
        points, pp = _f(points)
        idx, pidx = _i(idx)
        weight, pw = _f(weight)
        b, m, c = points.shape
        n = idx.shape[1]
        out = np.zeros((b, n, c), np.float32)
        rc = self.lib.ref_three_interpolate(b, m, c, n, pp, pidx, pw, out.ctypes.data_as(_fp))
        assert rc == 0
        return out

    def three_interpolate_grad(self, m, idx, weight, grad_out):
        """The reference's ThreeInterpolateGradOp::Compute (tf_interpolate.cpp:225-262), unmodified."""
        grad_out, pg = _f(grad_out)
        idx, pidx = _i(idx)
        weight, pw = _f(weight)
        b, n, c = grad_out.shape
        gp = np.zeros((b, m, c), np.float32)
        rc = self.lib.ref_three_interpolate_grad(b, n, c, int(m), pg, pidx, pw, gp.ctypes.data_as(_fp))
        assert rc == 0
        return gp

    def intersection2d(self, box1, box2):
        a1, p1 = _f(box1)
        a2, p2 = _f(box2)
        return float(self.lib.ref_intersection2d(p1, p2))

    def iou_greater(self, boxes, b, i, j, thr):
        boxes, pb = _f(boxes)
        return bool(self.lib.ref_iou_greater(pb, int(b), int(i), int(j), boxes.shape[1], C.c_float(thr)))

    def NMS3D(self, bboxes, scores, objectiveness, iou_threshold):
        bboxes, pb = _f(bboxes)
        scores, ps = _f(scores)
        objectiveness, po = _f(objectiveness)
        b, k = bboxes.shape[:2]
        out = np.zeros((max(b * k, 1), 2), np.int32)
        cnt = self.lib.ref_nms3d(b, k, pb, ps, po, C.c_float(float(iou_threshold)), out.ctypes.data_as(_ip))
        if cnt < 0:
            raise ValueError("reference op rejected its inputs (InvalidArgument)")
        return out[:cnt].copy()


ref = _Ref()

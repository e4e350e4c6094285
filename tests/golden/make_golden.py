"""Regenerates tests/golden/*.npz.  Run in the BUILD container (needs /root/reference mounted so that
oracle/_ref/libvotenet_ref_cpu.so — the reference's own tf_interpolate.cpp / tf_nms3d.cpp compiled unmodified — exists):

    python tests/golden/make_golden.py

Fixtures whose expected outputs come from the REAL reference code are named ref_*; fixtures produced by the C oracle
(FPS / ball query — the reference only has GPU kernels for those, pinned on the GPU box by tests/test_gpu_ref_kernels.py)
are named oracle_*.  Inputs are regenerated from the recorded seeds; only outputs (small) are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import random_boxes  # noqa: E402
from oracle import ops  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def interp_inputs():
    # BASELINE.json configs[0]: three_nn + three_interpolate on 1024 synthetic points (B=1, m=256 known, C=256)
    rng = np.random.default_rng(1)
    xyz1 = rng.random((1, 1024, 3), dtype=np.float32)
    xyz2 = rng.random((1, 256, 3), dtype=np.float32)
    pts = rng.standard_normal((1, 256, 256)).astype(np.float32)
    return xyz1, xyz2, pts


def fp_weights(dist):
    d = np.maximum(dist, np.float32(1e-10))
    r = (np.float32(1.0) / d).astype(np.float32)
    norm = r.sum(axis=2, keepdims=True, dtype=np.float32)
    return (r / norm).astype(np.float32)


def nms_inputs(seed, b, k, degenerate=False):
    rng = np.random.default_rng(seed)
    boxes = random_boxes(rng, b, k, spread=1.5, degenerate=degenerate)
    scores = rng.standard_normal((b, k)).astype(np.float32)
    obj = rng.standard_normal((b, k, 2)).astype(np.float32)
    return boxes, scores, obj


def fps_inputs():
    rng = np.random.default_rng(7)
    return rng.random((2, 3000, 3), dtype=np.float32)


def main():
    assert ops.ref.available, "build oracle/_ref first (make -C oracle) — needs /root/reference"
    xyz1, xyz2, pts = interp_inputs()
    dist, idx = ops.ref.three_nn(xyz1, xyz2)
    w = fp_weights(dist)
    out = ops.ref.three_interpolate(pts, idx, w)
    np.savez_compressed(os.path.join(HERE, "ref_interpolate_config1.npz"), dist=dist, idx=idx, out_rows=out[:, ::32],
                        out_sum=np.float64(out.astype(np.float64).sum()))
    for name, seed, b, k, deg, thr in (("ref_nms_random", 11, 2, 96, False, 0.25), ("ref_nms_degenerate", 12, 2, 64, True, 0.25),
                                       ("ref_nms_thr0", 13, 1, 48, False, 0.0)):
        boxes, scores, obj = nms_inputs(seed, b, k, deg)
        sel = ops.ref.NMS3D(boxes, scores, obj, thr)
        inter = np.array([ops.ref.intersection2d(boxes[0, i], boxes[0, j]) for i in range(16) for j in range(16)], np.float32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, b=b, k=k, degenerate=deg, thr=thr, selected=sel,
                            inter2d_16x16=inter)
    # ThreeInterpolateGrad from the reference's own op: same config-1 neighbours, seeded grad_out
    g = np.random.default_rng(2).standard_normal((1, 1024, 256)).astype(np.float32)
    gp = ops.ref.three_interpolate_grad(256, idx, w, g)
    np.savez_compressed(os.path.join(HERE, "ref_interpolate_grad_config1.npz"), grad_rows=gp[:, ::16],
                        grad_sum=np.float64(gp.astype(np.float64).sum()))
    x = fps_inputs()
    f = ops.farthest_point_sample(64, x)
    nx = ops.gather_point(x, f)
    bi, bc = ops.query_ball_point(0.12, 16, x, nx)
    np.savez_compressed(os.path.join(HERE, "oracle_fps_ballquery.npz"), fps=f, ball_idx=bi, ball_cnt=bc)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

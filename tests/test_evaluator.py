"""Evaluator: the numpy restatement of /root/reference/evaluator.py is pinned on hand-checkable
boxes (CPU), and the device evaluator (csrc/eval_ap.cu) is compared with it (`-m gpu`)."""
import numpy as np
import pytest

from conftest import random_boxes
from oracle import evalref as E


def _cube(cx=0.0, cy=0.0, cz=0.0, l=1.0, w=1.0, h=1.0, yaw=0.0):
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1]); sy = np.array([1, 1, 1, 1, -1, -1, -1, -1]); sz = np.array([1, -1, -1, 1, 1, -1, -1, 1])
    x, y, z = sx * l / 2, sy * h / 2, sz * w / 2
    c, s = np.cos(yaw), np.sin(yaw)
    return np.stack([c * x + s * z + cx, y + cy, -s * x + c * z + cz], -1).astype(np.float32)


def test_oracle_iou3d_known_answers():
    a = _cube()
    assert abs(E.iou_3d(a, a) - 1.0) < 1e-12
    assert abs(E.iou_3d(a, _cube(cx=0.5)) - (0.5 / 1.5)) < 1e-7          # half-shifted unit cubes: 0.5 / (1 + 1 - 0.5)
    assert abs(E.iou_3d(a, _cube(cy=0.5)) - (0.5 / 1.5)) < 1e-7          # height overlap only
    assert E.iou_3d(a, _cube(cx=3.0)) == 0.0
    oct_area = 2 * (np.sqrt(2) - 1)                                       # unit square x the same square turned 45 degrees
    assert abs(E.iou_3d(a, _cube(yaw=np.pi / 4)) - oct_area / (2 - oct_area)) < 1e-6
    from test_oracle import _demo_boxes                                   # tf_nms3d.py:30-46: IoU 0.491408676 in the reference's float clip
    bb = _demo_boxes()[0]
    assert abs(E.iou_3d(bb[0], bb[1]) - 0.491408676) < 2e-6


def test_oracle_voc_ap_and_matching():
    assert abs(E.voc_ap(np.array([0.5, 1.0]), np.array([1.0, 1.0])) - 1.0) < 1e-12
    assert abs(E.voc_ap(np.array([0.5, 0.5, 1.0]), np.array([1.0, 0.5, 2 / 3])) - (0.5 * 1.0 + 0.5 * 2 / 3)) < 1e-12
    g = _cube()
    pred = {0: [(_cube(cx=0.1), 0.9), (_cube(cx=0.05), 0.8), (_cube(cx=5.0), 0.7)]}   # 2nd hits an already-claimed box
    rec, prec, ap = E.eval_det_cls(pred, {0: [g]}, 0.25)
    assert rec.tolist() == [1.0, 1.0, 1.0] and np.allclose(prec, [1.0, 0.5, 1 / 3]) and abs(ap - 1.0) < 1e-12


def _random_eval_case(seed, nimg=12, ndet=40, ngt=6):
    rng = np.random.default_rng(seed)
    pred_all, gt_all = {}, {}
    names = ("bed", "table", "chair")
    for img in range(nimg):
        gtb = random_boxes(rng, 1, ngt, spread=2.0)[0]
        gt_all[img] = [(names[int(rng.integers(0, 3))], gtb[j]) for j in range(ngt)] if img != 3 else []
        det = random_boxes(rng, 1, ndet, spread=2.0)[0]
        for j in range(min(ngt, ndet) // 2):       # some detections near a ground truth
            det[j] = gtb[j] + rng.normal(0, 0.05, (1, 3)).astype(np.float32)
        sc = rng.random(ndet).astype(np.float32)
        sc[5] = sc[6]                               # an exact confidence tie
        pred_all[img] = [(names[int(rng.integers(0, 3))], det[j], float(sc[j])) for j in range(ndet)] if img != 7 else []
    for b in [v[1] for img in gt_all for v in gt_all[img]] + [v[1] for img in pred_all for v in pred_all[img]]:
        if not b[0, 1] > b[4, 1]:
            b[:, 1] = b[::-1, 1]
    return pred_all, gt_all


@pytest.mark.gpu
def test_device_iou3d_matches_oracle(cuda):
    from votenet_b200.evaluator import iou_3d

    rng = np.random.default_rng(3)
    a = random_boxes(rng, 1, 300, spread=1.0)[0]
    b = random_boxes(rng, 1, 300, spread=1.0)[0]
    b[:20] = a[:20]
    got = iou_3d(a, b)
    ref = np.array([E.iou_3d(a[i], b[i]) for i in range(300)])
    assert np.abs(got - ref).max() < 1e-9 and (ref > 0.05).sum() > 30 and np.allclose(got[:20], 1.0)
    assert abs(iou_3d(_cube(), _cube(yaw=np.pi / 4)) - E.iou_3d(_cube(), _cube(yaw=np.pi / 4))) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1])
def test_device_eval_det_matches_oracle(cuda, seed):
    from votenet_b200.evaluator import eval_det

    pred_all, gt_all = _random_eval_case(seed)
    rec, prec, ap = eval_det(pred_all, gt_all, 0.25)
    orec, oprec, oap = E.eval_det(pred_all, gt_all, 0.25)
    assert set(ap) == set(oap) and len(ap) == 3
    for c in oap:
        assert np.allclose(rec[c], orec[c], rtol=0, atol=1e-12) and np.allclose(prec[c], oprec[c], rtol=0, atol=1e-12), c
        assert abs(ap[c] - oap[c]) < 1e-12, (c, ap[c], oap[c])
    assert 0.0 < np.mean(list(ap.values())) < 1.0


@pytest.mark.gpu
def test_evaluate_detections_from_forward_outputs(cuda):
    """The forward's output gathers feed the evaluator directly (evaluator.py:214-229)."""
    import torch

    from votenet_b200 import synth
    from votenet_b200.evaluator import TYPE_WHITELIST, evaluate_detections
    from votenet_b200.model import decode_nms3d

    rng = np.random.default_rng(5)
    b, k = 4, 128
    pxyz = torch.as_tensor(rng.uniform(-2, 2, (b, k, 3)).astype(np.float32), device=cuda)
    pout = torch.as_tensor(rng.standard_normal((b, k, 79)).astype(np.float32), device=cuda)
    o = decode_nms3d(pxyz, pout, torch.as_tensor(np.asarray(synth.CLASS_MEAN_SIZE, np.float32), device=cuda), 0.25)
    n = int(o["nms_count"].item())
    bb = o["bboxes_pred"][:n].cpu().numpy()
    gt_all = {i: [] for i in range(b)}
    cls = o["class_scores_pred"][:n].cpu().numpy().argmax(-1)
    bi = o["batch_idx"][:n].cpu().numpy()
    for r in range(0, n, 3):   # every third detection is "correct": its own box is a ground truth of its class
        gt_all[int(bi[r])].append((TYPE_WHITELIST[int(cls[r])], bb[r]))
    m, ap = evaluate_detections(o["bboxes_pred"], o["class_scores_pred"], o["batch_idx"], n, gt_all)
    pred_all = {}
    sc = o["class_scores_pred"][:n].cpu().numpy().max(-1)
    for r in range(n):
        pred_all.setdefault(int(bi[r]), []).append((TYPE_WHITELIST[int(cls[r])], bb[r], float(sc[r])))
    _, _, oap = E.eval_det(pred_all, gt_all, 0.25)
    assert set(ap) == set(oap) and all(abs(ap[c] - oap[c]) < 1e-12 for c in oap) and abs(m - np.mean(list(oap.values()))) < 1e-12

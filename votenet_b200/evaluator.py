"""Detection evaluator on the GPU — the host-side mirror of /root/reference/evaluator.py (same function names, argument
meaning and return values: iou_3d :26-39, eval_det_cls :77-151, eval_det :154-200, the AP@0.25 the Evaluator callback
reports :203-231), calling csrc/eval_ap.cu through the C ABI.  SURVEY.md §8(f) rank 3.

The reference loops in Python with one shapely call per (detection, ground truth) pair; here one launch matches every
detection of a class against the ground truths of its image and a second one produces recall / precision / AP.
`evaluate_detections` consumes the forward's output gathers (bboxes_pred, class_scores_pred, batch_idx; model.py:135-137)
directly, as Evaluator._trigger does (:214-229)."""
import numpy as np
import torch

from ._lib import check, dptr, lib, stream_ptr

TYPE_WHITELIST = ("bed", "table", "sofa", "chair", "toilet", "desk", "dresser", "night_stand", "bookshelf", "bathtub")  # :19-20


def _dev(device=None):
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def iou_3d(bbox1, bbox2, device=None):
    """IoU of two (8,3) corner boxes (or two (n,8,3) stacks -> (n,) array).  evaluator.py:26-39."""
    a = torch.as_tensor(np.asarray(bbox1, np.float32), device=_dev(device)).reshape(-1, 8, 3).contiguous()
    b = torch.as_tensor(np.asarray(bbox2, np.float32), device=_dev(device)).reshape(-1, 8, 3).contiguous()
    if a.shape != b.shape:
        raise ValueError("iou_3d: both arguments must hold the same number of (8,3) boxes")
    if not bool((a[:, 0, 1] > a[:, 4, 1]).all() and (b[:, 0, 1] > b[:, 4, 1]).all()):
        raise AssertionError("iou_3d: corner 0 must lie above corner 4 (evaluator.py:33)")
    out = torch.empty((a.shape[0],), dtype=torch.float64, device=a.device)
    check(lib.vnb_iou3d_pairs(a.shape[0], dptr(a), dptr(b), dptr(out), stream_ptr()))
    o = out.cpu().numpy()
    return float(o[0]) if np.asarray(bbox1).ndim == 2 else o


def eval_det_cls_tensors(det_boxes, det_scores, det_img, gt_boxes, gt_offsets, ovthresh=0.25):
    """Flat-tensor form (all CUDA): det_boxes (nd,8,3) f32, det_scores (nd) f32, det_img (nd) i32, gt_boxes (ng,8,3) f32
    grouped by image, gt_offsets (nimg+1) i32 -> rec (nd) f64, prec (nd) f64, ap (1) f64 in descending-confidence order."""
    nd, ng, nimg = det_boxes.shape[0], gt_boxes.shape[0], gt_offsets.shape[0] - 1
    dev = det_scores.device
    rec = torch.empty((max(nd, 1),), dtype=torch.float64, device=dev)
    prec = torch.empty((max(nd, 1),), dtype=torch.float64, device=dev)
    ap = torch.zeros((1,), dtype=torch.float64, device=dev)
    ws = torch.empty((lib.vnb_eval_det_cls_workspace_bytes(nd, ng),), dtype=torch.uint8, device=dev)
    check(lib.vnb_eval_det_cls(nd, ng, nimg, dptr(det_boxes, torch.float32), dptr(det_scores, torch.float32),
                               dptr(det_img, torch.int32), dptr(gt_boxes, torch.float32) if ng else None,
                               dptr(gt_offsets, torch.int32), float(ovthresh), dptr(rec), dptr(prec), dptr(ap), dptr(ws),
                               stream_ptr()))
    return rec[:nd], prec[:nd], ap


def eval_det_cls(pred, gt, ovthresh=0.25, device=None):
    """pred {img_id: [(bbox (8,3), score)]}, gt {img_id: [bbox]} -> rec, prec (numpy, length nd), ap (float).
    evaluator.py:77-151; equal confidences are ordered as they are listed (the reference's argsort is not stable)."""
    dev = _dev(device)
    img_ids = list(gt.keys()) + [i for i in pred.keys() if i not in gt]      # :95-103
    index = {img: n for n, img in enumerate(img_ids)}
    gtb, offs = [], [0]
    for img in img_ids:
        for bb in gt.get(img, []):
            gtb.append(np.asarray(bb, np.float32).reshape(8, 3))
        offs.append(len(gtb))
    db, ds, di = [], [], []
    for img in pred.keys():                                                     # :106-113
        for box, score in pred[img]:
            db.append(np.asarray(box, np.float32).reshape(8, 3)); ds.append(float(score)); di.append(index[img])
    nd = len(db)
    T = lambda a, dt: torch.as_tensor(np.asarray(a, dt), device=dev).contiguous()  # noqa: E731
    det_boxes = T(np.stack(db) if nd else np.zeros((0, 8, 3)), np.float32)
    gt_boxes = T(np.stack(gtb) if gtb else np.zeros((0, 8, 3)), np.float32)
    rec, prec, ap = eval_det_cls_tensors(det_boxes, T(ds, np.float32), T(di, np.int32), gt_boxes, T(offs, np.int32), ovthresh)
    return rec.cpu().numpy(), prec.cpu().numpy(), float(ap.item())


def eval_det(pred_all, gt_all, ovthresh=0.25, device=None):
    """pred_all {img_id: [(classname, bbox, score)]}, gt_all {img_id: [(classname, bbox)]} -> rec, prec, ap dicts by
    class.  evaluator.py:154-200."""
    pred, gt = {}, {}
    for img_id in pred_all.keys():
        for classname, bbox, score in pred_all[img_id]:
            pred.setdefault(classname, {}).setdefault(img_id, [])
            gt.setdefault(classname, {}).setdefault(img_id, [])
            pred[classname][img_id].append((bbox, score))
    for img_id in gt_all.keys():
        for classname, bbox in gt_all[img_id]:
            gt.setdefault(classname, {}).setdefault(img_id, [])
            pred.setdefault(classname, {})
            gt[classname][img_id].append(bbox)
    rec, prec, ap = {}, {}, {}
    for classname in gt.keys():
        rec[classname], prec[classname], ap[classname] = eval_det_cls(pred[classname], gt[classname], ovthresh, device)
    return rec, prec, ap


def evaluate_detections(bboxes_pred, class_scores_pred, batch_idx, count, gt_all, ovthresh=0.25, class_names=TYPE_WHITELIST):
    """mAP from the forward's output gathers (model.py:135-137), as Evaluator._trigger does (evaluator.py:214-229):
    class = argmax of the class scores, confidence = their max.  bboxes_pred (n,8,3), class_scores_pred (n,10),
    batch_idx (n) device tensors with `count` valid rows; gt_all {batch index: [(classname, bbox)]}.
    -> (mAP, {classname: ap})."""
    n = int(count)
    bb = bboxes_pred[:n].cpu().numpy()
    cs = class_scores_pred[:n].cpu().numpy()
    bi = batch_idx[:n].cpu().numpy()
    cls = cs.argmax(-1) if n else np.zeros((0,), np.int64)
    sc = cs.max(-1) if n else np.zeros((0,), np.float32)
    pred_all = {}
    for r in range(n):
        pred_all.setdefault(int(bi[r]), []).append((class_names[int(cls[r])], bb[r], float(sc[r])))
    _, _, ap = eval_det(pred_all, gt_all, ovthresh, device=bboxes_pred.device)
    return (float(np.mean([ap[c] for c in ap])) if ap else 0.0), ap

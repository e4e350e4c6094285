"""TEST INFRASTRUCTURE — not product code (only tests/ may import this).

CPU restatement of the reference evaluator (/root/reference/evaluator.py): iou_3d (:26-39), voc_ap (:42-74),
eval_det_cls (:77-151), eval_det (:154-200), in plain numpy/Python loops.  shapely (the reference's polygon engine) is
absent from this image: the BEV intersection is restated as exact convex clipping (Sutherland-Hodgman) in float64 —
what GEOS computes for two convex quadrilaterals up to rounding.  Pinned in tests/test_oracle.py on hand-checkable
boxes (identical boxes, half-shifted cubes, the 45-degree octagon, disjoint boxes) and on the reference's own NMS demo
pair (tf_nms3d.py:30-46, IoU 0.4914).  Parity unpinned against shapely itself (not installable here)."""
import numpy as np


def _signed_area(p):
    x, z = p[:, 0], p[:, 1]
    return 0.5 * float(np.sum(x * np.roll(z, -1) - np.roll(x, -1) * z))


def polygon_intersection_area(a, b):
    """a, b: (4,2) convex quadrilaterals (any winding) -> area of their intersection."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if _signed_area(b) < 0:
        b = b[::-1]
    poly = [tuple(p) for p in a]
    for e in range(4):
        c0, c1 = b[e], b[(e + 1) % 4]
        out = []
        n = len(poly)
        for i in range(n):
            p, q = poly[i], poly[(i + 1) % n]
            sp = (c1[0] - c0[0]) * (p[1] - c0[1]) - (c1[1] - c0[1]) * (p[0] - c0[0])
            sq = (c1[0] - c0[0]) * (q[1] - c0[1]) - (c1[1] - c0[1]) * (q[0] - c0[0])
            pin, qin = sp >= 0.0, sq >= 0.0
            if pin:
                out.append(p)
            if pin != qin:
                t = sp / (sp - sq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
        poly = out
        if not poly:
            return 0.0
    return abs(_signed_area(np.asarray(poly, np.float64)))


def iou_3d(bbox1, bbox2):
    """evaluator.py:26-39."""
    bbox1 = np.asarray(bbox1, np.float64)
    bbox2 = np.asarray(bbox2, np.float64)
    p1 = np.stack([bbox1[:4, 0], bbox1[:4, 2]], -1)
    p2 = np.stack([bbox2[:4, 0], bbox2[:4, 2]], -1)
    inter_area = polygon_intersection_area(p1, p2)
    inter_vol = inter_area * max(0.0, min(bbox1[0, 1], bbox2[0, 1]) - max(bbox1[4, 1], bbox2[4, 1]))
    return inter_vol / (abs(_signed_area(p1)) * (bbox1[0, 1] - bbox1[4, 1]) + abs(_signed_area(p2)) * (bbox2[0, 1] - bbox2[4, 1]) - inter_vol)


def voc_ap(rec, prec):
    """evaluator.py:58-74 (the non-07 metric)."""
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def eval_det_cls(pred, gt, ovthresh=0.25):
    """evaluator.py:77-151.  pred {img_id: [(bbox, score)]}, gt {img_id: [bbox]} -> rec, prec, ap.  Equal confidences
    keep their input order (the reference's np.argsort is not stable; the product documents the stable order)."""
    class_recs = {}
    npos = 0
    for img_id in gt.keys():
        bbox = np.array(gt[img_id])
        class_recs[img_id] = {"bbox": bbox, "det": [False] * len(bbox)}
        npos += len(bbox)
    for img_id in pred.keys():
        if img_id not in gt:
            class_recs[img_id] = {"bbox": np.array([]), "det": []}
    image_ids, confidence, BB = [], [], []
    for img_id in pred.keys():
        for box, score in pred[img_id]:
            image_ids.append(img_id); confidence.append(score); BB.append(box)
    confidence = np.array(confidence)
    BB = np.array(BB)
    sorted_ind = np.argsort(-confidence, kind="stable")
    BB = BB[sorted_ind, ...]
    image_ids = [image_ids[x] for x in sorted_ind]
    nd = len(image_ids)
    tp, fp = np.zeros(nd), np.zeros(nd)
    for d in range(nd):
        R = class_recs[image_ids[d]]
        bb = BB[d, :].astype(float)
        ovmax, jmax = -np.inf, -1
        BBGT = R["bbox"].astype(float)
        if BBGT.size > 0:
            for j in range(BBGT.shape[0]):
                iou = iou_3d(bb, BBGT[j, ...])
                if iou > ovmax:
                    ovmax, jmax = iou, j
        if ovmax > ovthresh:
            if not R["det"][jmax]:
                tp[d] = 1.
                R["det"][jmax] = 1
            else:
                fp[d] = 1.
        else:
            fp[d] = 1.
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec)


def eval_det(pred_all, gt_all, ovthresh=0.25):
    """evaluator.py:154-200."""
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
        rec[classname], prec[classname], ap[classname] = eval_det_cls(pred[classname], gt[classname], ovthresh)
    return rec, prec, ap


def prepare_input(raw_upright_depth, draws):
    """/root/reference/dataset.py:185-190,302-308 + sunutils.py:70-77,133-139 for a batch: subsample, flip axes to the
    upright camera frame, then (when drawn) flip x / flip z / rotate about y / scale, in float64 like the reference;
    returned as float32 (what the TF placeholder receives)."""
    out = []
    for b in range(raw_upright_depth.shape[0]):
        pc = np.asarray(raw_upright_depth[b], np.float64)
        pc = pc[draws["choice"][b], :]                       # :185-186
        cam = np.copy(pc)
        cam[:, [0, 1, 2]] = cam[:, [0, 2, 1]]                # sunutils.py:75
        cam[:, 1] *= -1                                      # :76
        if draws.get("flip_x") is not None and draws["flip_x"][b]:
            cam[..., 0] = -cam[..., 0]                       # dataset.py:303-304
        if draws.get("flip_z") is not None and draws["flip_z"][b]:
            cam[..., 2] = -cam[..., 2]                       # :305-306
        if draws.get("roty_angle") is not None:
            c, s = np.cos(draws["roty_angle"][b]), np.sin(draws["roty_angle"][b])
            R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
            cam[:, :3] = (R @ cam[:, :3].T).T                # :307
        if draws.get("scale") is not None:
            cam[:, :3] = cam[:, :3] * draws["scale"][b]      # :308
        out.append(cam.astype(np.float32))
    return np.stack(out)


def votenet_losses(seeds_xyz, votes_xyz, proposals_xyz, proposals_output, bboxes_xyz, bboxes_lwh, bboxes_roty, semantic_labels,
                   heading_labels, heading_residuals, size_labels, size_residuals, positive_thres=0.3, negative_thres=0.6,
                   NH=12, NS=10, NC=10):
    """/root/reference/model.py:62-84,141-231 restated in numpy (float32 geometry like the TF graph, float64 sums).
    Returns the 14 numbers votenet_b200.losses.NAMES names."""
    f = np.float32
    seeds_xyz, votes_xyz, proposals_xyz, po = (np.asarray(a, f) for a in (seeds_xyz, votes_xyz, proposals_xyz, proposals_output))
    bxyz, blwh, broty = np.asarray(bboxes_xyz, f), np.asarray(bboxes_lwh, f), np.asarray(bboxes_roty, f)
    B, N, _ = seeds_xyz.shape
    BB = bxyz.shape[1]

    def huber(x):
        a = np.abs(x)
        return np.where(a <= 1.0, 0.5 * x * x, a - 0.5)

    def ce(logits, labels):
        logits = np.asarray(logits, np.float64)
        mx = logits.max(-1, keepdims=True)
        return (np.log(np.exp(logits - mx).sum(-1)) + mx[..., 0] - np.take_along_axis(logits, labels[..., None], -1)[..., 0])

    def top1(logits, labels):
        t = np.take_along_axis(logits, labels[..., None], -1)
        return ~(logits > t).any(-1)

    d = np.abs(seeds_xyz[:, :, None, :] - bxyz[:, None, :, :])                       # :62
    c, s = np.cos(-broty).astype(f), np.sin(-broty).astype(f)                        # :64-75 (angle -roty)
    rx = c[:, None, :] * d[..., 0] + s[:, None, :] * d[..., 2]
    ry = d[..., 1]
    rz = -s[:, None, :] * d[..., 0] + c[:, None, :] * d[..., 2]
    r = np.stack([rx, ry, rz], -1).astype(f)
    half = (blwh / f(2.0))[:, None, :, :]
    surface = ((r < half).sum(-1) == 3).sum(-1) >= 1                                 # :76-78
    nrm = np.sqrt((r * r).sum(-1, dtype=f))                                          # :80
    assign = nrm.argmin(-1)                                                          # :81
    gt_c = np.take_along_axis(bxyz, assign[..., None].repeat(3, -1), 1)              # :82-85
    vote = (np.abs(votes_xyz - gt_c).sum(-1).astype(np.float64) * surface).mean()    # :86

    diff = proposals_xyz[:, :, None, :] - bxyz[:, None, :, :]
    dist = np.sqrt((diff * diff).sum(-1, dtype=f))                                   # :147
    ba, md = dist.argmin(-1), dist.min(-1)                                           # :148-149
    pos, neg = np.nonzero(md < f(positive_thres)), np.nonzero(md > f(negative_thres))    # :152-154
    pg = (pos[0], ba[pos])                                                           # :155
    obj = po[..., 0:2]
    ones, zeros = np.ones(len(pos[0]), np.int64), np.zeros(len(neg[0]), np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        obj_loss = ce(obj[pos], ones).mean() + ce(obj[neg], zeros).mean()            # :158-163
        obj_acc = np.concatenate([top1(obj[pos], ones), top1(obj[neg], zeros)]).astype(np.float64).mean()   # :164-166
        delta_gt = bxyz[pg] - proposals_xyz[pos]                                     # :169-171
        center = huber(po[..., 2:5][pos].astype(np.float64) - delta_gt).sum(-1).mean()   # :172
        dual = dist.argmin(1)                                                        # :175 (B, BB)
        bi = np.arange(B)[:, None].repeat(BB, 1)
        center_dual = huber(po[..., 2:5][bi, dual].astype(np.float64) - (bxyz - proposals_xyz[bi, dual])).sum(-1).mean()   # :176-179
        center = center + center_dual                                                # :182
        hl = np.asarray(heading_labels)[pg].astype(np.int64)
        hcls = ce(po[..., 5:5 + NH][pos], hl).mean()                                 # :185-187
        hres_p = np.take_along_axis(po[..., 5 + NH:5 + 2 * NH][pos], hl[:, None], -1)[:, 0]
        hres = huber(hres_p.astype(np.float64) - np.asarray(heading_residuals, f)[pg]).mean()   # :189-193
        sl = np.asarray(size_labels)[pg].astype(np.int64)
        scls = ce(po[..., 5 + 2 * NH:5 + 2 * NH + NS][pos], sl).mean()               # :196-198
        sp = po[..., 5 + 2 * NH + NS:5 + 2 * NH + 4 * NS][pos].reshape(-1, NS, 3)
        sres_p = np.take_along_axis(sp, sl[:, None, None].repeat(3, -1), 1)[:, 0, :]
        sres = huber(sres_p.astype(np.float64) - np.asarray(size_residuals, f)[pg]).sum(-1).mean()   # :200-205
        box = center + 0.1 * hcls + hres + 0.1 * scls + sres                         # :207
        sem_l = np.asarray(semantic_labels)[pg].astype(np.int64)
        sem = ce(po[..., -NC:][pos], sem_l).mean()                                   # :210-214
        sem_acc = top1(po[..., -NC:][pos], sem_l).astype(np.float64).mean()          # :216-217
    total = vote + 0.5 * obj_loss + box + 0.1 * sem                                  # :231
    return np.array([total, vote, obj_loss, box, center, hcls, hres, scls, sres, sem, obj_acc, sem_acc, len(pos[0]), len(neg[0])], np.float64)

"""Drop-in for the reference's tf_ops/3d_nms/tf_nms3d.py over CUDA tensors (the reference runs NMS on ONE CPU thread,
tf_nms3d.cpp:308)."""
import torch

from ._lib import check, dptr, lib, stream_ptr


def nms3d_raw(bboxes, scores, objectiveness, iou_threshold):
    """Fixed-shape, sync-free form: -> (keep (B,K) uint8, idx (B*K,2) i32 [first `count` rows valid], count (1,) i32)."""
    if bboxes.dim() != 4 or tuple(bboxes.shape[2:]) != (8, 3):
        raise ValueError("3D NMS expects (batch_size, nbbox, 8, 3) bbox shape.")        # tf_nms3d.cpp:287
    b, k = bboxes.shape[:2]
    if tuple(scores.shape) != (b, k):
        raise ValueError("3D NMS expects (batch_size, nbbox) scores shape.")            # tf_nms3d.cpp:292
    if tuple(objectiveness.shape) != (b, k, 2):
        raise ValueError("3D NMS expects (batch_size, nbbox, 2) objectiveness shape.")  # tf_nms3d.cpp:295
    thr = float(iou_threshold)
    dev = bboxes.device
    keep = torch.empty((b, k), dtype=torch.uint8, device=dev)
    idx = torch.zeros((max(b * k, 1), 2), dtype=torch.int32, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws = torch.empty((lib.vnb_nms3d_workspace_bytes(b, k),), dtype=torch.uint8, device=dev)
    check(lib.vnb_nms3d(b, k, dptr(bboxes, torch.float32, "bboxes"), dptr(scores, torch.float32, "scores"),
                        dptr(objectiveness, torch.float32, "objectiveness"), thr, dptr(keep), dptr(idx), dptr(count),
                        dptr(ws), stream_ptr()))
    nms3d_raw.last_workspace = (ws, lib.vnb_nms3d_near_threshold_offset(b, k))
    return keep, idx, count


def near_threshold_pairs():
    """Number of clipped pairs of the LAST nms3d_raw call whose IoU lay within 1e-5 of the threshold (one host sync): 0
    means no pair was close enough for a last-bit arithmetic difference to change the keep mask."""
    ws, off = nms3d_raw.last_workspace
    return int(ws[off:off + 4].view(torch.int32).item())


def NMS3D(bboxes, scores, objectiveness, iou_threshold):
    """-> (Nnms,2) i32 rows (batch, box) in global descending-score order.   Reference: tf_nms3d.py:11-12.
    The output length is data dependent (tf_nms3d.cpp:267-272), so this wrapper reads the count back (one host sync);
    use nms3d_raw for the sync-free fixed-shape form."""
    keep, idx, count = nms3d_raw(bboxes, scores, objectiveness, iou_threshold)
    return idx[: int(count.item())]

"""Training losses + label assignment on the GPU — forward values of /root/reference/model.py:62-84,141-231 (SURVEY.md
§8(f) rank 2), computed by csrc/losses.cu in two launches from the forward's tensors and the dense ground-truth arrays
the reference's build_graph receives (model.py:34)."""
import torch

from ._lib import check, dptr, lib, stream_ptr

POSITIVE_THRES, NEGATIVE_THRES = 0.3, 0.6   # config.py:4-5
NAMES = ("total_cost", "vote_reg_loss", "obj_cls_loss", "box_loss", "center_loss", "heading_cls_loss", "heading_residual_loss",
         "size_cls_loss", "size_residual_loss", "sem_cls_loss", "obj_accuracy", "sem_accuracy", "n_positive", "n_negative")


def votenet_losses(seeds_xyz, votes_xyz, proposals_xyz, proposals_output, bboxes_xyz, bboxes_lwh, bboxes_roty,
                   semantic_labels, heading_labels, heading_residuals, size_labels, size_residuals,
                   positive_thres=POSITIVE_THRES, negative_thres=NEGATIVE_THRES):
    """All CUDA tensors (argument names as in model.py:34).  -> (14,) float64 device tensor, entries named by NAMES;
    asynchronous on the current stream."""
    b, n_seed, _ = seeds_xyz.shape
    n_prop, n_box = proposals_xyz.shape[1], bboxes_xyz.shape[1]
    if proposals_output.shape != (b, n_prop, 79) or bboxes_lwh.shape != (b, n_box, 3) or size_residuals.shape != (b, n_box, 3):
        raise ValueError("votenet_losses: inconsistent shapes")
    f32, i32 = torch.float32, torch.int32
    out = torch.empty((14,), dtype=torch.float64, device=seeds_xyz.device)
    ws = torch.empty((128,), dtype=torch.uint8, device=seeds_xyz.device)
    check(lib.vnb_votenet_losses(b, n_seed, n_prop, n_box, dptr(seeds_xyz, f32), dptr(votes_xyz, f32), dptr(proposals_xyz, f32),
                                 dptr(proposals_output, f32), dptr(bboxes_xyz, f32), dptr(bboxes_lwh, f32), dptr(bboxes_roty, f32),
                                 dptr(semantic_labels, i32), dptr(heading_labels, i32), dptr(heading_residuals, f32),
                                 dptr(size_labels, i32), dptr(size_residuals, f32), float(positive_thres),
                                 float(negative_thres), dptr(out), dptr(ws), stream_ptr()))
    return out


def losses_dict(out):
    v = out.cpu().tolist()
    return dict(zip(NAMES, v))

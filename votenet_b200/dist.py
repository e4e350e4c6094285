"""Data-parallel sharding of clouds over ranks (one process per GPU) — SURVEY.md §8(e).

Clouds are independent in inference (every reference op indexes by batch, tf_sampling_g.cu:113, tf_grouping_g.cu:4,
tf_interpolate.cpp:61; NMS suppresses within a cloud only, tf_nms3d.cpp:250), so the path shards with NO data-path
collective: rank r owns clouds [r*B, (r+1)*B).  The only exchange is ONE all-gather of the fixed-size per-rank
detection record (engine.DetectionRecord, written in wire layout by the decode/NMS kernels), after which the
reference's global score-ordered (Nnms,2) list is rebuilt over the whole batch by vnb_merge_detections.

The record/merge logic is backend-agnostic (NCCL on GPUs; gloo for the CPU tests of the host logic).
"""
import numpy as np
import torch
import torch.distributed as dist

from .engine import DetectionRecord


def shard_range(rank, world, clouds_per_rank):
    """Cloud ids owned by `rank`."""
    return range(rank * clouds_per_rank, (rank + 1) * clouds_per_rank)


def all_gather_records(rec_buf, world, group=None):
    """rec_buf: (nbytes,) uint8 on the rank's device -> (world, nbytes) uint8, one all-gather (NCCL over NVLink)."""
    out = torch.empty((world, rec_buf.numel()), dtype=torch.uint8, device=rec_buf.device)
    if world == 1:
        out[0].copy_(rec_buf)
        return out
    dist.all_gather_into_tensor(out.view(-1), rec_buf, group=group)
    return out


def merge_gathered(gathered, b, k, gather_outputs=False):
    """Device merge: (world, nbytes) gathered records -> (idx (world*b*k,2) i32 rows (global_batch, box) in descending
    score order, count (1,) i32); with gather_outputs also the reference's output gathers (model.py:135-137):
    bboxes_pred (world*b*k,8,3), class_scores_pred (world*b*k,10), batch_idx (world*b*k) — first `count` rows valid."""
    dg = DetectionGather(gathered.shape[0], b, k, gathered.device, slots=1, gather_outputs=gather_outputs)
    dg.gathered[0] = gathered
    return dg.merge(0)


class DetectionGather:
    """Pre-allocated all-gather + merge for a stream of forwards (bench.py's hot loop): one gather buffer and one merged
    index list per in-flight slot, record offsets resolved once — per step this is ONE collective and ONE kernel launch,
    with no allocation, fill or Python-side view construction on the host path.  The merge is a k-way merge by rank of
    the ranks' already-sorted detection lists (vnb_merge_detections), and can emit the reference's three output gathers
    (bboxes_pred, class_scores_pred, batch_idx; model.py:135-137) in the same launch."""

    def __init__(self, world, b, k, device, slots=1, group=None, gather_outputs=False):
        from ._lib import lib

        self.world, self.b, self.k, self.group = world, b, k, group
        lay = DetectionRecord(b, k, device="cpu")
        self.nbytes = lay.nbytes
        self.offs = tuple(lay.offsets[f][0] for f in ("nms_idx", "nms_key", "nms_count", "bboxes", "class_scores"))
        n = world * b * k
        E = lambda shape, dt: torch.empty(shape, dtype=dt, device=device)  # noqa: E731
        self.gathered = [E((world, self.nbytes), torch.uint8) for _ in range(slots)]
        self.idx = [E((n, 2), torch.int32) for _ in range(slots)]
        self.cnt = [torch.zeros((1,), dtype=torch.int32, device=device) for _ in range(slots)]
        self.gather_outputs = gather_outputs
        if gather_outputs:
            self.bboxes_pred = [E((n, 8, 3), torch.float32) for _ in range(slots)]
            self.class_scores_pred = [E((n, 10), torch.float32) for _ in range(slots)]
            self.batch_idx = [E((n,), torch.int32) for _ in range(slots)]
        self._merge = lib.vnb_merge_detections

    def merge(self, slot=0):
        """Merge self.gathered[slot] on the current stream -> (idx, count[, bboxes_pred, class_scores_pred, batch_idx])."""
        from ._lib import check, dptr, stream_ptr

        g = self.gathered[slot]
        go = self.gather_outputs
        check(self._merge(self.world, self.b, self.k, dptr(g), self.nbytes, *self.offs, dptr(self.idx[slot]),
                          dptr(self.cnt[slot]), dptr(self.bboxes_pred[slot]) if go else None,
                          dptr(self.class_scores_pred[slot]) if go else None, dptr(self.batch_idx[slot]) if go else None,
                          stream_ptr()))
        if go:
            return self.idx[slot], self.cnt[slot], self.bboxes_pred[slot], self.class_scores_pred[slot], self.batch_idx[slot]
        return self.idx[slot], self.cnt[slot]

    def __call__(self, rec_buf, slot=0):
        """rec_buf (nbytes,) uint8 of this rank -> merged outputs of the whole batch; asynchronous on the current stream."""
        g = self.gathered[slot]
        if self.world == 1:
            g[0].copy_(rec_buf, non_blocking=True)
        else:
            dist.all_gather_into_tensor(g.view(-1), rec_buf, group=self.group)
        return self.merge(slot)


def merge_gathered_host(gathered, b, k):
    """Host (numpy) statement of the same merge — used by the gloo tests and as the checker of the device merge."""
    g = gathered.cpu()
    world = g.shape[0]
    rows = []
    for r in range(world):
        rec = DetectionRecord(b, k, buf=g[r].contiguous())
        keep = rec.keep.numpy().astype(bool)
        sc = rec.scores.numpy()
        for bi, ki in zip(*np.nonzero(keep)):
            rows.append((-float(sc[bi, ki]), r * b + int(bi), int(ki)))
    rows.sort()
    return np.array([[gb, ki] for _, gb, ki in rows], dtype=np.int32).reshape(-1, 2)

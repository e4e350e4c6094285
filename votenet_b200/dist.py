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


def merge_gathered(gathered, b, k):
    """Device merge: (world, nbytes) gathered records -> (idx (world*b*k,2) i32 rows (global_batch, box) in descending
    score order, count (1,) i32)."""
    from ._lib import check, dptr, lib, stream_ptr

    world = gathered.shape[0]
    lay = DetectionRecord(b, k, buf=gathered[0])
    idx = torch.zeros((world * b * k, 2), dtype=torch.int32, device=gathered.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=gathered.device)
    check(lib.vnb_merge_detections(world, b, k, dptr(gathered), gathered.shape[1], lay.offsets["scores"][0],
                                   lay.offsets["keep"][0], dptr(idx), dptr(cnt), stream_ptr()))
    return idx, cnt


class DetectionGather:
    """Pre-allocated all-gather + merge for a stream of forwards (bench.py's hot loop): one gather buffer and one merged
    index list per in-flight slot, record offsets resolved once — per step this is ONE collective and ONE kernel launch,
    with no allocation, fill or Python-side view construction on the host path."""

    def __init__(self, world, b, k, device, slots=1, group=None):
        from ._lib import lib

        self.world, self.b, self.k, self.group = world, b, k, group
        lay = DetectionRecord(b, k, device="cpu")
        self.nbytes = lay.nbytes
        self.off_scores, self.off_keep = lay.offsets["scores"][0], lay.offsets["keep"][0]
        self.gathered = [torch.empty((world, self.nbytes), dtype=torch.uint8, device=device) for _ in range(slots)]
        self.idx = [torch.empty((world * b * k, 2), dtype=torch.int32, device=device) for _ in range(slots)]
        self.cnt = [torch.zeros((1,), dtype=torch.int32, device=device) for _ in range(slots)]
        self._merge = lib.vnb_merge_detections

    def __call__(self, rec_buf, slot=0):
        """rec_buf (nbytes,) uint8 of this rank -> (idx, count) of the whole batch; asynchronous on the current stream."""
        from ._lib import check, dptr, stream_ptr

        g = self.gathered[slot]
        if self.world == 1:
            g[0].copy_(rec_buf, non_blocking=True)
        else:
            dist.all_gather_into_tensor(g.view(-1), rec_buf, group=self.group)
        check(self._merge(self.world, self.b, self.k, dptr(g), self.nbytes, self.off_scores, self.off_keep,
                          dptr(self.idx[slot]), dptr(self.cnt[slot]), stream_ptr()))
        return self.idx[slot], self.cnt[slot]


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

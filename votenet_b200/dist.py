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


class PeerGather(DetectionGather):
    """DetectionGather whose all-gather is ONE-SIDED over NVLink peer memory (csrc/peer.cu) instead of an NCCL
    collective: after a forward, one kernel pushes the rank's record into every peer's inbox and raises a sequence flag;
    a one-warp kernel waits for the slot's `world` flags in front of the merge.  No rendezvous kernel holds SMs while
    it waits for the slowest rank, and collectives of different in-flight forwards do not serialise on one communicator.

    The inboxes are dedicated CUDA allocations shared through CUDA IPC (vnb_peer_alloc / vnb_peer_open; the 64-byte
    handles are exchanged once with all_gather_object) and opened in the CONSUMER device's context, which is what maps
    them peer-to-peer over NVLink.  Every slot has TWO inbox buffers used alternately: rank r pushes use u+2 of a slot only after
    its own merge of use u+1, which saw every peer's push u+1, which each peer issued after ITS merge of use u (stream
    order on the slot's stream) — so nobody overwrites a buffer a peer has not merged yet, with no acknowledgement
    traffic.  `slot` must be bound to one stream, as in bench.py."""

    def __init__(self, world, rank, b, k, device, slots=1, group=None, gather_outputs=False):
        super().__init__(world, b, k, device, slots=slots, group=group, gather_outputs=gather_outputs)
        import ctypes as C

        from ._lib import check, lib

        self.rank, self.device = rank, torch.device(device)
        nb = self.nbytes
        self.depth = 2
        # one dedicated allocation per rank: [slots][depth][world][nbytes] records, then [slots][depth][world] int32 flags
        n_inbox = slots * self.depth * world * nb
        n_flags = slots * self.depth * world * 4
        total = n_inbox + (n_flags + 255) // 256 * 256
        with torch.cuda.device(self.device):
            own = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            check(lib.vnb_peer_alloc(total, C.byref(own), handle))
            self._own = own.value
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle), group=group)
            bases = []
            self._opened = []
            for r in range(world):
                if r == rank:
                    bases.append(self._own)
                else:
                    p_ = C.c_void_p()
                    check(lib.vnb_peer_open((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p_)))
                    self._opened.append(p_.value)
                    bases.append(p_.value)
        inbox_ptr = bases
        flags_ptr = [p_ + n_inbox for p_ in bases]

        class _Raw:   # zero-copy torch views of this rank's own allocation
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}

        self.inbox = torch.as_tensor(_Raw(self._own, (slots, self.depth, world, nb), "|u1"), device=self.device)
        self.flags = torch.as_tensor(_Raw(self._own + n_inbox, (slots, self.depth, world), "<i4"), device=self.device)
        assert self.inbox.data_ptr() == self._own and self.flags.data_ptr() == self._own + n_inbox
        # per (slot, depth): host arrays of `world` device pointers
        self._ptrs = {}
        for s in range(slots):
            for d in range(self.depth):
                off_i = (s * self.depth + d) * world * nb
                off_f = (s * self.depth + d) * world * 4
                self._ptrs[(s, d)] = ((C.c_void_p * world)(*[p_ + off_i for p_ in inbox_ptr]),
                                      (C.c_void_p * world)(*[p_ + off_f for p_ in flags_ptr]))
        self._uses = [0] * slots
        from ._lib import lib
        self._push, self._wait = lib.vnb_peer_push_record, lib.vnb_peer_wait

    def __call__(self, rec_buf, slot=0):
        from ._lib import check, dptr, stream_ptr

        u = self._uses[slot]
        self._uses[slot] = u + 1
        d, seq = u % self.depth, u // self.depth + 1
        ip, fp = self._ptrs[(slot, d)]
        st = stream_ptr()
        check(self._push(self.world, self.rank, dptr(rec_buf), self.nbytes, ip, fp, seq, st))
        check(self._wait(self.world, dptr(self.flags[slot, d]), seq, st))
        self.gathered[slot] = self.inbox[slot, d]
        return self.merge(slot)


def make_gather(world, rank, b, k, device, slots=1, group=None, gather_outputs=False, transport="peer"):
    """The detection exchange of the multi-GPU path: `peer` = one-sided pushes over NVLink peer memory (PeerGather),
    `nccl` = one all_gather_into_tensor per forward.  Returns (gather, transport actually in use); `peer` falls back to
    NCCL — loudly, the caller reports it — only when CUDA IPC between the ranks cannot be set up."""
    if world == 1 or transport == "nccl":
        return DetectionGather(world, b, k, device, slots=slots, group=group, gather_outputs=gather_outputs), "nccl" if world > 1 else "local"
    ok, err = 1, ""
    g = None
    try:
        g = PeerGather(world, rank, b, k, device, slots=slots, group=group, gather_outputs=gather_outputs)
    except Exception as e:  # noqa: BLE001 — any IPC / peer-access failure
        ok, err = 0, f"{type(e).__name__}: {e}"
    t = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    if int(t.item()) == 1:
        return g, "peer"
    import sys
    print(f"[votenet_b200.dist] rank {rank}: peer-memory transport unavailable ({err or 'a peer failed'}); using NCCL all-gather",
          file=sys.stderr, flush=True)
    return DetectionGather(world, b, k, device, slots=slots, group=group, gather_outputs=gather_outputs), "nccl (peer transport unavailable)"


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

"""CPU tests of the N>1 host logic with the gloo backend, world_size 2: sharding, one all-gather of the fixed-size
detection records, and the global score-ordered merge."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from votenet_b200.dist import all_gather_records, merge_gathered_host, shard_range
    from votenet_b200.engine import DetectionRecord

    b, k = 3, 16
    ids = list(shard_range(rank, world, b))
    rec = DetectionRecord(b, k, device="cpu")
    rng = np.random.default_rng(rank)
    rec.scores.copy_(torch.as_tensor(rng.standard_normal((b, k)).astype(np.float32)))
    rec.keep.copy_(torch.as_tensor((rng.random((b, k)) < 0.3).astype(np.uint8)))
    rec.bboxes.fill_(float(rank))
    g = all_gather_records(rec.buf, world)
    merged = merge_gathered_host(g, b, k)
    q.put((rank, ids, g.clone(), merged))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_allgather_and_merge():
    from votenet_b200.engine import DetectionRecord

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4, 5]           # contiguous shards, no overlap
    assert torch.equal(res[0][2], res[1][2])                            # every rank holds the same gathered records
    assert np.array_equal(res[0][3], res[1][3])                         # ... and the same merged list
    b, k = 3, 16
    g, merged = res[0][2], res[0][3]
    recs = [DetectionRecord(b, k, buf=g[r].contiguous()) for r in range(world)]
    assert all(float(recs[r].bboxes[0, 0, 0, 0]) == float(r) for r in range(world))  # rank r's payload at row r
    nkeep = sum(int(r.keep.sum()) for r in recs)
    assert merged.shape == (nkeep, 2)
    sc = [float(recs[gb // b].scores[gb % b, ki]) for gb, ki in merged]
    assert all(sc[i] >= sc[i + 1] for i in range(len(sc) - 1))         # global descending score
    assert all(int(recs[gb // b].keep[gb % b, ki]) == 1 for gb, ki in merged)
    assert merged[:, 0].max() >= b                                      # rows of rank 1 carry global batch ids

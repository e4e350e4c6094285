"""torchrun -N ranks: the one-sided peer-memory exchange (votenet_b200.dist.PeerGather) against the NCCL all-gather
(DetectionGather): same gathered records, same merged list as the host merge; then a timing of both under a stream of
forward-less steps.    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/gpu_peer_test.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from votenet_b200 import synth
from votenet_b200.dist import DetectionGather, PeerGather, make_gather, merge_gathered_host
from votenet_b200.engine import DetectionRecord
from votenet_b200.model import decode_nms3d

b, k, SLOTS, STEPS = 8, 256, 4, 24
ms = torch.as_tensor(np.asarray(synth.CLASS_MEAN_SIZE, np.float32), device=dev)


def make_record(seed):
    rng = np.random.default_rng(seed)
    pxyz = torch.as_tensor(rng.uniform(-2, 2, (b, k, 3)).astype(np.float32), device=dev)
    pout = torch.as_tensor(rng.standard_normal((b, k, 79)).astype(np.float32), device=dev)
    o = decode_nms3d(pxyz, pout, ms, 0.25)
    rec = DetectionRecord(b, k, device=dev)
    for f, key in (("bboxes", "dec_bboxes"), ("scores", "dec_scores"), ("class_scores", "dec_class_scores"), ("objectness", "dec_objectness"),
                   ("keep", "nms_keep"), ("nms_idx", "nms_idx"), ("nms_key", "nms_key"), ("nms_count", "nms_count")):
        getattr(rec, f).copy_(o[key])
    return rec


recs = [make_record(1000 * rank + i) for i in range(STEPS)]
torch.cuda.synchronize()
nccl = DetectionGather(world, b, k, dev, slots=SLOTS)
peer, transport = make_gather(world, rank, b, k, dev, slots=SLOTS, transport="peer")
if rank == 0:
    print("transport:", transport, flush=True)
assert isinstance(peer, PeerGather), transport
streams = [torch.cuda.Stream(device=dev) for _ in range(SLOTS)]
# ---- equality: peer == nccl == host merge, over several uses of every slot, no sync in between
res_p, res_n = [], []
for i in range(STEPS):
    st = streams[i % SLOTS]
    with torch.cuda.stream(st):
        idx, cnt = peer(recs[i].buf, slot=i % SLOTS)
        res_p.append((idx.clone(), cnt.clone(), peer.gathered[i % SLOTS].clone()))
torch.cuda.synchronize()
for i in range(STEPS):
    st = streams[i % SLOTS]
    with torch.cuda.stream(st):
        idx, cnt = nccl(recs[i].buf, slot=i % SLOTS)
        res_n.append((idx.clone(), cnt.clone(), nccl.gathered[i % SLOTS].clone()))
torch.cuda.synchronize()
for i in range(STEPS):
    assert torch.equal(res_p[i][2], res_n[i][2]), f"step {i}: gathered records differ"
    n = int(res_n[i][1].item())
    assert int(res_p[i][1].item()) == n and torch.equal(res_p[i][0][:n], res_n[i][0][:n]), f"step {i}: merged lists differ"
    if i < 3:
        assert np.array_equal(res_p[i][0][:n].cpu().numpy(), merge_gathered_host(res_p[i][2], b, k))
dist.barrier()
if rank == 0:
    print(f"peer == nccl == host merge over {STEPS} steps / {SLOTS} slots: ok", flush=True)
# ---- timing (exchange + merge only)
for name, g in (("nccl", nccl), ("peer", peer)):
    for rep in range(2):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.time()
        for i in range(200):
            st = streams[i % SLOTS]
            with torch.cuda.stream(st):
                g(recs[i % STEPS].buf, slot=i % SLOTS)
        torch.cuda.synchronize()
        dt = time.time() - t0
    if rank == 0:
        print(f"{name}: {1e6 * dt / 200:.1f} us per exchange+merge (host-enqueue bound or device bound, {SLOTS} streams)", flush=True)
dist.barrier()
dist.destroy_process_group()

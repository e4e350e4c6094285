"""Step-by-step diagnosis of the peer-memory transport (torchrun, 2 ranks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from votenet_b200._lib import check, dptr, lib, stream_ptr
from votenet_b200.dist import PeerGather
from votenet_b200.engine import DetectionRecord
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
b, k = 8, 256
g = PeerGather(world, rank, b, k, dev, slots=2)
say("PeerGather built; own", hex(g._own), "opened", [hex(p) for p in g._opened], "can_access", [torch.cuda.can_device_access_peer(local, r) for r in range(world) if r != local])
rec = DetectionRecord(b, k, device=dev); rec.buf.fill_(rank + 1); torch.cuda.synchronize()
dist.barrier()
ip, fp = g._ptrs[(0, 0)]
say("ptrs", [hex(int(p or 0)) for p in ip], [hex(int(p or 0)) for p in fp])
check(lib.vnb_peer_push_record(world, rank, dptr(rec.buf), g.nbytes, ip, fp, 1, stream_ptr())); torch.cuda.synchronize(); say("push ok")
check(lib.vnb_peer_wait(world, dptr(g.flags[0, 0]), 1, stream_ptr())); torch.cuda.synchronize(); say("wait ok", g.flags[0, 0].tolist(), "inbox rows", [int(g.inbox[0, 0, r, 0]) for r in range(world)])
dist.barrier()
dist.destroy_process_group()

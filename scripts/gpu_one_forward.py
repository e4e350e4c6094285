"""One or a few eager forwards at the bench shape (for compute-sanitizer runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from votenet_b200 import synth
from votenet_b200.config import VoteNetConfig
from votenet_b200.engine import Engine
from votenet_b200.weights import make_synthetic_weights

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
inflight = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
cfg = VoteNetConfig()
eng = Engine(cfg, make_synthetic_weights(cfg, 0), 8, device=dev, use_graph=False, slots=inflight)
xyz = torch.as_tensor(synth.synthetic_batch(0, 8, cfg.num_points), device=dev)
feat = torch.as_tensor(synth.height_feature(xyz.cpu().numpy()), device=dev)
streams = [torch.cuda.Stream() for _ in range(inflight)]
from votenet_b200._lib import check, lib
trap = torch.zeros(8, dtype=torch.int32).pin_memory()
check(lib.vnb_debug_trap_buffer(trap.data_ptr()))
try:
    for i in range(steps):
        rec = eng.infer_device(xyz, feat, stream=streams[i % inflight])
    torch.cuda.synchronize()
except Exception:
    print("trap record {line, blockDim, blockIdx, threadIdx, gridDim}:", trap[:5].tolist(), flush=True)
    raise
print("ok: kept", int(rec.nms_count.item()))

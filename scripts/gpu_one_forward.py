"""One or a few eager forwards (for compute-sanitizer / ncu runs):  gpu_one_forward.py [steps] [inflight] [bench|small]
`bench` = the BASELINE shape (8 clouds x 20 000 points); `small` = 2 clouds x 6144 points through the same kernels
(every kernel of the forward runs, the sanitizers' 10-100x slow-down stays within minutes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from votenet_b200 import synth
from votenet_b200.config import SAParams, VoteNetConfig
from votenet_b200.engine import Engine
from votenet_b200.weights import make_synthetic_weights

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
inflight = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shape = sys.argv[3] if len(sys.argv) > 3 else "bench"
dev = torch.device("cuda:0")
if shape == "small":
    cfg = VoteNetConfig(num_points=6144,   # > 4096: the ball query's summary-bitmap mode, like the bench shape
                        sa=(SAParams(512, 0.3, 64, (64, 64, 128)), SAParams(256, 0.5, 64, (128, 128, 256)),
                            SAParams(128, 0.9, 64, (128, 128, 256)), SAParams(64, 1.4, 64, (128, 128, 256))),
                        proposal=SAParams(64, 0.4, 64, (128, 128, 128), (128, 128, 79)))
    B = 2
else:
    cfg, B = VoteNetConfig(), 8
eng = Engine(cfg, make_synthetic_weights(cfg, 0), B, device=dev, use_graph=False, slots=inflight)
xyz = torch.as_tensor(synth.synthetic_batch(0, B, cfg.num_points), device=dev)
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
print("ok: kept", int(rec.nms_count.item()), "launches/forward", eng.launches_per_forward)

"""Gantt-style timeline of overlapped steps (eager mode, CUDA events at stage boundaries)."""
import sys, os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", os.environ.get("MAXCONN", "32"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.config import VoteNetConfig
from votenet_b200.engine import Engine
from votenet_b200.weights import make_synthetic_weights

inflight = int(sys.argv[1]) if len(sys.argv) > 1 else 4
if len(sys.argv) > 3:
    check(lib.vnb_set_tuning(b"fps_threads", int(sys.argv[2]))); check(lib.vnb_set_tuning(b"fps_cluster", int(sys.argv[3])))
dev = torch.device("cuda:0")
cfg = VoteNetConfig()
eng = Engine(cfg, make_synthetic_weights(cfg, 0), 8, device=dev, use_graph=False, slots=inflight)
xyz = torch.as_tensor(synth.synthetic_batch(0, 8, cfg.num_points), device=dev)
feat = torch.as_tensor(synth.height_feature(xyz.cpu().numpy()), device=dev)
streams = [torch.cuda.Stream() for _ in range(inflight)]
for i in range(2 * inflight):
    eng.infer_device(xyz, feat, stream=streams[i % inflight])
torch.cuda.synchronize()
eng.timeline = []
t0 = torch.cuda.Event(enable_timing=True); t0.record(torch.cuda.current_stream())
for st in streams:
    st.wait_event(t0)
K = 3 * inflight
import time
w0 = time.time()
for i in range(K):
    eng.infer_device(xyz, feat, stream=streams[i % inflight])
cpu_ms = 1e3 * (time.time() - w0)
torch.cuda.synchronize()
rows = {}
for step, name, e in eng.timeline:
    rows.setdefault(step, {})[name] = t0.elapsed_time(e)
names = ["start", "fps1", "bq1", "sa1_begin", "sa1", "fps2", "sa2", "sa3", "sa4", "fps_prop", "fp", "vote", "proposal", "nms"]
print(f"inflight={inflight}  CPU enqueue time for {K} steps: {cpu_ms:.1f} ms ({cpu_ms / K:.3f} ms/step)")
print("step " + " ".join(f"{n:>9s}" for n in names))
for step in sorted(rows):
    print(f"{step:4d} " + " ".join(f"{rows[step].get(n, float('nan')):9.3f}" for n in names))
ends = [rows[s]["nms"] for s in sorted(rows)]
print("steady-state ms/step:", (ends[-1] - ends[inflight]) / (len(ends) - 1 - inflight))

"""Stage timeline of CTA 0 of the fused SA kernel with hoisted layer 1 (sa2..sa4, proposal): S2_STAMP roles in csrc/sa_ws2.cu.
usage: gpu_trace_sa2.py [level 1..3 (sa2..sa4)]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.config import VoteNetConfig
from votenet_b200.engine import Engine
from votenet_b200.weights import make_synthetic_weights

li = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for kv in os.environ.get("VNB_TUNE", "").split(","):   # e.g. VNB_TUNE=sa_wait_ns=100,sa_variant=3
    if kv:
        check(lib.vnb_set_tuning(kv.split("=")[0].encode(), int(kv.split("=")[1])))
dev = torch.device("cuda:0")
cfg = VoteNetConfig()
B = 8
eng = Engine(cfg, make_synthetic_weights(cfg, 0), B, device=dev, use_graph=False, slots=1)
xyz = torch.as_tensor(synth.synthetic_batch(0, B, cfg.num_points), device=dev)
feat = torch.as_tensor(synth.height_feature(xyz.cpu().numpy()), device=dev)
for _ in range(2):
    eng.infer_device(xyz, feat)
torch.cuda.synchronize()
s = eng.slots[0]
st = torch.cuda.current_stream()
l = s.lv[li]
src_xyz, src_feat, c = s.lv[li - 1].xyz, s.lv[li - 1].feat, cfg.sa[li - 1].mlp[-1]


def run():
    eng._sa(li, src_xyz, src_feat, l.n, c, l.xyz, l.idx, l.m, l.q, l.feat, st, s.sa_ws, l.cnt)


run(); torch.cuda.synchronize()
tr = torch.zeros(12 * 64 * 2, dtype=torch.int64, device=dev)
check(lib.vnb_debug_sa_trace(tr.data_ptr()))
run(); torch.cuda.synchronize()
check(lib.vnb_debug_sa_trace(None))
t = tr.cpu().numpy().reshape(12, 64, 2)
t0 = t[t > 0].min()
cols = [("P", t[0]), ("M2", t[1]), ("E2", t[2]), ("M3", t[3]), ("E3", t[4])]
ntile = int((t[0][:, 0] > 0).sum())
print(f"sa{li + 1}: CTA 0 ran {ntile} tiles (start-end cycles relative to the first stamp)")
print("tile | " + " | ".join(f"{n:>11s}" for n, _ in cols))
for k in list(range(0, min(6, ntile))) + list(range(max(6, ntile - 8), ntile)):
    print(f"{k:4d} | " + " | ".join(f"{a[k,0]-t0:5d}-{a[k,1]-t0:5d}" if a[k, 0] > 0 else " " * 11 for _, a in cols))
lo, hi = min(4, ntile - 2), max(ntile - 2, 5)
for n, a in cols:
    v = a[lo:hi]
    print(f"{n:4s} period {np.diff(v[:,0]).mean():7.1f}  busy {np.mean(v[:,1]-v[:,0]):7.1f}")

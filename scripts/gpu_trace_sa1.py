"""Stage timeline of CTA 0 of the fused SA1 kernel (debug trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.config import VoteNetConfig
from votenet_b200.tf_grouping import query_ball_point
from votenet_b200.tf_sampling import farthest_point_sample, gather_point
from votenet_b200.utils import WeightStore, sa_group_mlp_max
from votenet_b200.weights import make_synthetic_weights
for kv in os.environ.get("VNB_TUNE", "").split(","):   # e.g. VNB_TUNE=sa_wait_ns=100,sa_variant=3
    if kv:
        check(lib.vnb_set_tuning(kv.split("=")[0].encode(), int(kv.split("=")[1])))
dev = torch.device("cuda:0")
B, N = 8, 20000
cfg = VoteNetConfig(); w = make_synthetic_weights(cfg, 0); store = WeightStore(w, device=dev, precision=1)
xyz = torch.as_tensor(synth.synthetic_batch(0, B, N), device=dev)
feat = torch.as_tensor(synth.height_feature(xyz.cpu().numpy()), device=dev)
sa = cfg.sa[0]
f = farthest_point_sample(sa.npoint, xyz); nx = gather_point(xyz, f); idx, _ = query_ball_point(sa.radius, 64, xyz, nx)
L = [store.layer(f"sa1/conv{i}") for i in range(3)]
for _ in range(2): sa_group_mlp_max(xyz, feat, nx, idx, L, 1, store, "sa1")
tr = torch.zeros(12 * 64 * 2, dtype=torch.int64, device=dev)
check(lib.vnb_debug_sa_trace(tr.data_ptr()))
sa_group_mlp_max(xyz, feat, nx, idx, L, 1, store, "sa1"); torch.cuda.synchronize()
check(lib.vnb_debug_sa_trace(None))
t = tr.cpu().numpy().reshape(12, 64, 2)
t0 = t[t > 0].min()
# merge even/odd roles: E1 = roles 4|8, E2 = 5|9, E3 = 6|7
E1 = np.where(t[4] > 0, t[4], t[8]); E2 = np.where(t[5] > 0, t[5], t[9]); E3 = np.where(t[6] > 0, t[6], t[7])
cols = [("P", t[0]), ("M1", t[1]), ("E1", E1), ("M2", t[2]), ("E2", E2), ("M3", t[3]), ("E3", E3)]
print("tile | " + " | ".join(f"{n:>11s}" for n, _ in cols) + "   (start-end cycles relative to first stamp)")
for k in list(range(0, 8)) + list(range(36, 52)):
    print(f"{k:4d} | " + " | ".join(f"{a[k,0]-t0:5d}-{a[k,1]-t0:5d}" if a[k,0] > 0 else " " * 11 for _, a in cols))
for n, a in cols:
    v = a[16:50]
    print(f"{n:4s} period {np.diff(v[:,0]).mean():7.1f}  busy {np.mean(v[:,1]-v[:,0]):7.1f}")

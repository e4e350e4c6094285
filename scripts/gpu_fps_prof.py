import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, numpy as np, torch
from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.tf_sampling import farthest_point_sample
dev = torch.device("cuda:0")
xyz = torch.as_tensor(synth.synthetic_batch(0, 8, 20000), device=dev)
buf = torch.zeros(128, dtype=torch.int64, device=dev)
ref = farthest_point_sample(2048, xyz)
check(lib.vnb_debug_fps_profile(C.c_void_p(buf.data_ptr())))
out = farthest_point_sample(2048, xyz)
torch.cuda.synchronize()
check(lib.vnb_debug_fps_profile(C.c_void_p(0)))
assert torch.equal(out, ref)
t = buf.cpu().numpy().reshape(16, 8)[:8, :7] / 2047.0
names = ["scan", "warp-argmax", "bar.sync", "cta-champ", "push", "wait-peers", "final"]
print("cycles per round, warp 0 of each CTA (rows = cluster rank)")
print("      " + " ".join(f"{n:>11s}" for n in names) + "       total")
for r in range(8):
    print(f"rank{r} " + " ".join(f"{v:11.1f}" for v in t[r]) + f"  {t[r].sum():10.1f}")

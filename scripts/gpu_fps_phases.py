"""Per-phase cycle counts of the bucket-pruned FPS kernel (dependency-anchored clocks, PROF template variant) and its
plain timing.  usage: gpu_fps_phases.py [room|uniform]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.tf_sampling import farthest_point_sample, farthest_point_sample_ties

dev = torch.device("cuda:0")
B, N, M = 8, 20000, 2048
kind = sys.argv[1] if len(sys.argv) > 1 else "room"
xyz = torch.as_tensor(synth.synthetic_batch(0, B, N), device=dev) if kind == "room" else torch.rand(B, N, 3, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, iters=8, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


ms = timeit(lambda: farthest_point_sample(M, xyz))
ms_t = timeit(lambda: farthest_point_sample_ties(M, xyz, 1024))
ms_1k = timeit(lambda: farthest_point_sample(1024, xyz))
print(f"== {kind}: fps 20000->2048 B=8: {ms:.3f} ms ({1e6*ms/2047*1.965/1e3:.0f} cycles/round incl. setup), with tie tracking {ms_t:.3f} ms; 20000->1024: {ms_1k:.3f} ms")
prof = torch.zeros(16 * 10, dtype=torch.int64, device=dev)
check(lib.vnb_debug_fps_profile(prof.data_ptr()))
farthest_point_sample(M, xyz); torch.cuda.synchronize()
check(lib.vnb_debug_fps_profile(None))
p = prof.cpu().numpy().reshape(16, 10).astype(np.float64)
R = M - 1
print("warp | pick->mask  rescans  barrier  bar->key  key->pick  ties+store | rescans/round active-frac | setup  total/round")
for w in range(16):
    r = p[w]
    print(f"  {w:2d} | {r[0]/R:8.1f} {r[1]/R:8.1f} {r[2]/R:8.1f} {r[3]/R:8.1f} {r[6]/R:8.1f} {r[7]/R:8.1f} | {r[4]/R:5.2f} {r[5]/R:5.2f} | {int(r[8]):7d} {(r[9]-r[8])/R:8.1f}")

"""FPS timing on the GPU box: register/cluster kernel vs the bucket-pruned single-CTA kernel (CUDA events, warm-up,
L2 flush between iterations), plus the pruned kernel's per-warp phase cycles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.tf_sampling import farthest_point_sample, farthest_point_sample_nested, gather_point

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, iters=6, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def tune(k, v):
    check(lib.vnb_set_tuning(k.encode(), v))


B, N = 8, 20000
xyz = torch.as_tensor(synth.synthetic_batch(0, B, N), device=dev)
print("== FPS 20000->2048, B=8")
for (thr, cl) in ((256, 8), (256, 4), (512, 2)):
    tune("fps_variant", 0); tune("fps_threads", thr); tune("fps_cluster", cl)
    ms = timeit(lambda: farthest_point_sample(2048, xyz))
    print(f"  cluster kernel threads={thr} cluster={cl} ({8*cl} SMs): {ms:.3f} ms  {1e3*ms/2047:.3f} us/round  SM-ms={8*cl*ms:.1f}", flush=True)
tune("fps_variant", 1)
ms = timeit(lambda: farthest_point_sample(2048, xyz))
print(f"  pruned kernel (8 SMs): {ms:.3f} ms  {1e3*ms/2047:.3f} us/round  SM-ms={8*ms:.1f}", flush=True)
ms1 = timeit(lambda: farthest_point_sample(2048, xyz[:1]))
print(f"  pruned kernel, 1 cloud: {ms1:.3f} ms", flush=True)
ms = timeit(lambda: farthest_point_sample(1024, xyz))
print(f"  pruned kernel 20000->1024 (BASELINE configs[1]): {ms:.3f} ms", flush=True)
u = torch.rand(8, 20000, 3, device=dev)
ms = timeit(lambda: farthest_point_sample(2048, u))
print(f"  pruned kernel, uniform-volume cloud: {ms:.3f} ms", flush=True)

# phase profile (warp-level cycle counters, cloud 0)
prof = torch.zeros(16 * 8, dtype=torch.int64, device=dev)
check(lib.vnb_debug_fps_profile(prof.data_ptr()))
farthest_point_sample(2048, xyz); torch.cuda.synchronize()
check(lib.vnb_debug_fps_profile(None))
p = prof.cpu().numpy().reshape(16, 8)
print("== pruned FPS phase cycles per round (rows = warps): bound-test, rescans, publish+barrier, final-reduce | rescans/round, active-round frac | setup cycles, total cycles")
for w in range(16):
    r = p[w]
    print(f"  warp{w:2d} {r[0]/2047:7.1f} {r[1]/2047:7.1f} {r[2]/2047:7.1f} {r[3]/2047:7.1f} | {r[4]/2047:5.2f} {r[5]/2047:5.2f} | {r[6]:8d} {r[7]:9d}")
print(f"  total cycles/round (warp 0): {(p[0,7]-p[0,6])/2047:.1f}; setup {p[0,6]} cycles")

print("== small n")
f1 = farthest_point_sample(2048, xyz); x1 = gather_point(xyz, f1)
r1 = torch.rand(8, 2048, 3, device=dev)
for name, src, m in (("fps-ordered 2048->1024", x1, 1024), ("random 2048->1024", r1, 1024)):
    for v in (0, 2):
        tune("fps_variant", v); tune("fps_cluster", 0); tune("fps_threads", 256)
        ms = timeit(lambda: farthest_point_sample(m, src))
        print(f"  {name} variant={v}: {ms:.3f} ms  {1e3*ms/(m-1):.3f} us/round", flush=True)
    tune("fps_variant", 1)
    ms = timeit(lambda: farthest_point_sample_nested(m, src))
    print(f"  {name} nested: {ms:.3f} ms", flush=True)

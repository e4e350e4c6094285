"""Timing ablations of the bucket-pruned FPS kernel (results are WRONG with fps_ablate != 0; timing only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.tf_sampling import farthest_point_sample, farthest_point_sample_ties

dev = torch.device("cuda:0")
B, N, M = 8, 20000, 2048
xyz = torch.as_tensor(synth.synthetic_batch(0, B, N), device=dev)


def timeit(fn, iters=6, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def tune(k, v):
    check(lib.vnb_set_tuning(k.encode(), v))


setup_ms = timeit(lambda: farthest_point_sample(2, xyz))
print(f"setup only (m=2): {setup_ms:.4f} ms = {setup_ms*1.965e6:.0f} cycles")
for disp in (0,):
    tune("fps_dispatch", disp)
    for abl, what in ((0, "full"), (1, "no rescans"), (3, "no rescans, no bound test"), (7, "barrier + reduce, pick not consumed")):
        tune("fps_ablate", abl)
        ms = timeit(lambda: farthest_point_sample(M, xyz))
        mt = timeit(lambda: farthest_point_sample_ties(M, xyz, 1024))
        cyc = (ms - setup_ms) * 1.965e6 / (M - 1)
        cyt = (mt - setup_ms) * 1.965e6 / (M - 1)
        print(f"dispatch {disp} ablate {abl} ({what}): {ms:.3f} ms = {cyc:.0f} cycles/round; ties variant {mt:.3f} ms = {cyt:.0f}")
tune("fps_ablate", 0); tune("fps_dispatch", 0)
for tr in (0, 1, 512, 1024, 2048):
    mt = timeit(lambda: farthest_point_sample_ties(M, xyz, tr))
    print(f"ties variant, tie_rounds={tr}: {mt:.3f} ms = {(mt - setup_ms) * 1.965e6 / (M - 1):.0f} cycles/round")

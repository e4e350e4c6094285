"""SA-kernel timing on the GPU box (CUDA events, L2 flush between iterations): sa_variant 1 vs 2 on the BASELINE shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.config import VoteNetConfig
from votenet_b200.tf_grouping import query_ball_point
from votenet_b200.tf_sampling import farthest_point_sample, gather_point
from votenet_b200.utils import WeightStore, sa_group_mlp_max
from votenet_b200.weights import make_synthetic_weights

dev = torch.device("cuda:0")
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


def tune(k, v):
    check(lib.vnb_set_tuning(k.encode(), v))


B, N = 8, 20000
cfg = VoteNetConfig()
w = make_synthetic_weights(cfg, 0)
store = WeightStore(w, device=dev, precision=1)
xyz = torch.as_tensor(synth.synthetic_batch(0, B, N), device=dev)
feat = torch.as_tensor(synth.height_feature(xyz.cpu().numpy()), device=dev)
src_xyz, src_feat = xyz, feat
for li, sa in enumerate(cfg.sa):
    f = farthest_point_sample(sa.npoint, src_xyz)
    nx = gather_point(src_xyz, f)
    idx, _ = query_ball_point(sa.radius, 64, src_xyz, nx)
    L = [store.layer(f"sa{li + 1}/conv{i}") for i in range(3)]
    cin = 3 + src_feat.shape[-1]
    fl = 0
    for co in sa.mlp:
        fl += cin * co; cin = co
    gf = 2.0 * B * sa.npoint * 64 * fl / 1e9
    outs = {}
    for v in (1, 2):
        tune("sa_variant", v)
        ms = timeit(lambda: sa_group_mlp_max(src_xyz, src_feat, nx, idx, L, 1, store, f"sa{li + 1}"))
        outs[v] = sa_group_mlp_max(src_xyz, src_feat, nx, idx, L, 1, store, f"sa{li + 1}")
        print(f"sa{li + 1} variant={v}: {ms * 1e3:8.1f} us   {gf / ms:7.1f} TFLOP/s (nominal flops, incl. helper kernels)", flush=True)
    d = (outs[1] - outs[2]).abs().max().item() / outs[1].abs().max().item()
    print(f"   v1 vs v2: bit-identical {torch.equal(outs[1], outs[2])}, max rel diff {d:.2e}", flush=True)
    src_xyz, src_feat = nx, outs[2]

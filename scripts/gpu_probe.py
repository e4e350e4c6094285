"""Ad-hoc per-kernel timing on the GPU box (CUDA events, warm-up, L2 flush between iterations)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from votenet_b200 import synth
from votenet_b200.tf_sampling import farthest_point_sample, gather_point
from votenet_b200.tf_grouping import query_ball_point
from votenet_b200.utils import WeightStore, sa_group_mlp_max, linear, Layer
from votenet_b200.config import VoteNetConfig
from votenet_b200.weights import make_synthetic_weights

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


B, N = 8, 20000
xyz = torch.as_tensor(synth.synthetic_batch(0, B, N), device=dev)
feat = torch.as_tensor(synth.height_feature(xyz.cpu().numpy()), device=dev)
cfg = VoteNetConfig()
w = make_synthetic_weights(cfg, 0)
store = WeightStore(w, device=dev, precision=1)

for (n, m) in [(20000, 2048), (20000, 1024), (2048, 1024), (1024, 512), (512, 256), (1024, 256)]:
    x = xyz[:, :n].contiguous()
    med, mn = timeit(lambda: farthest_point_sample(m, x))
    print(f"fps {n}->{m} B={B}: {med:.3f} ms (min {mn:.3f})  {1e3*med/(m-1):.3f} us/round", flush=True)

f1 = farthest_point_sample(2048, xyz); x1 = gather_point(xyz, f1)
med, mn = timeit(lambda: query_ball_point(0.2, 64, xyz, x1)); print(f"ball query sa1 (20000,2048,r=.2): {med:.3f} ms (min {mn:.3f})", flush=True)
idx1, _ = query_ball_point(0.2, 64, xyz, x1)
f2 = farthest_point_sample(1024, x1); x2 = gather_point(x1, f2)
med, mn = timeit(lambda: query_ball_point(0.4, 64, x1, x2)); print(f"ball query sa2 (2048,1024,r=.4): {med:.3f} ms (min {mn:.3f})", flush=True)
idx2, _ = query_ball_point(0.4, 64, x1, x2)

for prec in (1, 0):
    store = WeightStore(w, device=dev, precision=prec)
    L1 = [store.layer(f"sa1/conv{i}") for i in range(3)]
    try:
        med, mn = timeit(lambda: sa_group_mlp_max(xyz, feat, x1, idx1, L1, prec, store, "sa1"), iters=5, warm=2)
        gf = B * 2048 * 64 * 2 * (4 * 64 + 64 * 64 + 64 * 128) / 1e9
        print(f"sa1 group+mlp+max precision={prec}: {med:.3f} ms (min {mn:.3f})  {gf/mn:.1f} TFLOP/s", flush=True)
        p1 = sa_group_mlp_max(xyz, feat, x1, idx1, L1, prec, store, "sa1")
        L2 = [store.layer(f"sa2/conv{i}") for i in range(3)]
        med, mn = timeit(lambda: sa_group_mlp_max(x1, p1, x2, idx2, L2, prec, store, "sa2"), iters=5, warm=2)
        gf = B * 1024 * 64 * 2 * (131 * 128 + 128 * 128 + 128 * 256) / 1e9
        print(f"sa2 group+mlp+max precision={prec} (incl. hoisted pre-GEMM): {med:.3f} ms (min {mn:.3f})  {gf/mn:.1f} TFLOP/s", flush=True)
    except Exception as e:
        print("sa kernel failed:", e, flush=True)

for rows, cin, cout in [(8192, 512, 256), (8192, 259, 256), (16384, 128, 128)]:
    x = torch.randn(rows, cin, device=dev)
    lay = Layer(torch.randn(cin, cout) * 0.05, torch.zeros(cout), dev)
    for prec in (1, 0):
        try:
            med, mn = timeit(lambda: linear(x, lay, True, prec))
            print(f"linear {rows}x{cin}->{cout} precision={prec}: {med:.3f} ms (min {mn:.3f}) {2*rows*cin*cout/mn/1e9:.1f} TFLOP/s", flush=True)
        except Exception as e:
            print("linear failed:", e, flush=True)

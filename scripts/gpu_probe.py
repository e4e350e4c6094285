"""Ad-hoc per-kernel timing on the GPU box (CUDA events, warm-up, L2 flush between iterations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.tf_sampling import farthest_point_sample, farthest_point_sample_nested, gather_point
from votenet_b200.tf_grouping import query_ball_point
from votenet_b200.utils import WeightStore, sa_group_mlp_max, linear, Layer
from votenet_b200.config import VoteNetConfig
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
    return float(np.median(ts)), float(np.min(ts))


def tune(k, v):
    check(lib.vnb_set_tuning(k.encode(), v))


B, N = 8, 20000
xyz = torch.as_tensor(synth.synthetic_batch(0, B, N), device=dev)
feat = torch.as_tensor(synth.height_feature(xyz.cpu().numpy()), device=dev)
cfg = VoteNetConfig()
w = make_synthetic_weights(cfg, 0)

print("== FPS 20000->2048, B=8: threads x cluster size (mode 1)")
for (thr, cl) in ((256, 8), (256, 4), (512, 4), (512, 2), (1024, 2), (512, 8), (256, 16)):
    tune("fps_threads", thr); tune("fps_cluster", cl)
    try:
        med, mn = timeit(lambda: farthest_point_sample(2048, xyz), iters=5)
        print(f"  threads={thr} cluster={cl} ({8*cl} SMs): {med:.3f} ms  {1e3*med/2047:.3f} us/round   SM-ms={8*cl*med:.1f}", flush=True)
    except Exception as e:
        print(f"  threads={thr} cluster={cl}: FAILED {e}", flush=True)
tune("fps_mode", 0); tune("fps_threads", 256); tune("fps_cluster", 8)
med, mn = timeit(lambda: farthest_point_sample(2048, xyz), iters=5)
print(f"  mode 0 (cluster barrier) threads=256 cluster=8: {med:.3f} ms", flush=True)
tune("fps_mode", 1); tune("fps_cluster", 0)
f1 = farthest_point_sample(2048, xyz); x1 = gather_point(xyz, f1)
f2 = farthest_point_sample(1024, x1); x2 = gather_point(x1, f2)
print("== small-n FPS (single CTA) and nested (parallel proof)")
for (src, n, m) in [(x1, 2048, 1024), (x2, 1024, 512), (x2, 1024, 256)]:
    for cl in (0, 4):
        tune("fps_cluster", cl)
        med, mn = timeit(lambda: farthest_point_sample(m, src))
        print(f"  fps {n}->{m} cluster={'auto(1)' if cl == 0 else cl}: {med:.3f} ms  {1e3*med/(m-1):.3f} us/round", flush=True)
    tune("fps_cluster", 0)
    med, mn = timeit(lambda: farthest_point_sample_nested(m, src))
    print(f"  fps_nested {n}->{m}: {med:.3f} ms", flush=True)

print("== ball query")
for v in (0, 1):
    tune("ball_query_variant", v)
    med, mn = timeit(lambda: query_ball_point(0.2, 64, xyz, x1)); print(f"  sa1 (20000,2048,r=.2) variant={v}: {med:.3f} ms", flush=True)
tune("ball_query_variant", 1)
idx1, _ = query_ball_point(0.2, 64, xyz, x1)
idx2, _ = query_ball_point(0.4, 64, x1, x2)
med, mn = timeit(lambda: query_ball_point(0.4, 64, x1, x2)); print(f"  sa2 (2048,1024,r=.4) scan: {med:.3f} ms", flush=True)

print("== fused SA kernels (tensor cores)")
store = WeightStore(w, device=dev, precision=1)
L1 = [store.layer(f"sa1/conv{i}") for i in range(3)]
med, mn = timeit(lambda: sa_group_mlp_max(xyz, feat, x1, idx1, L1, 1, store, "sa1"))
gf = B * 2048 * 64 * 2 * (4 * 64 + 64 * 64 + 64 * 128) / 1e9
print(f"  sa1: {med:.3f} ms  {gf/med:.1f} TFLOP/s", flush=True)
p1 = sa_group_mlp_max(xyz, feat, x1, idx1, L1, 1, store, "sa1")
L2 = [store.layer(f"sa2/conv{i}") for i in range(3)]
gf = B * 1024 * 64 * 2 * (131 * 128 + 128 * 128 + 128 * 256) / 1e9
for v in (0, 1):
    tune("sa_variant", v)
    med, mn = timeit(lambda: sa_group_mlp_max(x1, p1, x2, idx2, L2, 1, store, "sa2"))
    print(f"  sa2 variant={v} (incl. hoisted pre-GEMM): {med:.3f} ms  {gf/med:.1f} TFLOP/s", flush=True)
tune("sa_variant", 1)
lf = store.derived("sa2/conv0:feat", lambda: None)
med, mn = timeit(lambda: linear(p1.reshape(B * 2048, 128), lf, False, 1, out_f16=True))
print(f"  sa2 hoisted pre-GEMM alone (16384x128->128 f16): {med:.3f} ms", flush=True)

print("== linear")
for rows, cin, cout in [(8192, 512, 256), (8192, 259, 256), (8192, 256, 259), (2048, 128, 79)]:
    x = torch.randn(rows, cin, device=dev)
    lay = Layer(torch.randn(cin, cout) * 0.05, torch.zeros(cout), dev)
    med, mn = timeit(lambda: linear(x, lay, True, 1))
    print(f"  linear {rows}x{cin}->{cout}: {med:.3f} ms {2*rows*cin*cout/med/1e9:.1f} TFLOP/s", flush=True)

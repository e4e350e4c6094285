"""Stage timeline of CTA 0 of the fused FP (+ vote) kernel (debug trace, FP_STAMP slots in csrc/fp_chain.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from votenet_b200 import synth
from votenet_b200._lib import check, lib
from votenet_b200.config import VoteNetConfig
from votenet_b200.engine import Engine
from votenet_b200.utils import fp_module_fused
from votenet_b200.weights import make_synthetic_weights

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
f1, f2 = s.fp
st = torch.cuda.current_stream()


def run(which):
    if which == 1:
        fp_module_fused(f1.dist, f1.idx, s.lv[2].feat, s.lv[3].feat, [eng.store.layer(f"fp1/conv_{i}") for i in range(2)], f1.h[-1], stream=st)
    else:
        fp_module_fused(f2.dist, f2.idx, s.lv[1].feat, f1.h[-1].view(B, f1.n, -1), [eng.store.layer(f"fp2/conv_{i}") for i in range(2)],
                        f2.h[-1], vote=(eng.vote_fused, eng.vote_x0, s.lv[1].xyz, s.votes_xyz, s.votes_feat), stream=st)


for which, L in ((1, 2), (2, 5)):
    tr = torch.zeros(12 * 64 * 2, dtype=torch.int64, device=dev)
    run(which); torch.cuda.synchronize()
    check(lib.vnb_debug_sa_trace(tr.data_ptr()))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); run(which); e1.record(); torch.cuda.synchronize()
    check(lib.vnb_debug_sa_trace(None))
    t = tr.cpu().numpy()
    t0 = int(t[0])
    r = lambda i: int(t[i]) - t0 if t[i] > 0 else -1
    print(f"fp{which}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us (warm, eager); cycles relative to CTA 0's first stamp")
    print("  prologue done", r(1))
    print("  loader chunks (loaded, stored):", [(r(8 + 2 * c), r(9 + 2 * c)) for c in range(8)])
    for l in range(L):
        print(f"  layer {l}: first MMA {r(32 + 4 * l)}, last MMA issued {r(33 + 4 * l)}, accumulator seen {r(34 + 4 * l)}, epilogue done {r(35 + 4 * l)}")

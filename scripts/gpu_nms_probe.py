"""Timing of the fused decode + NMS kernel at the bench shape, per cluster size, and the pair statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from votenet_b200 import synth
from votenet_b200._lib import check, dptr, lib, stream_ptr
from votenet_b200.config import VoteNetConfig
from votenet_b200.engine import Engine
from votenet_b200.weights import make_synthetic_weights

dev = torch.device("cuda:0")
cfg = VoteNetConfig()
B = 8
eng = Engine(cfg, make_synthetic_weights(cfg, 0), B, device=dev, use_graph=False, slots=1)
xyz = torch.as_tensor(synth.synthetic_batch(0, B, cfg.num_points), device=dev)
feat = torch.as_tensor(synth.height_feature(xyz.cpu().numpy()), device=dev)
rec = eng.infer_device(xyz, feat)
torch.cuda.synchronize()
s = eng.slots[0]
p = cfg.proposal
obj = rec.objectness.cpu().numpy()
nc = (obj[..., 1] > obj[..., 0]).sum(1)
print("candidates per cloud:", nc.tolist(), "kept:", int(rec.nms_count.item()), "pairs per cloud:", (nc * (nc - 1) // 2).tolist())


def run():
    check(lib.vnb_decode_nms3d(B, p.npoint, dptr(s.p_xyz), dptr(s.p_h[-1]), dptr(eng.mean_size), float(cfg.nms_iou),
                               dptr(rec.bboxes), dptr(rec.scores), dptr(rec.objectness), dptr(rec.class_scores), dptr(rec.keep),
                               dptr(rec.nms_idx), dptr(rec.nms_key), dptr(rec.nms_count), dptr(s.bboxes_pred),
                               dptr(s.class_scores_pred), dptr(s.batch_idx), dptr(s.nms_ws), stream_ptr()))


for cl in (1, 2, 4, 8):
    check(lib.vnb_set_tuning(b"nms_cluster", cl))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(20):
        run()
    b_.record(); torch.cuda.synchronize()
    print(f"cluster {cl}: {a.elapsed_time(b_) / 20 * 1e3:.1f} us per call (back to back)")

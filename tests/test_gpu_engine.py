"""GPU tests (`-m gpu`) of the pre-allocated multi-stream / CUDA-graph engine: it must reproduce, bit for bit, what the
readable `VoteNetB200.forward` computes with the same kernels, across slots, replays and inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_graph", [False, True])
def test_engine_matches_model_forward(cuda, use_graph):
    from votenet_b200 import synth
    from votenet_b200.config import SAParams, VoteNetConfig
    from votenet_b200.engine import Engine
    from votenet_b200.model import VoteNetB200
    from votenet_b200.weights import make_synthetic_weights

    cfg = VoteNetConfig(num_points=8192,
                        sa=(SAParams(1024, 0.25, 64, (64, 64, 128)), SAParams(512, 0.45, 64, (128, 128, 256)),
                            SAParams(256, 0.8, 64, (128, 128, 256)), SAParams(128, 1.2, 64, (128, 128, 256))),
                        proposal=SAParams(128, 0.35, 64, (128, 128, 128), (128, 128, 79)))
    B = 4
    w = make_synthetic_weights(cfg, 0)
    net = VoteNetB200(cfg, w, device=cuda)
    eng = Engine(cfg, w, B, device=cuda, use_graph=use_graph)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for it in range(5):  # alternates slots and streams, replays the graphs
        xyz = torch.as_tensor(synth.synthetic_batch(100 + it * B, B, cfg.num_points), device=cuda)
        feat = torch.as_tensor(synth.height_feature(xyz.cpu().numpy()), device=cuda)
        ref = net.forward(xyz, feat)
        torch.cuda.synchronize()
        st = streams[it % 2]
        rec = eng.infer_device(xyz, feat, stream=st)
        st.synchronize()
        assert torch.equal(rec.bboxes, ref["dec_bboxes"])
        assert torch.equal(rec.scores, ref["dec_scores"])
        assert torch.equal(rec.keep, ref["nms_keep"])
        n = int(ref["nms_count"].item())
        assert int(rec.nms_count.item()) == n and n > 0
        assert torch.equal(rec.nms_idx[:n], ref["nms_idx"][:n])
    assert eng.launches_per_forward and eng.launches_per_forward > 30


def test_host_api_and_merge(cuda):
    from votenet_b200 import synth
    from votenet_b200.config import SAParams, VoteNetConfig
    from votenet_b200.dist import all_gather_records, merge_gathered, merge_gathered_host
    from votenet_b200.engine import DetectionRecord, Engine
    from votenet_b200.weights import make_synthetic_weights

    cfg = VoteNetConfig(num_points=4096,
                        sa=(SAParams(512, 0.3, 64, (64, 64, 128)), SAParams(256, 0.5, 64, (128, 128, 256)),
                            SAParams(128, 0.9, 64, (128, 128, 256)), SAParams(64, 1.4, 64, (128, 128, 256))),
                        proposal=SAParams(64, 0.4, 64, (128, 128, 128), (128, 128, 79)))
    B = 2
    eng = Engine(cfg, make_synthetic_weights(cfg, 0), B, device=cuda)
    xyz = synth.synthetic_batch(7, B, cfg.num_points)
    hx = torch.as_tensor(xyz).pin_memory(); hf = torch.as_tensor(synth.height_feature(xyz)).pin_memory()
    out = torch.empty((eng.record_nbytes,), dtype=torch.uint8).pin_memory()
    eng.infer_host(hx, hf, out).synchronize()
    host = DetectionRecord(B, cfg.proposal.npoint, buf=out)
    dev = eng.infer_device(hx.to(cuda), hf.to(cuda))
    torch.cuda.synchronize()
    assert torch.equal(host.bboxes, dev.bboxes.cpu()) and torch.equal(host.keep, dev.keep.cpu())
    # single-rank "all-gather" + device merge == host merge == the per-rank NMS output order
    g = all_gather_records(dev.buf, 1)
    idx, cnt = merge_gathered(g, B, cfg.proposal.npoint)
    n = int(cnt.item())
    assert np.array_equal(idx[:n].cpu().numpy(), merge_gathered_host(g, B, cfg.proposal.npoint))
    assert torch.equal(idx[:n], dev.nms_idx[:n]) and n == int(dev.nms_count.item())
    # the pre-allocated gather + merge used by bench.py's loop gives the same list
    from votenet_b200.dist import DetectionGather
    dg = DetectionGather(1, B, cfg.proposal.npoint, cuda, slots=2)
    i1, c1 = dg(dev.buf, slot=1)
    torch.cuda.synchronize()
    assert int(c1.item()) == n and torch.equal(i1[:n], idx[:n])
    # two fake ranks: the merged list interleaves both ranks by score with global batch ids
    g2 = torch.stack([dev.buf, dev.buf.clone()], 0)
    idx2, cnt2 = merge_gathered(g2, B, cfg.proposal.npoint)
    assert int(cnt2.item()) == 2 * n
    assert np.array_equal(idx2[: 2 * n].cpu().numpy(), merge_gathered_host(g2, B, cfg.proposal.npoint))


def test_engine_reference_shapes_batch1(cuda):
    """The reference's own shapes (config.py:1 POINT_NUM = 20480, xyz re-used as the 3 input features model.py:35-36,
    evaluation batch 1, evaluator.py:222): the graph engine reproduces the readable forward bit for bit."""
    from votenet_b200 import synth
    from votenet_b200.config import VoteNetConfig
    from votenet_b200.engine import Engine
    from votenet_b200.model import VoteNetB200
    from votenet_b200.weights import make_synthetic_weights

    cfg = VoteNetConfig(num_points=20480, feature_dim=3)
    w = make_synthetic_weights(cfg, 1)
    net = VoteNetB200(cfg, w, device=cuda)
    eng = Engine(cfg, w, 1, device=cuda, slots=2)
    for it in range(3):
        xyz = torch.as_tensor(synth.synthetic_batch(500 + it, 1, cfg.num_points), device=cuda)
        ref = net.forward(xyz, xyz)
        rec = eng.infer_device(xyz, xyz.clone())
        torch.cuda.synchronize()
        assert torch.equal(rec.bboxes, ref["dec_bboxes"]) and torch.equal(rec.keep, ref["nms_keep"])
        n = int(ref["nms_count"].item())
        assert int(rec.nms_count.item()) == n and torch.equal(rec.nms_idx[:n], ref["nms_idx"][:n])

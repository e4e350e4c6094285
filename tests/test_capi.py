"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/votenet_b200.h declares,
and rejects bad arguments with the reference's error conventions before touching the GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "votenet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vnb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from votenet_b200 import _lib

    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(_lib.lib, n), f"libvotenet_b200.so does not export {n}"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes binding and header disagree"
    assert _lib.lib.vnb_abi_version() == 1


def test_argument_validation_needs_no_gpu():
    from votenet_b200._lib import lib

    null = C.c_void_p(0)
    # npoint > 0 (tf_sampling.cpp:99), radius > 0 / nsample > 0 (tf_grouping.cpp:71,74), 0 <= thr <= 1 (tf_nms3d.cpp:300)
    assert lib.vnb_farthest_point_sample(1, 8, 0, null, null, null) == 1
    assert b"npoint" in lib.vnb_last_error()
    assert lib.vnb_farthest_point_sample(1, 1 << 20, 4, null, null, null) == 1
    assert lib.vnb_query_ball_point(1, 8, 4, -1.0, 4, null, null, null, null, null) == 1
    assert lib.vnb_query_ball_point(1, 8, 4, 0.5, 0, null, null, null, null, null) == 1
    # the two halves of the workspace ball query validate like the whole; without a workspace they are the scan / a no-op
    assert lib.vnb_query_ball_point_prepare(1, 8, -1.0, null, null, null) == 1
    assert b"radius" in lib.vnb_last_error()
    assert lib.vnb_query_ball_point_prepare(1, 8, 0.5, null, null, null) == 0
    assert lib.vnb_query_ball_point_prepared(1, 8, 4, 0.5, 0, null, null, null, null, null, null) == 1
    assert b"nsample" in lib.vnb_last_error()
    assert lib.vnb_query_ball_point_ws(1, 8, 4, -2.0, 4, null, null, null, null, null, null) == 1
    assert lib.vnb_nms3d(1, 4, null, null, null, 1.5, null, null, null, null, null) == 1
    assert b"iou_threshold" in lib.vnb_last_error()
    assert lib.vnb_sa_group_mlp_max(*([1, 8, 1, 2, 32] + [null] * 4 + [64, 64, 128] + [null] * 11 + [1, null, null])) == 1
    assert lib.vnb_linear(4, 8, 8, null, null, null, null, null, 7, null, null, 1, null) == 1
    # fused FP kernel: 256-bit global accesses need 32-byte aligned feature buffers (checked before anything is launched)
    cout = (C.c_int * 2)(256, 256)
    ptrs = (C.c_void_p * 2)(0x1000, 0x2000)
    fake = lambda a: C.c_void_p(a)  # noqa: E731  (never dereferenced: validation fails first)
    rc = lib.vnb_fp_module_fused(1, 128, 64, 256, 256, fake(0x1000), fake(0x1000), fake(0x1010), fake(0x1000), 2, ptrs, ptrs, cout,
                                 fake(0x1000), 0, None, None, None, None, None, None, None, None)
    assert rc == 1 and b"32-byte aligned" in lib.vnb_last_error()
    rc = lib.vnb_fp_module_fused(1, 128, 64, 256, 256, fake(0x1000), fake(0x1000), fake(0x1000), fake(0x1000), 2, ptrs, ptrs, cout,
                                 None, 0, None, None, None, None, None, None, None, None)
    assert rc == 1 and b"fp_out may be NULL only" in lib.vnb_last_error()
    # sizes
    assert lib.vnb_weight_image_bytes(128, 128) == 2 * 128 * 128
    assert lib.vnb_weight_image_bytes(6, 64) == 64 * 128
    assert lib.vnb_nms3d_workspace_bytes(8, 256) >= 2 * 8 * 256 * 4 + 8 * 4 + 4   # kept keys + kept boxes + counts + counter


def test_python_wrappers_refuse_cpu_tensors():
    import torch

    from votenet_b200.tf_grouping import query_ball_point
    from votenet_b200.tf_sampling import farthest_point_sample

    with pytest.raises(TypeError):
        farthest_point_sample(4, torch.zeros(1, 8, 3))
    with pytest.raises(ValueError):
        farthest_point_sample(4, torch.zeros(1, 8, 2))  # (b,n,3) shape check, tf_sampling.cpp:105
    with pytest.raises(ValueError):
        query_ball_point(0.1, 4, torch.zeros(1, 8, 2), torch.zeros(1, 2, 3))


def test_host_weight_logic():
    import torch

    from votenet_b200.config import VoteNetConfig
    from votenet_b200.weights import fold_bn, layer_specs, make_synthetic_weights

    cfg = VoteNetConfig()
    specs = layer_specs(cfg)
    names = [s[0] for s in specs]
    assert names[:3] == ["sa1/conv0", "sa1/conv1", "sa1/conv2"] and "proposal/conv_post_2" in names and "voting2" in names
    d = {n: (ci, co) for n, ci, co, _ in specs}
    assert d["sa1/conv0"] == (4, 64) and d["sa2/conv0"] == (131, 128) and d["sa3/conv0"] == (259, 128)
    assert d["fp1/conv_0"] == (512, 256) and d["voting0"] == (259, 256) and d["voting2"] == (256, 259)
    assert d["proposal/conv_post_2"] == (128, 79)
    total = sum(ci * co + co for _, ci, co, _ in specs)
    assert 0.9e6 < total < 1.0e6  # ~0.96 M parameters
    w = make_synthetic_weights(cfg, 0)
    w2 = make_synthetic_weights(cfg, 0)
    assert all(torch.equal(w[k], w2[k]) for k in w)
    # BN folding == conv -> BN on EMA stats
    x = torch.randn(5, 4)
    W, b = fold_bn(w, "sa1/conv0", cfg.bn_eps)
    y = x @ w["sa1/conv0/W"] + w["sa1/conv0/b"]
    y = (y - w["sa1/conv0/bn/mean/EMA"]) / torch.sqrt(w["sa1/conv0/bn/variance/EMA"] + cfg.bn_eps) * w["sa1/conv0/bn/gamma"] + w["sa1/conv0/bn/beta"]
    assert torch.allclose(x @ W + b, y, atol=1e-5)
    Wl, bl = fold_bn(w, "voting2", cfg.bn_eps)  # last FC has no BN
    assert torch.equal(Wl, w["voting2/W"]) and torch.equal(bl, w["voting2/b"])


def test_synthetic_clouds():
    import numpy as np

    from votenet_b200 import synth

    a = synth.synthetic_cloud(3, 2000)
    assert a.shape == (2000, 3) and a.dtype == np.float32
    assert np.array_equal(a, synth.synthetic_cloud(3, 2000)) and not np.array_equal(a, synth.synthetic_cloud(4, 2000))
    assert a[:, 0].min() > -3.1 and a[:, 0].max() < 3.1 and a[:, 1].max() < 1.3
    h = synth.height_feature(a)
    assert h.shape == (2000, 1) and h.min() > -0.1

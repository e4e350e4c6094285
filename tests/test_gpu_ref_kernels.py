"""GPU tests (`-m gpu`) against the REFERENCE'S OWN GPU KERNELS: tf_sampling_g.cu / tf_grouping_g.cu compiled
unmodified into oracle/_ref/libvotenet_ref_gpu.so in the build container (oracle/Makefile).  This pins both the product
kernels and the CPU oracle's FPS / ball-query restatement to the real reference, bit for bit, at full size."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import ops as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_gpu(cuda):
    if not os.path.exists(O.REF_GPU_PATH):
        pytest.skip("oracle/_ref/libvotenet_ref_gpu.so not built (needs /root/reference at build time)")
    return C.CDLL(O.REF_GPU_PATH)


def P(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("n,m,kind", [(20000, 2048, "room"), (20480, 1024, "uniform"), (3000, 500, "lattice"), (2048, 1024, "room")])
def test_fps_vs_reference_kernel(cuda, ref_gpu, n, m, kind):
    from votenet_b200 import synth
    from votenet_b200.tf_sampling import farthest_point_sample

    b = 3
    rng = np.random.default_rng(n + m)
    if kind == "room":
        x = synth.synthetic_batch(40, b, n)
    elif kind == "uniform":
        x = rng.random((b, n, 3), dtype=np.float32)
    else:
        x = (rng.integers(0, 7, (b, n, 3)) / 4).astype(np.float32)
    tx = torch.as_tensor(x, device=cuda)
    temp = torch.empty((32, n), dtype=torch.float32, device=cuda)
    ref = torch.zeros((b, m), dtype=torch.int32, device=cuda)
    torch.cuda.synchronize()
    assert ref_gpu.ref_gpu_fps(b, n, m, P(tx), P(temp), P(ref)) == 0
    got = farthest_point_sample(m, tx)
    torch.cuda.synchronize()
    assert torch.equal(got, ref), "product FPS differs from the reference GPU kernel"
    assert np.array_equal(O.farthest_point_sample(m, x), ref.cpu().numpy()), "oracle FPS differs from the reference GPU kernel"


@pytest.mark.parametrize("n,m,r,ns", [(20000, 2048, 0.2, 64), (2048, 1024, 0.4, 64), (1024, 256, 0.3, 64), (4000, 300, 0.07, 16)])
def test_ball_query_and_group_vs_reference_kernel(cuda, ref_gpu, n, m, r, ns):
    from votenet_b200 import synth
    from votenet_b200.tf_grouping import group_point, query_ball_point
    from votenet_b200.tf_sampling import farthest_point_sample, gather_point

    b = 2
    x = synth.synthetic_batch(60, b, n)
    tx = torch.as_tensor(x, device=cuda)
    f = farthest_point_sample(m, tx)
    nx = gather_point(tx, f)
    ref_nx = torch.zeros_like(nx)
    assert ref_gpu.ref_gpu_gather_point(b, n, m, P(tx), P(f), P(ref_nx)) == 0
    assert torch.equal(nx, ref_nx)
    ref_idx = torch.zeros((b, m, ns), dtype=torch.int32, device=cuda)
    ref_cnt = torch.zeros((b, m), dtype=torch.int32, device=cuda)
    torch.cuda.synchronize()
    assert ref_gpu.ref_gpu_query_ball_point(b, n, m, C.c_float(r), ns, P(tx), P(nx), P(ref_idx), P(ref_cnt)) == 0
    idx, cnt = query_ball_point(r, ns, tx, nx)
    torch.cuda.synchronize()
    assert torch.equal(cnt, ref_cnt) and torch.equal(idx, ref_idx)
    oi, oc = O.query_ball_point(r, ns, x, nx.cpu().numpy())
    assert np.array_equal(oi, ref_idx.cpu().numpy()) and np.array_equal(oc, ref_cnt.cpu().numpy())
    feats = torch.randn((b, n, 5), device=cuda)
    ref_g = torch.zeros((b, m, ns, 5), device=cuda)
    assert ref_gpu.ref_gpu_group_point(b, n, 5, m, ns, P(feats), P(idx), P(ref_g)) == 0
    assert torch.equal(group_point(feats, idx), ref_g)

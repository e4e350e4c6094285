"""Input stage: the fused device kernel against the numpy restatement of
/root/reference/dataset.py:185-190,302-308 on identical random draws."""
import numpy as np
import pytest

from oracle import evalref as E


def test_draws_follow_the_reference_order():
    from votenet_b200.input_stage import draw_augmentation

    d = draw_augmentation(np.random.RandomState(0), 3, 1000, 64, training=True)
    r = np.random.RandomState(0)
    ch = np.stack([r.choice(1000, 64, replace=False) for _ in range(3)])
    assert np.array_equal(d["choice"], ch) and len(set(d["choice"][0].tolist())) == 64
    fx = r.rand() > 0.5; fz = r.rand() > 0.5; ang = (r.rand() * 2 - 1.) * 5. / 180 * np.pi; sc = (r.rand() * 2 - 1.) * 0.1 + 1.
    assert bool(d["flip_x"][0]) == fx and bool(d["flip_z"][0]) == fz and d["roty_angle"][0] == ang and d["scale"][0] == sc
    assert abs(d["roty_angle"]).max() <= 5 / 180 * np.pi and (abs(d["scale"] - 1) <= 0.1).all()
    e = draw_augmentation(np.random.RandomState(1), 2, 100, 10, training=False)
    assert e["flip_x"] is None and e["scale"] is None


def test_oracle_prepare_input_axis_convention():
    raw = np.array([[[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]], np.float32)   # depth (x right, y forward, z up)
    out = E.prepare_input(raw, dict(choice=np.array([[1, 0]])))
    assert out.tolist() == [[[4.0, -6.0, 5.0], [1.0, -3.0, 2.0]]]        # camera (x right, y down, z forward)


@pytest.mark.gpu
@pytest.mark.parametrize("training", [True, False])
def test_prepare_input_matches_oracle(cuda, training):
    import torch

    from votenet_b200.input_stage import draw_augmentation, prepare_input

    rng = np.random.default_rng(11)
    b, n_raw, n = 4, 50000, 20000
    raw = (rng.random((b, n_raw, 3)) * np.array([6.0, 6.0, 2.7]) - np.array([3.0, 0.0, 1.2])).astype(np.float32)
    draws = draw_augmentation(np.random.RandomState(5), b, n_raw, n, training=training)
    if training:
        draws["flip_x"][:] = [1, 0, 1, 0]; draws["flip_z"][:] = [1, 1, 0, 0]
    ref = E.prepare_input(raw, draws)
    xyz, height = prepare_input(torch.as_tensor(raw, device=cuda), draws, floor_y=1.2)
    got = xyz.cpu().numpy()
    assert got.shape == (b, n, 3)
    # double arithmetic rounded once: at most one float32 ulp from numpy's (BLAS) summation order
    assert np.abs(got - ref).max() <= 2.5e-7 * np.abs(ref).max()
    assert (got == ref).mean() > 0.99
    assert np.allclose(height.cpu().numpy()[..., 0], 1.2 - got[..., 1], atol=1e-6)
    if not training:
        assert np.array_equal(got, ref)   # pure gather + axis flip: exact

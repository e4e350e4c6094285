"""Losses + label assignment: csrc/losses.cu against the numpy restatement of
/root/reference/model.py:62-84,141-231, plus hand-checkable pins of that restatement."""
import numpy as np
import pytest

from oracle import evalref as E


def _case(seed, B=3, N=1024, PR=256, BB=7, near=True):
    rng = np.random.default_rng(seed)
    bxyz = rng.uniform(-2, 2, (B, BB, 3)).astype(np.float32)
    blwh = rng.uniform(0.4, 2.0, (B, BB, 3)).astype(np.float32)
    broty = rng.uniform(-np.pi, np.pi, (B, BB)).astype(np.float32)
    seeds = rng.uniform(-2.5, 2.5, (B, N, 3)).astype(np.float32)
    votes = (seeds + rng.normal(0, 0.3, (B, N, 3))).astype(np.float32)
    prop = rng.uniform(-2.5, 2.5, (B, PR, 3)).astype(np.float32)
    if near:   # some proposals close to a ground-truth centre (positives)
        for b in range(B):
            for j in range(BB):
                prop[b, 3 * j:3 * j + 3] = bxyz[b, j] + rng.normal(0, 0.08, (3, 3)).astype(np.float32)
    pout = rng.standard_normal((B, PR, 79)).astype(np.float32) * 1.5
    sem = rng.integers(0, 10, (B, BB)).astype(np.int32)
    hl = rng.integers(0, 12, (B, BB)).astype(np.int32)
    hr = rng.uniform(-1, 1, (B, BB)).astype(np.float32)
    sl = rng.integers(0, 10, (B, BB)).astype(np.int32)
    sr = rng.uniform(-0.5, 0.5, (B, BB, 3)).astype(np.float32)
    return seeds, votes, prop, pout, bxyz, blwh, broty, sem, hl, hr, sl, sr


def test_oracle_losses_hand_checkable():
    """One cloud, one axis-aligned unit box at the origin, one seed inside (vote off by (0.1,0.2,0.3)) and one outside,
    one positive proposal with zero logits, one negative: every term has a closed form."""
    z = np.zeros
    seeds = np.array([[[0.1, 0.0, 0.0], [3.0, 0.0, 0.0]]], np.float32)
    votes = np.array([[[0.1, 0.2, 0.3], [9.0, 9.0, 9.0]]], np.float32)
    prop = np.array([[[0.1, 0.0, 0.0], [2.0, 0.0, 0.0]]], np.float32)
    pout = z((1, 2, 79), np.float32)
    out = E.votenet_losses(seeds, votes, prop, pout, z((1, 1, 3), np.float32), np.ones((1, 1, 3), np.float32), z((1, 1), np.float32),
                           z((1, 1), np.int32), z((1, 1), np.int32), z((1, 1), np.float32), z((1, 1), np.int32), z((1, 1, 3), np.float32))
    assert abs(out[1] - (0.1 + 0.2 + 0.3) / 2) < 1e-6                       # vote: mean over 2 seeds, only the inside one counts
    assert out[12] == 1 and out[13] == 1                                     # one positive (0.1 < 0.3), one negative (2.0 > 0.6)
    assert abs(out[2] - 2 * np.log(2)) < 1e-9                                # objectness: CE of zero logits, pos + neg
    assert abs(out[4] - 2 * 0.5 * 0.1 ** 2) < 1e-6                           # centre: huber(0 - (-0.1)) for the positive + the dual term
    assert abs(out[5] - np.log(12)) < 1e-9 and abs(out[7] - np.log(10)) < 1e-9 and abs(out[9] - np.log(10)) < 1e-9
    assert out[6] == 0 and out[8] == 0 and out[10] == 1.0 and out[11] == 1.0
    assert abs(out[0] - (out[1] + 0.5 * out[2] + out[3] + 0.1 * out[9])) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("seed,near", [(0, True), (1, True), (2, False)])
def test_device_losses_match_oracle(cuda, seed, near):
    import torch

    from votenet_b200.losses import NAMES, votenet_losses

    args = _case(seed, near=near)
    ref = E.votenet_losses(*args)
    out = votenet_losses(*[torch.as_tensor(a, device=cuda) for a in args]).cpu().numpy()
    assert out[12] == ref[12] and out[13] == ref[13], "label assignment differs"
    if near:
        assert ref[12] > 10
    for i, nme in enumerate(NAMES):
        if np.isnan(ref[i]):
            assert np.isnan(out[i]), nme          # a mean over an empty set is NaN in the reference too
        else:
            assert abs(out[i] - ref[i]) <= 2e-5 * max(1.0, abs(ref[i])), (nme, out[i], ref[i])

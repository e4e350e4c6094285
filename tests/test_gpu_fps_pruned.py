"""GPU parity tests (`-m gpu`) of the bucket-pruned FPS kernel (csrc/fps_pruned.cu) through the C ABI: the index
sequence must be bit-identical to the CPU oracle (oracle/oracle.c restating tf_sampling_g.cu:105-170) and to the
register/cluster kernel for every input — random, clustered, tie-laden, duplicated, padded, m > n."""
import numpy as np
import pytest
import torch

from oracle import ops as O

pytestmark = pytest.mark.gpu


def T(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


@pytest.fixture()
def variant():
    """Sets the FPS variant for one test and restores the default afterwards."""
    from votenet_b200._lib import check, lib

    def set_variant(v):
        check(lib.vnb_set_tuning(b"fps_variant", int(v)))

    yield set_variant
    set_variant(1)


def fps(m, x, dev):
    from votenet_b200.tf_sampling import farthest_point_sample

    return farthest_point_sample(m, T(x, dev)).cpu().numpy()


@pytest.mark.parametrize("n,m", [(1, 1), (2, 2), (3, 5), (31, 31), (33, 7), (160, 160), (161, 50), (512, 128), (513, 64),
                                 (1024, 256), (2048, 1024), (2561, 300), (3000, 64), (4096, 300), (5000, 200),
                                 (8192, 64), (16384, 333), (20000, 256), (20479, 100), (20480, 128)])
def test_pruned_matches_oracle_uniform(cuda, variant, n, m):
    variant(2)  # pruned kernel for every n
    rng = np.random.default_rng(n * 31 + m)
    x = rng.random((3, n, 3), dtype=np.float32) * 4 - 2
    got = fps(m, x, cuda)
    assert got.dtype == np.int32 and got.shape == (3, m)
    assert np.array_equal(got, O.farthest_point_sample(m, x))


def test_pruned_ties_duplicates_lattice(cuda, variant):
    variant(2)
    x = np.zeros((2, 700, 3), np.float32)
    x[0, 600] = (1, 0, 0); x[0, 100] = (1, 0, 0)          # 600 mod 512 = 88 < 100 -> 600 wins
    x[1, 612] = (1, 0, 0); x[1, 100] = (1, 0, 0)          # same slot -> lower k wins
    got = fps(5, x, cuda)
    assert np.array_equal(got, O.farthest_point_sample(5, x))
    assert got[0, 1] == 600 and got[1, 1] == 100
    rng = np.random.default_rng(0)
    # lattice clouds: massive exact ties in every round, at several sizes (ties inside a lane, a warp, across warps)
    for n, m, levels in ((3000, 400, 6), (20000, 700, 12), (8000, 500, 3), (640, 640, 4)):
        lat = (rng.integers(0, levels, (2, n, 3)) / 4).astype(np.float32)
        assert np.array_equal(fps(m, lat, cuda), O.farthest_point_sample(m, lat)), (n, m, levels)
    # every point identical; and more samples than points
    same = np.full((2, 5000, 3), 0.25, np.float32)
    assert np.array_equal(fps(64, same, cuda), O.farthest_point_sample(64, same))
    few = rng.random((2, 40, 3), dtype=np.float32)
    assert np.array_equal(fps(100, few, cuda), O.farthest_point_sample(100, few))
    # planar / collinear clouds (degenerate bounding boxes: zero extent along one or two axes)
    flat = rng.random((2, 6000, 3), dtype=np.float32); flat[:, :, 1] = 1.2
    assert np.array_equal(fps(300, flat, cuda), O.farthest_point_sample(300, flat))
    line = np.zeros((2, 4000, 3), np.float32); line[:, :, 2] = rng.random((2, 4000), dtype=np.float32)
    assert np.array_equal(fps(300, line, cuda), O.farthest_point_sample(300, line))


def test_pruned_clustered_and_far_outliers(cuda, variant):
    variant(2)
    rng = np.random.default_rng(7)
    # tight clusters + a few far outliers: most sub-buckets skip from the very first rounds
    c = rng.normal(0, 0.01, (2, 12000, 3)).astype(np.float32) + rng.integers(0, 3, (2, 12000, 3)).astype(np.float32) * 5
    c[:, :7] = rng.uniform(-100, 100, (2, 7, 3)).astype(np.float32)
    assert np.array_equal(fps(512, c, cuda), O.farthest_point_sample(512, c))
    # negative / large magnitude coordinates
    big = (rng.random((2, 9000, 3), dtype=np.float32) - 0.5) * 2000
    assert np.array_equal(fps(256, big, cuda), O.farthest_point_sample(256, big))


def test_pruned_full_size_equals_oracle_and_cluster_kernel(cuda, variant):
    """BASELINE shape: 8 clouds x 20000 -> 2048 (and the reference's POINT_NUM 20480) on SUN-RGB-D-shaped clouds."""
    from votenet_b200 import synth

    xyz = synth.synthetic_batch(0, 8, 20000)
    variant(1)
    a = fps(2048, xyz, cuda)          # default path (pruned for n > 2048)
    variant(0)
    b = fps(2048, xyz, cuda)          # register/cluster kernel
    assert np.array_equal(a, b)
    assert np.array_equal(a[:3], O.farthest_point_sample(2048, xyz[:3]))
    for row in a:
        assert len(set(row.tolist())) == 2048 and row[0] == 0  # a permutation prefix starting at 0
    xyz2 = synth.synthetic_batch(100, 2, 20480)
    variant(1)
    assert np.array_equal(fps(2048, xyz2, cuda), O.farthest_point_sample(2048, xyz2))


def test_pruned_nested_fallback_flags(cuda, variant):
    """vnb_farthest_point_sample_nested falls back to the sequential sampler (with per-cloud done flags) when the
    identity-prefix proof fails: exercise that path with the pruned kernel."""
    from votenet_b200.tf_sampling import farthest_point_sample_nested

    variant(2)
    rng = np.random.default_rng(3)
    x = rng.random((4, 3000, 3), dtype=np.float32)          # not FPS-ordered: the proof fails for every cloud
    ordered = O.gather_point(x, O.farthest_point_sample(3000, x))[:2]   # FPS-ordered: the proof succeeds
    mix = np.concatenate([ordered, x[2:]], 0)
    got = farthest_point_sample_nested(500, T(mix, cuda)).cpu().numpy()
    assert np.array_equal(got, O.farthest_point_sample(500, mix))
    assert np.array_equal(got[0], np.arange(500))


# ------------------------------------------------------------------------------------------------ tie tracking / provenance hint
def _fps_first_tie_numpy(x, m):
    """Float32-exact emulation of the reference's FPS (tf_sampling_g.cu:105-170, distances contracted as nvcc does:
    fmaf(dz,dz,fmaf(dx,dx,dy*dy))) that also reports the first round whose arg-max was shared by two points."""
    f32, f64 = np.float32, np.float64
    n = x.shape[0]
    k = np.arange(n)
    temp = np.full(n, 1e38, f32)
    picks = [0]
    first_tie = 0x7FFFFFFF
    last = 0
    for r in range(1, m):
        d = (x - x[last]).astype(f32)
        dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
        t = (dx.astype(f64) * dx.astype(f64) + (dy * dy).astype(f32).astype(f64)).astype(f32)
        d2 = (dz.astype(f64) * dz.astype(f64) + t.astype(f64)).astype(f32)
        temp = np.minimum(temp, d2)
        mx = temp.max()
        cand = np.flatnonzero(temp == mx)
        if cand.size > 1 and first_tie == 0x7FFFFFFF:
            first_tie = r
        # reference tie rule: smallest (k mod 512), then smallest k  
        last = int(cand[np.lexsort((cand, cand % 512))[0]])
        picks.append(last)
    return np.asarray(picks, np.int32), first_tie


@pytest.mark.parametrize("kind", ["sun", "uniform", "dup_of_pick", "dup_same_cell", "lattice", "mirror"])
def test_fps_tie_round_is_exact(cuda, kind):
    """vnb_farthest_point_sample_ties reports exactly the first round with a non-unique arg-max (what makes the
    provenance shortcut of vnb_farthest_point_sample_nested_hint sound), on inputs with and without engineered ties."""
    from votenet_b200 import synth
    from votenet_b200.tf_sampling import farthest_point_sample_ties

    rng = np.random.default_rng(42)
    n, m = 20000, 300
    if kind == "sun":
        x = synth.synthetic_cloud(3, n)
    else:
        x = rng.random((n, 3), dtype=np.float32) * np.asarray([6, 2, 6], np.float32)
    if kind == "dup_of_pick":       # an exact copy of the point picked in round 7, far away in index
        p, _ = _fps_first_tie_numpy(x, 8)
        x[15000] = x[p[7]]
    elif kind == "dup_same_cell":   # copies of the round-5 pick right next to it in index order
        p, _ = _fps_first_tie_numpy(x, 6)
        j = int(p[5])
        x[(j + 1) % n] = x[j]
        x[(j + 2) % n] = x[j]
    elif kind == "lattice":
        x = (rng.integers(0, 12, (n, 3)) / 4).astype(np.float32)
    elif kind == "mirror":          # every point has a mirror image across x = 0 and the start point lies on the plane
        half = x[: n // 2].copy()
        half[0, 0] = 0.0
        x = np.concatenate([half, half * np.asarray([-1, 1, 1], np.float32)], 0)
    want_idx, want_tie = _fps_first_tie_numpy(x, m)
    assert np.array_equal(want_idx, O.farthest_point_sample(m, x[None])[0])  # the emulation is the oracle's sequence
    got_idx, got_tie = farthest_point_sample_ties(m, T(x[None], cuda))
    assert np.array_equal(got_idx.cpu().numpy()[0], want_idx)
    assert int(got_tie.item()) == want_tie, (kind, int(got_tie.item()), want_tie)
    if kind in ("sun", "uniform"):
        assert want_tie == 0x7FFFFFFF
    if kind == "dup_of_pick":  # only rounds < track_rounds are examined
        assert want_tie == 7
        for tr, expect in ((7, 0x7FFFFFFF), (8, 7)):
            _, t2 = farthest_point_sample_ties(m, T(x[None], cuda), track_rounds=tr)
            assert int(t2.item()) == expect


@pytest.mark.parametrize("kind", ["sun", "lattice", "dup"])
def test_fps_nested_hint_matches_oracle(cuda, kind):
    """Nested levels through the provenance hint: identical to the oracle whether the parent was tie-free (no proof
    runs) or not (parallel proof / sequential fallback)."""
    from votenet_b200 import synth
    from votenet_b200.tf_sampling import farthest_point_sample_nested_hint, farthest_point_sample_ties, gather_point

    rng = np.random.default_rng(7)
    b, n = 3, 20000
    x = np.stack([synth.synthetic_cloud(10 + i, n) for i in range(b)], 0)
    if kind == "lattice":
        x[1] = (rng.integers(0, 12, (n, 3)) / 4).astype(np.float32)
    elif kind == "dup":
        p = O.farthest_point_sample(40, x[2:3])[0]
        x[2, 12345] = x[2, p[33]]
    tx = T(x, cuda)
    f1, ties = farthest_point_sample_ties(2048, tx)
    assert np.array_equal(f1.cpu().numpy(), O.farthest_point_sample(2048, x))
    src, src_np = gather_point(tx, f1), O.gather_point(x, f1.cpu().numpy())
    for m in (1024, 512, 256):
        got = farthest_point_sample_nested_hint(m, src, ties).cpu().numpy()
        assert np.array_equal(got, O.farthest_point_sample(m, src_np)), (kind, m, ties.tolist())
        src, src_np = src[:, :m].contiguous(), src_np[:, :m]   # == gather of the identity prefix when it holds
        if not np.array_equal(got, np.tile(np.arange(m, dtype=np.int32), (b, 1))):
            break  # a level that is not the identity prefix ends the provenance chain
    if kind == "sun":
        assert (ties.cpu().numpy() == 0x7FFFFFFF).all()


@pytest.mark.parametrize("kind", ["sun", "lattice", "dup", "dup_late"])
def test_fps_nested_proof_chain_matches_oracle(cuda, kind):
    """The engine's sampling chain: plain FPS on the raw cloud, ONE parallel proof at the first nested level
    (2048 -> 1024) whose per-cloud result is the hint of every deeper level and of the proposal module's sampling
    (1024 -> 512, 512 -> 256, 1024 -> 256).  Identical to the oracle with and without engineered ties — including a
    tie beyond the proven rounds and clouds whose proof fails (hint 0: proof / sequential fallback at every level)."""
    from votenet_b200 import synth
    from votenet_b200.tf_sampling import (farthest_point_sample, farthest_point_sample_nested_hint,
                                          farthest_point_sample_nested_proof, gather_point)

    rng = np.random.default_rng(8)
    b, n = 3, 20000
    x = np.stack([synth.synthetic_cloud(20 + i, n) for i in range(b)], 0)
    if kind == "lattice":
        x[1] = (rng.integers(0, 12, (n, 3)) / 4).astype(np.float32)
    elif kind in ("dup", "dup_late"):
        p = O.farthest_point_sample(1500, x[2:3])[0]
        x[2, 12345] = x[2, p[33 if kind == "dup" else 1400]]   # an exact duplicate of an early / a late pick
    tx = T(x, cuda)
    f1 = farthest_point_sample(2048, tx)
    assert np.array_equal(f1.cpu().numpy(), O.farthest_point_sample(2048, x))
    l1, l1_np = gather_point(tx, f1), O.gather_point(x, f1.cpu().numpy())
    f2, proven = farthest_point_sample_nested_proof(1024, l1)
    assert np.array_equal(f2.cpu().numpy(), O.farthest_point_sample(1024, l1_np)), (kind, proven.tolist())
    ident = np.array_equal(f2.cpu().numpy(), np.tile(np.arange(1024, dtype=np.int32), (b, 1)))
    pv = proven.cpu().numpy()
    assert set(pv.tolist()) <= {0, 1024}
    if kind == "sun":
        assert (pv == 1024).all() and ident
    l2, l2_np = gather_point(l1, f2), O.gather_point(l1_np, f2.cpu().numpy())
    # deeper levels + the proposal sampling take the proof's result as their hint (valid while the chain is a prefix chain)
    for (src, src_np, m) in ((l2, l2_np, 512), (l2, l2_np, 256)):
        got = farthest_point_sample_nested_hint(m, src, proven).cpu().numpy()
        assert np.array_equal(got, O.farthest_point_sample(m, src_np)), (kind, m, pv.tolist())
    f3 = farthest_point_sample_nested_hint(512, l2, proven)
    l3, l3_np = gather_point(l2, f3), O.gather_point(l2_np, f3.cpu().numpy())
    got = farthest_point_sample_nested_hint(256, l3, proven).cpu().numpy()
    assert np.array_equal(got, O.farthest_point_sample(256, l3_np)), (kind, pv.tolist())

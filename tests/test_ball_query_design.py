"""CPU model of the grid + sparse-bitmap ball query (csrc/ball_query_grid.cu) against the oracle: the design's exactness
argument — a point within r of the query lies in one of the 27 cells around it for a cell edge >= 1.01 r, and reading an
index bitmap in ascending bit order is the index-ordered scan (tf_grouping_g.cu:16-17) — and the host-side arithmetic of
the bitmap ownership (words per lane, the multiply-shift division) are checked without a GPU."""
import numpy as np
import pytest

from oracle import ops as O

GRID_MAX_DIM = 32


def _d2_max(radius):
    """Largest float32 t with max(sqrtf(t), 1e-20f) < radius: what the kernels compare d2 against."""
    r = np.float32(radius)
    t = np.float32(r * r)
    while np.sqrt(t, dtype=np.float32) >= r:
        t = np.nextafter(t, np.float32(-np.inf), dtype=np.float32)
    while np.sqrt(np.nextafter(t, np.float32(np.inf), dtype=np.float32), dtype=np.float32) < r:
        t = np.nextafter(t, np.float32(np.inf), dtype=np.float32)
    return t


def _d2(q, p):
    """fmaf(dz,dz, fmaf(dx,dx, dy*dy)) in float32, emulated in float64 with one rounding per step."""
    d = (q.astype(np.float32) - p.astype(np.float32)).astype(np.float32)
    dx, dy, dz = d[..., 0].astype(np.float64), d[..., 1].astype(np.float64), d[..., 2].astype(np.float64)
    t = np.float32(dy * dy).astype(np.float64)
    t = np.float32(dx * dx + t).astype(np.float64)     # exact product + one rounding = fmaf for float32 inputs
    return np.float32(dz * dz + t)


def grid_bitmap_query(radius, nsample, xyz1, xyz2):
    """The device algorithm, one cloud: grid build (grid_build_kernel), 27-cell candidate set, exact predicate, bitmap."""
    n, m = xyz1.shape[0], xyz2.shape[0]
    lo, hi = xyz1.min(0), xyz1.max(0)
    cell = max(np.float32(radius) * np.float32(1.01), np.float32((hi - lo).max()) / np.float32(GRID_MAX_DIM - 1))
    inv = np.float32(1.0) / np.float32(cell)
    dims = np.minimum(GRID_MAX_DIM, np.floor((hi - lo) * inv).astype(np.int64) + 1)

    def coord(p):
        c = np.floor((p.astype(np.float32) - lo) * inv).astype(np.int64)
        return np.clip(c, 0, dims - 1)

    cells = coord(xyz1)
    key = (cells[:, 2] * dims[1] + cells[:, 1]) * dims[0] + cells[:, 0]
    by_cell = {}
    for k, c in enumerate(key):
        by_cell.setdefault(int(c), []).append(k)          # order inside a cell is irrelevant (the bitmap restores it)
    t = _d2_max(radius)
    idx = np.zeros((m, nsample), np.int32)
    cnt = np.zeros((m,), np.int32)
    for j in range(m):
        cq = coord(xyz2[j])
        bitmap = np.zeros(n, bool)
        for z in range(max(cq[2] - 1, 0), min(cq[2] + 1, dims[2] - 1) + 1):
            for y in range(max(cq[1] - 1, 0), min(cq[1] + 1, dims[1] - 1) + 1):
                for x in range(max(cq[0] - 1, 0), min(cq[0] + 1, dims[0] - 1) + 1):
                    cand = by_cell.get(int((z * dims[1] + y) * dims[0] + x), [])
                    if cand:
                        cand = np.asarray(cand)
                        bitmap[cand[_d2(xyz2[j][None, :], xyz1[cand]) <= t]] = True
        hits = np.flatnonzero(bitmap)                      # ascending bit order == ascending index order
        c = min(len(hits), nsample)
        cnt[j] = c
        if c:
            idx[j, :c] = hits[:c]
            idx[j, c:] = hits[0]                           # pad with the first hit (tf_grouping_g.cu:26-29)
    return idx, cnt


@pytest.mark.parametrize("n,m,r,ns,kind", [(1500, 60, 0.2, 16, "uniform"), (2000, 40, 0.45, 64, "uniform"), (1200, 50, 0.3, 32, "flat"),
                                           (1000, 40, 0.25, 8, "outside")])
def test_grid_bitmap_model_equals_the_index_ordered_scan(n, m, r, ns, kind):
    rng = np.random.default_rng(n + m)
    x = (rng.random((n, 3), dtype=np.float32) * 2).astype(np.float32)
    if kind == "flat":
        x[:, 1] = 0.5
    q = x[rng.permutation(n)[:m]].copy()
    if kind == "outside":
        q = (q + rng.normal(0, 0.3, q.shape)).astype(np.float32)
        q[:5] += 20.0
    want_i, want_c = O.query_ball_point(r, ns, x[None], q[None])
    got_i, got_c = grid_bitmap_query(r, ns, x, q)
    assert np.array_equal(got_c, want_c[0])
    live = want_c[0] > 0                                   # rows of empty balls are never written by either
    assert np.array_equal(got_i[live], want_i[0][live])


def test_bitmap_ownership_arithmetic():
    """Lane l owns the words [l * per, (l + 1) * per), per = ceil(words / 32); the kernel finds the owner of word w with
    (w * ceil(65536 / per)) >> 16 — exact for every word of every supported cloud size (n <= 32768)."""
    for n in list(range(1024, 32769, 997)) + [20000, 20480, 32768]:
        nw = (n + 31) // 32
        per = (nw + 31) // 32
        assert 1 <= per <= 32 and 32 * per >= nw
        inv = (65536 + per - 1) // per
        w = np.arange(nw, dtype=np.int64)
        owner = (w * inv) >> 16
        assert np.array_equal(owner, w // per) and owner.max() < 32
        assert np.all((w - owner * per >= 0) & (w - owner * per < per))

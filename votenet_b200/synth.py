"""Synthetic SUN-RGB-D-shaped clouds (BASELINE.md §3 / SURVEY.md §8(d)) — there is no dataset in this environment.

Upright-camera frame as produced by the reference's loader (/root/reference/dataset.py:188,310 and
sunutils.py:70-77): x right, y DOWN, z forward.  A room x in [-3,3], z in [0.5,6.5], floor y=+1.2, ceiling
y=-1.5; 40 % of the points on the floor, 25 % on the back wall, 10 % on each side wall, 15 % on the surfaces of
six furniture-sized boxes (sizes from the reference's class_mean_size table, dataset.py:36-45) resting on the
floor; N(0, 0.005) jitter; shuffled; float32.  Height feature = 1.2 - y.
"""
import numpy as np

# (l, w, h) per class, class order of /root/reference/dataset.py:30-49 (type2class)
CLASS_MEAN_SIZE = np.array(
    [
        [2.114256, 1.620300, 0.927272],  # bed
        [0.791118, 1.279516, 0.718182],  # table
        [0.923508, 1.867419, 0.845495],  # sofa
        [0.591958, 0.552978, 0.827272],  # chair
        [0.699104, 0.454178, 0.756250],  # toilet
        [0.695190, 1.346299, 0.736364],  # desk
        [0.528526, 1.002642, 1.172878],  # dresser
        [0.500618, 0.632163, 0.683424],  # night_stand
        [0.404671, 1.071108, 1.688889],  # bookshelf
        [0.765840, 1.398258, 0.472728],  # bathtub
    ],
    dtype=np.float32,
)

FLOOR_Y = 1.2


def synthetic_cloud(cloud_id, n=20000):
    """One cloud: returns xyz (n,3) float32.  Deterministic in `cloud_id` (seed 1000 + cloud_id)."""
    rng = np.random.default_rng(1000 + int(cloud_id))
    n_floor = int(0.40 * n)
    n_back = int(0.25 * n)
    n_side = int(0.10 * n)
    n_obj = n - n_floor - n_back - 2 * n_side
    parts = []
    # floor (y = +1.2)
    p = np.empty((n_floor, 3))
    p[:, 0] = rng.uniform(-3, 3, n_floor)
    p[:, 1] = FLOOR_Y
    p[:, 2] = rng.uniform(0.5, 6.5, n_floor)
    parts.append(p)
    # back wall (z = 6.5)
    p = np.empty((n_back, 3))
    p[:, 0] = rng.uniform(-3, 3, n_back)
    p[:, 1] = rng.uniform(-1.5, FLOOR_Y, n_back)
    p[:, 2] = 6.5
    parts.append(p)
    # side walls (x = -3, x = +3)
    for xw in (-3.0, 3.0):
        p = np.empty((n_side, 3))
        p[:, 0] = xw
        p[:, 1] = rng.uniform(-1.5, FLOOR_Y, n_side)
        p[:, 2] = rng.uniform(0.5, 6.5, n_side)
        parts.append(p)
    # six furniture boxes resting on the floor, yawed about y
    per = [n_obj // 6 + (1 if i < n_obj % 6 else 0) for i in range(6)]
    for k in range(6):
        cls = int(rng.integers(0, 10))
        l, w, h = CLASS_MEAN_SIZE[cls].astype(np.float64) * rng.uniform(0.9, 1.1)
        cx, cz = rng.uniform(-2.2, 2.2), rng.uniform(1.3, 5.7)
        yaw = rng.uniform(0, 2 * np.pi)
        m = per[k]
        # sample points on the 5 visible faces (top + 4 sides), area-weighted
        areas = np.array([l * w, l * h, l * h, w * h, w * h])
        face = rng.choice(5, size=m, p=areas / areas.sum())
        u, v = rng.uniform(-0.5, 0.5, m), rng.uniform(-0.5, 0.5, m)
        loc = np.empty((m, 3))
        # local frame: x along l, z along w, y up-negative (y down): top face at y = FLOOR_Y - h
        top = face == 0
        loc[top] = np.stack([u[top] * l, np.full(top.sum(), -h), v[top] * w], 1)
        for f, (sx, sz) in zip((1, 2), ((0, 0.5), (0, -0.5))):
            s = face == f
            loc[s] = np.stack([u[s] * l, (v[s] - 0.5) * h, np.full(s.sum(), sz * w)], 1)
        for f, sx in zip((3, 4), (0.5, -0.5)):
            s = face == f
            loc[s] = np.stack([np.full(s.sum(), sx * l), (v[s] - 0.5) * h, u[s] * w], 1)
        c, s_ = np.cos(yaw), np.sin(yaw)
        wx = c * loc[:, 0] + s_ * loc[:, 2] + cx
        wz = -s_ * loc[:, 0] + c * loc[:, 2] + cz
        wy = FLOOR_Y + loc[:, 1]
        parts.append(np.stack([wx, wy, wz], 1))
    pts = np.concatenate(parts, 0)
    pts += rng.normal(0.0, 0.005, pts.shape)
    pts = pts[rng.permutation(pts.shape[0])]
    return pts.astype(np.float32)


def synthetic_batch(first_cloud_id, batch, n=20000):
    """(batch, n, 3) float32 xyz for cloud ids first_cloud_id .. first_cloud_id+batch-1."""
    return np.stack([synthetic_cloud(first_cloud_id + i, n) for i in range(batch)], 0)


def height_feature(xyz):
    """(…,n,3) -> (…,n,1): height above the synthetic floor (y is down)."""
    return (FLOOR_Y - xyz[..., 1:2]).astype(np.float32)

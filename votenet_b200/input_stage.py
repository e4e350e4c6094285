"""Input stage on the GPU — the host-side mirror of the reference's per-scene preparation in
/root/reference/dataset.py:185-190,219-231,302-308: subsample POINT_NUM points, flip the axes to
the upright camera frame, and (training) flip / rotate / scale — one fused kernel over the whole batch
(csrc/input_stage.cu).  The random draws are made on the host exactly as the reference makes them (`draw_augmentation`),
so a given numpy RandomState reproduces the reference's clouds."""
import numpy as np
import torch

from ._lib import check, dptr, lib, stream_ptr


def draw_augmentation(rng, batch, n_raw, n, training=True):
    """The reference's draws, in its order, per cloud: choice without replacement (dataset.py:185-186), then — training
    only — flip_x, flip_z, rotation angle, scale (:219-231).  `rng` is a numpy RandomState-like (choice / rand)."""
    choice = np.stack([rng.choice(n_raw, n, replace=False) for _ in range(batch)]).astype(np.int32)
    if not training:
        return dict(choice=choice, flip_x=None, flip_z=None, roty_angle=None, scale=None)
    fx, fz, ang, sc = [], [], [], []
    for _ in range(batch):
        fx.append(rng.rand() > 0.5)
        fz.append(rng.rand() > 0.5)
        ang.append((rng.rand() * 2 - 1.) * 5. / 180 * np.pi)
        sc.append((rng.rand() * 2 - 1.) * 0.1 + 1.)
    return dict(choice=choice, flip_x=np.array(fx, np.uint8), flip_z=np.array(fz, np.uint8),
                roty_angle=np.array(ang, np.float64), scale=np.array(sc, np.float64))


def prepare_input(raw_upright_depth, draws, floor_y=None):
    """raw_upright_depth (B,n_raw,3) CUDA f32 + `draws` (draw_augmentation) -> xyz (B,n,3) f32 in the upright camera frame
    (and height (B,n,1) f32 = floor_y - y when floor_y is given).  Asynchronous on the current stream."""
    if raw_upright_depth.dim() != 3 or raw_upright_depth.shape[2] != 3:
        raise ValueError("prepare_input expects (batch, n_raw, 3) points")
    b, n_raw, _ = raw_upright_depth.shape
    dev = raw_upright_depth.device
    choice = draws.get("choice")
    n = choice.shape[1] if choice is not None else n_raw
    up = lambda a, dt: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dt), device=dev)  # noqa: E731
    t_choice, t_fx, t_fz = up(choice, np.int32), up(draws.get("flip_x"), np.uint8), up(draws.get("flip_z"), np.uint8)
    t_ang, t_sc = up(draws.get("roty_angle"), np.float64), up(draws.get("scale"), np.float64)
    xyz = torch.empty((b, n, 3), dtype=torch.float32, device=dev)
    height = torch.empty((b, n, 1), dtype=torch.float32, device=dev) if floor_y is not None else None
    check(lib.vnb_prepare_input(b, n_raw, n, dptr(raw_upright_depth, torch.float32), dptr(t_choice), dptr(t_fx), dptr(t_fz),
                                dptr(t_ang), dptr(t_sc), dptr(xyz), dptr(height), float(floor_y or 0.0), stream_ptr()))
    return (xyz, height) if floor_y is not None else xyz

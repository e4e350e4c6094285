"""Drop-in for the reference's tf_ops/sampling/tf_sampling.py (same function names, argument order, output
dtypes/shapes) over CUDA tensors.  prob_sample is off the VoteNet path and not provided."""
import torch

from ._lib import check, dptr, lib, stream_ptr


def _xyz3(t, name):
    if t.dim() != 3 or t.shape[2] != 3:
        raise ValueError(f"{name} expects (batch_size,num_points,3) inp shape")  # tf_sampling.cpp:105,131
    return t


def farthest_point_sample(npoint, inp):
    """inp (B,N,3) f32 -> (B,npoint) i32.   Reference: tf_sampling.py:48-56 (FarthestPointSample, tf_sampling.cpp:95-123)."""
    inp = _xyz3(inp, "FarthestPointSample")
    b, n, _ = inp.shape
    out = torch.empty((b, int(npoint)), dtype=torch.int32, device=inp.device)
    check(lib.vnb_farthest_point_sample(b, n, int(npoint), dptr(inp, torch.float32, "inp"), dptr(out), stream_ptr()))
    return out


def farthest_point_sample_nested(npoint, inp):
    """Bit-identical to farthest_point_sample; fast when `inp` is itself FPS-ordered (nested SA levels): proves in
    parallel that the answer is the identity prefix and only falls back to the sequential sampler where it is not."""
    inp = _xyz3(inp, "FarthestPointSample")
    b, n, _ = inp.shape
    out = torch.empty((b, int(npoint)), dtype=torch.int32, device=inp.device)
    ws = torch.empty((lib.vnb_fps_nested_workspace_bytes(b, int(npoint)),), dtype=torch.uint8, device=inp.device)
    check(lib.vnb_farthest_point_sample_nested(b, n, int(npoint), dptr(inp, torch.float32, "inp"), dptr(out), dptr(ws),
                                               stream_ptr()))
    return out


def farthest_point_sample_ties(npoint, inp, track_rounds=0):
    """farthest_point_sample that also returns, per cloud, the first round whose arg-max was not unique (int32 (B,);
    0x7fffffff = every round had a unique winner).  Feeds farthest_point_sample_nested(..., parent_first_tie=...)."""
    inp = _xyz3(inp, "FarthestPointSample")
    b, n, _ = inp.shape
    out = torch.empty((b, int(npoint)), dtype=torch.int32, device=inp.device)
    ties = torch.empty((b,), dtype=torch.int32, device=inp.device)
    check(lib.vnb_farthest_point_sample_ties(b, n, int(npoint), dptr(inp, torch.float32, "inp"), dptr(out), dptr(ties),
                                             int(track_rounds), stream_ptr()))
    return out, ties


def farthest_point_sample_nested_hint(npoint, inp, parent_first_tie):
    """farthest_point_sample_nested for an `inp` that is gather_point(parent_xyz, parent_fps_idx[:, :N]) of a
    farthest_point_sample_ties call: clouds whose parent was tie-free for `npoint` rounds need no proof at all."""
    inp = _xyz3(inp, "FarthestPointSample")
    b, n, _ = inp.shape
    out = torch.empty((b, int(npoint)), dtype=torch.int32, device=inp.device)
    ws = torch.empty((lib.vnb_fps_nested_workspace_bytes(b, int(npoint)),), dtype=torch.uint8, device=inp.device)
    check(lib.vnb_farthest_point_sample_nested_hint(b, n, int(npoint), dptr(inp, torch.float32, "inp"), dptr(out), dptr(ws),
                                                    dptr(parent_first_tie, torch.int32, "parent_first_tie"), stream_ptr()))
    return out


def farthest_point_sample_nested_proof(npoint, inp):
    """farthest_point_sample_nested that also returns what it proved: (idx (B,npoint) i32, proven (B,) i32) with
    proven[b] = npoint where the identity prefix was proven, else 0 — a valid `parent_first_tie` hint for
    farthest_point_sample_nested_hint on any prefix inp[:, :n'] and npoint' <= npoint (include/votenet_b200.h)."""
    inp = _xyz3(inp, "FarthestPointSample")
    b, n, _ = inp.shape
    out = torch.empty((b, int(npoint)), dtype=torch.int32, device=inp.device)
    proven = torch.empty((b,), dtype=torch.int32, device=inp.device)
    ws = torch.empty((lib.vnb_fps_nested_workspace_bytes(b, int(npoint)),), dtype=torch.uint8, device=inp.device)
    check(lib.vnb_farthest_point_sample_nested_proof(b, n, int(npoint), dptr(inp, torch.float32, "inp"), dptr(out), dptr(ws),
                                                     dptr(proven), stream_ptr()))
    return out, proven


class _GatherPointFn(torch.autograd.Function):
    """GatherPoint with its registered gradient (@tf.RegisterGradient('GatherPoint'), tf_sampling.py:43-47)."""

    @staticmethod
    def forward(ctx, inp, idx):
        ctx.save_for_backward(inp, idx)
        return _gather_point_fwd(inp, idx)

    @staticmethod
    def backward(ctx, out_g):
        inp, idx = ctx.saved_tensors
        return gather_point_grad(inp, idx, out_g.contiguous()), None


def gather_point(inp, idx):
    """inp (B,N,3) f32, idx (B,M) i32 -> (B,M,3) f32.   Reference: tf_sampling.py:29-37 (GatherPoint, tf_sampling.cpp:126-148).
    Differentiable w.r.t. inp (GatherPointGrad) when inp requires grad."""
    if torch.is_grad_enabled() and inp.requires_grad:
        return _GatherPointFn.apply(inp, idx)
    return _gather_point_fwd(inp, idx)


def _gather_point_fwd(inp, idx):
    inp = _xyz3(inp, "GatherPoint")
    if idx.dim() != 2 or idx.shape[0] != inp.shape[0]:
        raise ValueError("GatherPoint expects (batch_size,num_result) idx shape")  # tf_sampling.cpp:136
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
    check(lib.vnb_gather_point(b, n, m, dptr(inp, torch.float32, "inp"), dptr(idx, torch.int32, "idx"), dptr(out),
                               stream_ptr()))
    return out


def gather_point_grad(inp, idx, out_g):
    """Gradient of gather_point w.r.t. inp: (B,N,3) <- scatter-add of out_g (B,M,3) at idx (B,M).
    Reference: GatherPointGrad, tf_sampling.py:43-47 / tf_sampling.cpp:150-178 / tf_sampling_g.cu:183-192."""
    inp = _xyz3(inp, "GatherPointGrad")
    if idx.dim() != 2 or idx.shape[0] != inp.shape[0]:
        raise ValueError("GatherPointGradGpuOp expects (batch_size,num_result) idx shape")        # tf_sampling.cpp:161
    b, n, _ = inp.shape
    m = idx.shape[1]
    if out_g.dim() != 3 or tuple(out_g.shape) != (b, m, 3):
        raise ValueError("GatherPointGradGpuOp expects (batch_size,num_result,3) out_g shape")    # tf_sampling.cpp:167
    inp_g = torch.empty((b, n, 3), dtype=torch.float32, device=inp.device)
    check(lib.vnb_gather_point_grad(b, n, m, dptr(out_g, torch.float32, "out_g"), dptr(idx, torch.int32, "idx"), dptr(inp_g),
                                    stream_ptr()))
    return inp_g

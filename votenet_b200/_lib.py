"""ctypes binding of libvotenet_b200.so (the C ABI declared in include/votenet_b200.h).

PyTorch is used for device memory and streams only.  There is NO CPU fallback: if the CUDA library has not been built
the import fails loudly; if a tensor is not a contiguous CUDA tensor of the expected dtype the call raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VNB_LIB", os.path.join(_HERE, "libvotenet_b200.so"))   # VNB_LIB: an instrumented debug build

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the sm_100a CUDA library first (python -m votenet_b200.build, or "
        f"__graft_entry__.build()).  votenet_b200 has no CPU / PyTorch fallback."
    )

lib = C.CDLL(LIB_PATH)

_i, _f, _p, _sz = C.c_int, C.c_float, C.c_void_p, C.c_size_t
_SIGS = {
    "vnb_abi_version": ([], _i),
    "vnb_last_error": ([], C.c_char_p),
    "vnb_launch_count": ([], C.c_ulonglong),
    "vnb_set_tuning": ([C.c_char_p, _i], _i),
    "vnb_debug_fps_profile": ([_p], _i),
    "vnb_debug_sa_trace": ([_p], _i),
    "vnb_debug_trap_buffer": ([_p], _i),
    "vnb_fps_nested_workspace_bytes": ([_i, _i], _sz),
    "vnb_farthest_point_sample_nested": ([_i, _i, _i, _p, _p, _p, _p], _i),
    "vnb_farthest_point_sample_ties": ([_i, _i, _i, _p, _p, _p, _i, _p], _i),
    "vnb_farthest_point_sample_nested_hint": ([_i, _i, _i, _p, _p, _p, _p, _p], _i),
    "vnb_farthest_point_sample_nested_proof": ([_i, _i, _i, _p, _p, _p, _p, _p], _i),
    "vnb_query_ball_point_workspace_bytes": ([_i, _i], _sz),
    "vnb_query_ball_point_ws": ([_i, _i, _i, _f, _i, _p, _p, _p, _p, _p, _p], _i),
    "vnb_query_ball_point_prepare": ([_i, _i, _f, _p, _p, _p], _i),
    "vnb_query_ball_point_prepared": ([_i, _i, _i, _f, _i, _p, _p, _p, _p, _p, _p], _i),
    "vnb_merge_detections": ([_i, _i, _i, _p, _sz, _sz, _sz, _sz, _sz, _sz, _p, _p, _p, _p, _p, _p], _i),
    "vnb_sa_group_mlp_max_backward": ([_i, _i, _i, _i, _i, _p, _p, _p, _p, _i, _i, _i] + [_p] * 18 + [_p], _i),
    "vnb_votenet_losses": ([_i, _i, _i, _i] + [_p] * 12 + [_f, _f, _p, _p, _p], _i),
    "vnb_prepare_input": ([_i, _i, _i] + [_p] * 8 + [C.c_double, _p], _i),
    "vnb_iou3d_pairs": ([_i, _p, _p, _p, _p], _i),
    "vnb_eval_det_cls_workspace_bytes": ([_i, _i], _sz),
    "vnb_eval_det_cls": ([_i, _i, _i, _p, _p, _p, _p, _p, C.c_double, _p, _p, _p, _p, _p], _i),
    "vnb_peer_push_record": ([_i, _i, _p, _sz, _p, _p, _i, _p], _i),
    "vnb_peer_wait": ([_i, _p, _i, _p], _i),
    "vnb_peer_enable_access": ([_i], _i),
    "vnb_peer_alloc": ([_sz, _p, _p], _i),
    "vnb_peer_open": ([_p, _p], _i),
    "vnb_peer_close": ([_p], _i),
    "vnb_peer_free": ([_p], _i),
    "vnb_decode_nms3d": ([_i, _i, _p, _p, _p, _f] + [_p] * 13, _i),
    "vnb_farthest_point_sample": ([_i, _i, _i, _p, _p, _p], _i),
    "vnb_gather_point": ([_i, _i, _i, _p, _p, _p, _p], _i),
    "vnb_query_ball_point": ([_i, _i, _i, _f, _i, _p, _p, _p, _p, _p], _i),
    "vnb_group_point": ([_i, _i, _i, _i, _i, _p, _p, _p, _p], _i),
    "vnb_three_nn": ([_i, _i, _i, _p, _p, _p, _p, _p], _i),
    "vnb_three_interpolate": ([_i, _i, _i, _i, _p, _p, _p, _p, _p], _i),
    "vnb_nms3d_workspace_bytes": ([_i, _i], _sz),
    "vnb_nms3d_near_threshold_offset": ([_i, _i], _sz),
    "vnb_nms3d": ([_i, _i, _p, _p, _p, _f, _p, _p, _p, _p, _p], _i),
    "vnb_weight_image_bytes": ([_i, _i], _sz),
    "vnb_pack_weight_f16": ([_i, _i, _p, _p, _p], _i),
    "vnb_linear": ([_i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p, _i, _p], _i),
    "vnb_sa_group_mlp_max": ([_i] * 5 + [_p] * 4 + [_i] * 3 + [_p] * 11 + [_i, _p, _p], _i),
    "vnb_sa_group_mlp_max_counted": ([_i] * 5 + [_p] * 5 + [_i] * 3 + [_p] * 11 + [_i, _p, _p], _i),
    "vnb_sa_workspace_bytes": ([_i, _i, _i], _sz),
    "vnb_fp_interpolate_concat": ([_i] * 5 + [_p] * 6, _i),
    "vnb_fp_module_fused": ([_i] * 5 + [_p] * 4 + [_i] + [_p] * 4 + [_i] + [_p] * 8, _i),
    "vnb_gather_point_grad": ([_i, _i, _i, _p, _p, _p, _p], _i),
    "vnb_group_point_grad": ([_i] * 5 + [_p] * 4, _i),
    "vnb_three_interpolate_grad": ([_i] * 4 + [_p] * 5, _i),
    "vnb_concat2": ([_i, _i, _i, _p, _p, _p, _p], _i),
    "vnb_split2": ([_i, _i, _i, _p, _p, _p, _p], _i),
    "vnb_decode_boxes": ([_i, _i] + [_p] * 8, _i),
}
for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not export a declared symbol
    _fn.argtypes = _args
    _fn.restype = _res

EXPORTED_SYMBOLS = tuple(_SIGS)
VNB_ERR_INVALID, VNB_ERR_CUDA = 1, 2

if lib.vnb_abi_version() != 1:
    raise ImportError("libvotenet_b200.so: ABI version mismatch (rebuild)")


def check(rc):
    """0 -> ok; VNB_ERR_INVALID -> ValueError (the reference raises InvalidArgument); VNB_ERR_CUDA -> RuntimeError."""
    if rc == 0:
        return
    msg = lib.vnb_last_error().decode("utf-8", "replace")
    if rc == VNB_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(f"votenet_b200 CUDA error: {msg}")


def stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def dptr(t, dtype=None, name="tensor"):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not torch.is_tensor(t) or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor (votenet_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    if t.device.index != torch.cuda.current_device():   # the C side launches on the CURRENT device
        raise ValueError(f"{name}: tensor lives on {t.device} but the current CUDA device is {torch.cuda.current_device()} "
                         "(wrap the call in `with torch.cuda.device(...)`)")
    return C.c_void_p(t.data_ptr())

"""votenet_b200 — B200 (sm_100a) implementation of the VoteNet / PointNet++ inference hot path behind the reference's
own Python op signatures (tf_ops/{sampling,grouping,3d_interpolation,3d_nms}) and layer functions.

    from votenet_b200.tf_sampling import farthest_point_sample, gather_point
    from votenet_b200.tf_grouping import query_ball_point, group_point
    from votenet_b200.tf_interpolate import three_nn, three_interpolate
    from votenet_b200.tf_nms3d import NMS3D
    from votenet_b200.utils import pointnet_sa_module, pointnet_fp_module

Sub-modules that launch kernels import `votenet_b200._lib`, which fails loudly when libvotenet_b200.so has not been
built; `config`, `weights` and `synth` are pure host code.  There is no CPU fallback anywhere in this package.
"""
__version__ = "0.1.0"

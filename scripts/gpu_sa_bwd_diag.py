import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_train import _params
from votenet_b200.train import pointnet_sa_module_fused_trainable, pointnet_sa_module_trainable
cuda = torch.device("cuda:0")
for (b, n, c, m, dims, r) in [(2, 1500, 16, 128, (64, 64, 128), 0.25), (1, 900, 16, 64, (128, 128, 256), 0.3), (1, 900, 128, 64, (64, 64, 128), 0.3),
                              (1, 900, 128, 64, (128, 128, 256), 0.3), (1, 900, 61, 64, (128, 128, 128), 0.3)]:
    rng = np.random.default_rng(c)
    xyz_np = rng.random((b, n, 3)).astype(np.float32); pts_np = rng.standard_normal((b, n, c)).astype(np.float32)
    g_np = None; grads = []
    for fn in (pointnet_sa_module_trainable, pointnet_sa_module_fused_trainable):
        layers = _params(np.random.default_rng(1), [3 + c] + list(dims), cuda)
        xyz = torch.as_tensor(xyz_np, device=cuda).requires_grad_(True); pts = torch.as_tensor(pts_np, device=cuda).requires_grad_(True)
        _, out, idx = fn(xyz, pts, m, r, 64, layers)
        if g_np is None: g_np = rng.standard_normal(tuple(out.shape)).astype(np.float32)
        out.backward(torch.as_tensor(g_np, device=cuda))
        grads.append((out.detach(), [pts.grad, xyz.grad] + [t.grad for Wb in layers for t in Wb]))
    (o_ref, g_ref), (o_got, g_got) = grads
    print(f"c={c} dims={dims}: fwd err {(o_ref-o_got).abs().max().item():.2e}", " ".join(f"{nm}:{(a-bb).abs().max().item()/max(a.abs().max().item(),1e-9):.1e}" for nm, a, bb in zip(["dP","dX","dW1","db1","dW2","db2","dW3","db3"], g_ref, g_got)), flush=True)

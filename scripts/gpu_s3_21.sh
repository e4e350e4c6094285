#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
python - <<'PY'
import torch, numpy as np, sys
sys.path.insert(0,'.')
from votenet_b200.tf_grouping import group_point_grad
from votenet_b200.tf_interpolate import three_interpolate_grad
from votenet_b200.tf_sampling import gather_point_grad
dev='cuda'
def t(fn, it=20):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(True),torch.cuda.Event(True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/it
B=8
# SA2 shapes: grad_out (8,1024,64,128) -> points (8,2048,128)
idx=torch.randint(0,2048,(B,1024,64),dtype=torch.int32,device=dev); g=torch.randn(B,1024,64,128,device=dev); p=torch.zeros(B,2048,128,device=dev)
ms=t(lambda: group_point_grad(p,idx,g)); byt=g.numel()*4+idx.numel()*4+p.numel()*4
print(f"group_point_grad sa2 shape: {ms*1e3:.1f} us  {byt/ms/1e6:.0f} GB/s (read grad_out + idx, write grad_points)")
idx3=torch.randint(0,512,(B,1024,3),dtype=torch.int32,device=dev); w=torch.rand(B,1024,3,device=dev); g3=torch.randn(B,1024,256,device=dev); p3=torch.zeros(B,512,256,device=dev)
ms=t(lambda: three_interpolate_grad(p3,idx3,w,g3)); byt=g3.numel()*4+p3.numel()*4
print(f"three_interpolate_grad fp2 shape: {ms*1e3:.1f} us  {byt/ms/1e6:.0f} GB/s")
xi=torch.randint(0,20000,(B,2048),dtype=torch.int32,device=dev); xg=torch.randn(B,2048,3,device=dev); xp=torch.zeros(B,20000,3,device=dev)
ms=t(lambda: gather_point_grad(xp,xi,xg)); print(f"gather_point_grad sa1 shape: {ms*1e3:.1f} us")
PY

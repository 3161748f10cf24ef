import sys, torch
sys.path.insert(0, "/root/repo")
from ctrlora_b200 import ops
from tools.profile_kernels import rnd
b,h,w,c,n,ks = 8,64,64,320,320,1
a, wt = rnd(b,h,w,c), rnd(n, ks*ks, c, scale=(ks*ks*c)**-0.5)
bias, res = torch.randn(n, device="cuda"), rnd(b*h*w, n)
out = torch.empty(b,h,w,n, device="cuda", dtype=torch.float16)
for _ in range(3): ops.gemm(a, wt, ksize=ks, bias=bias, residual=res, out=out)
torch.cuda.synchronize()

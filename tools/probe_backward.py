import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere
dev="cuda:0"
for n in (1000, 5000):
    gen=torch.Generator().manual_seed(0)
    x=torch.nn.functional.normalize(torch.randn(n,11,generator=gen),dim=-1).to(dev)
    k=kernels_sphere.SphereRiemannianMaternKernel(dim=10,nu=2.5).to(dev); k.lengthscale=0.5
    k.raw_lengthscale.requires_grad_(True)
    G=torch.randn(n,n,generator=gen).to(dev)
    def step(xg=False):
        xx = x.clone().requires_grad_(xg)
        K=k.forward(xx,x)
        return torch.autograd.grad((K*G).sum(), [k.raw_lengthscale]+([xx] if xg else []))
    for xg in (False, True):
        for _ in range(3): step(xg)
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(10): step(xg)
        torch.cuda.synchronize(); print("n",n,"input grad",xg,"fwd+bwd ms",(time.perf_counter()-t0)/10*1e3)

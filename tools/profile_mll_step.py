import sys, os
sys.path.insert(0, os.getcwd())
import torch
torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere
from materngabo_b200.gpytorch_compat import ScaleKernel
from materngabo_b200.graphs import exact_mll
from torch.profiler import profile, ProfilerActivity
dev="cuda:0"
gen=torch.Generator().manual_seed(0)
x=torch.nn.functional.normalize(torch.randn(100,4,generator=gen),dim=-1).to(dev); y=torch.sin(3*x[:,0])
base=kernels_sphere.SphereRiemannianMaternKernel(dim=3,nu=2.5); sk=ScaleKernel(base).to(dev)
params=[base.raw_lengthscale, sk.raw_outputscale]
def step():
    loss,info=exact_mll(sk,x,y,1e-2,0.0)
    return torch.autograd.grad(loss,params)
for _ in range(5): step()
torch.cuda.synchronize()
# phase split: coefficients only
def coef_only():
    spec,coef=base.series()
    return torch.autograd.grad(coef.sum(),[base.raw_lengthscale])
for _ in range(3): coef_only()
for name,fn in (("mll",step),("coef",coef_only)):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
        fn(); torch.cuda.synchronize()
    evs=[e for e in p.events() if e.device_type==torch.autograd.DeviceType.CUDA]
    tot=sum(e.device_time for e in evs) if hasattr(evs[0],'device_time') else sum(e.cuda_time for e in evs)
    print(name,"cuda kernels",len(evs),"total device us",tot)
    import collections
    c=collections.Counter(e.name[:60] for e in evs)
    for k,v in c.most_common(12): print("   ",v,k)

import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import torch
torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere, _lib, _ops
from materngabo_b200.kernels_sphere import _series_on
dev="cuda:0"
lib=_lib.load()
for n in (100,160,30):
    gen=torch.Generator().manual_seed(0)
    x=torch.nn.functional.normalize(torch.randn(n,4,generator=gen),dim=-1).to(dev); y=torch.sin(3*x[:,0])
    base=kernels_sphere.SphereRiemannianMaternKernel(dim=3,nu=2.5).to(dev)
    with torch.no_grad(): spec,coef=_series_on(base,x.device)
    t=coef.shape[-1]
    out=torch.zeros(4+t+16,device=dev); st=torch.zeros(1,dtype=torch.int32,device=dev)
    hyper=torch.tensor([0.7,1e-2,0.0],device=dev)
    d=spec.desc(t)
    def run():
        rc=lib.mgb_series_mll(C.byref(d),_lib.ptr(x),n,x.stride(0),_lib.ptr(y),_lib.ptr(coef),_lib.ptr(hyper),_lib.ptr(out),_lib.ptr(st),_lib.stream_ptr(x.device)); assert rc==0
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print("n",n,"kernel us",e0.elapsed_time(e1)/20*1e3, "phase cycles", [int(v) for v in out[4+t:4+t+8].tolist()])

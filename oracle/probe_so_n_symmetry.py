"""TEST INFRASTRUCTURE ONLY (dev container, needs /root/reference): is the reference's SO(n), n >= 4, kernel a
well-defined function of the pair?

kernels_so.py:329-346 feeds one eigenvalue of each conjugate pair of R1 R2^T, in the order torch.linalg.eigvals
returns them, into a character polynomial written in simple-root coordinates (kernels_so.py:234-239,269).  This
script evaluates that polynomial at permuted / conjugated arguments.  Result (recorded in DESIGN.md section 8):
  n = 4: invariant under swapping the two arguments, NOT under conjugating one of them (value depends on LAPACK
         listing the positive-imaginary eigenvalue of a pair first); the Cartan matrix used for D_2 is that of A_2
         (math_utils/so_root_systems.py:21-25 with rank 2);
  n = 5: NOT invariant under swapping the two arguments -> the kernel value depends on the order in which the
         eigen-solver happens to deflate the two conjugate pairs; there is no function of (R1, R2) to be at parity with.
Usage: python -m oracle.probe_so_n_symmetry 4|5|6
"""
import torch, warnings, itertools, sys
warnings.simplefilter("ignore")
torch.set_default_dtype(torch.float64)
from oracle import reference_loader as rl
kso = rl.load("BoManifolds.kernel_utils.kernels_so")
n=int(sys.argv[1])
with rl.quiet():
    k = kso.SORiemannianMaternKernel(dim=n, nu=2.5)
k.lengthscale=0.5
r=n//2
torch.manual_seed(0)
th=torch.rand(r)*3.1
def ev(angles):
    xs=[torch.exp(torch.complex(torch.zeros(1), a.reshape(1))) for a in angles]
    return k.series_approx(*xs, k.lengthscale, k.nu)
base=ev(th)
print("base", base)
for perm in itertools.permutations(range(r)):
    print("perm",perm, (ev(th[list(perm)])-base).abs().max().item())
for signs in itertools.product([1,-1],repeat=r):
    v=ev(th*torch.tensor(signs,dtype=torch.float64))
    print("signs",signs,(v.real-base.real).abs().max().item(), v.imag.abs().max().item())

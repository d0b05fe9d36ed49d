"""TEST INFRASTRUCTURE ONLY - golden vectors for SO(n), n in {2, 4, 5, 6}, from the UNMODIFIED reference.

Run in the dev container:  ``python -m oracle.make_golden_son``

The reference's value for n >= 4 depends on the order in which torch.linalg.eigvals lists the conjugate pairs of
R1 R2^T (kernels_so.py:329-346).  Besides K this script therefore records, per pair, the angles
the reference actually substituted (its pairing loop re-run on the same eigvals call) and the mask ``canonical`` of
the pairs where those angles are all positive and in DESCENDING order - the order csrc/son_gram.cu fixes.  Parity
tests compare on the canonical pairs; the gradient with respect to raw_lengthscale is taken of sum(K * G) with G
zeroed outside the mask.
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch

from oracle import reference_loader as rl

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rotations(cnt, n, gen):
    """pymanopt SpecialOrthogonalGroup.rand recipe: QR of a Gaussian matrix, sign and determinant fixed."""
    q, r = torch.linalg.qr(torch.randn(cnt, n, n, generator=gen))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)
    q[:, :, 0] = q[:, :, 0] * torch.sign(torch.linalg.det(q)).unsqueeze(-1)
    return q.reshape(cnt, n * n)


def reference_angles(x1, x2, n):
    """The arguments the reference feeds to its polynomial: first member of each conjugate pair, LAPACK order."""
    m, k = x1.shape[0], x2.shape[0]
    a = x1.reshape(m, 1, n, n).expand(m, k, n, n).reshape(-1, n, n)
    b = x2.reshape(1, k, n, n).expand(m, k, n, n).reshape(-1, n, n)
    ev = torch.linalg.eigvals(torch.bmm(a, b.transpose(-1, -2)))
    ang = np.zeros((m * k, n // 2))
    for p in range(ev.shape[0]):
        lst = list(ev[p])
        chosen = []
        while lst:
            e = lst.pop(0)
            ec = torch.conj(e)
            for i, other in enumerate(lst):
                if torch.abs(ec - other) < 1e-10:
                    chosen.append(e)
                    del lst[i]
                    break
        assert len(chosen) == n // 2
        ang[p] = [float(torch.angle(e)) for e in chosen]
    return ang.reshape(m, k, n // 2)


def main():
    warnings.simplefilter("ignore")
    torch.set_default_dtype(torch.float64)
    kso = rl.load("BoManifolds.kernel_utils.kernels_so")
    torch.set_default_dtype(torch.float64)
    gen = torch.Generator().manual_seed(20261020)
    for n in (2, 4, 5, 6):
        x1, x2 = rotations(9, n, gen), rotations(11, n, gen)
        ang = reference_angles(x1, x2, n)
        canonical = np.all(ang > 0, axis=-1) & np.all(np.diff(ang, axis=-1) <= 0, axis=-1)
        g = torch.randn(9, 11, generator=gen) * torch.from_numpy(canonical)
        for name, make in (("matern", lambda: kso.SORiemannianMaternKernel(dim=n, nu=2.5)),
                           ("gaussian", lambda: kso.SORiemannianGaussianKernel(dim=n))):
            with rl.quiet():
                k = make()
            k.lengthscale = 0.5
            k.raw_lengthscale.requires_grad_(True)
            kk = k.forward(x1, x2)
            kk = kk.real if torch.is_complex(kk) else kk
            (gl,) = torch.autograd.grad((kk * g).sum(), [k.raw_lengthscale])
            out = dict(x1=x1.numpy(), x2=x2.numpy(), K=kk.detach().numpy(), angles=ang, canonical=canonical, G=g.numpy(),
                       g_raw_lengthscale=gl.numpy(), lengthscale=0.5, nu=2.5, dim=n)
            np.savez_compressed(os.path.join(OUT, "son_%s_SO%d.npz" % (name, n)), **out)
            print("son_%s_SO%d  K%s canonical pairs %d / %d  max|K| %.4f" % (name, n, tuple(kk.shape), canonical.sum(),
                                                                          canonical.size, float(kk.abs().max())))


if __name__ == "__main__":
    main()

"""Who owns the posterior error at the headline size?  CUDA path and float64 oracle against the long-double truth leg
(oracle/truth_ld.c) at n = 5000, S^10, noise 1e-2 / 1e-4 / 1e-6, incl. candidates 1e-3 away from training points.
Run on the GPU box: python tools/probe_a19_parity.py > gpurun_out/a19_parity.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from materngabo_b200 import kernels_sphere  # noqa: E402
from materngabo_b200.posterior import ExactGPPosterior  # noqa: E402
from oracle import restated as R, truth  # noqa: E402

dev = "cuda:0"
n = int(os.environ.get("N", 5000))
m, mt = 4096, 512
gen = torch.Generator().manual_seed(5)
xt = torch.nn.functional.normalize(torch.randn(n, 11, generator=gen), dim=-1)
xc = torch.nn.functional.normalize(torch.randn(m, 11, generator=gen), dim=-1)
xc[:256] = torch.nn.functional.normalize(xt[:256] + 1e-3 * torch.randn(256, 11, generator=gen), dim=-1)
xc[256:384] = xt[256:384]
y = torch.sin(3.0 * xt[:, 0]) + 0.5 * xt[:, 1] * xt[:, 2]
sel = torch.cat([torch.arange(0, 384), torch.arange(m - (mt - 384), m)])
out = []
for noise in (1e-2, 1e-4, 1e-6):
    k = kernels_sphere.SphereRiemannianMaternKernel(dim=10, nu=2.5)
    k.lengthscale = 0.5
    t0 = time.time()
    gp = ExactGPPosterior(k, xt.to(dev), y.to(dev), outputscale=1.0, noise=noise).fit()
    torch.cuda.synchronize()
    t_fit = time.time() - t0
    mg, vg = gp.posterior(xc.to(dev))
    mg, vg = mg.cpu().numpy(), vg.cpu().numpy()
    t0 = time.time()
    mtr, vtr = truth.sphere_matern_posterior(xt.numpy(), y.numpy(), xc[sel].numpy(), 10, 2.5, 0.5, 1.0, noise)
    t_truth = time.time() - t0
    kfn = lambda a, b: R.sphere_matern(a, b, 10, 2.5, 0.5)  # noqa: E731
    L, alpha = R.gp_fit(kfn, xt, y, 1.0, noise, 0.0)
    mo, vo = R.gp_predict(kfn, xt, L, alpha, xc, outputscale=1.0, mean_const=0.0, chunk=1024)
    mo, vo = mo.numpy(), vo.numpy()
    s = sel.numpy()
    rec = {"n": n, "noise": noise, "fit_s": t_fit, "truth_s": t_truth,
           "max_abs_mean_truth": float(np.abs(mtr).max()), "max_var_truth": float(vtr.max()), "min_var_truth": float(vtr.min()),
           "gpu_vs_truth_mean_abs": float(np.abs(mg[s] - mtr).max()), "oracle_vs_truth_mean_abs": float(np.abs(mo[s] - mtr).max()),
           "gpu_vs_truth_var_abs": float(np.abs(vg[s] - vtr).max()), "oracle_vs_truth_var_abs": float(np.abs(vo[s] - vtr).max()),
           "gpu_vs_oracle_mean_abs_all": float(np.abs(mg - mo).max()), "gpu_vs_oracle_var_abs_all": float(np.abs(vg - vo).max()),
           "gpu_vs_truth_var_abs_near": float(np.abs(vg[s][:384] - vtr[:384]).max()),
           "gpu_vs_truth_var_abs_far": float(np.abs(vg[s][384:] - vtr[384:]).max())}
    out.append(rec)
    print(json.dumps(rec), flush=True)

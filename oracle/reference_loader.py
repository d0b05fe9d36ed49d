"""TEST INFRASTRUCTURE ONLY - imports the UNMODIFIED reference from /root/reference on CPU.

Used in the dev container to (a) validate ``oracle/restated.py`` and (b) generate the golden
vectors under ``tests/golden`` (``oracle/make_golden.py``).  /root/reference does not exist on
the GPU box: nothing that runs there may import this module (it raises if the tree is absent).

Three shims are needed to import the reference on a CPU-only box without gpytorch
(SURVEY.md section 8c):
  1. ``torch.cuda.current_device`` is called unconditionally at import
     (BoManifolds/kernel_utils/kernels_sphere.py:20, kernels_hyperbolic.py:14, kernels_spd.py:20,
     Riemannian_utils/spd_utils_torch.py:13) -> patched to return ``'cpu'`` while importing;
  2. ``gpytorch`` -> ``materngabo_b200.gpytorch_compat`` registered as ``sys.modules['gpytorch']``;
  3. ``more_itertools`` / ``pymanopt.manifolds`` (kernels_so.py:12,18) -> tiny stand-ins.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import itertools
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MGB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "BoManifolds"))


def _install_shims():
    from materngabo_b200 import gpytorch_compat as gc

    if "gpytorch" not in sys.modules:
        g = types.ModuleType("gpytorch")
        g.kernels = types.ModuleType("gpytorch.kernels")
        g.constraints = types.ModuleType("gpytorch.constraints")
        for name in ("Kernel", "ScaleKernel", "ProductKernel"):
            setattr(g.kernels, name, getattr(gc, name))
        for name in ("Positive", "GreaterThan", "Interval"):
            setattr(g.constraints, name, getattr(gc, name))
        g.Module = gc.Module
        sys.modules["gpytorch"] = g
        sys.modules["gpytorch.kernels"] = g.kernels
        sys.modules["gpytorch.constraints"] = g.constraints
    if "more_itertools" not in sys.modules:
        m = types.ModuleType("more_itertools")

        def padded(iterable, fillvalue=None, n=None):
            it = iter(iterable)
            count = 0
            for x in it:
                count += 1
                yield x
            while n is not None and count < n:
                count += 1
                yield fillvalue

        def distinct_permutations(iterable):
            return sorted(set(itertools.permutations(iterable)))

        m.padded = padded
        m.distinct_permutations = distinct_permutations
        sys.modules["more_itertools"] = m
    if "pymanopt" not in sys.modules:
        p = types.ModuleType("pymanopt")
        p.manifolds = types.ModuleType("pymanopt.manifolds")
        sys.modules["pymanopt"] = p
        sys.modules["pymanopt.manifolds"] = p.manifolds


_cache = {}


def load(module: str):
    """Import e.g. ``BoManifolds.kernel_utils.kernels_sphere`` from the reference tree."""
    if module in _cache:
        return _cache[module]
    if not available():
        raise RuntimeError("reference tree not present at %s (only exists in the dev container)" % REFERENCE_ROOT)
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    saved_dtype = torch.get_default_dtype()
    real = torch.cuda.current_device
    torch.cuda.current_device = lambda: "cpu"
    try:
        mod = importlib.import_module(module)
    finally:
        torch.cuda.current_device = real
        torch.set_default_dtype(saved_dtype)  # reference modules force float32 at import
    _cache[module] = mod
    return mod


@contextlib.contextmanager
def quiet():
    """The SO(n) constructors print every weight (kernels_so.py:279-286)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield

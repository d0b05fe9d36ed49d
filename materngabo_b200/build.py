"""Builds materngabo_b200/lib/libmgb.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  ``python -m
materngabo_b200.build`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmgb.so")
SOURCES = ["runtime.cu", "series_gram.cu", "son_gram.cu", "spd_eigh.cu", "quad_gram.cu", "quad_profile.cu", "posterior.cu", "mll.cu", "dense_spd.cu", "acquisition.cu", "host_api.cu",
           "int8_posterior.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _cutlass_flags():
    """Include flags for the CUTLASS / CuTe header tree vendored in the image (flashinfer wheel), used by
    int8_posterior.cu only (a library int8 GEMM for the opt-in INT8 path).  Without the headers that file builds its
    stub: mgb_int8_available() == 0 and the FP64 tensor-core path stays the only one."""
    roots = []
    for base in sys.path:
        roots.append(os.path.join(base, "flashinfer", "data", "cutlass"))
    roots.append(os.environ.get("MGB_CUTLASS_DIR", ""))
    for root in roots:
        if root and os.path.exists(os.path.join(root, "include", "cutlass", "gemm", "collective", "builders", "sm100_umma_builder.inl")):
            return ["-I" + os.path.join(root, "include"), "-DMGB_HAVE_CUTLASS=1", "-diag-suppress", "20012", "--expt-extended-lambda"]
    return []


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libmgb.so cannot be built (there is no CPU fallback)")
    return exe


STAMP = os.path.join(LIBDIR, "libmgb.stamp")


def _source_digest() -> str:
    """Content hash of every source, the header and the compiler flags: a tree copied to another machine (gpurun
    snapshot: fresh mtimes) must not rebuild a library that matches its sources."""
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "mgb.h")]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + SOURCES + _cutlass_flags()).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    # one builder at a time (several ranks of a torchrun launch may import the package at once)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        extra = _cutlass_flags() if src == "int8_posterior.cu" else []
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)                         # atomic: a concurrent loader sees the old or the new library
    with open(STAMP, "w") as fh:
        fh.write(_source_digest() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

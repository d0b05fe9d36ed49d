"""Builds materngabo_b200/lib/libmgb.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  ``python -m
materngabo_b200.build`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import fcntl
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmgb.so")
SOURCES = ["runtime.cu", "series_gram.cu", "son_gram.cu", "spd_eigh.cu", "quad_gram.cu", "quad_profile.cu", "posterior.cu", "mll.cu", "dense_spd.cu", "acquisition.cu", "host_api.cu",
           "int8_posterior.cu", "int8_fused.cu", "series_i8.cu"]
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
INFO = os.path.join(LIBDIR, "build_info.json")     # how the library on disk came to be (compiled here / reused)
LAST_MODE = None                                   # "compiled" or "reused": what the last build() call in this process did


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


def build_info() -> dict:
    """{"compiled_at", "seconds", "nvcc", "digest", "sources", "int8_cutlass"} of the library on disk + "last_call"."""
    try:
        with open(INFO) as fh:
            info = json.load(fh)
    except Exception:
        info = {}
    info["last_call"] = LAST_MODE
    return info


def build(force: bool = False, verbose: bool = False) -> str:
    global LAST_MODE
    if not force and not needs_build():
        LAST_MODE = "reused"
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    # one builder at a time (several ranks of a torchrun launch may import the package at once)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                LAST_MODE = "reused"
                return LIB
            LAST_MODE = "compiled"
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    t0 = time.time()
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
    try:
        version = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    except Exception:
        version = "unknown"
    with open(INFO, "w") as fh:
        json.dump({"compiled_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "seconds": round(time.time() - t0, 1),
                   "nvcc": version, "flags": NVCC_FLAGS, "digest": _source_digest(), "sources": SOURCES,
                   "int8_cutlass": bool(_cutlass_flags())}, fh)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(json.dumps(build_info()))

"""Prints the FP64 micro-benchmarks of libmgb (DFMA chains, DMMA chains, both interleaved) as one JSON line."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from materngabo_b200 import _lib  # noqa: E402

lib = _lib.load()
out = {}
for kind, name in ((0, "dfma_tflops"), (1, "dmma_tflops"), (2, "dfma_plus_dmma_interleaved_tflops")):
    tf, ms = C.c_double(), C.c_double()
    _lib.check(lib.mgb_fp64_peak(kind, 3, C.byref(tf), C.byref(ms)), "mgb_fp64_peak")
    out[name] = tf.value
print(json.dumps(out))

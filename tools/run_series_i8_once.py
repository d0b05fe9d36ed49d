"""One launch of the INT8 tensor-core series Gram kernel (for ncu): python tools/run_series_i8_once.py [m] [n] [w] [path]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from materngabo_b200 import _ops  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
w = int(sys.argv[3]) if len(sys.argv) > 3 else 11
os.environ["MGB_GRAM_PATH"] = sys.argv[4] if len(sys.argv) > 4 else "int8"
dev = torch.device("cuda")
g = torch.Generator().manual_seed(1)
x1 = torch.nn.functional.normalize(torch.randn(m, w, generator=g, dtype=torch.float64), dim=-1).to(dev)
x2 = torch.nn.functional.normalize(torch.randn(n, w, generator=g, dtype=torch.float64), dim=-1).to(dev)
c = torch.tensor([0.11, 0.2, 0.15, 0.1, 0.09, 0.08, 0.07, 0.06, 0.05, 0.04], dtype=torch.float64, device=dev)
spec = _ops.SeriesSpec(1, w, 0, 1.0, 0.0, -1.0 + 1e-15, 1.0 - 1e-15)
for _ in range(4):
    out = _ops._series_fwd_raw(spec, x1, x2, c)
torch.cuda.synchronize()
print("ok", float(out[0, 0]))

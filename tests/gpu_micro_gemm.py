"""GEMM-only timings (see gpu_micro.py)"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "clustered-low-rank-sdp-solver_b200"))
sys.path.insert(0, os.path.dirname(__file__))
import numpy as np  # noqa: E402
from clrsdp import solver  # noqa: E402
from clrsdp.wire import MpArray  # noqa: E402
from gpu_micro import table  # noqa: E402

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = solver.product_handle(prec, 0)
shapes = [(64, 64, 64, 64), (128, 32, 32, 32), (1, 256, 256, 8192), (64, 256, 128, 128), (64, 128, 128, 64)]
if len(sys.argv) > 2 and sys.argv[2] == "cfg5":   # the products of one GPU's share of BASELINE config 5 (block 128, K = 256, n_y = 1024)
    shapes = [(64, 128, 128, 128), (64, 256, 256, 128), (16, 1024, 256, 256), (1, 1024, 1024, 16384)]
for batch, M, N, K in shapes:
    nrng = np.random.default_rng(1)
    A = MpArray.from_double(nrng.uniform(-1, 1, size=batch * M * K), h.nlimb)
    B = MpArray.from_double(nrng.uniform(-1, 1, size=batch * K * N), h.nlimb)
    h.op_gemm(batch, M, N, K, A, B)
    h.profile_reset(True)
    for _ in range(3):
        h.op_gemm(batch, M, N, K, A, B)
    table(h, f"gemm batch={batch} M={M} N={N} K={K}", 3)
    h.profile_reset(False)

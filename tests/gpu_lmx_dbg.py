"""Measuring aid: lambda_min on 128 symmetric n x n matrices (build with CLRSDP_EXTRA_NVCC_FLAGS=-DCLRSDP_LMX_TIMING to
get the kernel's per-phase clock table).   python tests/gpu_lmx_dbg.py [prec] [n]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "clustered-low-rank-sdp-solver_b200"))
import numpy as np
from clrsdp import solver
from clrsdp.wire import MpArray
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
h = solver.product_handle(prec, 0)
rng = np.random.default_rng(n)
mats = []
for _ in range(128):
    G = rng.uniform(-1, 1, size=(n, n))
    mats.append((G + G.T) / 2)
A = MpArray.from_double(np.array(mats).reshape(-1), h.nlimb)
h.op_lambda_min(128, n, A)

"""Measuring aid: one 32 x 32 panel factorisation with CLRSDP_PANEL_DEBUG=1 (per-step clock table printed by the kernel)."""
import os, random, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "clustered-low-rank-sdp-solver_b200"))
sys.path.insert(0, os.path.dirname(__file__))
from clrsdp import solver
from gpu_common import spd_batch
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
h = solver.product_handle(prec, 0)
A = spd_batch(random.Random(1), batch, 32, h.nlimb)
h.op_cholesky(batch, 32, A)

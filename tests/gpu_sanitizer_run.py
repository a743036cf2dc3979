"""Measuring aid: a small pass over the kernels for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tests/gpu_sanitizer_run.py
two IPM iterations of a small general-structure problem (direct launches) and one blocked factorisation with inverse."""
import os, random, sys
os.environ.setdefault("CLRSDP_GRAPH", "0")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "clustered-low-rank-sdp-solver_b200"))
sys.path.insert(0, os.path.dirname(__file__))
from clrsdp import instances, solver
from gpu_common import spd_batch
prec = 256
cons, b, _ = instances.synthetic_clustered_sdp(J=3, delta=8, K=12, n_y=7, prec=prec)
bi = solver.get_block_info(cons)
h = solver.product_handle(prec, 0)
solver.load_problem(h, cons, b, bi)
h.set_params(solver.real_params(h.nlimb))
h.init_point()
h.prepare()
for _ in range(2):
    r = h.iterate()
    assert r.status == 0
A = spd_batch(random.Random(3), 2, 70, h.nlimb)
h.op_cholesky(2, 70, A)
h.op_signed_factor(2, 70, A)
print("sanitizer pass done: mu", r.mu)

"""Regenerates tests/golden/sphere_packing_512.json: examples/SpherePacking.jl (n = 3, radii 1 and sqrt(2)-1) at
degree d = 8 (BASELINE config 1), 12 and 16, solved by the CPU oracle at the example's own precision, 512 bits
(ex:29-31, 117-119). Run from the repo root: python tests/golden/make_sphere_golden.py  (about 3 minutes)."""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "clustered-low-rank-sdp-solver_b200"))
import mpmath  # noqa: E402
from clrsdp import instances, solver  # noqa: E402
from oracle.ref import oracle_handle  # noqa: E402

prec = 512
solver.set_precision(prec)
cases = []
for d in (8, 12, 16):
    cons, b, _ = instances.sphere_packing_2point(n=3, d=d, prec=prec)
    bi = solver.get_block_info(cons)
    t = time.time()
    out, rows = solver.solverank1sdp(cons, b, bi, handle=oracle_handle(prec, 8), verbose=False, return_info=True,
                                     omega_p=100, omega_d=100)
    with mpmath.workprec(prec):
        cases.append({"d": d, "prec": prec, "iterations": len(rows), "terminate": rows[-1].terminate,
                      "primal_obj": mpmath.nstr(out[8], 60), "dual_obj": mpmath.nstr(out[9], 60),
                      "gap": mpmath.nstr(out[7], 10), "oracle_seconds": time.time() - t})
json.dump({"generator": "tests/golden/make_sphere_golden.py", "cases": cases},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sphere_packing_512.json"), "w"), indent=1)

"""Generates tests/golden/*.json from the CPU oracle (the reference itself cannot run here: no Julia, no Arb).
    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "clustered-low-rank-sdp-solver_b200"))
import mpmath  # noqa: E402

from clrsdp import instances, solver  # noqa: E402
from oracle.ref import oracle_handle  # noqa: E402


def synthetic(name, **inst):
    cons, b, _ = instances.synthetic_clustered_sdp(**inst)
    bi = solver.get_block_info(cons)
    out, rows = solver.solverank1sdp(cons, b, bi, handle=oracle_handle(inst["prec"], 2), verbose=False, return_info=True)
    with mpmath.workprec(inst["prec"]):
        g = dict(instance=inst, iterations=len(rows), primal_obj=mpmath.nstr(out[8], 70), dual_obj=mpmath.nstr(out[9], 70),
                 gap=mpmath.nstr(out[7], 20),
                 rows=[{k: getattr(r, k) for k in ("iter", "mu", "alpha_p", "alpha_d", "beta_c", "p_obj", "d_obj", "gap")} for r in rows])
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(g, f, indent=1)
    print(name, len(rows), g["primal_obj"][:30])


if __name__ == "__main__":
    synthetic("synthetic_J3_d6_K10_ny5_p256.json", J=3, delta=6, K=10, n_y=5, prec=256, seed=20261018)
    synthetic("synthetic_J4_d8_K12_ny6_p384.json", J=4, delta=8, K=12, n_y=6, prec=384, seed=3)

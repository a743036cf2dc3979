"""Regenerates tests/golden/sphere_packing_lowprec.json: the sphere-packing instance of BASELINE config 1
(examples/SpherePacking.jl, n = 3, radii 1 and sqrt(2)-1) BELOW the example's own 512 bits, where the Schur complements
are singular to working precision (cond(S') > 2^240 near the optimum at d = 8).

What the file records, per (d, prec):
  * "lu": the oracle proper (the reference's algorithm: pivoted LU of S_j and Q, MPMP.jl:1436,1501) for several thread
    counts and for its block fixed-point product mode. Its summation order (the Q product is chunked by the number of
    threads, :1467-1495) and its product mode change rounding errors only - and they change the iteration count: the LU
    trajectory at these precisions is dominated by rounding noise, so "the reference's iteration count" is a range.
  * "ldl": the oracle with CLRSDP_REF_FACTOR=ldl - the factorisation of the GPU path (equilibrated signed Cholesky) in
    MPFR arithmetic. This is what the GPU solve is compared with iteration for iteration.
Run from the repo root: python tests/golden/make_sphere_lowprec_golden.py   (about 10 minutes on 8 cores)."""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")

WORKER = r'''
import json, os, sys, time
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "clustered-low-rank-sdp-solver_b200"))
import mpmath
from clrsdp import instances, solver
from oracle.ref import oracle_handle
d, prec, nt = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
solver.set_precision(prec)
cons, b, _ = instances.sphere_packing_2point(n=3, d=d, prec=prec)
bi = solver.get_block_info(cons)
t = time.time()
try:
    out, rows = solver.solverank1sdp(cons, b, bi, handle=oracle_handle(prec, nt), verbose=False, return_info=True,
                                     omega_p=100, omega_d=100)
    with mpmath.workprec(prec):
        print(json.dumps({"iterations": len(rows), "terminate": rows[-1].terminate, "primal_obj": mpmath.nstr(out[8], 40),
                          "dual_obj": mpmath.nstr(out[9], 40), "gap": mpmath.nstr(out[7], 10),
                          "max_p_err": max(r.p_err for r in rows), "seconds": time.time() - t}))
except Exception as e:
    print(json.dumps({"error": str(e), "seconds": time.time() - t}))
'''


def run(d, prec, nt, env):
    e = dict(os.environ)
    e.pop("CLRSDP_REF_FACTOR", None)
    e.pop("CLRSDP_REF_GEMM", None)
    e.update(env)
    out = subprocess.run([sys.executable, "-c", WORKER, ROOT, str(d), str(prec), str(nt)], env=e, capture_output=True, text=True)
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    rec.update({"threads": nt, **{k.lower(): v for k, v in env.items()}})
    return rec


JOBS = []
for d, prec, nts in ((8, 256, (1, 3, 5, 8)), (8, 384, (1, 4, 8)), (12, 256, (2,))):
    for nt in nts:
        JOBS.append((d, prec, "lu", nt, {}))
    if d == 8:
        JOBS.append((d, prec, "lu", 8, {"CLRSDP_REF_GEMM": "fixed"}))
    JOBS.append((d, prec, "ldl", 2, {"CLRSDP_REF_FACTOR": "ldl"}))
    JOBS.append((d, prec, "ldl", 5, {"CLRSDP_REF_FACTOR": "ldl"}))

if __name__ == "__main__":
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    with ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) // 2)) as ex:
        res = list(ex.map(lambda j: run(j[0], j[1], j[3], j[4]), JOBS))
    cases = {}
    for (d, prec, kind, nt, env), r in zip(JOBS, res):
        cases.setdefault(f"d{d}_p{prec}", {"d": d, "prec": prec, "lu": [], "ldl": []})[kind].append(r)
    json.dump({"generator": "tests/golden/make_sphere_lowprec_golden.py", "cases": list(cases.values())},
              open(os.path.join(HERE, "sphere_packing_lowprec.json"), "w"), indent=1)
    for c in cases.values():
        print(c["d"], c["prec"], "lu:", [(r.get("iterations"), r.get("terminate")) for r in c["lu"]],
              "ldl:", [(r.get("iterations"), r.get("terminate")) for r in c["ldl"]])

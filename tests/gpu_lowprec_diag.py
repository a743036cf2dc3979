"""Diagnostic driver (not a test): GPU solve of sphere packing at (d, prec) pairs, rows dumped as JSON lines.
  python tests/gpu_lowprec_diag.py 8:256 12:256 16:256 > gpurun_out/lowprec_rows.jsonl"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "clustered-low-rank-sdp-solver_b200"))
from clrsdp import instances, solver
from clrsdp.capi import ClrsdpError

use_oracle = os.environ.get("ORACLE", "") != ""
for spec in sys.argv[1:]:
    d, prec = map(int, spec.split(":"))
    solver.set_precision(prec)
    cons, b, _ = instances.sphere_packing_2point(n=3, d=d, prec=prec)
    bi = solver.get_block_info(cons)
    h = None
    if use_oracle:
        from oracle.ref import oracle_handle
        h = oracle_handle(prec, os.cpu_count())
    t0 = time.time()
    rows = []
    try:
        out, rows = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True, omega_p=100, omega_d=100, handle=h,
                                         maxiterations=int(os.environ.get("MAXIT", "500")))
        res = dict(d=d, prec=prec, terminate=rows[-1].terminate, iterations=len(rows), p_obj=str(out[8]), d_obj=str(out[9]),
                   gap=float(out[7]), wall=time.time() - t0)
    except ClrsdpError as e:
        res = dict(d=d, prec=prec, error=str(e), wall=time.time() - t0)
    print(json.dumps(dict(result=res, rows=[dict(it=r.iter, mu=r.mu, p_obj=r.p_obj, d_obj=r.d_obj, gap=r.gap, P_err=r.P_err, p_err=r.p_err,
                                                  d_err=r.d_err, ap=r.alpha_p, ad=r.alpha_d, beta=r.beta_c) for r in rows])), flush=True)

"""Diagnostic driver (not a test): full solve of a synthetic instance on the GPU and on the oracle, rows side by side.
  python tests/gpu_solve_diag.py J delta K n_y prec"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "clustered-low-rank-sdp-solver_b200"))
from clrsdp import instances, solver
from oracle.ref import oracle_handle
J, delta, K, ny, prec = [int(a) for a in sys.argv[1:6]]
cons, b, _ = instances.synthetic_clustered_sdp(J=J, delta=delta, K=K, n_y=ny, prec=prec, seed=20261018)
bi = solver.get_block_info(cons)
solver.set_precision(prec)
t0 = time.time()
og, rg = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True)
tg = time.time() - t0
if os.environ.get("NO_ORACLE"):
    print(f"gpu: {len(rg)} iterations terminate {rg[-1].terminate} in {tg:.1f}s")
    for a in rg[::6] + [rg[-1]]:
        print(f"{a.iter:3d} mu {a.mu:.3e} | ap {a.alpha_p:.6f} | P {a.P_err:.1e} | p {a.p_err:.1e} | d {a.d_err:.1e}")
    sys.exit(0)
os.environ["CLRSDP_REF_GEMM"] = "fixed"
ho = oracle_handle(prec, os.cpu_count())
os.environ.pop("CLRSDP_REF_GEMM", None)
t0 = time.time()
oo, ro = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True, handle=ho)
print(f"gpu: {len(rg)} iterations terminate {rg[-1].terminate} in {tg:.1f}s; oracle: {len(ro)} iterations terminate {ro[-1].terminate} in {time.time() - t0:.1f}s")
print("p_obj", og[8], oo[8]); print("d_obj", og[9], oo[9])
for a, o in zip(rg, ro):
    print(f"{a.iter:3d} mu {a.mu:.3e} {o.mu:.3e} | ap {a.alpha_p:.6f} {o.alpha_p:.6f} | ad {a.alpha_d:.6f} {o.alpha_d:.6f} | P {a.P_err:.1e} {o.P_err:.1e} | p {a.p_err:.1e} {o.p_err:.1e} | d {a.d_err:.1e} {o.d_err:.1e}")

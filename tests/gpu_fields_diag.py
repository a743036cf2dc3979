"""Diagnostic driver (not a test): one or two iterations GPU vs oracle vs oracle at p+64 bits on a synthetic instance,
matching bits of every field printed (no assertions).
  python tests/gpu_fields_diag.py J delta K n_y prec [iters]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "clustered-low-rank-sdp-solver_b200"))
sys.path.insert(0, os.path.dirname(__file__))
from clrsdp import instances, solver
from clrsdp.wire import rel_err_bits
from oracle.ref import oracle_handle
from test_gpu_solver import widen_problem

J, delta, K, ny, prec = [int(a) for a in sys.argv[1:6]]
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 1
cons, b, _ = instances.synthetic_clustered_sdp(J=J, delta=delta, K=K, n_y=ny, prec=prec, seed=int(os.environ.get("SEED", "20261018")))
bi = solver.get_block_info(cons)
wc, wb = widen_problem(cons, b, 2)
os.environ["CLRSDP_REF_GEMM"] = "fixed"
hg, ho, ht = solver.product_handle(prec), oracle_handle(prec, os.cpu_count()), oracle_handle(prec + 64, os.cpu_count())
os.environ.pop("CLRSDP_REF_GEMM", None)
for h, (c_, b_) in ((hg, (cons, b)), (ho, (cons, b)), (ht, (wc, wb))):
    solver.load_problem(h, c_, b_, bi); h.set_params(solver.real_params(h.nlimb)); h.init_point(); h.prepare()
for it in range(iters):
    t0 = time.time()
    rg, ro, rt = hg.iterate(), ho.iterate(), ht.iterate()
    print(f"iter {it + 1} ({time.time() - t0:.0f}s): status {rg.status} {ro.status}; mu {rg.mu:.6e} {ro.mu:.6e}; alpha {rg.alpha_p:.6f} {ro.alpha_p:.6f} {rg.alpha_d:.6f} {ro.alpha_d:.6f}", flush=True)
    for name in ("d", "p", "dx_pred", "dy_pred", "dx", "dy", "x", "y", "Q"):
        t = ht.fetch(name)
        print(f"  {name}: gpu {rel_err_bits(hg.fetch(name), t):.0f} oracle {rel_err_bits(ho.fetch(name), t):.0f}", end="")
    print()
    for name in ("Xinv", "Px", "Py", "P", "Z", "dX_pred", "dY_pred", "dX", "dY", "X", "Y"):
        g = min(rel_err_bits(hg.fetch(name, j, 0), ht.fetch(name, j, 0)) for j in range(min(bi.J, 2)))
        o = min(rel_err_bits(ho.fetch(name, j, 0), ht.fetch(name, j, 0)) for j in range(min(bi.J, 2)))
        print(f"  {name}: gpu {g:.0f} oracle {o:.0f}", end="")
    print()
    print("  S:", [(round(min(999, rel_err_bits(hg.fetch("S", j), ht.fetch("S", j)))), round(min(999, rel_err_bits(ho.fetch("S", j), ht.fetch("S", j))))) for j in range(min(bi.J, 4))], flush=True)
    W = hg.fetch("W"); Wt = ht.fetch("W") if False else None

# ---- the factorisation by itself: S_0 of the last iteration through op_signed_factor (GPU) and op_cholesky (oracle) ----
if os.environ.get("FACTOR_CHECK", "1") == "1":
    import mpmath
    import numpy as np
    n = bi.dim_S[0]
    S = ho.fetch("S", 0)
    M, sg = hg.op_signed_factor(1, n, S)
    try:
        Lo, Lio = ho.op_cholesky(1, n, S)
    except Exception as e:
        Lio = None
        print("oracle cholesky failed:", e)
    with mpmath.workprec(prec + 128):
        Sm = mpmath.matrix(n, n)
        sv = S.to_mpfs()
        for i in range(n):
            for j in range(n):
                Sm[i, j] = sv[i * n + j]
        def resid(Mi, signs):
            mv = Mi.reshape(n * n).to_mpfs()
            Mm = mpmath.matrix(n, n)
            for i in range(n):
                for j in range(n):
                    Mm[i, j] = mv[i * n + j]
            Sg = mpmath.diag([int(s) for s in signs])
            Sinv = Mm.T * Sg * Mm
            R = Sinv * Sm - mpmath.eye(n)
            return float(mpmath.log(max(abs(R[i, j]) for i in range(n) for j in range(n)), 2)), Sinv
        rg_, Sinv_g = resid(M, sg[0])
        print(f"factor check n={n}: negative pivots {int((sg < 0).sum())}; log2 max|M^T Sigma M S - I|: gpu {rg_:.1f}", end="")
        if Lio is not None:
            ro_, Sinv_o = resid(Lio, [1] * n)
            d = max(abs(Sinv_g[i, j] - Sinv_o[i, j]) for i in range(n) for j in range(n)) / max(abs(Sinv_o[i, j]) for i in range(n) for j in range(n))
            print(f" oracle {ro_:.1f}; log2 |Sinv_gpu - Sinv_oracle| / max|Sinv| {float(mpmath.log(d, 2)):.1f}")
        else:
            print()

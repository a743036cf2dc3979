"""Diagnostic driver (not a test): the sphere-packing instances of BASELINE configs 1/2 on the GPU path.
  python tests/gpu_sphere_diag.py compare D ITERS   # GPU vs oracle, field by field, for ITERS iterations at degree D
  python tests/gpu_sphere_diag.py solve D           # GPU-only full solve at degree D (log rows, bound)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "clustered-low-rank-sdp-solver_b200"))
from clrsdp import instances, solver
from clrsdp.capi import ClrsdpError
from clrsdp.wire import rel_err_bits

mode, d = sys.argv[1], int(sys.argv[2])
prec = int(os.environ.get("PREC", "256"))
quiet = os.environ.get("QUIET", "0") == "1"
solver.set_precision(prec)
t0 = time.time()
cons, b, _ = instances.sphere_packing_2point(n=3, d=d, prec=prec)
bi = solver.get_block_info(cons)
print(f"d={d}: J={bi.J} n_y={bi.n_y} dimS={list(bi.dim_S)} generated in {time.time() - t0:.1f}s", flush=True)
if mode == "solve":
    t0 = time.time()
    try:
        out, rows = solver.solverank1sdp(cons, b, bi, verbose=True, return_info=True, omega_p=100, omega_d=100)
        print(f"terminate={rows[-1].terminate} iterations={len(rows)} bound={float(-out[8]):.12f} gap={float(out[7]):.3e} "
              f"wall={time.time() - t0:.1f}s", flush=True)
    except ClrsdpError as e:
        print("GPU solve failed:", e, flush=True)
else:
    from oracle.ref import oracle_handle
    iters = int(sys.argv[3])
    hg, ho = solver.product_handle(prec), oracle_handle(prec, os.cpu_count())
    for h in (hg, ho):
        solver.load_problem(h, cons, b, bi)
        h.set_params(solver.real_params(h.nlimb, omega_p=100, omega_d=100))
        h.init_point(); h.prepare()
    for it in range(iters):
        t0 = time.time(); ro = ho.iterate(); to = time.time() - t0
        try:
            rg = hg.iterate()
        except ClrsdpError as e:
            print("gpu iterate failed:", e, flush=True)
            rg = None
        if rg is not None:
            print(f"iter {it + 1} (oracle {to:.1f}s) mu {rg.mu:.6e} {ro.mu:.6e} alpha {rg.alpha_p:.6f} {ro.alpha_p:.6f} "
                  f"{rg.alpha_d:.6f} {ro.alpha_d:.6f}")
        for k in ("d", "p", "dx_pred", "dy_pred", "dx", "dy", "x", "y", "Q"):
            print(f"  {k}: {rel_err_bits(hg.fetch(k), ho.fetch(k)):.0f}", end="")
        print()
        if quiet:
            if rg is None or rg.terminate or ro.terminate:
                print("terminate", rg.terminate if rg else None, ro.terminate, "p_obj", hg.scalar("p_obj"), ho.scalar("p_obj"))
                break
            continue
        for name in ("Xinv", "Px", "Py", "P", "Z", "dX_pred", "dY_pred", "dX", "dY", "X", "Y"):
            w = min(rel_err_bits(hg.fetch(name, j, l), ho.fetch(name, j, l)) for j in range(bi.J) for l in range(bi.L[j]))
            print(f"  {name}: {w:.0f}", end="")
        print()
        print("  S:", [min(999, round(min(1e9, rel_err_bits(hg.fetch("S", j), ho.fetch("S", j))))) for j in range(bi.J)],
              flush=True)
        if rg is not None and rg.terminate:
            break
        if rg is None:
            break

"""Diagnostic run on a GPU box (not a pytest file): exercises every phase-level op and one IPM iteration
against exact arithmetic / the oracle and prints error magnitudes. Usage: python tests/gpu_diag.py"""
import os
import sys
import time
import traceback
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "clustered-low-rank-sdp-solver_b200"))
import random

from clrsdp import instances, solver
from clrsdp.wire import MpArray, rel_err_bits
from oracle.ref import oracle_handle


def rand_mp(rng, n, nlimb, erange=6):
    p = 32 * nlimb
    m = [(rng.getrandbits(p) | (1 << (p - 1))) * rng.choice([-1, 1]) for _ in range(n)]
    e = [rng.randint(-erange, erange) - p for _ in range(n)]
    return MpArray.from_ints(m, e, nlimb)


def section(name):
    print("\n==== " + name, flush=True)


def guarded(f):
    try:
        f()
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()


def diag_elementwise(h, ho, prec):
    rng = random.Random(1)
    a, b = rand_mp(rng, 500, h.nlimb), rand_mp(rng, 500, h.nlimb)
    for op in "+-*/s":
        aa = a
        if op == "s":
            aa = a.view()
            aa.sign = np.abs(a.sign)
        g = h.op_elementwise(op, aa, b)
        o = ho.op_elementwise(op, aa, b)
        same = int(np.sum((g.limb == o.limb).all(axis=0) & (g.exp == o.exp) & (g.sign == o.sign)))
        print(f"op {op}: identical to MPFR in {same}/500, rel err bits {rel_err_bits(g, o):.1f}")


def exact_planes(A, B, batch, M, N, K, rexp, cexp, S):
    """python big-int model of slice + plane products"""
    def digits(x: MpArray, i, rowexp):
        m, e = x.get_int(i)
        if m == 0:
            return [0] * S
        # F = trunc(x * 2^(8S-2-rowexp))
        sh = e + 8 * S - 2 - rowexp
        F = (abs(m) << sh) if sh >= 0 else (abs(m) >> (-sh))
        if m < 0:
            F = -F
        ds = []
        for _ in range(S):  # balanced digits, least significant first
            d = F & 0xFF
            if d >= 128:
                d -= 256
            ds.append(d)
            F = (F - d) >> 8
        assert F == 0, F
        return ds[::-1]
    planes = np.zeros((S, batch, M, N), dtype=object)
    for b in range(batch):
        Ad = np.array([[digits(A, (b * M + i) * K + k, int(rexp[b, i])) for k in range(K)] for i in range(M)], dtype=object)
        Bd = np.array([[digits(B, (b * K + k) * N + j, int(cexp[b, j])) for k in range(K)] for j in range(N)], dtype=object)
        for t in range(S):
            acc = np.zeros((M, N), dtype=object)
            for a in range(t + 1):
                acc = acc + Ad[:, :, a].dot(Bd[:, :, t - a].T)
            planes[t, b] = acc
    return planes


def diag_gemm(h, ho, prec):
    rng = random.Random(2)
    for (batch, M, N, K) in [(1, 8, 8, 8), (2, 20, 12, 40), (1, 130, 70, 64), (3, 64, 64, 64), (1, 16, 200, 130)]:
        A, B = rand_mp(rng, batch * M * K, h.nlimb), rand_mp(rng, batch * K * N, h.nlimb)
        # a zero row / zero entries
        A.sign[:K] = 0
        A.limb[:, :K] = 0
        t0 = time.time()
        planes, rexp, cexp = h.op_gemm_planes(batch, M, N, K, A, B)
        S = planes.shape[0]
        if M * N * K * batch <= 20000:
            ex = exact_planes(A, B, batch, M, N, K, rexp, cexp, S)
            bad = int(np.sum(planes.astype(object) != ex))
            print(f"planes {batch}x{M}x{N}x{K}: S={S} mismatches={bad} (of {planes.size})", flush=True)
            if bad:
                idx = np.argwhere(planes.astype(object) != ex)[:5]
                for t, b, i, j in idx:
                    print("   first bad", t, b, i, j, planes[t, b, i, j], ex[t, b, i, j])
        C = h.op_gemm(batch, M, N, K, A, B)
        Co = ho.op_gemm(batch, M, N, K, A, B)
        print(f"gemm {batch}x{M}x{N}x{K}: vs oracle rel-to-max err bits {rel_err_bits(C, Co):.1f}  ({time.time()-t0:.2f}s)",
              flush=True)


def spd_batch(rng, batch, n, nlimb):
    mats = []
    for b in range(batch):
        G = np.array([[rng.uniform(-1, 1) for _ in range(n)] for _ in range(n)])
        A = G @ G.T + n * np.eye(n) * 0.1
        A = (A + A.T) / 2
        mats.append(A)
    return MpArray.from_double(np.array(mats).reshape(-1), nlimb)


def diag_chol(h, ho, prec):
    rng = random.Random(3)
    for (batch, n) in [(2, 5), (3, 33), (2, 64), (1, 130)]:
        A = spd_batch(rng, batch, n, h.nlimb)
        L, Li = h.op_cholesky(batch, n, A)
        Lo, Lio = ho.op_cholesky(batch, n, A)
        print(f"chol {batch}x{n}: L err bits {rel_err_bits(L, Lo):.1f}  Linv err bits {rel_err_bits(Li, Lio):.1f}", flush=True)


def diag_lambda(h, ho, prec):
    rng = random.Random(4)
    for (batch, n) in [(3, 1), (3, 2), (2, 7), (2, 40), (2, 64)]:
        mats = []
        for b in range(batch):
            G = np.array([[rng.uniform(-1, 1) for _ in range(n)] for _ in range(n)])
            mats.append((G + G.T) / 2)
        A = MpArray.from_double(np.array(mats).reshape(-1), h.nlimb)
        t0 = time.time()
        lg = h.op_lambda_min(batch, n, A)
        t1 = time.time()
        lo = ho.op_lambda_min(batch, n, A)
        ref = [np.linalg.eigvalsh(m)[0] for m in mats]
        print(f"lambda_min {batch}x{n}: err bits {rel_err_bits(lg, lo):.1f} gpu {lg.to_double()} numpy {ref} ({t1-t0:.3f}s)",
              flush=True)


def diag_iterate(prec):
    for cfg in [dict(J=2, delta=4, K=6, n_y=3), dict(J=3, delta=8, K=12, n_y=7)]:
        cons, b, info = instances.synthetic_clustered_sdp(prec=prec, **cfg)
        bi = solver.get_block_info(cons)
        hs = [solver.product_handle(prec), oracle_handle(prec, 4)]
        for hh in hs:
            solver.load_problem(hh, cons, b, bi)
            hh.set_params(solver.real_params(hh.nlimb))
            hh.init_point()
            hh.prepare()
        for it in range(3):
            t0 = time.time()
            rg = hs[0].iterate()
            t1 = time.time()
            ro = hs[1].iterate()
            print(f"cfg {cfg} iter {it+1}: gpu {t1-t0:.3f}s alpha_p {rg.alpha_p:.6e}/{ro.alpha_p:.6e} alpha_d {rg.alpha_d:.6e}/{ro.alpha_d:.6e} "
                  f"mu {rg.mu:.6e}/{ro.mu:.6e} beta {rg.beta_c:.3e}/{ro.beta_c:.3e} status {rg.status}", flush=True)
            for name, per in [("XY", "b"), ("R", "b"), ("Xinv", "b"), ("Px", "b"), ("Py", "b"), ("S", "c"), ("Q", "g"), ("P", "b"), ("p", "g"),
                              ("d", "g"), ("Z", "b"), ("dx_pred", "g"), ("dy_pred", "g"), ("dX_pred", "b"), ("dY_pred", "b"),
                              ("dx", "g"), ("dy", "g"), ("dX", "b"), ("dY", "b"), ("x", "g"), ("y", "g"), ("X", "b"), ("Y", "b")]:
                try:
                    if per == "g":
                        e = rel_err_bits(hs[0].fetch(name), hs[1].fetch(name))
                    elif per == "c":
                        e = min(rel_err_bits(hs[0].fetch(name, j), hs[1].fetch(name, j)) for j in range(bi.J))
                    else:
                        e = min(rel_err_bits(hs[0].fetch(name, j, l), hs[1].fetch(name, j, l)) for j in range(bi.J) for l in range(bi.L[j]))
                    print(f"     {name:8s} agree bits {e:7.1f}", flush=True)
                except Exception as ex:
                    print(f"     {name:8s} fetch failed: {ex}", flush=True)
            for sname in ["mu", "lambda_x", "lambda_y", "alpha_p", "alpha_d", "beta_c", "p_obj", "d_obj", "gap"]:
                import mpmath
                with mpmath.workprec(prec + 32):
                    a, bq = hs[0].scalar(sname), hs[1].scalar(sname)
                    rel = abs(a - bq) / abs(bq) if bq != 0 else abs(a)
                    print(f"     scalar {sname:9s} rel diff 2^{float(mpmath.log(rel, 2)) if rel != 0 else -999:.1f}")


def main():
    prec = int(os.environ.get("DIAG_PREC", "256"))
    h = solver.product_handle(prec)
    ho = oracle_handle(prec, 4)
    which = sys.argv[1:] or ["ew", "gemm", "chol", "lam", "iter"]
    if "ew" in which:
        section("elementwise"); guarded(lambda: diag_elementwise(h, ho, prec))
    if "gemm" in which:
        section("gemm"); guarded(lambda: diag_gemm(h, ho, prec))
    if "chol" in which:
        section("cholesky"); guarded(lambda: diag_chol(h, ho, prec))
    if "lam" in which:
        section("lambda_min"); guarded(lambda: diag_lambda(h, ho, prec))
    if "iter" in which:
        section("iterate"); guarded(lambda: diag_iterate(prec))
    print("launches:", h.launch_count())


if __name__ == "__main__":
    main()

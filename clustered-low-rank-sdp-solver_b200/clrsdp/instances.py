"""Instance generators: inputs of the hot path in the layout `prepareabc` produces (MPMP.jl:385-406).

* `synthetic_clustered_sdp` — BASELINE configs 3 and 5: a manufactured, strictly feasible
  clustered low-rank SDP (rank-1 constraints, m = 1, one block per cluster). All data are dyadic
  rationals built from a counter-based PRNG, so the instance is bit-identical on every machine and is
  converted to the working precision without rounding (for p >= 192).
* `sphere_packing_2point` — BASELINE configs 1 and 2: the binary sphere-packing bound of
  examples/SpherePacking.jl:28-114, restating the univariate `Pi = nothing` path of `prepareabc`
  (MPMP.jl:250-254, 283-312, 345-376, 387-400) with mpmath.
Neither touches the oracle or the GPU; they only produce `Constraint` objects and `b`.
"""
from __future__ import annotations

import numpy as np

from .solver import Constraint
from .wire import MpArray

FRAC_BITS = 40  # all synthetic data are integers / 2^40


def _rand_scaled(rng: np.random.Generator, shape, lo: float, hi: float) -> np.ndarray:
    """integers r with r/2^40 uniform in [lo, hi)"""
    return rng.integers(int(lo * 2 ** FRAC_BITS), int(hi * 2 ** FRAC_BITS), size=shape, dtype=np.int64)


def synthetic_clustered_sdp(J=64, delta=64, K=128, n_y=256, prec=256, seed=20261018, j_offset=0, j_total=None,
                            yrank=4):
    """Manufactured strictly feasible pair.

    Cluster j: sample points t_k = Chebyshev nodes on [-1,1] (MPMP.jl:184-191), vectors
    v_k = w_k * [T_0(t_k) .. T_{delta-1}(t_k)] with w_k in [0.5,1.5), H = +1, B_j in [-1,1)^{K x n_y}.
    A primal point x0 in [0.5,1.5) (so X0 = sum x0_k v_k v_k^T > 0 as K >= delta) and a dual point
    (y0 in [-1,1), Y0 = I + G G^T, G delta x yrank in [-1/8,1/8)) define b := B^T x0 and
    c := Tr(A_* Y0) + B y0. Both problems are then strictly feasible, so the gap closes.

    `j_offset/j_total` generate clusters j_offset .. j_offset+J-1 of a j_total-cluster problem (each
    cluster draws from its own PRNG stream), for sharding over ranks: b is the FULL problem's b.
    Returns (constraints, b, info) with info holding x0,y0 for tests.
    """
    nlimb = prec // 32
    j_total = j_total if j_total is not None else j_offset + J
    assert K >= delta
    kk = np.arange(1, K + 1)
    t = np.cos((2 * kk - 1) / (2.0 * K) * np.pi)            # create_sample_points_chebyshev(K-1)
    Tm = np.cos(np.outer(np.arccos(t), np.arange(delta)))     # T_i(t_k), K x delta
    y0 = _rand_scaled(np.random.default_rng([seed, 10 ** 6]), n_y, -1.0, 1.0)
    y0_obj = y0.astype(object)
    b_int = np.zeros(n_y, dtype=object)                       # scale 2^-80
    constraints = []
    x0_all = []
    for j in range(j_total):
        rng = np.random.default_rng([seed, j])
        w = rng.uniform(0.5, 1.5, size=K)
        V = np.round(Tm * w[:, None] * 2 ** FRAC_BITS).astype(np.int64)   # K x delta, scale 2^-40
        B = _rand_scaled(rng, (K, n_y), -1.0, 1.0)
        x0 = _rand_scaled(rng, K, 0.5, 1.5)
        G = _rand_scaled(rng, (delta, yrank), -0.125, 0.125)
        b_int += B.astype(object).T.dot(x0.astype(object))
        if not (j_offset <= j < j_offset + J):
            continue
        Vo, Go = V.astype(object), G.astype(object)
        vv = (Vo * Vo).sum(axis=1)                               # |v_k|^2, scale 2^-80
        gv = Vo.dot(Go)                                          # G^T v_k, scale 2^-80
        gg = (gv * gv).sum(axis=1)                               # scale 2^-160
        By = B.astype(object).dot(y0_obj)                        # scale 2^-80
        c_int = (vv + By) * (1 << 80) + gg                       # scale 2^-160
        constraints.append(Constraint(
            V=[MpArray.from_scaled_int64(V, -FRAC_BITS, nlimb).reshape(K, delta)],
            ranks=[np.ones(K, dtype=np.int32)],
            H=[MpArray.from_scaled_int64(np.ones(K, dtype=np.int64), 0, nlimb)],
            B=MpArray.from_scaled_int64(B, -FRAC_BITS, nlimb).reshape(K, n_y),
            c=MpArray.from_ints(list(c_int), -160, nlimb),
        ))
        x0_all.append(x0)
    b = MpArray.from_ints(list(b_int), -80, nlimb)
    info = dict(x0=x0_all, y0=y0, J=J, delta=delta, K=K, n_y=n_y, prec=prec, seed=seed)
    return constraints, b, info


# ----------------------------------------------------------------------------------------------------
# sphere packing (examples/SpherePacking.jl) — mpmath restatement of the input generation
# ----------------------------------------------------------------------------------------------------
def _poly_eval(coeffs, x):
    r = 0
    for c in reversed(coeffs):
        r = r * x + c
    return r


def laguerrebasis_polys(k, alpha, scale, mp):
    """Coefficient lists (ascending powers of x) of laguerrebasis(k, alpha, scale*x) (MPMP.jl:43-54)."""
    def padd(a, b):
        n = max(len(a), len(b))
        return [(a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0) for i in range(n)]

    def pscal(a, s):
        return [c * s for c in a]

    def pmulx(a, s):  # a(x) * (s*x)
        return [mp.mpf(0)] + [c * s for c in a]

    v = [[mp.mpf(1)]]
    if k == 0:
        return v
    v.append(padd([1 + alpha], pmulx([mp.mpf(-1)], scale)))
    for l in range(2, k + 1):
        # v[l+1] = 1/l * ((2l-1+alpha-x) v[l] - (l+alpha-1) v[l-1])
        t1 = padd(pscal(v[l - 1], 2 * l - 1 + alpha), pscal(pmulx(v[l - 1], scale), -1))
        t2 = pscal(v[l - 2], -(l + alpha - 1))
        v.append(pscal(padd(t1, t2), mp.mpf(1) / l))
    return v


def sphere_packing_2point(n=3, d=8, r=None, prec=512, N=2):
    """The SDP of examples/SpherePacking.jl:28-110 for N=2 radii, as (constraints, b, omega).

    Variables y = (M, a_{ij,k} for k=0:2d for i=1:N for j=1:i) (:54, :89); objective max -M (:88-89).
    The seven constraints (:56-66), their sample points (:69-72), weights G (:75-78), the rescaled
    Laguerre basis (:81-83), degrees (:86) and the cluster reordering (:99-105) follow the example;
    `prepareabc`'s default path turns each into (A,B,c,H).
    """
    import mpmath
    assert N == 2
    nlimb = prec // 32
    mp = mpmath.mp.clone()
    mp.prec = prec + 64
    if r is None:
        r = [mp.mpf(1), mp.sqrt(2) - 1]
    r = [mp.mpf(v) for v in r]
    pi = mp.pi
    pairs = [(i, j) for i in range(N) for j in range(i + 1)]            # for i=1:N for j=1:i
    n_a = (2 * d + 1) * len(pairs)
    n_y = 1 + n_a

    def a_index(k, pi_idx):  # column of a_{pair,k} in y (after M)
        return 1 + k * len(pairs) + pi_idx

    def spherevolume(nn, rr):
        return mp.sqrt(pi) ** nn / mp.gamma(mp.mpf(nn) / 2 + 1) * rr ** nn

    # basis: laguerrebasis(d, n/2-1, 2*pi*x), each divided by its max coefficient (:81-83)
    q = laguerrebasis_polys(d, mp.mpf(n) / 2 - 1, 2 * pi, mp)
    q = [[c / max(p) for c in p] for p in q]
    deg_q = list(range(d + 1))
    # Laguerre polynomials for the f constraints: k!/pi^k L_k^{n/2-1}(pi x)
    lag = laguerrebasis_polys(2 * d, mp.mpf(n) / 2 - 1, pi, mp)
    fcoef = [mp.factorial(k) / pi ** k for k in range(2 * d + 1)]

    def samples_1d(dd):  # create_sample_points_1d (MPMP.jl:173-182)
        const = -mp.sqrt(pi) / (64 * (dd + 1) * mp.log(3 - 2 * mp.sqrt(2)))
        return [const * (-1 + 4 * k) ** 2 for k in range(dd + 1)]

    def prepare(m, Mfun, G, deg_G, xs, delta):
        """prepareabc default path: Mfun(x) -> (M0 (m x m), list over y-index of m x m matrices)."""
        K = len(xs)
        last_deg = []
        for e in range(delta // 2 + 1):  # (:296-303)
            idx = [i for i, dg in enumerate(deg_q) if dg == e]
            last_deg.append(idx[-1] + 1 if idx else last_deg[-1])
        V, H, ranks = [], [], []
        for l, g in enumerate(G):
            nvec = last_deg[(delta - deg_G[l]) // 2]
            vl, hl = [], []
            for xk in xs:
                gv = g(xk)
                sq = mp.sqrt(abs(gv))
                vl.extend(_poly_eval(q[dd], xk) * sq for dd in range(nvec))       # (:365-374)
                hl.append(mp.sign(gv))                                            # (:307-312)
            # threshold pruning (:378-383) never triggers here: |H| = 1
            V.append(MpArray.from_mpf(vl, nlimb).reshape(K, nvec))
            H.append(MpArray.from_mpf(hl, nlimb))
            ranks.append(np.ones(K, dtype=np.int32))
        Brows, crows = [], []
        evals = [Mfun(xk) for xk in xs]
        for rr in range(m):          # for r = 1:m for s = 1:r for k  (:387-400)
            for ss in range(rr + 1):
                for k in range(K):
                    M0, Mi = evals[k]
                    Brows.extend(-Mi[i][rr][ss] for i in range(n_y))
                    crows.append(M0[rr][ss])
        dimS = m * (m + 1) // 2 * K
        return Constraint(V=V, ranks=ranks, H=H, B=MpArray.from_mpf(Brows, nlimb).reshape(dimS, n_y),
                          c=MpArray.from_mpf(crows, nlimb))

    zero2 = lambda: [[mp.mpf(0)] * N for _ in range(N)]

    def E(i, j, val):
        Z = zero2()
        Z[i][j] = val
        Z[j][i] = val
        return Z

    def M0fun(x):  # (:56-57)
        M0 = [[-mp.sqrt(spherevolume(n, r[i]) * spherevolume(n, r[j])) for j in range(N)] for i in range(N)]
        Mi = [zero2() for _ in range(n_y)]
        for pi_idx, (i, j) in enumerate(pairs):
            Mi[a_index(0, pi_idx)] = E(i, j, mp.mpf(1))
        return M0, Mi

    def M1fun(x):  # (:59)
        Mi = [zero2() for _ in range(n_y)]
        for k in range(2 * d + 1):
            for pi_idx, (i, j) in enumerate(pairs):
                Mi[a_index(k, pi_idx)] = E(i, j, x ** k)
        return zero2(), Mi

    def M2fun(pair_idx):  # (:61-62)
        def f(x):
            Mi = [[[mp.mpf(0)]] for _ in range(n_y)]
            for k in range(2 * d + 1):
                Mi[a_index(k, pair_idx)] = [[-fcoef[k] * _poly_eval(lag[k], x)]]
            return [[mp.mpf(0)]], Mi
        return f

    def M3fun(i):  # (:64-65)
        pair_idx = pairs.index((i, i))
        def f(x):
            Mi = [[[mp.mpf(0)]] for _ in range(n_y)]
            Mi[0] = [[mp.mpf(1)]]
            for k in range(2 * d + 1):
                Mi[a_index(k, pair_idx)] = [[-fcoef[k] * _poly_eval(lag[k], mp.mpf(0))]]
            return [[mp.mpf(0)]], Mi
        return f

    one = lambda x: mp.mpf(1)
    cons = [prepare(N, M0fun, [one], [0], [mp.mpf(0)], 0),
            prepare(N, M1fun, [one, lambda x: x], [0, 1], samples_1d(2 * d), 2 * d)]
    for pair_idx, (i, j) in enumerate(pairs):
        shift = (r[i] + r[j]) ** 2
        xs = [v + shift for v in samples_1d(2 * d)]
        cons.append(prepare(1, M2fun(pair_idx), [one, (lambda s: (lambda x: x - s))(shift)], [0, 1], xs, 2 * d))
    for i in range(N):
        cons.append(prepare(1, M3fun(i), [one], [0], [mp.mpf(0)], 0))
    ordering = [3, 6, 5, 7, 4, 1, 2]  # (:102), 1-based
    cons = [cons[o - 1] for o in ordering]
    b = MpArray.from_mpf([mp.mpf(-1)] + [mp.mpf(0)] * n_a, nlimb)  # (:89)
    return cons, b, dict(n=n, d=d, prec=prec, n_y=n_y, omega=mp.mpf(100))


def random_structured_sdp(spec, n_y=5, prec=256, seed=7, positive_H=True):
    """Random clustered SDP with arbitrary structure, for parity tests of the general code paths
    (m > 1, several blocks per cluster, rank > 1, samples of rank 0).

    spec: list over clusters of dict(m=.., K=.., blocks=[dict(delta=.., ranks=[.. K ints ..]), ...]).
    With positive_H the instance is manufactured strictly feasible (X0 = sum_k (sum_rnk H v v^T) (x) x0_k
    with each m x m matrix x0_k diagonally dominant), so a full solve converges; otherwise H has mixed
    signs and only single iterations from omega*I are meaningful.
    """
    nlimb = prec // 32
    rng = np.random.default_rng(seed)
    constraints, b_int = [], np.zeros(n_y, dtype=object)
    y0 = _rand_scaled(rng, n_y, -1.0, 1.0).astype(object)
    for cl in spec:
        m, K = cl["m"], cl["K"]
        npairs = m * (m + 1) // 2
        dimS = npairs * K
        B = _rand_scaled(rng, (dimS, n_y), -1.0, 1.0)
        # x0: for every sample a diagonally dominant symmetric m x m matrix, stored by pairs (r,s), s <= r
        x0 = np.zeros(dimS, dtype=np.int64)
        for r in range(m):
            for s in range(r + 1):
                pr = s + r * (r + 1) // 2
                x0[pr * K:(pr + 1) * K] = _rand_scaled(rng, K, 1.0, 2.0) if r == s else _rand_scaled(rng, K, -0.2, 0.2)
        V, H, ranks, Vint, Hint = [], [], [], [], []
        for bk in cl["blocks"]:
            dl, rk = bk["delta"], np.asarray(bk["ranks"], dtype=np.int32)
            Nv = int(rk.sum())
            v = _rand_scaled(rng, (Nv, dl), -1.0, 1.0)
            hsign = np.ones(Nv, dtype=np.int64) if positive_H else rng.choice([-1, 1], size=Nv)
            h = hsign * _rand_scaled(rng, Nv, 0.5, 2.0)
            V.append(MpArray.from_scaled_int64(v, -FRAC_BITS, nlimb).reshape(Nv, dl))
            H.append(MpArray.from_scaled_int64(h, -FRAC_BITS, nlimb))
            ranks.append(rk)
            Vint.append(v.astype(object))
            Hint.append(h.astype(object))
        # dual point Y0 = I (block identity): Tr(A_(r,s,k) Y0) = sum_l sum_rnk H |v|^2 [r == s]
        c_int = B.astype(object).dot(y0) * (1 << (2 * FRAC_BITS))          # scale 2^-120
        for r in range(m):
            pr = r + r * (r + 1) // 2
            for l, bk in enumerate(cl["blocks"]):
                rk = ranks[l]
                pos = 0
                for k in range(K):
                    for _ in range(rk[k]):
                        vv = int((Vint[l][pos] * Vint[l][pos]).sum())       # scale 2^-80
                        c_int[pr * K + k] += int(Hint[l][pos]) * vv          # scale 2^-120
                        pos += 1
        b_int += B.astype(object).T.dot(x0.astype(object))
        constraints.append(Constraint(V=V, ranks=ranks, H=H,
                                      B=MpArray.from_scaled_int64(B, -FRAC_BITS, nlimb).reshape(dimS, n_y),
                                      c=MpArray.from_ints(list(c_int), -3 * FRAC_BITS, nlimb)))
    b = MpArray.from_ints(list(b_int), -2 * FRAC_BITS, nlimb)
    return constraints, b


def bivariate_matrix_program(D=6, n_y=60, clusters=4, prec=384, seed=20261021, ball_radius2=3):
    """BASELINE config 4: `clusters` constraints, each a 2 x 2 polynomial matrix inequality in two
    variables of degree 2D on {G_l >= 0}, G = {1, ball_radius2 - x1^2 - x2^2}, sampled at the Padua points of degree 2D
    (K = C(2D+2, 2)) in the product-Chebyshev basis T_a(x1) T_b(x2), a + b <= D, with a fixed positive definite 2 x 2
    matrix Pi_l per weight, through `frontend.prepareabc` (the `Pi` path, all_of_Pi = true): every (l, k) has rank 2,
    vectors of length 2 C(D+2, 2) and 2 C(D+1, 2). D = 6: K = 91, blocks 112 and 84, dim_S = 273.

    ball_radius2 = 3 keeps G_2 positive at every sample, corners of the square included (H > 0, rank 2 everywhere),
    and the instance is manufactured strictly feasible like `random_structured_sdp` (so a full solve converges);
    ball_radius2 = 2 vanishes at the two corner Padua points (rank 0 there, pruned by the threshold); ball_radius2 = 1
    is SURVEY's G = 1 - x1^2 - x2^2, negative at the Padua points outside the disc (mixed-sign H: single iterations
    from omega*I only).
    The M_i are random symmetric polynomial matrices of degree D (seeded); M_0 enters only through c, which is set so
    that (y0, Y0 = I) is dual feasible; b = B^T x0.
    """
    import mpmath
    from . import frontend as fe
    nlimb = prec // 32
    mp = mpmath.mp.clone()
    mp.prec = prec + 64
    rng = np.random.default_rng(seed)
    x1, x2 = fe.Poly.var(2, 0), fe.Poly.var(2, 1)
    def cheb(k, x):
        t = [x * 0 + 1, x * 1]
        for i in range(2, k + 1):
            t.append(2 * x * t[i - 1] - t[i - 2])
        return t[:k + 1]
    c1, c2 = cheb(D, x1), cheb(D, x2)
    q = [c1[a] * c2[e - a] for e in range(D + 1) for a in range(e, -1, -1)]       # degree-monotone
    G = [fe.Poly.const(2, 1), ball_radius2 - x1 * x1 - x2 * x2]
    Pi = [[[fe.Poly.const(2, 2), fe.Poly.const(2, 1)], [fe.Poly.const(2, 1), fe.Poly.const(2, 2)]],
          [[fe.Poly.const(2, 3), fe.Poly.const(2, -1)], [fe.Poly.const(2, -1), fe.Poly.const(2, 2)]]]
    xs = fe.create_sample_points_2d(2 * D, prec)
    K = len(xs)
    mono = [e for k in range(D + 1) for e in fe.multiexponents(2, k)]
    monovals = [[xk[0] ** e[0] * xk[1] ** e[1] for e in mono] for xk in xs]          # [K][n_mono]
    qvals = [[fe.evaluate(p, xk) for p in q] for xk in xs]
    y0 = [mp.mpf(int(v)) / (1 << FRAC_BITS) for v in _rand_scaled(rng, n_y, -1.0, 1.0)]
    constraints = []
    b_acc = [mp.mpf(0)] * n_y
    m = 2
    for _ in range(clusters):
        # M_i[r][s](x_k) for i = 1..n_y: random coefficients on the monomials of degree <= D, evaluated by dot products
        coef = _rand_scaled(rng, (n_y, 3, len(mono)), -1.0, 1.0).astype(object)
        zero = fe.Sampled([mp.mpf(0)] * K)
        Ms = [[[zero] * m for _ in range(m)]]
        for i in range(n_y):
            ent = {}
            for pr, (r, s) in enumerate(((0, 0), (1, 0), (1, 1))):
                cf = [mp.mpf(int(v)) / (1 << FRAC_BITS) for v in coef[i, pr]]
                ent[(r, s)] = fe.Sampled([mp.fdot(cf, monovals[k]) for k in range(K)])
            Ms.append([[ent[(max(r, s), min(r, s))] for s in range(m)] for r in range(m)])
        con = fe.prepareabc(Ms, G, q, xs, 2 * D, Pi, prec=prec, qp_precomp=qvals)
        # c = Tr(A_* Y0) + B y0 with Y0 = I; b += B^T x0 with x0 diagonally dominant per sample
        dimS = 3 * K
        Bm = [[con.B.to_mpf(r * n_y + i) for i in range(n_y)] for r in range(dimS)]
        cvals = [mp.fdot(Bm[r], y0) for r in range(dimS)]
        for l in range(con.L):
            Vl, Hl, rk = con.V[l], con.H[l], con.ranks[l]
            width = Vl.shape[1]
            pos = 0
            for k in range(K):
                for _r in range(int(rk[k])):
                    vv = mp.fsum(Vl.to_mpf(pos * width + t) ** 2 for t in range(width))
                    for r in range(m):  # Tr(A_(r,r,k) Y0) = sum H |v|^2 for Y0 = I; the (r, s != r) blocks of I vanish
                        cvals[(r + r * (r + 1) // 2) * K + k] += Hl.to_mpf(pos) * vv
                    pos += 1
        x0 = [mp.mpf(0)] * dimS
        for r in range(m):
            for s in range(r + 1):
                pr = s + r * (r + 1) // 2
                draw = _rand_scaled(rng, K, 1.0, 2.0) if r == s else _rand_scaled(rng, K, -0.2, 0.2)
                for k in range(K):
                    x0[pr * K + k] = mp.mpf(int(draw[k])) / (1 << FRAC_BITS)
        for i in range(n_y):
            b_acc[i] += mp.fsum(Bm[r][i] * x0[r] for r in range(dimS))
        con.c = MpArray.from_mpf(cvals, nlimb)
        constraints.append(con)
    return constraints, MpArray.from_mpf(b_acc, nlimb)

"""Host-side mirror of the reference's input helpers: polynomial bases, sample points,
`prepareabc` and `solvempmp` (MPMP.jl:21-200, :225-407, :562-586), on mpmath numbers.

One-off input generation — nothing here is on the per-iteration hot path. It exists so that real (non-synthetic)
instances can be produced without Julia, in exactly the `(A, B, c, H)` layout `clrsdp_upload_cluster` takes
(SURVEY Appendix D). Polynomials are the small sparse class `Poly` below instead of AbstractAlgebra's; the bases are
generic in their argument like the reference's (a number evaluates the basis, a `Poly` variable builds it).

Working precision: every function computes in an mpmath context of `prec + 64` bits and rounds once to `prec` bits
when the `Constraint` is assembled (the reference evaluates in BigFloat at the global precision and converts to Arb,
:309, :374, :394, :399).
"""
from __future__ import annotations

import itertools
from math import comb

import mpmath
import numpy as np

from .solver import Constraint, get_block_info, precision, solverank1sdp
from .wire import MpArray


def _ctx(prec=None):
    mp = mpmath.mp.clone()
    mp.prec = (prec or precision()) + 64
    return mp


# ---- polynomials -------------------------------------------------------------------------------------------
class Poly:
    """Sparse multivariate polynomial: {exponent tuple: coefficient}. Stand-in for AbstractAlgebra's MPoly as far
    as `prepareabc` uses it: evaluation `p(*x)`, `total_degree`, `nvars`, ring arithmetic."""
    __slots__ = ("nvars", "terms")

    def __init__(self, nvars, terms=None):
        self.nvars = nvars
        self.terms = {e: c for e, c in (terms or {}).items() if c != 0}

    @staticmethod
    def var(nvars, i):
        return Poly(nvars, {tuple(1 if k == i else 0 for k in range(nvars)): 1})

    @staticmethod
    def const(nvars, c):
        return Poly(nvars, {(0,) * nvars: c})

    @staticmethod
    def monomial(exponents, c=1):
        return Poly(len(exponents), {tuple(exponents): c})

    def _lift(self, o):
        return o if isinstance(o, Poly) else Poly.const(self.nvars, o)

    def __add__(self, o):
        o = self._lift(o)
        t = dict(self.terms)
        for e, c in o.terms.items():
            t[e] = t.get(e, 0) + c
        return Poly(self.nvars, t)

    __radd__ = __add__

    def __neg__(self):
        return Poly(self.nvars, {e: -c for e, c in self.terms.items()})

    def __sub__(self, o):
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) + (-self)

    def __mul__(self, o):
        if not isinstance(o, Poly):
            return Poly(self.nvars, {e: c * o for e, c in self.terms.items()})
        t = {}
        for e1, c1 in self.terms.items():
            for e2, c2 in o.terms.items():
                e = tuple(a + b for a, b in zip(e1, e2))
                t[e] = t.get(e, 0) + c1 * c2
        return Poly(self.nvars, t)

    __rmul__ = __mul__

    def __truediv__(self, s):
        return Poly(self.nvars, {e: c / s for e, c in self.terms.items()})

    def __pow__(self, k):
        r = Poly.const(self.nvars, 1)
        for _ in range(int(k)):
            r = r * self
        return r

    def __call__(self, *x):
        if len(x) == 1 and isinstance(x[0], (list, tuple)):
            x = tuple(x[0])
        assert len(x) == self.nvars
        r = 0
        for e, c in self.terms.items():
            t = c
            for xi, ei in zip(x, e):
                if ei:
                    t = t * xi ** ei
            r = r + t
        return r

    def total_degree(self):
        return max((sum(e) for e in self.terms), default=0)

    def max_abs_coeff(self):
        return max((abs(c) for c in self.terms.values()), default=0)


def total_degree(p):
    return p.total_degree() if isinstance(p, Poly) else 0


class Sampled:
    """A function known only by its values at the sample points (values[k] = f(x_k)): accepted wherever `prepareabc`
    evaluates an entry of M at a sample (the reference evaluates polynomials; a generator that already has the
    values need not interpolate them first)."""
    __slots__ = ("values",)

    def __init__(self, values):
        self.values = list(values)


def evaluate(p, xk, k=None):
    """p(x_k...) for a Poly, a callable of the point, a `Sampled` (by sample index k), or a constant."""
    if isinstance(p, Sampled):
        return p.values[k]
    if isinstance(p, Poly):
        return p(*xk) if isinstance(xk, (list, tuple)) else p(xk)
    if callable(p):
        return p(*xk) if isinstance(xk, (list, tuple)) else p(xk)
    return p


def _one(x):
    return x * 0 + 1


# ---- bases (MPMP.jl:21-92) ------------------------------------------------------------------------------------
def multiexponents(n, k):
    """Exponent vectors of length n with sum k, in the order of Combinatorics.multiexponents (lexicographically
    decreasing: (k,0,..,0) first)."""
    if n == 1:
        yield (k,)
        return
    for first in range(k, -1, -1):
        for rest in multiexponents(n - 1, k - first):
            yield (first,) + rest


def make_monomial_basis(nvars, d):
    """All monomials of total degree <= d, by increasing degree (:22-40): binomial(n + d, d) polynomials."""
    return [Poly.monomial(e) for k in range(d + 1) for e in multiexponents(nvars, k)]


def laguerrebasis(k, alpha, x):
    """Generalised Laguerre polynomials L_0^alpha .. L_k^alpha at x (:43-54)."""
    v = [_one(x)]
    if k == 0:
        return v
    v.append(1 + alpha - x)
    for l in range(2, k + 1):
        v.append(((2 * l - 1 + alpha - x) * v[l - 1] - (l + alpha - 1) * v[l - 2]) / mpmath.mpf(l))
    return v


def jacobi_basis(d, alpha, beta, x, normalized=True):
    """Jacobi-type basis by the reference's three-term recurrence, LITERALLY as written at :56-75: the prefactor
    (2k+a+b-1) / (2k (k+a+b) (2k+a+b-2)) multiplies only the first term; the second term
    -2 (k+a-1)(k+b-1)(2k+a+b) q[k-1] is added undivided (operator precedence in the source). For alpha = beta the
    result still spans the same spaces degree by degree, which is all `prepareabc` needs of a basis."""
    q = [_one(x)]
    if d == 0:
        return q
    q.append(x * 1 if normalized else x * (alpha + 1))
    for k in range(2, d + 1):
        s = 2 * k + alpha + beta
        pref = (s - 1) / mpmath.mpf(2 * k * (k + alpha + beta) * (s - 2))
        q.append(pref * ((s * (s - 2)) * x + (beta ** 2 - alpha ** 2)) * q[k - 1]
                 + (-2 * (k + alpha - 1) * (k + beta - 1) * s) * q[k - 2])
    return q


def gegenbauer_basis(k, n, x):
    """Gegenbauer polynomials of dimension n (lambda = n/2 - 1) up to degree k, normalised to 1 at x = 1 (:82-92)."""
    v = [_one(x)]
    if k == 0:
        return v
    v.append(x * 1)
    for l in range(2, k + 1):
        v.append(mpmath.mpf(2 * l + n - 4) / (l + n - 3) * x * v[l - 1] - mpmath.mpf(l - 1) / (l + n - 3) * v[l - 2])
    return v


# ---- sample points (MPMP.jl:94-200) -------------------------------------------------------------------------
def create_sample_points(n, d, prec=None):
    """Rational points of the unit simplex with denominator d (:94-106): binomial(n + d, d) points. The reference
    walks CartesianIndices, whose FIRST index runs fastest."""
    mp = _ctx(prec)
    pts = []
    for idx in itertools.product(range(d + 1), repeat=n):
        I = idx[::-1]  # first coordinate fastest
        if sum(I) <= d:
            pts.append([mp.mpf(i) / d for i in I])
    assert len(pts) == comb(n + d, d)
    return pts


def create_sample_points_2d(d, prec=None):
    """Padua points (:108-122): binomial(d + 2, 2) points in [-1, 1]^2, unisolvent for total degree d."""
    mp = _ctx(prec)
    z = []
    for j in range(d + 1):
        delta_j = 1 if (j % 2 == 1 and d % 2 == 1) else 0
        mu_j = mp.cospi(mp.mpf(j) / d)
        for k in range(1, d // 2 + 1 + delta_j + 1):
            eta_k = mp.cospi(mp.mpf(2 * k - 2) / (d + 1)) if j % 2 == 1 else mp.cospi(mp.mpf(2 * k - 1) / (d + 1))
            z.append([mu_j, eta_k])
    assert len(z) == comb(d + 2, 2)
    return z


def create_sample_points_chebyshev(d, a=-1, b=1, prec=None):
    """Roots of the Chebyshev polynomial T_{d+1}, mapped to [a, b] (:184-191)."""
    mp = _ctx(prec)
    return [mp.mpf(a + b) / 2 + mp.mpf(b - a) / 2 * mp.cos(mp.mpf(2 * k - 1) / (2 * (d + 1)) * mp.pi) for k in range(1, d + 2)]


def create_sample_points_chebyshev_mod(d, a=-1, b=1, prec=None):
    """The same roots divided by cos(pi / (2(d+1))) (lower Lebesgue constant, :193-200)."""
    mp = _ctx(prec)
    s = mp.cos(mp.pi / (2 * (d + 1)))
    return [mp.mpf(a + b) / 2 + mp.mpf(b - a) / 2 * mp.cos(mp.mpf(2 * k - 1) / (2 * (d + 1)) * mp.pi) / s
            for k in range(1, d + 2)]


def create_sample_points_1d(d, prec=None):
    """'Rescaled Laguerre' points of Simmons-Duffin (:173-182): x_k = sqrt(pi) (4k - 1)^2 / (-64 (d+1) log(3 - 2 sqrt 2))."""
    mp = _ctx(prec)
    const = -mp.sqrt(mp.pi) / (64 * (d + 1) * mp.log(3 - 2 * mp.sqrt(2)))
    return [const * (-1 + 4 * k) ** 2 for k in range(d + 1)]


def create_sample_points_3d(d, pairs=((1, 3), (3, 2), (2, 1)), prec=None):
    """Padua x Chebyshev extension to three variables (:124-146): (d+1)(d+2)(d+3)/6 points, good for odd d."""
    pad = create_sample_points_2d(d, prec)
    ch = create_sample_points_chebyshev(d + 2, prec=prec)
    pad_div = [pad[i::3] for i in range(3)]
    cheb_div = [ch[i::3] for i in range(3)]
    pts = [list(p1) + [p2] for (a, b) in pairs for p1 in pad_div[a - 1] for p2 in cheb_div[b - 1]]
    return pts[:(d + 1) * (d + 2) * (d + 3) // 6]


def points_X_general(n, d, prec=None):
    """Recursive Padua/Chebyshev construction in n variables (:148-170; the reference notes it is only sometimes
    unisolvent)."""
    if n == 2:
        return create_sample_points_2d(d, prec)
    Xn_1 = points_X_general(n - 1, d, prec)
    cheb = create_sample_points_chebyshev(d + n - 1, prec=prec)
    X_div = [Xn_1[i::n] for i in range(n)]
    cheb_div = [cheb[i::n] for i in range(n)]
    pts = []
    for i in range(n):
        j = n - 1 if i == 0 else i - 1
        pts.extend(list(p1) + [p2] for p1 in X_div[i] for p2 in cheb_div[j])
    return pts[:comb(n + d, d)]


# ---- prepareabc (MPMP.jl:225-407) -----------------------------------------------------------
def _last_deg(q, delta):
    """last_deg[e] = number of basis polynomials of degree <= e for e = 0 .. delta/2, gaps filled from the left
    (:284-303)."""
    degs = [total_degree(p) for p in q]
    if any(degs[i] > degs[i + 1] for i in range(len(degs) - 1)):
        print("Degrees are not monotone. The program will (most probably) not be correct if you don't fix this")
    out = []
    for e in range(delta // 2 + 1):
        idx = [i for i, dg in enumerate(degs) if dg == e]
        out.append(idx[-1] + 1 if idx else out[-1])
    return out


def _sym_eig(A, mp):
    """Eigen-decomposition of the small symmetric matrix Pi[l](x_k), ordered like the reference's SVD (:256-270):
    by decreasing |lambda|; value = sign(<U_r, Vt_r>) sigma_r = lambda_r, vector = U[:, r]."""
    n = len(A)
    M = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            M[i, j] = (A[i][j] + A[j][i]) / 2
    ev, Q = mp.eigsy(M)
    order = sorted(range(n), key=lambda r: -abs(ev[r]))
    return [ev[r] for r in order], [[Q[i, r] for i in range(n)] for r in order]


def prepareabc(M, G, q, x, delta=-1, Pi=None, prec=None, all_of_Pi=True, threshold=None, qp_precomp=None) -> Constraint:
    """One polynomial matrix constraint M[0](x) + sum_i y_i M[i](x) >= 0 on {G_l >= 0}, sampled at the points x in the
    basis q, as the tuple (A, B, c, H) of the reference (returned as a `Constraint`, whose fields are that tuple stored
    densely).

    M: list of m x m nested lists (entries Poly, callables of the point, or constants); G: list of Poly; q: list of
    Poly, degree-monotone; x: list of points (lists of coordinates; bare numbers for one variable); Pi: None or one
    symmetric polynomial matrix per G. `qp_precomp[k][d]` may hold q[d](x_k) (:338, :354).
    With all_of_Pi (the only live path of the reference) row i of Pi gets its own degree budget
    (Pi index outer, basis index inner, :353-374); without it the vector is kron(q-part, Pi-vector) (:315-343).
    """
    prec = prec or precision()
    nlimb = prec // 32
    mp = _ctx(prec)
    threshold = mp.mpf(10) ** -70 if threshold is None else mp.mpf(threshold)
    m = len(M[0])
    K = len(x)
    if delta < 0:
        delta = 2 * total_degree(q[-1])
    last_deg = _last_deg(q, delta)
    qx = [[qp_precomp[k][d] if qp_precomp is not None else evaluate(q[d], x[k]) for d in range(len(q))]
          for k in range(K)]
    V, H, ranks = [], [], []
    for l, g in enumerate(G):
        deg_g = total_degree(g)
        if Pi is None:
            size_pi, deg_pi_vec, deg_pi = 1, [0], 0
        else:
            size_pi = len(Pi[l])
            deg_pi_vec = [total_degree(Pi[l][i][i]) for i in range(size_pi)]
            deg_pi = max(total_degree(Pi[l][i][j]) for i in range(size_pi) for j in range(size_pi))
        vl, hl, rl = [], [], []
        for k in range(K):
            gv = evaluate(g, x[k])
            sq = mp.sqrt(abs(gv))
            if Pi is None:
                vals, vecs = [mp.mpf(1)], [[mp.mpf(1)]]
            else:
                vals, vecs = _sym_eig([[evaluate(Pi[l][i][j], x[k]) for j in range(size_pi)] for i in range(size_pi)], mp)
            kept = 0
            for r in range(size_pi):
                h = vals[r] * mp.sign(gv)                                      # A_sign (:307-312)
                if not abs(h) > threshold:                                     # pruning (:378-383)
                    continue
                if all_of_Pi:
                    vec = [vecs[r][i] * qx[k][d] * sq for i in range(size_pi)
                           for d in range(last_deg[(delta - deg_g - deg_pi_vec[i]) // 2])]
                else:
                    vec = [qx[k][d] * sq * vecs[r][i] for d in range(last_deg[(delta - deg_g - deg_pi) // 2])
                           for i in range(size_pi)]
                vl.append(vec)
                hl.append(h)
                kept += 1
            rl.append(kept)
        width = len(vl[0]) if vl else 0
        V.append(MpArray.from_mpf([v for vec in vl for v in vec], nlimb).reshape(len(vl), width))
        H.append(MpArray.from_mpf(hl, nlimb))
        ranks.append(np.asarray(rl, dtype=np.int32))
    n_y = len(M) - 1
    Brows, crows = [], []
    for r in range(m):                  # rows (r, s, k), s <= r, k fastest (:387-400)
        for s in range(r + 1):
            for k in range(K):
                Brows.extend(-evaluate(M[i][r][s], x[k], k) for i in range(1, n_y + 1))
                crows.append(evaluate(M[0][r][s], x[k], k))
    dimS = m * (m + 1) // 2 * K
    return Constraint(V=V, ranks=ranks, H=H, B=MpArray.from_mpf(Brows, nlimb).reshape(dimS, n_y),
                      c=MpArray.from_mpf(crows, nlimb))


def solvempmp(M, G, q, x, delta, b, Pi=None, all_of_Pi=True, **kwargs):
    """`solvempmp` (:562-586): prepareabc per constraint, get_block_info, solverank1sdp. As in the reference
     precision, threshold and qp_precomp are not forwarded: the precision is the global one
    (`clrsdp.solver.set_precision`)."""
    if Pi is not None:
        abc = [prepareabc(M[j], G[j], q[j], x[j], delta[j], Pi[j], all_of_Pi=all_of_Pi) for j in range(len(M))]
    else:
        abc = [prepareabc(M[j], G[j], q[j], x[j], delta[j]) for j in range(len(M))]
    blockinfo = get_block_info(abc)
    if not isinstance(b, MpArray):
        b = MpArray.from_mpf(b, precision() // 32)
    return solverank1sdp(abc, b, blockinfo, **kwargs)

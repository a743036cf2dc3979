"""Front-end helpers: bases, sample points, prepareabc, solvempmp — CPU only.
The known answers are mathematical identities (orthogonal-polynomial values, unisolvence, reconstruction of
G * (q q^T (x) Pi) from the low-rank vectors) plus one cross-check against the independently written sphere-packing
generator and one tiny polynomial programme with an analytic optimum solved through `solvempmp` on the oracle."""
from math import comb

import mpmath
import numpy as np
import pytest

from clrsdp import frontend as fe
from clrsdp import instances, solver
from clrsdp.wire import MpArray, rel_err_bits
from oracle.ref import oracle_handle

mpmath.mp.prec = 320


def test_monomial_basis_order_and_count():
    q = fe.make_monomial_basis(2, 3)
    assert len(q) == comb(5, 3)
    assert [next(iter(p.terms)) for p in q[:6]] == [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (0, 2)]
    degs = [p.total_degree() for p in q]
    assert degs == sorted(degs)
    assert len(fe.make_monomial_basis(4, 5)) == comb(9, 5)


def test_laguerre_and_gegenbauer_match_mpmath():
    x = mpmath.mpf("0.37")
    alpha = mpmath.mpf("0.5")
    lag = fe.laguerrebasis(7, alpha, x)
    for k, v in enumerate(lag):
        assert abs(v - mpmath.laguerre(k, alpha, x)) < mpmath.mpf(2) ** -250
    n = 5
    geg = fe.gegenbauer_basis(6, n, x)
    lam = mpmath.mpf(n) / 2 - 1
    for k, v in enumerate(geg):
        assert abs(v - mpmath.gegenbauer(k, lam, x) / mpmath.gegenbauer(k, lam, 1)) < mpmath.mpf(2) ** -250
    assert all(abs(v - 1) < mpmath.mpf(2) ** -250 for v in fe.gegenbauer_basis(6, n, mpmath.mpf(1)))


def test_bases_are_generic_in_their_argument():
    """A `Poly` variable builds the basis polynomials; evaluating them gives the numeric basis."""
    X = fe.Poly.var(1, 0)
    x = mpmath.mpf("1.25")
    for build in (lambda a: fe.laguerrebasis(5, mpmath.mpf("0.5"), a), lambda a: fe.gegenbauer_basis(5, 4, a),
                  lambda a: fe.jacobi_basis(5, 1, 1, a), lambda a: fe.jacobi_basis(4, 0, 2, a, False)):
        polys, vals = build(X), build(x)
        assert [p.total_degree() for p in polys] == list(range(len(polys)))
        for p, v in zip(polys, vals):
            assert abs(p(x) - v) <= abs(v) * mpmath.mpf(2) ** -240


def test_jacobi_recurrence_is_the_literal_one():
    a, b, x = 1, 2, mpmath.mpf("0.3")
    q = fe.jacobi_basis(2, a, b, x)
    s = 2 * 2 + a + b
    expect = (s - 1) / mpmath.mpf(2 * 2 * (2 + a + b) * (s - 2)) * (s * (s - 2) * x + b * b - a * a) * x \
        - 2 * (2 + a - 1) * (2 + b - 1) * s * 1
    assert abs(q[2] - expect) < mpmath.mpf(2) ** -250


def test_sample_point_counts_and_formulas():
    assert len(fe.create_sample_points(3, 4)) == comb(7, 4)
    assert all(sum(p) <= 1 for p in fe.create_sample_points(3, 4))
    assert [float(v) for v in fe.create_sample_points(2, 2)[1]] == [0.5, 0.0]      # first coordinate runs fastest
    for d in (4, 5, 12):
        assert len(fe.create_sample_points_2d(d)) == comb(d + 2, 2)
    assert len(fe.create_sample_points_3d(5)) == 6 * 7 * 8 // 6
    assert len(fe.points_X_general(3, 3)) == comb(6, 3)
    d = 9
    ch = fe.create_sample_points_chebyshev(d)
    assert len(ch) == d + 1 and all(abs(mpmath.chebyt(d + 1, v)) < mpmath.mpf(2) ** -240 for v in ch)
    chm = fe.create_sample_points_chebyshev_mod(d)
    assert abs(chm[0] - 1) < mpmath.mpf(2) ** -240 and abs(chm[-1] + 1) < mpmath.mpf(2) ** -240
    ch2 = fe.create_sample_points_chebyshev(3, 2, 6)
    assert all(2 < v < 6 for v in ch2)
    x1 = fe.create_sample_points_1d(6)
    const = -mpmath.sqrt(mpmath.pi) / (64 * 7 * mpmath.log(3 - 2 * mpmath.sqrt(2)))
    assert len(x1) == 7 and abs(x1[0] - const) < mpmath.mpf(2) ** -240 and abs(x1[3] - const * 121) < mpmath.mpf(2) ** -240


@pytest.mark.parametrize("d", [4, 7, 12])
def test_padua_points_are_unisolvent(d):
    pts = fe.create_sample_points_2d(d)
    # product Chebyshev basis of total degree <= d: well conditioned at the Padua points
    V = np.array([[float(mpmath.chebyt(a, p[0]) * mpmath.chebyt(e - a, p[1])) for e in range(d + 1) for a in range(e + 1)]
                  for p in pts])
    assert V.shape[0] == V.shape[1] == comb(d + 2, 2)
    assert np.linalg.cond(V) < 1e4


def _poly1(coeffs):
    return fe.Poly(1, {(i,): c for i, c in enumerate(coeffs)})


def test_prepareabc_default_path_reproduces_the_sphere_packing_generator():
    """The example's seven constraints (ex:56-105) built as polynomial matrices and passed through `prepareabc`
    give the same (A, B, c, H) as `instances.sphere_packing_2point`, which samples them directly."""
    prec, n, d, N = 256, 3, 4, 2
    ref_cons, _, _ = instances.sphere_packing_2point(n=n, d=d, prec=prec)
    mp = mpmath.mp.clone()
    mp.prec = prec + 64
    pi = mp.pi
    r = [mp.mpf(1), mp.sqrt(2) - 1]
    pairs = [(i, j) for i in range(N) for j in range(i + 1)]
    n_y = 1 + (2 * d + 1) * len(pairs)
    col = lambda k, pidx: 1 + k * len(pairs) + pidx
    X = fe.Poly.var(1, 0)
    q = fe.laguerrebasis(d, mp.mpf(n) / 2 - 1, 2 * pi * X)
    q = [p / max(p.terms.values()) for p in q]                          # ex:81-83 (divide by the maximum coefficient)
    lag = fe.laguerrebasis(2 * d, mp.mpf(n) / 2 - 1, pi * X)
    f = [lag[k] * (mp.factorial(k) / pi ** k) for k in range(2 * d + 1)]
    vol = lambda rr: mp.sqrt(pi) ** n / mp.gamma(mp.mpf(n) / 2 + 1) * rr ** n
    Z = lambda m: [[0] * m for _ in range(m)]

    def E(i, j, v):
        M = Z(N)
        M[i][j] = M[j][i] = v
        return M

    M0 = [[[-mp.sqrt(vol(r[i]) * vol(r[j])) for j in range(N)] for i in range(N)]] + [Z(N) for _ in range(n_y)]
    M1 = [Z(N)] + [Z(N) for _ in range(n_y)]
    for pidx, (i, j) in enumerate(pairs):
        M0[1 + col(0, pidx)] = E(i, j, 1)
        for k in range(2 * d + 1):
            M1[1 + col(k, pidx)] = E(i, j, X ** k)
    cons = [fe.prepareabc(M0, [fe.Poly.const(1, 1)], q, [[mp.mpf(0)]], 0, prec=prec),
            fe.prepareabc(M1, [fe.Poly.const(1, 1), X], q, [[v] for v in fe.create_sample_points_1d(2 * d, prec)], 2 * d,
                          prec=prec)]
    for pidx, (i, j) in enumerate(pairs):
        shift = (r[i] + r[j]) ** 2
        M2 = [[[0]]] + [[[0]] for _ in range(n_y)]
        for k in range(2 * d + 1):
            M2[1 + col(k, pidx)] = [[-f[k]]]
        xs = [[v + shift] for v in fe.create_sample_points_1d(2 * d, prec)]
        cons.append(fe.prepareabc(M2, [fe.Poly.const(1, 1), X - shift], q, xs, 2 * d, prec=prec))
    for i in range(N):
        pidx = pairs.index((i, i))
        M3 = [[[0]]] + [[[0]] for _ in range(n_y)]
        M3[1] = [[1]]
        for k in range(2 * d + 1):
            M3[1 + col(k, pidx)] = [[-f[k](mp.mpf(0))]]
        cons.append(fe.prepareabc(M3, [fe.Poly.const(1, 1)], q, [[mp.mpf(0)]], 0, prec=prec))
    cons = [cons[o - 1] for o in [3, 6, 5, 7, 4, 1, 2]]                 # ex:102
    assert len(cons) == len(ref_cons)
    for a, b in zip(cons, ref_cons):
        assert a.L == b.L and a.n_samples == b.n_samples
        assert rel_err_bits(a.B, b.B) >= prec - 6 and rel_err_bits(a.c, b.c) >= prec - 6
        for l in range(a.L):
            assert list(a.ranks[l]) == list(b.ranks[l]) and a.V[l].shape == b.V[l].shape
            assert rel_err_bits(a.V[l], b.V[l]) >= prec - 6 and rel_err_bits(a.H[l], b.H[l]) >= prec - 6
    bi_a, bi_b = solver.get_block_info(cons), solver.get_block_info(ref_cons)
    assert (bi_a.J, bi_a.n_y, list(bi_a.dim_S)) == (bi_b.J, bi_b.n_y, list(bi_b.dim_S))


@pytest.mark.parametrize("all_of_Pi", [True, False])
def test_prepareabc_pi_path_reconstructs_the_weighted_kronecker_form(all_of_Pi):
    """sum_r H_r v_r v_r^T = sign(G) |G| (Pi(x_k) (x) q q^T) restricted to the rows each Pi row may use
    (:256-281, :345-376), for a 2 x 2 polynomial Pi with rows of different degree."""
    prec = 256
    mp = mpmath.mp.clone()
    mp.prec = prec + 64
    X = fe.Poly.var(1, 0)
    q = fe.make_monomial_basis(1, 3)
    Pi = [[[X * X + 1, X], [X, fe.Poly.const(1, 2)]]]
    G = [2 - X]                                                        # negative at the last sample
    xs = [[mp.mpf(v)] for v in ("-0.5", "0.25", "1.5", "3")]
    M = [[[X]], [[fe.Poly.const(1, 1)]]]
    con = fe.prepareabc(M, G, q, xs, 6, Pi, prec=prec, all_of_Pi=all_of_Pi)
    assert list(con.ranks[0]) == [2, 2, 2, 2]
    # degree budgets: all_of_Pi: row 0 (degree 2) may use q up to degree (6-1-2)//2 = 1, row 1 up to (6-1-0)//2 = 2;
    # otherwise every row uses q up to degree (6-1-2)//2 = 1
    if all_of_Pi:
        idx = [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2)]
    else:
        idx = [(0, 0), (1, 0), (0, 1), (1, 1)]                         # basis index outer, Pi index inner
    assert con.V[0].shape == (8, len(idx))
    w = len(idx)
    for k, (xk,) in enumerate(xs):
        g = 2 - xk
        piv = [[xk * xk + 1, xk], [xk, mp.mpf(2)]]
        for a, (i, da) in enumerate(idx):
            for b, (j, db) in enumerate(idx):
                got = sum(con.H[0].to_mpf(2 * k + r) * con.V[0].to_mpf((2 * k + r) * w + a) * con.V[0].to_mpf((2 * k + r) * w + b)
                          for r in range(2))
                want = g * piv[i][j] * xk ** da * xk ** db
                assert abs(got - want) <= mpmath.mpf(2) ** -(prec - 12) * max(1, abs(want))
    assert rel_err_bits(con.c, type(con.c).from_mpf([v[0] for v in xs], prec // 32)) >= prec - 2
    assert all(con.B.to_mpf(k) == -1 for k in range(4))


def test_prepareabc_prunes_vanishing_ranks():
    """|H| <= threshold is dropped (:378-383): a singular Pi loses a rank everywhere, G(x_k) = 0 loses the sample."""
    prec = 256
    X = fe.Poly.var(1, 0)
    q = fe.make_monomial_basis(1, 1)
    one = fe.Poly.const(1, 1)
    Pi = [[[one, one], [one, one]], [[one, one * 0], [one * 0, one]]]
    G = [one, 1 - X]
    xs = [[mpmath.mpf(0)], [mpmath.mpf(1)], [mpmath.mpf(-1)]]
    con = fe.prepareabc([[[one]], [[X]]], G, q, xs, 2, Pi, prec=prec)
    assert list(con.ranks[0]) == [1, 1, 1]
    assert list(con.ranks[1]) == [2, 0, 2]
    assert con.V[0].shape[0] == 3 and con.V[1].shape[0] == 4
    bi = solver.get_block_info([con])
    assert bi.J == 1 and bi.n_y == 1


def test_solvempmp_univariate_minimum():
    """max t s.t. x^4 - x^2 + 1 - t >= 0 on [-1, 1]: the minimum 3/4 (at x^2 = 1/2), through solvempmp on the oracle."""
    prec = 256
    solver.set_precision(prec)
    X = fe.Poly.var(1, 0)
    M = [[[X ** 4 - X ** 2 + 1]], [[fe.Poly.const(1, -1)]]]
    G = [fe.Poly.const(1, 1), 1 - X * X]
    q = fe.make_monomial_basis(1, 2)
    xs = [[v] for v in fe.create_sample_points_chebyshev(4)]
    out = fe.solvempmp([M], [G], [q], [xs], [4], [mpmath.mpf(1)], handle=oracle_handle(prec, 2), verbose=False)
    assert abs(out[9] - mpmath.mpf(3) / 4) < mpmath.mpf(10) ** -12     # dual objective <b, y> = t
    assert abs(out[8] - mpmath.mpf(3) / 4) < mpmath.mpf(10) ** -12
    assert out[7] < mpmath.mpf(10) ** -15


def test_bivariate_matrix_program_structure_and_oracle_solve():
    """The BASELINE config 4 generator at reduced degree: structure (m = 2, L = 2, rank 2 everywhere, Pi index outer)
    and a converging oracle solve; the full-size instance is a GPU test."""
    prec = 256
    cons, b = instances.bivariate_matrix_program(D=2, n_y=6, clusters=2, prec=prec)
    bi = solver.get_block_info(cons)
    K = comb(6, 2)
    assert list(bi.dim_S) == [3 * K] * 2 and [list(r) for r in bi.Y_blocksizes] == [[2 * 2 * 6, 2 * 2 * 3]] * 2
    assert all(list(c.ranks[l]) == [2] * K for c in cons for l in range(2))
    solver.set_precision(prec)
    out, rows = solver.solverank1sdp(cons, b, bi, handle=oracle_handle(prec, 4), verbose=False, return_info=True)
    assert rows[-1].terminate == 3 and out[7] < mpmath.mpf(10) ** -15


def test_sdpb_format_roundtrip_and_limits(tmp_path):
    """clrsdp.sdpb_io (row f4; the example's missing WriteFilesSDPB.write_files, ex:97): the sphere-packing constraints
    (rank-1 samples, weights {1, x}: exactly SDPB's block form) survive the SDPB directory bit for bit; what SDPB's format
    cannot express (rank > 1 samples, L > 2, negative signs) is refused by name, a positive H != 1 is folded into the basis."""
    import json
    from clrsdp import instances, sdpb_io
    prec = 192
    solver.set_precision(prec)
    try:
        cons, b, _ = instances.sphere_packing_2point(n=3, d=3, prec=prec)
        bi = solver.get_block_info(cons)
        sdpb_io.write_sdpb(tmp_path / "sdp", cons, bi, b, b0="0")
        with open(tmp_path / "sdp" / "control.json") as f:
            assert json.load(f)["num_blocks"] == bi.J
        with open(tmp_path / "sdp" / "block_info_1.json") as f:
            info = json.load(f)
        assert info == dict(dim=bi.m[1], num_points=bi.n_samples[1])
        cons2, b2, b0 = sdpb_io.read_sdpb(tmp_path / "sdp")
        same = lambda a, o: np.array_equal(a.sign, o.sign) and np.array_equal(a.exp, o.exp) and np.array_equal(a.limb, o.limb)
        assert b0 == "0" and same(b2, b) and len(cons2) == len(cons)
        for a, o in zip(cons2, cons):
            assert same(a.B, o.B) and same(a.c, o.c) and a.L == o.L
            for l in range(o.L):
                assert same(a.V[l], o.V[l]) and same(a.H[l], o.H[l]) and list(a.ranks[l]) == list(o.ranks[l])
        bi2 = solver.get_block_info(cons2)
        assert (bi2.J, bi2.n_y, list(bi2.dim_S)) == (bi.J, bi.n_y, list(bi.dim_S))
        # H = 4 is folded into the basis as sqrt(H) = 2: the block read back has H = 1 and vectors twice as long
        c0 = cons[1]
        four = MpArray.from_mpf([mpmath.mpf(4)] * c0.H[0].n, prec // 32)
        scaled = [solver.Constraint(V=c0.V, ranks=c0.ranks, H=[four] + list(c0.H[1:]), B=c0.B, c=c0.c)]
        sdpb_io.write_sdpb(tmp_path / "h4", scaled, solver.get_block_info(scaled), b)
        back = sdpb_io.read_sdpb(tmp_path / "h4")[0][0]
        assert [2 * v for v in c0.V[0].reshape(c0.V[0].n).to_fractions()] == back.V[0].reshape(back.V[0].n).to_fractions()
        # not SDPB's block form
        gen, bg = instances.random_structured_sdp([dict(m=1, K=3, blocks=[dict(delta=3, ranks=[2, 1, 1])])], n_y=2, prec=prec)
        with pytest.raises(ValueError, match="rank"):
            sdpb_io.write_sdpb(tmp_path / "bad", gen, solver.get_block_info(gen), bg)
    finally:
        solver.set_precision(256)

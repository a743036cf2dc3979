"""GPU parity tests of the phase-level entry points, called through the C ABI.

Bar (north star): bit-exact for the integer work (digit planes of the sliced GEMM), and within
2^-(p-16) of the oracle / exact arithmetic for the floating-point results — the tolerance is written
next to every assertion."""
import os
import random
from fractions import Fraction

import numpy as np
import pytest

from clrsdp import solver
from clrsdp.wire import MpArray, rel_err_bits
from gpu_common import exact_planes, rand_mp, spd_batch
from oracle.ref import oracle_handle

pytestmark = pytest.mark.gpu
PRECS = [128, 256, 384, 512]


@pytest.fixture(scope="module", params=PRECS)
def handles(request):
    prec = request.param
    return prec, solver.product_handle(prec), oracle_handle(prec, 8)


def test_elementwise_ops_within_half_ulp(handles):
    prec, h, ho = handles
    rng = random.Random(prec)
    a, b = rand_mp(rng, 600, h.nlimb, zero_frac=0.03), rand_mp(rng, 600, h.nlimb)
    fa, fb = a.to_fractions(), b.to_fractions()
    for op, f in (("+", lambda x, y: x + y), ("-", lambda x, y: x - y), ("*", lambda x, y: x * y), ("/", lambda x, y: x / y)):
        g = h.op_elementwise(op, a, b)
        for i in range(a.n):
            ex = f(fa[i], fb[i])
            got = g.to_fraction(i)
            if ex == 0:
                assert got == 0
            else:
                assert abs(got - ex) <= abs(ex) * Fraction(1, 2 ** (prec - 1))   # <= 1 ulp (device ops are 0.5 ulp)
    sq = a.view()
    sq.sign = np.abs(a.sign)
    g, o = h.op_elementwise("s", sq, b), ho.op_elementwise("s", sq, b)
    assert rel_err_bits(g, o) >= prec - 1


@pytest.mark.parametrize("shape", [(1, 8, 8, 8), (2, 20, 12, 40), (1, 3, 5, 33), (2, 1, 1, 1)])
def test_sliced_gemm_digit_planes_are_bit_exact(handles, shape):
    """tcgen05 kind::i8 planes == python big-int model of slice + digit products (exact, no tolerance)."""
    prec, h, _ = handles
    batch, M, N, K = shape
    rng = random.Random(sum(shape) + prec)
    A, B = rand_mp(rng, batch * M * K, h.nlimb, zero_frac=0.05), rand_mp(rng, batch * K * N, h.nlimb, erange=20)
    if M > 1:
        A.sign[:K] = 0          # an all-zero row
        A.limb[:, :K] = 0
    planes, rexp, cexp = h.op_gemm_planes(batch, M, N, K, A, B)
    assert planes.shape[0] == prec // 8 + 2
    ex = exact_planes(A, B, batch, M, N, K, rexp, cexp, planes.shape[0])
    assert np.array_equal(planes.astype(object), ex)


@pytest.mark.parametrize("shape", [(1, 130, 70, 64), (3, 64, 64, 64), (1, 16, 200, 130), (2, 33, 17, 300), (1, 40, 300, 96)])
def test_gemm_matches_oracle(handles, shape):
    prec, h, ho = handles
    batch, M, N, K = shape
    rng = random.Random(prec + K)
    A, B = rand_mp(rng, batch * M * K, h.nlimb), rand_mp(rng, batch * K * N, h.nlimb)
    C, Co = h.op_gemm(batch, M, N, K, A, B), ho.op_gemm(batch, M, N, K, A, B)
    # normwise: error relative to the largest entry of the product, tolerance 2^-(p-16)
    assert rel_err_bits(C, Co) >= prec - 16


def test_gemm_split_k_and_exponent_spread(handles):
    """K large enough to need several K-splits (int32 exactness bound) and rows whose entries span
    hundreds of binades (block fixed point must stay normwise accurate)."""
    prec, h, ho = handles
    rng = random.Random(3)
    batch, M, N, K = 1, 24, 24, 4500
    A, B = rand_mp(rng, batch * M * K, h.nlimb, erange=3), rand_mp(rng, batch * K * N, h.nlimb, erange=3)
    C, Co = h.op_gemm(batch, M, N, K, A, B), ho.op_gemm(batch, M, N, K, A, B)
    assert rel_err_bits(C, Co) >= prec - 16
    M = N = K = 20
    A, B = rand_mp(rng, M * K, h.nlimb, erange=300), rand_mp(rng, K * N, h.nlimb, erange=300)
    C, Co = h.op_gemm(1, M, N, K, A, B), ho.op_gemm(1, M, N, K, A, B)
    fa, fb = A.to_fractions(), B.to_fractions()
    fc, fo = C.to_fractions(), Co.to_fractions()
    for i in range(M):
        for j in range(N):
            # bound: row-max(A) * col-max(B) * K * 2^-(p-8)
            scale = max(abs(fa[i * K + k]) for k in range(K)) * max(abs(fb[k * N + j]) for k in range(K)) * K
            assert abs(fc[i * N + j] - fo[i * N + j]) <= scale / 2 ** (prec - 8)


def test_gemm_of_zero_matrix(handles):
    prec, h, _ = handles
    Z = MpArray(16 * 16, h.nlimb)
    rng = random.Random(1)
    B = rand_mp(rng, 16 * 16, h.nlimb)
    C = h.op_gemm(1, 16, 16, 16, Z, B)
    assert not C.sign.any()


@pytest.mark.parametrize("shape", [(2, 1), (2, 5), (3, 33), (2, 64), (1, 130), (1, 260)])
def test_cholesky_and_inverse_factor(handles, shape):
    prec, h, ho = handles
    batch, n = shape
    rng = random.Random(n)
    A = spd_batch(rng, batch, n, h.nlimb)
    L, Li = h.op_cholesky(batch, n, A)
    Lo, Lio = ho.op_cholesky(batch, n, A)
    assert rel_err_bits(L, Lo) >= prec - 16
    assert rel_err_bits(Li, Lio) >= prec - 16
    # upper triangle is exactly zero
    for b in range(batch):
        for i in range(n):
            for j in range(i + 1, n):
                assert L.sign[(b * n + i) * n + j] == 0 and Li.sign[(b * n + i) * n + j] == 0


def test_cholesky_reports_not_positive_definite(handles):
    from clrsdp.capi import ClrsdpError
    prec, h, _ = handles
    A = MpArray.from_double(np.array([[1.0, 2.0], [2.0, 1.0]]).reshape(-1), h.nlimb)
    with pytest.raises(ClrsdpError) as e:
        h.op_cholesky(1, 2, A)
    assert e.value.code == -10 and "higher precision" in str(e.value)


@pytest.mark.parametrize("shape", [(3, 1), (3, 2), (2, 7), (2, 40), (2, 64), (1, 100)])
def test_lambda_min_matches_oracle(handles, shape):
    prec, h, ho = handles
    batch, n = shape
    rng = random.Random(n + 1)
    mats = []
    for _ in range(batch):
        G = np.array([[rng.uniform(-1, 1) for _ in range(n)] for _ in range(n)])
        mats.append((G + G.T) / 2)
    A = MpArray.from_double(np.array(mats).reshape(-1), h.nlimb)
    lg, lo = h.op_lambda_min(batch, n, A), ho.op_lambda_min(batch, n, A)
    ref = np.array([np.linalg.eigvalsh(m)[0] for m in mats])
    assert np.allclose(lg.to_double(), ref, rtol=1e-10)
    assert rel_err_bits(lg, lo) >= prec - 16


def test_lambda_min_degenerate_spectra(handles):
    """repeated / clustered smallest eigenvalues and an already-diagonal matrix (Newton polish must fall back)"""
    prec, h, ho = handles
    n = 6
    D = np.diag([-2.0, -2.0, -2.0, 1.0, 3.0, 3.0])
    rng = np.random.default_rng(0)
    Qm, _ = np.linalg.qr(rng.normal(size=(n, n)))
    mats = [D, Qm @ D @ Qm.T, np.eye(n) * 0.5, np.diag([1e-30, 1.0, 2.0, 3.0, 4.0, 5.0])]
    mats = [(m + m.T) / 2 for m in mats]
    A = MpArray.from_double(np.array(mats).reshape(-1), h.nlimb)
    lg = h.op_lambda_min(len(mats), n, A)
    got = lg.to_double()
    assert got[0] == pytest.approx(-2.0, abs=1e-12)
    assert got[1] == pytest.approx(-2.0, abs=1e-12)
    assert got[2] == pytest.approx(0.5, abs=1e-12)
    assert got[3] == pytest.approx(1e-30, rel=1e-9)


@pytest.mark.parametrize("shape", [(2, 7), (2, 33), (1, 70), (1, 130)])
def test_signed_factor_inverts_indefinite_matrices(handles, shape):
    """The factorisation of S_j and Q (stand-in for the reference's pivoted LU, MPMP.jl:1436,1501): symmetric INDEFINITE
    matrices - what a Schur complement that is singular to working precision looks like - are factored as
    A = D U^T Sigma U D without any pivot being reported or perturbed; M^T Sigma M is the inverse, checked against the
    identity in exact rational arithmetic relative to cond(A): |M^T Sigma M A - I| <= cond * 2^-(p-16)."""
    prec, h, _ = handles
    batch, n = shape
    rng = np.random.default_rng(n)
    mats = []
    for _ in range(batch):
        Qm, _ = np.linalg.qr(rng.normal(size=(n, n)))
        ev = rng.uniform(0.5, 2.0, size=n) * np.where(np.arange(n) % 3 == 1, -1.0, 1.0)   # a third of the spectrum is negative
        mats.append((Qm * ev) @ Qm.T)
        mats[-1] = (mats[-1] + mats[-1].T) / 2
    A = MpArray.from_double(np.array(mats).reshape(-1), h.nlimb)
    M, sg = h.op_signed_factor(batch, n, A)
    assert set(np.unique(sg)) <= {-1, 1}
    import mpmath
    with mpmath.workprec(prec + 64):
        for b in range(batch):
            # inertia (Sylvester): the number of negative pivots equals the number of negative eigenvalues
            assert int((sg[b] < 0).sum()) == int((np.linalg.eigvalsh(mats[b]) < 0).sum())
            Mm = mpmath.matrix(n, n)
            vals = M.take(np.arange(b * n * n, (b + 1) * n * n)).to_mpfs()
            for i in range(n):
                for j in range(n):
                    Mm[i, j] = vals[i * n + j]
                    if j > i:
                        assert vals[i * n + j] == 0          # lower triangular
            Am = mpmath.matrix(mats[b].tolist())
            Sg = mpmath.diag([int(s) for s in sg[b]])
            R = Mm.T * Sg * Mm * Am - mpmath.eye(n)
            err = max(abs(R[i, j]) for i in range(n) for j in range(n))
            # unpivoted LDL^T of an indefinite matrix: element growth enters the bound (measured < 2^30 on these seeds)
            assert err <= mpmath.mpf(2) ** -(prec - 16 - 40), (b, mpmath.nstr(err, 5))


@pytest.mark.parametrize("signed", [False, True])
def test_lookahead_and_cluster_panels_do_not_change_the_factorisation(signed):
    """The blocked factorisation with the trailing update split into (next row block | remainder on a helper stream)
    applies the same contributions to every row block, in the same order, as the unsplit update: the factors agree to
    rounding (the CUDA-core products choose their K-split by the size of the product, so the summation order inside one
    contribution may differ - not bit-identical) and match the oracle to p - 16 bits. The panel kernel on a thread-block
    cluster applies exactly the same operations per entry as one CTA per matrix: BIT-IDENTICAL. 5 panels, so that the helper
    stream of the remainders, the helper stream of the inverse panels and the pivot chain all overlap."""
    from clrsdp import solver
    from oracle.ref import oracle_handle
    prec, batch, n = 256, 2, 160
    A = spd_batch(random.Random(5), batch, n, prec // 32)
    keys = ("CLRSDP_LOOKAHEAD", "CLRSDP_LOOKAHEAD_RATIO", "CLRSDP_LOOKAHEAD_MIN_PMAC", "CLRSDP_PANEL_CLUSTER")
    saved = {k: os.environ.get(k) for k in keys}
    results = []
    try:
        for env in ({"CLRSDP_LOOKAHEAD": "0"},
                    {"CLRSDP_LOOKAHEAD": "1", "CLRSDP_LOOKAHEAD_RATIO": "0", "CLRSDP_LOOKAHEAD_MIN_PMAC": "0"},
                    {"CLRSDP_LOOKAHEAD": "1", "CLRSDP_LOOKAHEAD_RATIO": "0", "CLRSDP_LOOKAHEAD_MIN_PMAC": "0",
                     "CLRSDP_PANEL_CLUSTER": "8"}):
            for k in keys:
                os.environ.pop(k, None)
            os.environ.update(env)
            h = solver.product_handle(prec, 0)      # the switches are read when the handle is created / per launch
            results.append(h.op_signed_factor(batch, n, A) if signed else h.op_cholesky(batch, n, A))
            del h
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    plain, ahead, ahead_cluster = results
    for x, y in zip(ahead, ahead_cluster):          # cluster panels: bit-identical
        if isinstance(x, np.ndarray):
            assert np.array_equal(x, y)
        else:
            assert np.array_equal(x.limb, y.limb) and np.array_equal(x.exp, y.exp) and np.array_equal(x.sign, y.sign)
    for x, y in zip(plain, ahead):                  # lookahead: the same factors to rounding (2^-(p-16) relative)
        if isinstance(x, np.ndarray):
            assert np.array_equal(x, y)             # pivot signs
        else:
            assert rel_err_bits(x, y) >= prec - 16
    if not signed:
        Lo, Lio = oracle_handle(prec, 2).op_cholesky(batch, n, A)
        for r in (plain, ahead):
            assert rel_err_bits(r[0], Lo) >= prec - 16 and rel_err_bits(r[1], Lio) >= prec - 16


def test_malformed_wire_tensors_and_short_buffers_are_rejected():
    """Boundary hygiene (advisor, round 1): an exponent outside the header word's range or a mantissa whose top bit is
    clear is CLRSDP_ERR_BAD_ARG, not a silently wrapped number; result buffers that are too small are refused before
    anything is written."""
    from clrsdp.capi import ClrsdpError
    h = solver.product_handle(256)
    a = MpArray.from_double(np.array([1.5, -2.25, 3.0]), h.nlimb)
    big = MpArray.from_double(np.array([1.5, -2.25, 3.0]), h.nlimb)
    big.exp[1] = 1 << 40
    with pytest.raises(ClrsdpError) as e:
        h.op_elementwise("+", big, a)
    assert e.value.code == -1
    den = MpArray.from_double(np.array([1.5, -2.25, 3.0]), h.nlimb)
    den.limb[h.nlimb - 1, 2] &= 0x7FFFFFFF
    with pytest.raises(ClrsdpError) as e:
        h.op_elementwise("*", a, den)
    assert e.value.code == -1
    with pytest.raises(ClrsdpError) as e:                       # 2 matrices of 4 x 4 announced, 3 numbers given
        h.op_lambda_min(2, 4, a)
    assert e.value.code == -1

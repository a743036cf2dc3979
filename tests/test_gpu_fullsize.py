"""BASELINE config 3 at full size (64 clusters, block 64, 128 samples, n_y = 256, 256 bit) on the GPU: the oracle
would need minutes per iteration here, so the checks are size-independent properties of the quantities the
iteration produces (the small-instance parity against the oracle is in test_gpu_solver.py):

  * X^-1 X = I and U^T-free reconstruction L^-1 X L^-T = I for sampled blocks (host mpmath, p - 40 bits: the blocks
    are well conditioned at the start of the solve),
  * Q is symmetric (it is computed as a general product, MPMP.jl:1467-1495) to p - 20 bits,
  * the sliced GEMM is linear: A (B + C) = A B + A C at 64 x (64 x 64 x 64), exact rational check of sampled entries,
  * the iteration makes progress: mu decreases, 0 < alpha <= 1, errors finite, status 0."""
import random

import mpmath
import numpy as np
import pytest

from clrsdp import instances, solver
from clrsdp.wire import MpArray, rel_err_bits
from gpu_common import rand_mp

pytestmark = pytest.mark.gpu
PREC = 256


@pytest.fixture(scope="module")
def cfg3():
    cons, b, _ = instances.synthetic_clustered_sdp(J=64, delta=64, K=128, n_y=256, prec=PREC, seed=20261018)
    bi = solver.get_block_info(cons)
    h = solver.product_handle(PREC)
    solver.load_problem(h, cons, b, bi)
    h.set_params(solver.real_params(h.nlimb))
    h.init_point()
    h.prepare()
    rows = [h.iterate() for _ in range(3)]
    return h, bi, rows


def _mat(a: MpArray, n):
    return mpmath.matrix([[a.to_mpf(r * n + c) for c in range(n)] for r in range(n)])


def test_iterations_make_progress(cfg3):
    h, bi, rows = cfg3
    assert all(r.status == 0 for r in rows)
    assert rows[0].mu > rows[1].mu > rows[2].mu > 0
    for r in rows:
        assert 0 < r.alpha_p <= 1 and 0 < r.alpha_d <= 1
        assert np.isfinite(r.P_err) and np.isfinite(r.p_err) and np.isfinite(r.d_err)
    assert h.launch_count() > 0


def test_inverse_and_factor_identities_on_sampled_blocks(cfg3):
    h, bi, rows = cfg3
    mpmath.mp.prec = PREC + 64
    n = 64
    # X^-1 and L^-1 of the last iteration were made from the X before its update: X_old = X - alpha_p dX
    for j in (0, 37):
        X = _mat(h.fetch("X", j, 0), n)
        dX = _mat(h.fetch("dX", j, 0), n)
        alpha = h.scalar("alpha_p")
        Xold = X - alpha * dX
        Xinv = _mat(h.fetch("Xinv", j, 0), n)
        E = Xinv * Xold - mpmath.eye(n)
        err = max(abs(E[r, c]) for r in range(n) for c in range(n))
        assert err < mpmath.mpf(2) ** -(PREC - 40), mpmath.nstr(err, 5)
        Li = _mat(h.fetch("Linvx", j, 0), n)
        F = Li * Xold * Li.T - mpmath.eye(n)
        err = max(abs(F[r, c]) for r in range(n) for c in range(n))
        assert err < mpmath.mpf(2) ** -(PREC - 40), mpmath.nstr(err, 5)


def test_Q_is_symmetric(cfg3):
    h, bi, rows = cfg3
    q = h.fetch("Q")
    n = bi.n_y
    idx = np.arange(n * n).reshape(n, n)
    assert rel_err_bits(q.take(idx.reshape(-1)), q.take(idx.T.reshape(-1))) >= PREC - 20


def test_sliced_gemm_is_linear_at_block_size():
    h = solver.product_handle(PREC)
    rng = random.Random(7)
    batch, n = 64, 64
    A = rand_mp(rng, batch * n * n, h.nlimb)
    B = rand_mp(rng, batch * n * n, h.nlimb)
    C = rand_mp(rng, batch * n * n, h.nlimb)
    BC = h.op_elementwise("+", B, C)
    lhs = h.op_gemm(batch, n, n, n, A, BC)
    rhs = h.op_elementwise("+", h.op_gemm(batch, n, n, n, A, B), h.op_gemm(batch, n, n, n, A, C))
    # block fixed point: errors are relative to rowmax * colmax * K, so compare normwise per sampled entry
    scale = max(abs(lhs.to_fraction(i)) for i in range(0, lhs.n, 4099))
    for i in range(0, lhs.n, 4099):
        assert abs(lhs.to_fraction(i) - rhs.to_fraction(i)) <= scale / 2 ** (PREC - 16)

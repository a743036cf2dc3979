"""BASELINE config 3 at full size (64 clusters, block 64, 128 samples, n_y = 256, 256 bit) on the GPU.

Oracle parity AT FULL SIZE: `test_cfg3_full_size_iterations_match_oracle` runs two whole iterations of the headline
bench workload on the GPU and on the oracle (a few seconds per iteration on the box's host cores) and compares every
field; `test_cfg5_shape_iterations_match_oracle_512bit` does the same on a cluster-subsampled instance of BASELINE
config 5's shape (block 128, 256 samples) at 512 bits. Beside them, size-independent properties of the
quantities the iteration produces:

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


def _pair(cons, b, bi, prec, env=None):
    """GPU handle, oracle handle, and the oracle at p + 64 bits on the same (exactly widened) instance: the arbiter for
    quantities whose conditioning exceeds 2^16 (test_gpu_solver.agree)."""
    import os
    from oracle.ref import oracle_handle
    from test_gpu_solver import widen_problem
    hg = solver.product_handle(prec)
    for k, v in (env or {}).items():
        os.environ[k] = v
    try:
        ho = oracle_handle(prec, os.cpu_count() or 1)
        ht = oracle_handle(prec + 64, os.cpu_count() or 1)
    finally:
        for k in (env or {}):
            os.environ.pop(k, None)
    wc, wb = widen_problem(cons, b, 2)
    for h, (c_, b_) in ((hg, (cons, b)), (ho, (cons, b)), (ht, (wc, wb))):
        solver.load_problem(h, c_, b_, bi)
        h.set_params(solver.real_params(h.nlimb))
        h.init_point()
        h.prepare()
    return hg, ho, ht


def test_cfg3_full_size_iterations_match_oracle():
    """The headline bench workload itself (BASELINE config 3: J = 64, block 64, K = 128, n_y = 256, 256 bit, the bench's
    seed): two iterations on the GPU and on the oracle, compared field by field at relative 2^-(p-16) (where the conditioning
    of the instance pushes both arithmetics past that: GPU no worse than the oracle against the same iteration at p + 64
    bits, test_gpu_solver.agree): all vectors
    (d, dx, dy, x, y of both the predictor and the corrector), p, Q (65 536 entries), the driver scalars and objectives
    in full; S_j and the thirteen block fields for the clusters 0, 21, 42 and 63."""
    from test_gpu_solver import compare_iteration
    cons, b, _ = instances.synthetic_clustered_sdp(J=64, delta=64, K=128, n_y=256, prec=PREC, seed=20261018)
    bi = solver.get_block_info(cons)
    hg, ho, ht = _pair(cons, b, bi, PREC)
    for it in range(2):
        rg, ro, rt = hg.iterate(), ho.iterate(), ht.iterate()
        assert rg.status == 0 and ro.status == 0 and rt.status == 0
        compare_iteration(hg, ho, bi, PREC, PREC - 16, ht, clusters=(0, 21, 42, 63))
        assert rg.pd_feasible == ro.pd_feasible and rg.terminate == ro.terminate
        assert rg.alpha_p == pytest.approx(ro.alpha_p, rel=1e-13) and rg.alpha_d == pytest.approx(ro.alpha_d, rel=1e-13)


def test_cfg5_shape_iterations_match_oracle_512bit():
    """BASELINE config 5's block and cluster shape (block 128, K = 256, dim_S = 256) at its precision, 512 bits, on a
    cluster subsample: J = 3 of the 512 clusters (the configuration's own seed) with n_y = 512 - the free variables are
    halved with the clusters so that sum dim_S = 768 > n_y and Q = B^T S^-1 B stays non-singular (with J = 2 and
    n_y = 1024 it has rank 512 and the oracle's own dy is noise). Two iterations GPU against oracle, every field at
    relative 2^-(p-16), with the oracle at p + 64 bits as the arbiter where the conditioning exceeds 2^16. The oracle
    runs its block fixed-point products here (CLRSDP_REF_GEMM=fixed: exact integer dot products, one rounding per entry -
    as accurate as its fma loops, test_oracle_pin.py, and 2-3x faster: ~20 s per iteration at this shape)."""
    from test_gpu_solver import compare_iteration
    prec = 512
    cons, b, _ = instances.synthetic_clustered_sdp(J=3, delta=128, K=256, n_y=512, prec=prec, seed=20261019)
    bi = solver.get_block_info(cons)
    assert list(bi.dim_S) == [256] * 3 and bi.n_y == 512 and bi.Y_blocksizes[0][0] == 128
    hg, ho, ht = _pair(cons, b, bi, prec, env={"CLRSDP_REF_GEMM": "fixed"})
    for it in range(2):
        rg, ro, rt = hg.iterate(), ho.iterate(), ht.iterate()
        assert rg.status == 0 and ro.status == 0 and rt.status == 0
        compare_iteration(hg, ho, bi, prec, prec - 16, ht)

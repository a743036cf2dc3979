"""GPU parity of the interior-point iteration against the oracle, through the C ABI (host-side mirror
`clrsdp.solver`). Tolerance (north star): relative 2^-(p-16) on single iterations; identical iteration
counts and objectives to 2^-(p-16) * condition slack on full solves (slack stated per test)."""
import mpmath
import numpy as np
import pytest

from clrsdp import instances, solver
from clrsdp.capi import ClrsdpError
from clrsdp.wire import rel_err_bits
from oracle.ref import oracle_handle

pytestmark = pytest.mark.gpu

GENERAL_SPEC = [dict(m=2, K=5, blocks=[dict(delta=3, ranks=[2, 1, 0, 2, 1]), dict(delta=2, ranks=[1, 1, 1, 1, 1])]),
                dict(m=1, K=4, blocks=[dict(delta=3, ranks=[1, 2, 1, 1])]),
                dict(m=3, K=3, blocks=[dict(delta=2, ranks=[2, 2, 1])]),
                dict(m=2, K=5, blocks=[dict(delta=3, ranks=[1, 1, 2, 0, 2]), dict(delta=2, ranks=[1, 1, 1, 1, 1])])]

BLOCK_FIELDS = ["XY", "R", "Xinv", "Px", "Py", "P", "Z", "dX_pred", "dY_pred", "dX", "dY", "X", "Y"]
VEC_FIELDS = ["d", "dx_pred", "dy_pred", "dx", "dy", "x", "y"]


def pair(cons, b, bi, prec, **params):
    hs = [solver.product_handle(prec), oracle_handle(prec, 8)]
    for h in hs:
        solver.load_problem(h, cons, b, bi)
        h.set_params(solver.real_params(h.nlimb, **params))
        h.init_point()
        h.prepare()
    return hs


def widen_problem(cons, b, extra_limbs):
    """the same instance at a higher working precision (exact)."""
    from clrsdp.solver import Constraint
    nl = b.nlimb + extra_limbs
    wc = [Constraint(V=[v.widen(nl) for v in c.V], ranks=c.ranks, H=[h.widen(nl) for h in c.H], B=c.B.widen(nl),
                     c=c.c.widen(nl)) for c in cons]
    return wc, b.widen(nl)


def agree(a, o, tol_bits, truth=None, **kw):
    """GPU value `a` against the oracle's `o`: 2^-tol_bits relative; where the conditioning of the instance
    pushes BOTH arithmetics past that, the GPU must be no worse than 4x the MPFR error, both measured
    against the same computation at p+64 bits (`truth`) - the oracle's own distance from the arbiter is what the
    instance allows at this precision (BASELINE config 3 has exactly singular Schur complements: the oracle itself is
    140 bits from the arbiter there) - and never below a quarter of the working precision."""
    bits = rel_err_bits(a, o, **kw)
    LAST.clear()
    LAST["gpu_vs_oracle_bits"] = round(bits, 1)
    if bits >= tol_bits or truth is None:
        return bits >= tol_bits
    eg, eo = rel_err_bits(a, truth, **kw), rel_err_bits(o, truth, **kw)
    LAST.update(gpu_vs_truth_bits=round(eg, 1), oracle_vs_truth_bits=round(eo, 1))
    return eg >= eo - 2 and eg >= (tol_bits + 16) / 4


LAST = {}   # the numbers behind the last agree() verdict, shown in assertion messages


def compare_iteration(hg, ho, bi, prec, tol_bits, ht=None, clusters=None):
    """Every quantity one iteration produces, GPU against oracle. `clusters`: restrict the per-block / per-cluster fields
    to these clusters (full-size instances: the vectors, Q and the scalars are always compared in full)."""
    for name in VEC_FIELDS:
        assert agree(hg.fetch(name), ho.fetch(name), tol_bits, ht.fetch(name) if ht else None), (name, dict(LAST))
    # p = b - B^T x cancels to rounding level once the primal step is complete: its error scale is
    # |b| + |B|^T |x| (all generators draw |B_ij| < 1), not |p| itself
    from fractions import Fraction
    xs = ho.fetch("x").to_fractions()
    pscale = max(abs(v) for v in ho.fetch("b").to_fractions()) + sum(abs(v) for v in xs)
    assert agree(hg.fetch("p"), ho.fetch("p"), tol_bits, ht.fetch("p") if ht else None, scale=pscale), ("p", dict(LAST))
    for j in (range(bi.J) if clusters is None else clusters):
        assert agree(hg.fetch("S", j), ho.fetch("S", j), tol_bits, ht.fetch("S", j) if ht else None), ("S", j, dict(LAST))
        for l in range(bi.L[j]):
            for name in BLOCK_FIELDS:
                a, o = hg.fetch(name, j, l), ho.fetch(name, j, l)
                if name == "P":
                    xs = max(abs(v) for v in ho.fetch("X", j, l).to_double().reshape(-1))
                    if max(abs(v) for v in o.to_double().reshape(-1)) < xs * 2.0 ** -(prec - 40):
                        continue   # residual at rounding level of X: nothing to compare
                assert agree(a, o, tol_bits, ht.fetch(name, j, l) if ht else None), (name, j, l, dict(LAST))
    assert agree(hg.fetch("Q"), ho.fetch("Q"), tol_bits, ht.fetch("Q") if ht else None), ("Q", dict(LAST))
    with mpmath.workprec(prec + 32):
        for s in ("mu", "lambda_x", "lambda_y", "alpha_p", "alpha_d", "beta_c"):
            a, o = hg.scalar(s), ho.scalar(s)
            if abs(a - o) <= abs(o) * mpmath.mpf(2) ** -tol_bits:
                continue
            assert ht is not None, (s, a, o)            # beyond 2^-tol: the GPU must be as close to the arbiter as the oracle
            t = ht.scalar(s)
            assert abs(a - t) <= 4 * abs(o - t) and abs(a - t) <= abs(t) * mpmath.mpf(2) ** -((tol_bits + 16) // 4), (s, a, o, t)
        # objectives are inner products with cancellation: the error scale is sum |c_i x_i| (resp. |b_i y_i|)
        for s, u, v in (("p_obj", "c", "x"), ("d_obj", "b", "y")):
            scale = sum(abs(pp * qq) for pp, qq in zip(ho.fetch(u).to_mpfs(), ho.fetch(v).to_mpfs()))
            a, o = hg.scalar(s), ho.scalar(s)
            if abs(a - o) <= max(scale, abs(o)) * mpmath.mpf(2) ** -tol_bits:
                continue
            assert ht is not None, (s, a, o)
            t = ht.scalar(s)
            assert abs(a - t) <= 4 * abs(o - t), (s, a, o, t)


@pytest.mark.parametrize("prec", [128, 256, 512])
def test_iterations_match_oracle_rank1(prec):
    cons, b, _ = instances.synthetic_clustered_sdp(J=3, delta=8, K=12, n_y=7, prec=prec)
    bi = solver.get_block_info(cons)
    hg, ho = pair(cons, b, bi, prec)
    for it in range(3):
        rg, ro = hg.iterate(), ho.iterate()
        assert rg.status == 0 and ro.status == 0
        compare_iteration(hg, ho, bi, prec, prec - 16)          # relative 2^-(p-16)
        assert rg.pd_feasible == ro.pd_feasible and rg.terminate == ro.terminate


@pytest.mark.parametrize("positive_H", [True, False])
def test_iterations_match_oracle_general_structure(positive_H):
    """m > 1, several blocks per cluster, rank > 1, a sample of rank 0, mixed-sign H, ragged shapes."""
    prec = 256
    cons, b = instances.random_structured_sdp(GENERAL_SPEC, n_y=4, prec=prec, positive_H=positive_H)
    bi = solver.get_block_info(cons)
    hg, ho = pair(cons, b, bi, prec)
    # the same iteration at p+64 bits: arbiter for the quantities whose conditioning exceeds 2^16
    wc, wb = widen_problem(cons, b, 2)
    ht = oracle_handle(prec + 64, 8)
    solver.load_problem(ht, wc, wb, bi)
    ht.set_params(solver.real_params(ht.nlimb))
    ht.init_point()
    ht.prepare()
    for it in range(2):
        rg, ro, rt = hg.iterate(), ho.iterate(), ht.iterate()
        assert rg.status == 0 and ro.status == 0 and rt.status == 0
        compare_iteration(hg, ho, bi, prec, prec - 16, ht)


def test_prepare_matches_oracle():
    prec = 256
    cons, b = instances.random_structured_sdp(GENERAL_SPEC, n_y=4, prec=prec)
    bi = solver.get_block_info(cons)
    hs = [solver.product_handle(prec), oracle_handle(prec, 4)]
    infos = []
    for h in hs:
        solver.load_problem(h, cons, b, bi, b0=3)
        h.set_params(solver.real_params(h.nlimb, omega_p=100, omega_d=7))
        h.init_point()
        infos.append(h.prepare())
    g, o = infos
    for k in ("mu", "p_obj", "d_obj", "gap", "P_err", "p_err", "d_err", "primal_err_new", "dual_err_new"):
        assert getattr(g, k) == pytest.approx(getattr(o, k), rel=1e-13), k
    assert rel_err_bits(hs[0].fetch("d"), hs[1].fetch("d")) >= prec - 16
    assert rel_err_bits(hs[0].fetch("p"), hs[1].fetch("p")) >= prec - 16


@pytest.mark.parametrize("case", ["rank1", "general"])
def test_full_solve_same_iteration_count_and_objectives(case):
    prec = 256
    if case == "rank1":
        cons, b, _ = instances.synthetic_clustered_sdp(J=3, delta=6, K=10, n_y=5, prec=prec)
    else:
        cons, b = instances.random_structured_sdp(GENERAL_SPEC, n_y=4, prec=prec)
    bi = solver.get_block_info(cons)
    og, rg = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True)
    oo, ro = solver.solverank1sdp(cons, b, bi, handle=oracle_handle(prec, 8), verbose=False, return_info=True)
    assert len(rg) == len(ro)                                    # identical iteration count
    assert rg[-1].terminate == ro[-1].terminate == 3
    for a, o in zip(rg, ro):
        assert a.alpha_p == pytest.approx(o.alpha_p, rel=1e-12) and a.alpha_d == pytest.approx(o.alpha_d, rel=1e-12)
    with mpmath.workprec(prec):
        # objectives after ~60-80 iterations: rounding differences are amplified by the conditioning of X, S
        # (up to ~2^60 near the optimum); tolerance 2^-(p-16-64)
        tol = mpmath.mpf(2) ** -(prec - 16 - 64)
        assert abs(og[8] - oo[8]) <= abs(oo[8]) * tol
        assert abs(og[9] - oo[9]) <= abs(oo[9]) * tol
        assert og[7] < mpmath.mpf(10) ** -15 and oo[7] < mpmath.mpf(10) ** -15


def test_objective_matrix_C_matches_oracle():
    """C != 0 (MPMP.jl:599): residual P = sum x_i A_i - X - C (:1108-1118), dual objective <C,Y> + <b,y> + b0
    (:1031-1034), through clrsdp_upload_C: three iterations GPU against oracle on the general structure (m > 1, L > 1,
    rank > 1), every field at 2^-(p-16), then a full solve with the same iteration count and objectives."""
    from test_oracle_pin import random_C
    prec = 256
    cons, b = instances.random_structured_sdp(GENERAL_SPEC, n_y=4, prec=prec)
    bi = solver.get_block_info(cons)
    C = random_C(bi, prec // 32)
    hs = [solver.product_handle(prec), oracle_handle(prec, 8)]
    infos = []
    for h in hs:
        solver.load_problem(h, cons, b, bi, b0=1, C=C)
        h.set_params(solver.real_params(h.nlimb))
        h.init_point()
        infos.append(h.prepare())
    assert infos[0].d_obj == pytest.approx(infos[1].d_obj, rel=1e-13) and abs(infos[0].d_obj) > 1e6   # <C, omega_d I>
    assert infos[0].gap == pytest.approx(infos[1].gap, rel=1e-13)
    hg, ho = hs
    wc, wb = widen_problem(cons, b, 2)
    ht = oracle_handle(prec + 64, 8)
    solver.load_problem(ht, wc, wb, bi, b0=1, C=[[blk.widen(prec // 32 + 2) for blk in row] for row in C])
    ht.set_params(solver.real_params(ht.nlimb))
    ht.init_point()
    ht.prepare()
    for it in range(3):
        rg, ro, rt = hg.iterate(), ho.iterate(), ht.iterate()
        assert rg.status == 0 and ro.status == 0 and rt.status == 0
        compare_iteration(hg, ho, bi, prec, prec - 16, ht)
    og, rg = solver.solverank1sdp(cons, b, bi, C=C, b0=1, verbose=False, return_info=True)
    oo, ro = solver.solverank1sdp(cons, b, bi, C=C, b0=1, handle=oracle_handle(prec, 8), verbose=False, return_info=True)
    assert len(rg) == len(ro) and rg[-1].terminate == ro[-1].terminate
    with mpmath.workprec(prec):
        tol = mpmath.mpf(2) ** -(prec - 16 - 64)
        assert abs(og[8] - oo[8]) <= max(1, abs(oo[8])) * tol and abs(og[9] - oo[9]) <= max(1, abs(oo[9])) * tol


def test_warm_start_roundtrip():
    """download_point -> upload_point reproduces the same next iteration (initial_solutions, MPMP.jl:613,689)."""
    prec = 256
    cons, b, _ = instances.synthetic_clustered_sdp(J=2, delta=4, K=6, n_y=3, prec=prec)
    bi = solver.get_block_info(cons)
    h1 = solver.product_handle(prec)
    solver.load_problem(h1, cons, b, bi)
    h1.set_params(solver.real_params(h1.nlimb))
    h1.init_point()
    h1.prepare()
    h1.iterate()
    n_x, n_X = sum(bi.dim_S), sum(s * s for row in bi.Y_blocksizes for s in row)
    st = h1.download_point(n_x, n_X, bi.n_y)
    r1 = h1.iterate()
    h2 = solver.product_handle(prec)
    solver.load_problem(h2, cons, b, bi)
    h2.set_params(solver.real_params(h2.nlimb))
    h2.upload_point(*st)
    h2.prepare()
    r2 = h2.iterate()
    assert r1.alpha_p == r2.alpha_p and r1.alpha_d == r2.alpha_d and r1.mu == r2.mu
    a, o = h1.fetch("x"), h2.fetch("x")
    assert np.array_equal(a.limb, o.limb) and np.array_equal(a.exp, o.exp)     # bit-identical


def test_host_resident_iterate_loop_replays_prepare_bit_identically():
    """The end-to-end pattern of bench.py: the iterate lives on the host, every step is upload_point + prepare + iterate +
    download_point on ONE handle (the iteration replayed from its CUDA graph, the global size of X uploaded once); the
    log row of prepare, the iteration's row and the new iterate must be bit-identical to a fresh handle that runs the
    same step with direct launches."""
    prec = 256
    cons, b, _ = instances.synthetic_clustered_sdp(J=3, delta=5, K=7, n_y=4, prec=prec, seed=5)
    bi = solver.get_block_info(cons)
    n_x, n_X = sum(bi.dim_S), sum(s * s for row in bi.Y_blocksizes for s in row)
    h = solver.product_handle(prec)
    solver.load_problem(h, cons, b, bi)
    h.set_params(solver.real_params(h.nlimb))
    h.init_point()
    h.prepare()
    h.iterate()
    st = h.download_point(n_x, n_X, bi.n_y)
    same = lambda a, o: np.array_equal(a.limb, o.limb) and np.array_equal(a.exp, o.exp) and np.array_equal(a.sign, o.sign)
    for step in range(4):
        h.upload_point(*st)
        p1 = h.prepare()
        r1 = h.iterate()
        out = h.download_point(n_x, n_X, bi.n_y)
        f = solver.product_handle(prec)                       # fresh handle: everything launched directly
        solver.load_problem(f, cons, b, bi)
        f.set_params(solver.real_params(f.nlimb))
        f.upload_point(*st)
        p2 = f.prepare()
        r2 = f.iterate()
        ref = f.download_point(n_x, n_X, bi.n_y)
        for k in ("mu", "p_obj", "d_obj", "gap", "P_err", "p_err", "d_err", "pd_feasible", "terminate"):
            assert getattr(p1, k) == getattr(p2, k), (step, k)
        for k in ("mu", "alpha_p", "alpha_d", "beta_c", "p_obj_new", "d_obj_new", "P_err", "p_err", "d_err", "terminate"):
            assert getattr(r1, k) == getattr(r2, k), (step, k)
        assert all(same(a, o) for a, o in zip(out, ref)), step
        st = out


def test_pinned_host_buffers_roundtrip_is_bit_identical():
    """clrsdp_pin_host: the direct-DMA path (limb planes by strided copy, header words converted on the device) moves
    exactly the same bits as the staged path, in both directions, including zero entries."""
    prec = 256
    cons, b, _ = instances.synthetic_clustered_sdp(J=2, delta=48, K=50, n_y=5, prec=prec)
    bi = solver.get_block_info(cons)
    h = solver.product_handle(prec)
    solver.load_problem(h, cons, b, bi)
    h.set_params(solver.real_params(h.nlimb))
    h.init_point()                       # omega * I: the off-diagonal entries are exact zeros
    h.prepare()
    h.iterate()
    n_x, n_X = sum(bi.dim_S), sum(s * s for row in bi.Y_blocksizes for s in row)
    assert n_X >= 4096                   # large enough for the direct path
    staged = h.download_point(n_x, n_X, bi.n_y)
    pinned = h.download_point(n_x, n_X, bi.n_y)
    h.pin(*pinned)
    pinned = h.download_point(n_x, n_X, bi.n_y, out=pinned)
    for a, o in zip(staged, pinned):
        assert np.array_equal(a.sign, o.sign) and np.array_equal(a.exp, o.exp) and np.array_equal(a.limb, o.limb)
    # upload from pinned buffers (with some exact zeros), read back through the staged path
    pinned[1].sign[:7] = 0
    h.upload_point(*pinned)
    back = h.download_point(n_x, n_X, bi.n_y)
    assert np.all(back[1].sign[:7] == 0) and np.all(back[1].limb[:, :7] == 0)
    assert np.array_equal(back[1].sign[7:], staged[1].sign[7:]) and np.array_equal(back[1].limb[:, 7:], staged[1].limb[:, 7:])
    assert np.array_equal(back[3].limb, staged[3].limb) and np.array_equal(back[3].exp, staged[3].exp)


def test_call_order_and_argument_errors():
    h = solver.product_handle(256)
    with pytest.raises(ClrsdpError) as e:
        h.iterate()
    assert e.value.code == -15
    with pytest.raises(ClrsdpError):
        h.set_structure(3, [1], [1], [0], [2], [1])            # n_samples = 0
    cons, b, _ = instances.synthetic_clustered_sdp(J=2, delta=3, K=4, n_y=2, prec=256)
    bi = solver.get_block_info(cons)
    solver.load_problem(h, cons, b, bi)
    with pytest.raises(ClrsdpError):
        h.upload_cluster(0, cons[0].V[0], cons[0].H[0], cons[0].B, cons[1].c.take(range(3)))   # wrong size
    with pytest.raises(ClrsdpError):
        h.fetch("nonsense")


def test_not_positive_definite_point_is_reported():
    prec = 256
    cons, b, _ = instances.synthetic_clustered_sdp(J=2, delta=3, K=4, n_y=2, prec=prec)
    bi = solver.get_block_info(cons)
    h = solver.product_handle(prec)
    solver.load_problem(h, cons, b, bi)
    h.set_params(solver.real_params(h.nlimb, omega_p=-1))       # X = -I is not PD
    h.init_point()
    h.prepare()
    with pytest.raises(ClrsdpError) as e:
        h.iterate()
    assert e.value.code == -10 and "higher precision" in str(e.value)


def test_launch_counter_and_profile():
    prec = 256
    cons, b, _ = instances.synthetic_clustered_sdp(J=2, delta=4, K=6, n_y=3, prec=prec)
    bi = solver.get_block_info(cons)
    h = solver.product_handle(prec)
    solver.load_problem(h, cons, b, bi)
    h.set_params(solver.real_params(h.nlimb))
    h.init_point()
    h.prepare()
    n0 = h.launch_count()
    h.profile_reset(True)
    r = h.iterate()
    prof = h.profile_dump()
    assert h.launch_count() - n0 == sum(v["launches"] for v in prof.values()) > 100
    assert any(k.startswith("mma_planes") for k in prof)
    assert r.seconds > 0 and sum(r.timings) > 0


def _sphere_golden(d):
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "sphere_packing_512.json")) as f:
        return [c for c in json.load(f)["cases"] if c["d"] == d][0]


def test_sphere_packing_config1_matches_oracle():
    """BASELINE config 1 (examples/SpherePacking.jl, n = 3, d = 8) at the example's own precision, 512 bits (ex:29-31).
    The Schur complements of this instance are graded over 2^60 and have condition numbers 2^126 (first iteration) to
    beyond 2^240 (near the optimum) after equilibration, so two backward-stable methods agree to p - log2(cond) bits,
    not to p - 16: per iteration the GPU result must be as close to the same iteration at p + 64 bits as the MPFR
    oracle's (LU) is; over the whole solve: the same number of iterations as the oracle (live and golden), the same
    log rows, and the objective to 2^-128."""
    prec = 512
    solver.set_precision(prec)
    try:
        cons, b, _ = instances.sphere_packing_2point(n=3, d=8, prec=prec)
        bi = solver.get_block_info(cons)
        assert (bi.J, bi.n_y) == (7, 52)
        kw = dict(omega_p=100, omega_d=100)
        hg, ho = pair(cons, b, bi, prec, **kw)
        for it in range(2):
            rg, ro = hg.iterate(), ho.iterate()
            assert rg.alpha_p == pytest.approx(ro.alpha_p, rel=1e-12) and rg.alpha_d == pytest.approx(ro.alpha_d, rel=1e-12)
            for name in ("dx", "dy", "x", "y"):
                assert rel_err_bits(hg.fetch(name), ho.fetch(name)) >= prec - 16 - 130, name   # cond(S') = 2^126
            for j in range(bi.J if it == 0 else 0):                               # the Schur complements of the common
                assert rel_err_bits(hg.fetch("S", j), ho.fetch("S", j)) >= prec - 16   # starting point: full accuracy
        g = _sphere_golden(8)
        og, rg = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True, **kw)
        oo, ro = solver.solverank1sdp(cons, b, bi, handle=oracle_handle(prec, 8), verbose=False, return_info=True, **kw)
        assert len(rg) == len(ro) == g["iterations"]
        assert rg[-1].terminate == ro[-1].terminate == 3
        for a, o in zip(rg, ro):
            assert a.alpha_p == pytest.approx(o.alpha_p, rel=1e-9) and a.alpha_d == pytest.approx(o.alpha_d, rel=1e-9)
            assert a.mu == pytest.approx(o.mu, rel=1e-9)
        with mpmath.workprec(prec):
            tol = mpmath.mpf(2) ** -128
            assert abs(og[8] - oo[8]) <= abs(oo[8]) * tol and abs(og[9] - oo[9]) <= abs(oo[9]) * tol
            assert abs(og[8] - mpmath.mpf(g["primal_obj"])) <= tol
            assert og[7] < mpmath.mpf(10) ** -15
    finally:
        solver.set_precision(256)


def test_sphere_packing_higher_degree_known_answers():
    """BASELINE config 2 (SpherePacking at higher polynomial degree on one B200). As specified (d = 40 at 256 bits) the
    instance is out of reach of 256-bit arithmetic for ANY method: at the first iteration the equilibrated Schur
    complement of the first cluster has condition number > 2^257 and the oracle's own S is indefinite to working
    precision (measured with the oracle) - the example itself never runs below 512 bits (ex:29-31).
    What is checked: (i) the higher degrees that 512 bits do support, against the oracle's golden runs: d = 12 takes
    the same 91 iterations and gives the same objective; d = 16 (where the oracle needs 168 iterations, wandering at
    its precision limit) reaches the same optimum; the bounds decrease with the degree towards ~0.813 and stay above
    the literature value 0.793 (ex:125-126). (ii) d = 40 at 256 bits either raises the reference's "higher precision"
    error or ends without claiming optimality - never a silent wrong bound."""
    prec = 512
    solver.set_precision(prec)
    try:
        bounds = {8: -mpmath.mpf(_sphere_golden(8)["primal_obj"])}
        for d in (12, 16):
            g = _sphere_golden(d)
            cons, b, _ = instances.sphere_packing_2point(n=3, d=d, prec=prec)
            bi = solver.get_block_info(cons)
            og, rg = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True, omega_p=100, omega_d=100)
            assert rg[-1].terminate == 3
            if d == 12:
                assert len(rg) == g["iterations"]
            with mpmath.workprec(prec):
                assert abs(og[8] - mpmath.mpf(g["primal_obj"])) < mpmath.mpf(10) ** -15   # the duality-gap threshold
                assert og[7] < mpmath.mpf(10) ** -15
                bounds[d] = -og[8]
        assert mpmath.mpf("0.793") <= bounds[16] < bounds[12] < bounds[8] < mpmath.mpf("0.8151")
    finally:
        solver.set_precision(256)
    prec = 256
    cons, b, _ = instances.sphere_packing_2point(n=3, d=40, prec=prec)
    bi = solver.get_block_info(cons)
    assert bi.n_y == 244 and max(max(r) for r in bi.Y_blocksizes) == 82
    try:
        og, rg = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True, omega_p=100, omega_d=100,
                                      maxiterations=60)
        assert rg[-1].terminate != 3
    except ClrsdpError as e:
        assert "higher precision" in str(e)



def _sphere_lowprec(d, prec):
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "sphere_packing_lowprec.json")) as f:
        return [c for c in json.load(f)["cases"] if c["d"] == d and c["prec"] == prec][0]


@pytest.mark.parametrize("d,prec", [(8, 256), (8, 384), (12, 256)])
def test_sphere_packing_below_the_examples_precision(d, prec):
    """BASELINE config 1 (sphere packing, n = 3, d = 8) at 256 and 384 bits, and d = 12 at 256 bits: precisions at which
    the Schur complements are singular to working precision near the optimum. Round 1 lost the iterate here (pivots
    clamped by a Cholesky that cannot take a negative pivot); the signed factorisation does not.

    What "parity with the reference" means at these precisions (tests/golden/sphere_packing_lowprec.json, generator
    beside it): the reference's own algorithm (pivoted LU; the oracle proper) walks a trajectory dominated by rounding
    noise - its iteration count changes with the number of threads (the summation order of the Q product) and with the
    product mode: 93/94/114 at (8, 256), 130/131/170 at (8, 384), and at (12, 256) it does not converge at all (maxiter).
    The oracle run with the GPU's factorisation in MPFR arithmetic (CLRSDP_REF_FACTOR=ldl) takes 75 / 82 / 85 iterations,
    82 being the count of the noise-free trajectory (every method takes 82 at 512 bits). So the GPU solve is held to:
    (i) row for row the trajectory of the ldl oracle (alpha and mu to 1e-6) until the rounding noise takes over - the first
    50 iterations at (8, 256) - and the identical iteration count where the whole solve stays above it (384 bits: 82 = 82); (ii) "Optimal" with the objective of the oracle (ldl, and LU wherever that converges), to the
    duality-gap threshold; (iii) no more iterations than the ldl oracle and than the best LU run."""
    import os
    g = _sphere_lowprec(d, prec)
    solver.set_precision(prec)
    try:
        cons, b, _ = instances.sphere_packing_2point(n=3, d=d, prec=prec)
        bi = solver.get_block_info(cons)
        kw = dict(omega_p=100, omega_d=100)
        og, rg = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True, **kw)
        os.environ["CLRSDP_REF_FACTOR"] = "ldl"
        try:
            ho = oracle_handle(prec, 8)
        finally:
            del os.environ["CLRSDP_REF_FACTOR"]
        oo, ro = solver.solverank1sdp(cons, b, bi, handle=ho, verbose=False, return_info=True, **kw)
        assert rg[-1].terminate == ro[-1].terminate == 3 and rg[-1].status == 0
        assert len(ro) in {r["iterations"] for r in g["ldl"]}   # the live MPFR run reproduces the golden one
        same = 0
        for a, o in zip(rg, ro):                                 # (i): rows agree until the rounding noise takes over
            if not (a.alpha_p == pytest.approx(o.alpha_p, rel=1e-6) and a.alpha_d == pytest.approx(o.alpha_d, rel=1e-6)
                    and a.mu == pytest.approx(o.mu, rel=1e-6)):
                break
            same += 1
        # measured: 50 common rows at (8, 256), all 82 at (8, 384); d = 12 is harder (cond(S') grows faster)
        assert same >= (len(ro) if prec >= 384 else 30), same
        if prec >= 384:
            assert len(rg) == len(ro)
        assert len(rg) <= len(ro)                                # (iii)
        with mpmath.workprec(prec):
            assert abs(og[8] - oo[8]) <= mpmath.mpf(10) ** -14 and og[7] < mpmath.mpf(10) ** -15    # (ii)
            lu_ok = [r for r in g["lu"] if r.get("terminate") == 3]
            for r in lu_ok:
                assert abs(og[8] - mpmath.mpf(r["primal_obj"])) <= mpmath.mpf(10) ** -14
            if lu_ok:
                assert len(rg) <= min(r["iterations"] for r in lu_ok)
            else:
                assert all(r.get("terminate") == 4 or "error" in r for r in g["lu"])
    finally:
        solver.set_precision(256)


@pytest.mark.parametrize("d,prec,known", [(16, 256, "-0.813595526098925642586"), (8, 128, "-0.815009706442796514379"),
                                          (12, 128, "-0.813660572002088522099")])
def test_runs_beyond_the_precision_never_return_garbage(d, prec, known):
    """Instances whose conditioning exceeds the working precision (sphere packing d = 16 at 256 bits - the MPFR run with
    the same factorisation dies with "X not positive definite" - and d = 8, 12 at 128 bits; measured in round 2: maxiter with
    the optimum to 1e-11 / 4e-11, and DIVERGED at (12, 128)): the run must either end with
    a non-zero status (NOT_PD_X/Y, or DIVERGED once mu / a step length / an objective is zero, negative or not finite:
    the reference's "higher precision" error) or stay on the central path and stop at maxiterations with the known
    optimum to the accuracy the precision allows. Round 1 returned `terminate = 4` with a ZERO bound here."""
    solver.set_precision(prec)
    try:
        cons, b, _ = instances.sphere_packing_2point(n=3, d=d, prec=prec)
        bi = solver.get_block_info(cons)
        try:
            og, rg = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True, omega_p=100, omega_d=100,
                                          maxiterations=300)
        except ClrsdpError as e:
            assert e.code in (-10, -11, -16) and "higher precision" in str(e)
            return
        assert all(r.status == 0 for r in rg)
        for r in rg:
            assert np.isfinite(r.mu) and r.mu > 0 and r.alpha_p > 0 and r.alpha_d > 0
        with mpmath.workprec(prec):
            tol = mpmath.mpf(10) ** (-8 if prec >= 256 else -4)      # what is left of the optimum at this precision
            assert abs(og[8] - mpmath.mpf(known)) < tol and abs(og[9] - mpmath.mpf(known)) < tol
    finally:
        solver.set_precision(256)


def test_diverged_status_is_mapped_to_the_higher_precision_error():
    from clrsdp import capi
    e = capi.ClrsdpError(-16, "clrsdp_iterate")
    assert "higher precision" in str(e) and "lost" in str(e)


def test_failed_iteration_leaves_the_iterate_intact():
    """When a factorisation fails inside iterate() the update must not run: the point on the device is still the last
    good one (the reference raises before its update, MPMP.jl:793 precedes :877), so it can be downloaded and resumed at
    a higher precision."""
    prec = 256
    cons, b, _ = instances.synthetic_clustered_sdp(J=2, delta=3, K=4, n_y=2, prec=prec)
    bi = solver.get_block_info(cons)
    h = solver.product_handle(prec)
    solver.load_problem(h, cons, b, bi)
    h.set_params(solver.real_params(h.nlimb))
    h.init_point()
    h.prepare()
    h.iterate()
    n_x, n_X = sum(bi.dim_S), sum(s * s for row in bi.Y_blocksizes for s in row)
    x, X, y, Y = h.download_point(n_x, n_X, bi.n_y)
    X.sign[0] = -1                      # X[0][0,0] < 0: not positive definite
    h.upload_point(x, X, y, Y)
    h.prepare()
    with pytest.raises(ClrsdpError) as e:
        h.iterate()
    assert e.value.code == -10
    after = h.download_point(n_x, n_X, bi.n_y)
    for a, o in zip(after, (x, X, y, Y)):
        assert np.array_equal(a.sign, o.sign) and np.array_equal(a.exp, o.exp) and np.array_equal(a.limb, o.limb)


CFG4_SPEC = [dict(m=2, K=91, blocks=[dict(delta=56, ranks=[2] * 91), dict(delta=42, ranks=[2] * 91)])] * 4


def test_config4_structure_384bit_matches_oracle():
    """BASELINE config 4 by its structure: 4 clusters of 2 x 2 polynomial-matrix constraints, K = 91 sample
    points, every (l, k) of rank 2, vector lengths 56 and 42 (blocks 112 and 84), dim_S = 273, n_y = 60, 384 bits.
    The constraint vectors are random instead of Padua-point evaluations of a bivariate basis (manufactured strictly
    feasible): what the hot path sees - m > 1, L > 1, rank > 1 at these sizes - is the same. Two iterations against the
    oracle to 2^-(p-16) (p + 64-bit arbiter where the conditioning exceeds 2^16), then the GPU solve converges."""
    prec = 384
    cons, b = instances.random_structured_sdp(CFG4_SPEC, n_y=60, prec=prec, seed=20261020)
    bi = solver.get_block_info(cons)
    assert list(bi.dim_S) == [273] * 4 and [list(r) for r in bi.Y_blocksizes] == [[112, 84]] * 4
    hg, ho = pair(cons, b, bi, prec)
    wc, wb = widen_problem(cons, b, 2)
    ht = oracle_handle(prec + 64, 8)
    solver.load_problem(ht, wc, wb, bi)
    ht.set_params(solver.real_params(ht.nlimb))
    ht.init_point()
    ht.prepare()
    for it in range(2):
        rg, ro, rt = hg.iterate(), ho.iterate(), ht.iterate()
        assert rg.status == 0 and ro.status == 0 and rt.status == 0
        compare_iteration(hg, ho, bi, prec, prec - 16, ht)
    solver.set_precision(prec)
    try:
        og, rows = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True)
        assert rows[-1].terminate == 3
        with mpmath.workprec(prec):
            assert og[7] < mpmath.mpf(10) ** -15
    finally:
        solver.set_precision(256)


def test_config4_bivariate_polynomial_program_384bit():
    """BASELINE config 4 with real polynomial data (`instances.bivariate_matrix_program`: 2 x 2 polynomial-matrix
    constraints in two variables sampled at the 91 Padua points of degree 12 through the `Pi` path of `prepareabc`,
    product-Chebyshev basis; sizes exactly SURVEY §8d's: dim_S = 273, blocks 112 and 84, Nv = 182, n_y = 60, 384 bits).
    Two iterations against the oracle: every field agrees to 2^-(p-16), or - where the conditioning of the sampled
    polynomial data exceeds 2^16 - the GPU result is as close (within 2 bits) to the same iteration at p + 64 bits as
    the oracle's is. Then the GPU solve converges to the oracle-independent optimality conditions."""
    prec = 384
    cons, b = instances.bivariate_matrix_program(prec=prec)
    bi = solver.get_block_info(cons)
    assert list(bi.dim_S) == [273] * 4 and [list(r) for r in bi.Y_blocksizes] == [[112, 84]] * 4 and bi.n_y == 60
    hg, ho = pair(cons, b, bi, prec)
    wc, wb = widen_problem(cons, b, 2)
    ht = oracle_handle(prec + 64, 8)
    solver.load_problem(ht, wc, wb, bi)
    ht.set_params(solver.real_params(ht.nlimb))
    ht.init_point()
    ht.prepare()

    def check(name, *idx):
        a, o, t = hg.fetch(name, *idx), ho.fetch(name, *idx), ht.fetch(name, *idx)
        if rel_err_bits(a, o) >= prec - 16:
            return
        eg, eo = rel_err_bits(a, t), rel_err_bits(o, t)
        assert eg >= eo - 2, (name, idx, eg, eo)

    for it in range(2):
        rg, ro, rt = hg.iterate(), ho.iterate(), ht.iterate()
        assert rg.status == 0 and ro.status == 0 and rt.status == 0
        assert rg.alpha_p == pytest.approx(ro.alpha_p, rel=1e-12) and rg.alpha_d == pytest.approx(ro.alpha_d, rel=1e-12)
        for name in ("d", "dx", "dy", "x", "y"):
            check(name)
        for j in range(bi.J):
            for l in range(bi.L[j]):
                for name in ("Xinv", "Px", "Py", "Z", "dX", "dY", "X", "Y"):
                    check(name, j, l)
    solver.set_precision(prec)
    try:
        og, rows = solver.solverank1sdp(cons, b, bi, verbose=False, return_info=True)
        assert rows[-1].terminate == 3
        with mpmath.workprec(prec):
            assert og[7] < mpmath.mpf(10) ** -15
    finally:
        solver.set_precision(256)


def test_solvempmp_front_end_on_the_gpu():
    """The reference's user entry `solvempmp` (MPMP.jl:562-586) end to end on the product path: polynomial input ->
    `prepareabc` -> `get_block_info` -> `solverank1sdp` on the GPU. max t s.t. x^4 - x^2 + 1 - t >= 0 on [-1, 1]: 3/4;
    and the same call on the oracle takes the same number of iterations."""
    from clrsdp import frontend as fe
    prec = 256
    solver.set_precision(prec)
    X = fe.Poly.var(1, 0)
    M = [[[X ** 4 - X ** 2 + 1]], [[fe.Poly.const(1, -1)]]]
    G = [fe.Poly.const(1, 1), 1 - X * X]
    q = fe.make_monomial_basis(1, 2)
    xs = [[v] for v in fe.create_sample_points_chebyshev(4)]
    args = ([M], [G], [q], [xs], [4], [mpmath.mpf(1)])
    og, rg = fe.solvempmp(*args, verbose=False, return_info=True)
    oo, ro = fe.solvempmp(*args, verbose=False, return_info=True, handle=oracle_handle(prec, 2))
    assert len(rg) == len(ro) and rg[-1].terminate == ro[-1].terminate == 3
    with mpmath.workprec(prec):
        assert abs(og[9] - mpmath.mpf(3) / 4) < mpmath.mpf(10) ** -12 and abs(og[8] - mpmath.mpf(3) / 4) < mpmath.mpf(10) ** -12
        assert abs(og[8] - oo[8]) <= mpmath.mpf(2) ** -(prec - 16 - 64)


def test_checkpoint_resume_on_the_gpu(tmp_path):
    """save_checkpoint / load_checkpoint through the product handle: the resumed solve continues bit-identically."""
    prec = 256
    cons, b, _ = instances.synthetic_clustered_sdp(J=3, delta=6, K=10, n_y=5, prec=prec)
    bi = solver.get_block_info(cons)

    def fresh():
        h = solver.product_handle(prec)
        solver.load_problem(h, cons, b, bi)
        h.set_params(solver.real_params(h.nlimb))
        return h

    h1 = fresh()
    h1.init_point()
    h1.prepare()
    for _ in range(3):
        h1.iterate()
    solver.save_checkpoint(tmp_path / "it3", h1, bi, iteration=3)
    h1.prepare()                      # the resumed handle starts from a prepare(): do the same here
    want = [h1.iterate() for _ in range(2)]
    h2 = fresh()
    assert solver.load_checkpoint(tmp_path / "it3", h2, bi) == 3
    h2.prepare()
    got = [h2.iterate() for _ in range(2)]
    for a, o in zip(got, want):
        assert (a.mu, a.alpha_p, a.alpha_d, a.p_obj_new, a.d_obj_new) == (o.mu, o.alpha_p, o.alpha_d, o.p_obj_new, o.d_obj_new)
    n_x, n_X = int(sum(bi.dim_S)), int(sum(s * s for row in bi.Y_blocksizes for s in row))
    for a, o in zip(h2.download_point(n_x, n_X, bi.n_y), h1.download_point(n_x, n_X, bi.n_y)):
        assert np.array_equal(a.sign, o.sign) and np.array_equal(a.exp, o.exp) and np.array_equal(a.limb, o.limb)


def _reference_files():
    import glob
    import os
    return sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.clrsdp")))


@pytest.mark.parametrize("path", _reference_files() or [None])
def test_gpu_against_reference_problem_files(path):
    """The GPU path against the REAL reference's result, for every CLRSDP1 file under tests/golden/ (written by
    julia/ClrsdpB200.jl where Julia + Arblib exist; see test_oracle_pin.py). Skipped while there is none."""
    if path is None:
        pytest.skip("no tests/golden/*.clrsdp reference file present (needs Julia + Arblib to produce)")
    from clrsdp import problem_io
    cons, b, meta = problem_io.load_problem_file(path)
    prec = meta["prec"]
    bi = solver.get_block_info(cons)
    solver.set_precision(prec)
    try:
        out, rows = solver.solverank1sdp(cons, b, bi, b0=mpmath.mpf(meta["b0"]), verbose=False, return_info=True)
    finally:
        solver.set_precision(256)
    sol = meta["solution"]
    assert sol is not None and len(rows) == sol["iterations"]
    with mpmath.workprec(prec):
        tol = mpmath.mpf(2) ** -(prec - 16 - 64)
        assert abs(out[8] - mpmath.mpf(sol["primal_obj"])) <= abs(mpmath.mpf(sol["primal_obj"])) * tol
        assert abs(out[9] - mpmath.mpf(sol["dual_obj"])) <= abs(mpmath.mpf(sol["dual_obj"])) * tol

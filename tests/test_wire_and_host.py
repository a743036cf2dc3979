"""Wire format conversions and the host-side mirror of BlockInfo / get_block_info (MPMP.jl:467-560)."""
import random
from fractions import Fraction

import mpmath
import numpy as np
import pytest

from clrsdp import instances, solver
from clrsdp.wire import MpArray


@pytest.mark.parametrize("nlimb", [4, 8, 16])
def test_int_roundtrip_and_rounding(nlimb):
    rng = random.Random(0)
    p = 32 * nlimb
    vals = [0, 1, -1, 3, (1 << p) - 1, -(1 << (p - 1)), rng.getrandbits(p + 40), -rng.getrandbits(p - 7)]
    a = MpArray.from_ints(vals, -17, nlimb)
    for i, v in enumerate(vals):
        exact = Fraction(v) / 2 ** 17
        got = a.to_fraction(i)
        if v == 0:
            assert got == 0 and a.sign[i] == 0
            continue
        assert abs(got - exact) <= abs(exact) / 2 ** p           # round to nearest at p bits
        assert a.limb[nlimb - 1, i] >> 31 == 1                    # normalised
    # exactly representable values round-trip exactly
    assert a.to_fraction(1) == Fraction(1, 2 ** 17) and a.to_fraction(3) == Fraction(3, 2 ** 17)


def test_ties_to_even():
    a = MpArray.from_ints([(1 << 128) + (1 << 0) * 0 + (1 << 0), ((1 << 127) | 1) << 1 | 1], 0, 4)
    # 2^128 + 1 needs 129 bits: the dropped bit is exactly half -> ties to even (mantissa stays 2^127)
    assert a.to_fraction(0) == 1 << 128


def test_mpf_and_double_conversions():
    with mpmath.workprec(300):
        v = [mpmath.pi, -mpmath.e / 7, mpmath.mpf(0), mpmath.mpf(2) ** -200]
        a = MpArray.from_mpf(v, 8)
        for i, x in enumerate(v):
            assert abs(a.to_mpf(i) - x) <= abs(x) * mpmath.mpf(2) ** -256
    d = np.array([0.0, 1.5, -3.25e-30, 7e100])
    b = MpArray.from_double(d, 8)
    assert [b.to_fraction(i) for i in range(4)] == [Fraction(x) for x in d]
    assert np.array_equal(b.to_double(), d)
    c = MpArray.from_scaled_int64(np.array([5, -3, 0, 1 << 40]), -40, 4)
    assert [c.to_fraction(i) for i in range(4)] == [Fraction(5, 2 ** 40), Fraction(-3, 2 ** 40), 0, 1]
    f = MpArray.from_fraction([Fraction(3, 10), Fraction(-1, 3)], 8)
    assert abs(f.to_fraction(0) - Fraction(3, 10)) < Fraction(1, 2 ** 256)


def test_transpose_take_concat():
    a = MpArray.from_ints(list(range(1, 7)), 0, 4).reshape(2, 3)
    t = a.transpose2d()
    assert t.shape == (3, 2) and [int(t.to_fraction(i)) for i in range(6)] == [1, 4, 2, 5, 3, 6]
    c = MpArray.concat([a.reshape(6), t.reshape(6)])
    assert c.n == 12 and int(c.to_fraction(7)) == 4


def test_get_block_info_matches_reference_rules():
    spec = [dict(m=2, K=5, blocks=[dict(delta=3, ranks=[2, 1, 0, 2, 1]), dict(delta=2, ranks=[0, 1, 1, 1, 1])]),
            dict(m=1, K=4, blocks=[dict(delta=3, ranks=[1, 2, 1, 1])])]
    cons, b = instances.random_structured_sdp(spec, n_y=4, prec=128)
    bi = solver.get_block_info(cons)
    assert bi.J == 2 and bi.n_y == 4
    assert bi.m == [2, 1] and bi.L == [2, 1] and bi.n_samples == [5, 4]
    assert bi.dim_S == [15, 4] and bi.x_indices == [0, 15, 19]                 # m(m+1)/2 * K  (:511, :486)
    assert bi.Y_blocksizes == [[6, 4], [3]]                                    # m * len(vector)  (:550-551)
    assert bi.rank_sums[0][0] == [0, 2, 3, 3, 5, 6]                            # (:488)
    assert bi.nz_k[0] == [0, 1]                                                # first k with non-zero rank (:489-491)
    assert bi.jl_pairs == [(0, 0), (0, 1), (1, 0)]


def test_blockinfo_validates_lengths():
    with pytest.raises(ValueError):
        solver.BlockInfo(2, 3, [1], [1, 1], [2, 2], [[2], [2]], [2, 2], [[[1, 1]], [[1, 1]]])
    with pytest.raises(ValueError):
        solver.BlockInfo(1, 3, [1], [2], [2], [[2]], [2], [[[1, 1]]])


def test_real_params_defaults_are_the_reference_defaults():
    rp = solver.real_params(8)
    want = [Fraction(3, 10), Fraction(1, 10), Fraction(7, 10), Fraction(10) ** 10, Fraction(10) ** 10,
            Fraction(1, 10 ** 15), Fraction(1, 10 ** 30), Fraction(1, 10 ** 30)]
    for i, w in enumerate(want):
        assert abs(rp.to_fraction(i) - w) <= w / 2 ** 256


def test_synthetic_instance_is_deterministic_and_shardable():
    c1, b1, _ = instances.synthetic_clustered_sdp(J=4, delta=3, K=5, n_y=3, prec=128, seed=9)
    c2, b2, _ = instances.synthetic_clustered_sdp(J=2, delta=3, K=5, n_y=3, prec=128, seed=9, j_offset=2, j_total=4)
    assert np.array_equal(b1.limb, b2.limb) and np.array_equal(b1.exp, b2.exp)       # b is the full problem's b
    assert np.array_equal(c1[2].B.limb, c2[0].B.limb) and np.array_equal(c1[3].c.limb, c2[1].c.limb)


def test_precision_setter():
    solver.set_precision(384)
    assert solver.precision() == 384
    with pytest.raises(ValueError):
        solver.set_precision(100)
    solver.set_precision(256)


def test_checkpoint_roundtrip_resumes_the_same_trajectory(tmp_path):
    """save_checkpoint / load_checkpoint (row f4): a solve interrupted after 3 iterations and resumed from the file
    produces exactly the iterations the uninterrupted solve does (oracle handle; the product handle uses the same
    calls)."""
    from oracle.ref import oracle_handle
    prec = 128
    cons, b, _ = instances.synthetic_clustered_sdp(J=2, delta=3, K=4, n_y=2, prec=prec)
    bi = solver.get_block_info(cons)

    def fresh():
        h = oracle_handle(prec, 1)
        solver.load_problem(h, cons, b, bi)
        h.set_params(solver.real_params(h.nlimb))
        return h

    h1 = fresh()
    h1.init_point(); h1.prepare()
    for _ in range(3):
        h1.iterate()
    solver.save_checkpoint(tmp_path / "it3", h1, bi, iteration=3)
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
    h3 = oracle_handle(256, 1)
    with pytest.raises(ValueError):
        solver.load_checkpoint(tmp_path / "it3", h3, bi)


def test_problem_file_roundtrip_is_bit_exact(tmp_path):
    """clrsdp.problem_io (row f4): constraints, objective and an iterate survive the file bit for bit; a file written by
    the reference-side shim (julia/ClrsdpB200.jl: write_problem) has the same layout."""
    from clrsdp import problem_io
    spec = [dict(m=2, K=3, blocks=[dict(delta=3, ranks=[2, 0, 1]), dict(delta=2, ranks=[1, 1, 1])]),
            dict(m=1, K=4, blocks=[dict(delta=3, ranks=[1, 2, 1, 1])])]
    cons, b = instances.random_structured_sdp(spec, n_y=4, prec=192)
    bi = solver.get_block_info(cons)
    n_x, n_X = int(sum(bi.dim_S)), int(sum(s * s for row in bi.Y_blocksizes for s in row))
    rng = np.random.default_rng(3)
    point = tuple(MpArray.from_scaled_int64(rng.integers(-2 ** 40, 2 ** 40, size=n, dtype=np.int64), -37, 6)
                  for n in (n_x, n_X, 4, n_X))
    path = tmp_path / "p.clrsdp"
    problem_io.save_problem(path, cons, b, b0="3", solution=dict(iterations=17, primal_obj="-0.5", dual_obj="-0.5"),
                            point=point)
    cons2, b2, meta = problem_io.load_problem_file(path)
    same = lambda a, o: np.array_equal(a.sign, o.sign) and np.array_equal(a.exp, o.exp) and np.array_equal(a.limb, o.limb)
    assert meta["prec"] == 192 and meta["b0"] == "3" and meta["solution"]["iterations"] == 17 and same(b2, b)
    assert len(cons2) == len(cons)
    for a, o in zip(cons2, cons):
        assert same(a.B, o.B) and same(a.c, o.c) and a.B.shape == o.B.shape
        for l in range(o.L):
            assert same(a.V[l], o.V[l]) and a.V[l].shape == o.V[l].shape and same(a.H[l], o.H[l])
            assert list(a.ranks[l]) == list(o.ranks[l])
    bi2 = solver.get_block_info(cons2)
    assert (bi2.J, bi2.n_y, list(bi2.dim_S), [list(r) for r in bi2.Y_blocksizes]) == \
           (bi.J, bi.n_y, list(bi.dim_S), [list(r) for r in bi.Y_blocksizes])
    assert all(same(a, o) for a, o in zip(meta["point"], point))
    with open(path, "r+b") as f:
        f.write(b"X")
    with pytest.raises(ValueError):
        problem_io.load_problem_file(path)

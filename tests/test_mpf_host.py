"""csrc/mpf.cuh compiled for the host: every scalar op against exact rational arithmetic.

The device kernels compile the very same header, so this pins the arithmetic the GPU path uses
(rounding model: within 0.5 + 2^-20 ulp of the exact result)."""
import ctypes
import os
import random
import subprocess
from fractions import Fraction

import pytest

from clrsdp.wire import MpArray, clrsdp_mp

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM_SRC = os.path.join(HERE, "helpers", "mpf_host_shim.cpp")
SHIM = os.path.join(HERE, "helpers", "libmpf_host.so")


@pytest.fixture(scope="module")
def shim():
    hdr = os.path.join(HERE, "..", "clustered-low-rank-sdp-solver_b200", "csrc", "mpf.cuh")
    if not os.path.exists(SHIM) or os.path.getmtime(SHIM) < max(os.path.getmtime(SHIM_SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SHIM_SRC, "-o", SHIM])
    lib = ctypes.CDLL(SHIM)
    lib.mpf_host_op.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(clrsdp_mp), ctypes.POINTER(clrsdp_mp),
                                ctypes.POINTER(clrsdp_mp)]
    return lib


def run(lib, op, a, b=None):
    c = MpArray(a.n, a.nlimb)
    sa, sc = a.c_struct(), c.c_struct()
    sb = b.c_struct() if b is not None else None
    st = lib.mpf_host_op(a.nlimb, ord(op), ctypes.byref(sa), ctypes.byref(sb) if sb is not None else None,
                         ctypes.byref(sc))
    assert st == 0
    return c


def rand_nums(rng, n, nlimb, erange=40, special=True):
    p = 32 * nlimb
    mants, exps = [], []
    for i in range(n):
        kind = rng.random()
        if special and kind < 0.05:
            m = 0
        elif special and kind < 0.15:
            m = (1 << p) - 1 - rng.getrandbits(8)            # all-ones patterns (carry chains)
        elif special and kind < 0.25:
            m = (1 << (p - 1)) + rng.getrandbits(8)          # 1000...0xxx
        elif special and kind < 0.30:
            m = rng.getrandbits(20) + 1                      # short mantissas
        else:
            m = rng.getrandbits(p) | (1 << (p - 1))
        if rng.random() < 0.5:
            m = -m
        mants.append(m)
        exps.append(rng.randint(-erange, erange) - p)
    return MpArray.from_ints(mants, exps, nlimb)


def ulp_err(got: MpArray, exact, p):
    worst = Fraction(0)
    for i, ex in enumerate(exact):
        g = got.to_fraction(i)
        if ex == 0:
            assert g == 0, (i, g)
            continue
        e = abs(ex).numerator.bit_length() - abs(ex).denominator.bit_length()
        k = e if abs(ex) < Fraction(2) ** e else e + 1   # |ex| in [2^(k-1), 2^k)
        ulp = Fraction(2) ** (k - p)
        worst = max(worst, abs(g - ex) / ulp)
    return float(worst)


@pytest.mark.parametrize("nlimb", [4, 8, 12, 16])
@pytest.mark.parametrize("op", ["+", "-", "*", "/"])
def test_binary_ops(shim, nlimb, op):
    rng = random.Random(1234 + nlimb)
    n = 400
    a = rand_nums(rng, n, nlimb)
    b = rand_nums(rng, n, nlimb, erange=40 if op in "*/" else 12)
    if op == "/":
        for i in range(n):
            if b.sign[i] == 0:
                b.set_int(i, 3, 0)
    # cancellation cases for +/-: b = -a * (1 + tiny)
    if op in "+-":
        for i in range(0, n, 7):
            m, e = a.get_int(i)
            b.set_int(i, (-m if op == "+" else m) + rng.randint(-3, 3), e)
    fa, fb = a.to_fractions(), b.to_fractions()
    exact = [{"+": x + y, "-": x - y, "*": x * y}[op] if op != "/" else x / y for x, y in zip(fa, fb)]
    got = run(shim, op, a, b)
    assert ulp_err(got, exact, 32 * nlimb) <= 0.5001


@pytest.mark.parametrize("nlimb", [4, 8, 12, 16])
def test_fused_mul_sub_mul(shim, nlimb):
    """a*b - c*d with a single rounding (the elimination step of the panel factorisation), including
    near-total cancellation: c*d = a*b*(1 + tiny)."""
    rng = random.Random(77 + nlimb)
    n = 400
    a = rand_nums(rng, n, nlimb, erange=20)
    b = rand_nums(rng, n, nlimb, erange=20)
    for i in range(0, n - 1, 5):                     # entries i, i+1 multiply to nearly the same product
        ma, ea = a.get_int(i)
        mb, eb = b.get_int(i)
        if ma == 0 or mb == 0:
            continue
        a.set_int(i + 1, mb + rng.randint(-2, 2), eb + rng.randint(-1, 1))
        b.set_int(i + 1, ma, ea)
    fa, fb = a.to_fractions(), b.to_fractions()
    exact = [fa[i] * fb[i] - fa[(i + 1) % n] * fb[(i + 1) % n] for i in range(n)]
    got = run(shim, "m", a, b)
    p = 32 * nlimb
    # error: half an ulp of the result plus the truncation of the two products (NL units of limb NL+1)
    worst = Fraction(0)
    for i, ex in enumerate(exact):
        g = got.to_fraction(i)
        mag = max(abs(fa[i] * fb[i]), abs(fa[(i + 1) % n] * fb[(i + 1) % n]))
        bound = Fraction(2) ** -(p + 27) * mag
        if ex != 0:
            e = abs(ex).numerator.bit_length() - abs(ex).denominator.bit_length()
            k = e if abs(ex) < Fraction(2) ** e else e + 1
            bound += Fraction(2) ** (k - p) * Fraction(5001, 10000)
        assert abs(g - ex) <= bound, (i, float(abs(g - ex) / bound))


@pytest.mark.parametrize("nlimb", [4, 8, 12, 16])
def test_sqrt_and_rsqrt(shim, nlimb):
    import mpmath
    rng = random.Random(99 + nlimb)
    n = 200
    a = rand_nums(rng, n, nlimb, special=True)
    a.sign[:] = abs(a.sign)
    p = 32 * nlimb
    got = run(shim, "s", a)
    gotr = run(shim, "r", a)
    with mpmath.workprec(2 * p + 64):
        for i in range(n):
            if a.sign[i] == 0:
                assert got.sign[i] == 0
                continue
            x = a.to_mpf(i)
            s = mpmath.sqrt(x)
            rel = abs(got.to_mpf(i) - s) / s
            assert rel <= mpmath.mpf(2) ** (-p) * (1 + mpmath.mpf(2) ** -20), (i, rel)
            rel2 = abs(gotr.to_mpf(i) * s - 1)
            assert rel2 < mpmath.mpf(2) ** (-p + 1), (i, rel2)


@pytest.mark.parametrize("nlimb", [4, 8])
def test_cmp_and_int(shim, nlimb):
    rng = random.Random(5)
    n = 300
    a = rand_nums(rng, n, nlimb, erange=3)
    b = rand_nums(rng, n, nlimb, erange=3)
    for i in range(0, n, 5):
        m, e = a.get_int(i)
        b.set_int(i, m, e)
    got = run(shim, "c", a, b)
    fa, fb = a.to_fractions(), b.to_fractions()
    for i in range(n):
        want = (fa[i] > fb[i]) - (fa[i] < fb[i])
        assert got.to_fraction(i) == want
    # from_int
    vals = [0, 1, -1, 7, -12345678901234, (1 << 62) + 12345, -(1 << 63) + 1]
    src = MpArray(len(vals), nlimb)
    for i, v in enumerate(vals):
        src.sign[i] = -1 if v < 0 else 1
        src.exp[i] = abs(v)
    got = run(shim, "i", src, src)
    for i, v in enumerate(vals):
        assert got.to_fraction(i) == v

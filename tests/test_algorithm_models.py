"""CPU models of two identities the CUDA kernels rely on (exact integer / rational arithmetic, no GPU):
the balanced radix-256 digits of the slicer by one addition and one xor (csrc/gemm_i8.cu: fixed_point_digits), and the
algebra of the equilibrated Schur factorisation (csrc/solver.cu: chol_inverse with d_keep_scale, decomposition(),
search_direction())."""
import random
from fractions import Fraction

import pytest


def balanced_digits_serial(F, S):
    """digit-serial reference: F = sum d_i 256^i with d_i in [-128, 127], from the two's complement bytes of F."""
    W = F & ((1 << (8 * S + 64)) - 1)
    out, carry = [], 0
    for i in range(S):
        v = ((W >> (8 * i)) & 0xFF) + carry
        carry = 1 if v >= 128 else 0
        out.append(v - (carry << 8))
    return out


def balanced_digits_add_xor(F, S):
    """what the kernel does: bytes of ((F + M) xor M) read as int8, M = 0x80 in each of the S byte positions."""
    M = sum(0x80 << (8 * i) for i in range(S))
    W = ((F & ((1 << (8 * S + 64)) - 1)) + M) & ((1 << (8 * S + 64)) - 1)
    W ^= M
    return [((W >> (8 * i)) & 0xFF) - (256 if (W >> (8 * i)) & 0x80 else 0) for i in range(S)]


@pytest.mark.parametrize("S", [18, 34, 50, 66])
def test_balanced_digits_by_add_and_xor(S):
    rng = random.Random(S)
    lim = 1 << (8 * S - 2)                                    # |F| < 2^(8S-2): the slicer's window
    cases = [0, 1, -1, lim - 1, -(lim - 1), 127, 128, -128, -129, 0x7F7F7F, 0x808080, -0x808080]
    cases += [rng.randrange(-lim + 1, lim) for _ in range(300)]
    for F in cases:
        d = balanced_digits_add_xor(F, S)
        assert d == balanced_digits_serial(F, S)
        assert all(-128 <= v <= 127 for v in d)
        assert sum(v << (8 * i) for i, v in enumerate(d)) == F          # the digits represent F exactly


def _inv(A):
    n = len(A)
    M = [list(map(Fraction, r)) + [Fraction(int(i == j)) for j in range(n)] for i, r in enumerate(A)]
    for c in range(n):
        p = next(r for r in range(c, n) if M[r][c] != 0)
        M[c], M[p] = M[p], M[c]
        M[c] = [v / M[c][c] for v in M[c]]
        for r in range(n):
            if r != c and M[r][c] != 0:
                M[r] = [a - M[r][c] * b for a, b in zip(M[r], M[c])]
    return [r[n:] for r in M]


def _mul(A, B):
    return [[sum(a * b for a, b in zip(r, c)) for c in zip(*B)] for r in A]


def test_equilibrated_schur_algebra():
    """With S' = D^-1 S D^-1 (D a diagonal of powers of two) and B' = D^-1 B:
         Q  = B^T S^-1 B        = B'^T S'^-1 B'                      (the slicer's d_kshift on B)
         dx = S^-1 (rhs + B dy) = D^-1 S'^-1 (D^-1 rhs + B' dy)      (vec_scale on rhs and on dx)
    so the library never needs the factor of S itself. Exact rational arithmetic, graded S."""
    rng = random.Random(7)
    n, ny = 5, 3
    G = [[Fraction(rng.randint(-9, 9)) for _ in range(n)] for _ in range(n)]
    grade = [Fraction(2) ** e for e in (40, -3, 17, 0, -25)]           # a graded SPD matrix: diag(g) (G G^T + I) diag(g)
    S0 = _mul(G, [list(r) for r in zip(*G)])
    S = [[grade[i] * (S0[i][j] + (i == j)) * grade[j] for j in range(n)] for i in range(n)]
    B = [[Fraction(rng.randint(-5, 5)) * grade[i] for _ in range(ny)] for i in range(n)]
    rhs = [[Fraction(rng.randint(-5, 5)) * grade[i]] for i in range(n)]
    dy = [[Fraction(rng.randint(-5, 5))] for _ in range(ny)]
    # D = diag(2^ceil(e_ii / 2)) with e_ii the binary exponent of S_ii (equil_exponents)
    def exp2(x):
        e = 0
        while x >= 1:
            x /= 2; e += 1
        while x < Fraction(1, 2):
            x *= 2; e -= 1
        return e
    s = [-((-exp2(S[i][i])) // 2) for i in range(n)]
    D = [Fraction(2) ** e for e in s]
    Sp = [[S[i][j] / (D[i] * D[j]) for j in range(n)] for i in range(n)]
    assert all(Fraction(1, 4) <= Sp[i][i] < 1 for i in range(n))          # the equilibrated diagonal lies in [1/4, 1)
    Bp = [[B[i][j] / D[i] for j in range(ny)] for i in range(n)]
    Si, Spi = _inv(S), _inv(Sp)
    BT, BpT = [list(r) for r in zip(*B)], [list(r) for r in zip(*Bp)]
    assert _mul(_mul(BT, Si), B) == _mul(_mul(BpT, Spi), Bp)
    u = [[rhs[i][0] + sum(B[i][k] * dy[k][0] for k in range(ny))] for i in range(n)]
    dx = _mul(Si, u)
    up = [[rhs[i][0] / D[i] + sum(Bp[i][k] * dy[k][0] for k in range(ny))] for i in range(n)]
    dxp = _mul(Spi, up)
    assert dx == [[dxp[i][0] / D[i]] for i in range(n)]

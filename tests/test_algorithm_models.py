"""CPU models of what the CUDA kernels and their orchestration rely on (exact integer / rational arithmetic, no GPU):
the balanced radix-256 digits of the slicer by one addition and one xor (csrc/gemm_i8.cu: fixed_point_digits), the
algebra of the equilibrated Schur factorisation (csrc/solver.cu: chol_inverse with d_keep_scale, decomposition(),
search_direction()), the signed division-free elimination of panel_factor_kernel (csrc/linalg.cu), and the stream
schedule of the blocked factorisation with lookahead (every data hazard ordered, every update applied once)."""
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


# ---------------------------------------------------------------------------------------------------------------------
# the signed, division-free elimination of panel_factor_kernel (csrc/linalg.cu) in exact rational arithmetic
# ---------------------------------------------------------------------------------------------------------------------
def _signed_elimination(A):
    """[A | I] -> [R' | G'] by  row_r <- mu_k row_r - R'_k[r] row_k  (mu_k = the current pivot R'_kk; the kernel also
    divides both factors by 2^e_k, which cancels in everything below). Returns R', G' and tau (the common scale of the
    unpivoted rows: tau_0 = 1, tau_{k+1} = mu_k tau_k)."""
    n = len(A)
    R = [[Fraction(v) for v in row] for row in A]
    G = [[Fraction(int(i == j)) for j in range(n)] for i in range(n)]
    tau = [Fraction(1)]
    for k in range(n - 1):
        mu = R[k][k]
        assert mu != 0
        for r in range(k + 1, n):
            gam = R[k][r]                      # by symmetry of the trailing matrix R'[r][k] = R'[k][r] (what the kernel reads)
            R[r] = [mu * a - gam * b for a, b in zip(R[r], R[k])]
            G[r] = [mu * a - gam * b for a, b in zip(G[r], G[k])]
        tau.append(mu * tau[-1])
    return R, G, tau


@pytest.mark.parametrize("seed", range(6))
def test_signed_division_free_factorisation_identities(seed):
    """For symmetric A with non-zero leading minors of EITHER sign (what the Schur complements of a polynomial programme
    look like at the precision limit): G' A G'^T = diag(tau_k d'_k) with d'_k = R'_kk, exactly. Hence
    A^-1 = M^T Sigma M with M = diag(|tau_k d'_k|^-1/2) G', Sigma_k = sign(tau_k d'_k) - the factorisation
    clrsdp_op_signed_factor returns - without a division or a square root inside the elimination, and with negative
    pivots costing nothing."""
    rng = random.Random(100 + seed)
    n = 6
    while True:
        B = [[Fraction(rng.randint(-6, 6)) for _ in range(n)] for _ in range(n)]
        sg = [rng.choice([-1, 1]) for _ in range(n)]
        A = [[sum(B[i][k] * sg[k] * B[j][k] for k in range(n)) for j in range(n)] for i in range(n)]   # indefinite, symmetric
        try:
            R, G, tau = _signed_elimination(A)
        except AssertionError:
            continue
        if R[n - 1][n - 1] != 0:
            break
    # R' is upper triangular, G' lower triangular with diagonal tau
    assert all(R[i][j] == 0 for i in range(n) for j in range(i))
    assert all(G[i][j] == 0 for i in range(n) for j in range(i + 1, n)) and [G[i][i] for i in range(n)] == tau
    assert _mul(G, A) == R                                                                              # the row operations
    D = _mul(_mul(G, A), [list(r) for r in zip(*G)])
    assert all(D[i][j] == (tau[i] * R[i][i] if i == j else 0) for i in range(n) for j in range(n))
    # A^-1 = G'^T diag(1 / (tau_k d'_k)) G'  (= M^T Sigma M after pulling |.|^-1/2 into M)
    Dinv = [[(1 / (tau[i] * R[i][i]) if i == j else Fraction(0)) for j in range(n)] for i in range(n)]
    assert _mul(_mul([list(r) for r in zip(*G)], Dinv), G) == _inv(A)


# ---------------------------------------------------------------------------------------------------------------------
# the stream schedule of the blocked factorisation with lookahead (csrc/solver.cu: chol_inverse)
# ---------------------------------------------------------------------------------------------------------------------
def _chol_inverse_schedule(nb, lookahead, inverse_helper=True):
    """The launches chol_inverse issues for a matrix of nb panels, as (stream, name, reads, writes) with the event edges
    between the streams; regions are panel-sized blocks ('U', r, c) of the work matrix and ('L', r, c) of the inverse.
    Mirrors the control flow of Solver::chol_inverse (main = 'M', inverse-panel helper = 'H', remainder helper = 'T')."""
    ops, edges = [], []            # edges: (index of the op after which the event is recorded, index of the op that waits)
    last = {"M": None, "H": None, "T": None}
    pending_waits = {"M": [], "H": [], "T": []}

    def emit(stream, name, reads, writes):
        ops.append((stream, name, frozenset(reads), frozenset(writes)))
        i = len(ops) - 1
        for src in pending_waits[stream]:
            edges.append((src, i))
        pending_waits[stream] = []
        last[stream] = i
        return i

    def wait(stream, on_stream):                     # cudaEventRecord(on_stream) ; cudaStreamWaitEvent(stream)
        if last[on_stream] is not None:
            pending_waits[stream].append(last[on_stream])

    trail_pending = False
    for k in range(nb):
        emit("M", f"panel{k}", {("U", k, k)}, {("U", k, k), ("L", k, k)})
        if inverse_helper and nb > 2 and k >= 1:
            wait("H", "M")
            emit("H", f"inv{k}", {("U", r, k) for r in range(k)} | {("L", r, c) for r in range(k) for c in range(r + 1)} | {("L", k, k)},
                 {("L", k, c) for c in range(k)})
        rest = list(range(k + 1, nb))
        if not rest:
            continue
        emit("M", f"u12_{k}", {("L", k, k)} | {("U", k, c) for c in rest}, {("U", k, c) for c in rest})
        if trail_pending:
            wait("M", "T")
            trail_pending = False
        rem = rest[1:]
        if lookahead and rem:
            wait("T", "M")
            emit("T", f"rem{k}", {("U", k, c) for c in rem}, {("U", r, c) for r in rem for c in rem})
            trail_pending = True
            emit("M", f"row{k}", {("U", k, c) for c in rest}, {("U", rest[0], c) for c in rest})
        else:
            emit("M", f"trail{k}", {("U", k, c) for c in rest}, {("U", r, c) for r in rest for c in rest})
    wait("M", "T")
    wait("M", "H")
    if not (inverse_helper and nb > 2):
        for k in range(1, nb):
            emit("M", f"inv{k}", {("U", r, k) for r in range(k)} | {("L", r, c) for r in range(k) for c in range(r + 1)} | {("L", k, k)},
                 {("L", k, c) for c in range(k)})
    emit("M", "join", set(), set())
    return ops, edges


@pytest.mark.parametrize("nb", [2, 3, 5, 8])
@pytest.mark.parametrize("lookahead", [False, True])
def test_lookahead_schedule_orders_every_hazard_and_applies_every_update_once(nb, lookahead):
    ops, edges = _chol_inverse_schedule(nb, lookahead)
    n = len(ops)
    # happens-before: program order within a stream + event edges, transitively
    hb = [[False] * n for _ in range(n)]
    prev = {}
    for i, (s, *_rest) in enumerate(ops):
        if s in prev:
            hb[prev[s]][i] = True
        prev[s] = i
    for a, b in edges:
        hb[a][b] = True
    for k in range(n):
        for i in range(n):
            if hb[i][k]:
                row_k, row_i = hb[k], hb[i]
                for j in range(n):
                    if row_k[j]:
                        row_i[j] = True
    for i in range(n):
        for j in range(i + 1, n):
            _, ni, ri, wi = ops[i]
            _, nj, rj, wj = ops[j]
            if (wi & (rj | wj)) or (ri & wj):
                assert hb[i][j], (ni, nj)            # issued earlier => must complete earlier when they touch the same block
    # every trailing block (r, c), r <= c, receives the contribution of every earlier panel exactly once
    for r in range(1, nb):
        for c in range(r, nb):
            for k in range(r):
                hits = [name for _, name, _, w in ops if ("U", r, c) in w and name in (f"trail{k}", f"rem{k}", f"row{k}")]
                assert len(hits) == 1, (k, r, c, hits)
    # everything is joined back to the main stream before the call returns
    join = n - 1
    assert all(hb[i][join] for i in range(n - 1))

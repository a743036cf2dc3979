import random

import numpy as np

from clrsdp.wire import MpArray


def rand_mp(rng, n, nlimb, erange=6, zero_frac=0.0):
    p = 32 * nlimb
    m = [0 if rng.random() < zero_frac else (rng.getrandbits(p) | (1 << (p - 1))) * rng.choice([-1, 1]) for _ in range(n)]
    e = [rng.randint(-erange, erange) - p for _ in range(n)]
    return MpArray.from_ints(m, e, nlimb)


def exact_planes(A, B, batch, M, N, K, rexp, cexp, S):
    """python big-int model of the slicing (balanced radix-256 digits relative to the row exponent,
    F = trunc(x * 2^(8S-2-e_row))) and of the plane products  D_t = sum_{a+b=t} A_a B_b^T."""
    def digits(x, i, rowexp):
        m, e = x.get_int(i)
        if m == 0:
            return [0] * S
        sh = e + 8 * S - 2 - rowexp
        F = (abs(m) << sh) if sh >= 0 else (abs(m) >> (-sh))
        F = -F if m < 0 else F
        ds = []
        for _ in range(S):
            d = F & 0xFF
            d = d - 256 if d >= 128 else d
            ds.append(d)
            F = (F - d) >> 8
        assert F == 0
        return ds[::-1]
    planes = np.zeros((S, batch, M, N), dtype=object)
    for b in range(batch):
        Ad = np.array([[digits(A, (b * M + i) * K + k, int(rexp[b, i])) for k in range(K)] for i in range(M)], dtype=object)
        Bd = np.array([[digits(B, (b * K + k) * N + j, int(cexp[b, j])) for k in range(K)] for j in range(N)], dtype=object)
        for t in range(S):
            acc = np.zeros((M, N), dtype=object)
            for a in range(t + 1):
                acc = acc + Ad[:, :, a].dot(Bd[:, :, t - a].T)
            planes[t, b] = acc
    return planes


def spd_batch(rng, batch, n, nlimb):
    mats = []
    for _ in range(batch):
        G = np.array([[rng.uniform(-1, 1) for _ in range(n)] for _ in range(n)])
        A = G @ G.T + 0.1 * n * np.eye(n)
        mats.append((A + A.T) / 2)
    return MpArray.from_double(np.array(mats).reshape(-1), nlimb)

"""Problem files in the wire format: the tuple list `(A, B, c, H)` of `prepareabc`, the objective `b`
and, optionally, a reference solution, in one self-describing binary file.

Purpose: a maintainer with a Julia installation dumps the reference's own `prepareabc` output (and the result of its
`solverank1sdp`) with `ClrsdpB200.write_problem` (julia/ClrsdpB200.jl); dropped under tests/golden/ such a file pins both
the oracle and the GPU path against the real reference — the one thing this repository cannot produce itself (no Julia
in the build image: DESIGN.md §2, "parity unpinned").

Layout (little endian):  b"CLRSDP1\\n" | uint64 header_bytes | JSON header | arrays.
Header: {"prec", "n_y", "b0": "<decimal string>", "clusters": [{"m", "K", "L", "delta": [..], "ranks": [[..K..], ..]}],
         "solution": null | {"iterations", "primal_obj", "dual_obj", "has_point": bool}}.
Arrays, in this order, each as  uint64 n | int8 sign[n] | int64 exp[n] | uint32 limb[nlimb][n]  (nlimb = prec / 32):
  b;  for every cluster: V[0..L-1] (vectors of block l, k outer / rank inner, each vector contiguous), H[0..L-1], B
  (dim_S x n_y row-major, rows (r, s, k) with k fastest), c;  then, if has_point: x, X, y, Y (X, Y flat in block order).
"""
from __future__ import annotations

import json
import struct

import numpy as np

from .solver import Constraint
from .wire import MpArray

MAGIC = b"CLRSDP1\n"


def _put(f, a: MpArray):
    f.write(struct.pack("<Q", a.n))
    f.write(np.ascontiguousarray(a.sign, dtype="<i1").tobytes())
    f.write(np.ascontiguousarray(a.exp, dtype="<i8").tobytes())
    f.write(np.ascontiguousarray(a.limb, dtype="<u4").tobytes())


def _get(f, nlimb) -> MpArray:
    (n,) = struct.unpack("<Q", f.read(8))
    a = MpArray(n, nlimb)
    a.sign[:] = np.frombuffer(f.read(n), dtype="<i1")
    a.exp[:] = np.frombuffer(f.read(8 * n), dtype="<i8")
    a.limb[:] = np.frombuffer(f.read(4 * n * nlimb), dtype="<u4").reshape(nlimb, n)
    return a


def save_problem(path, constraints, b: MpArray, b0="0", solution=None, point=None):
    """solution: optional dict(iterations=.., primal_obj="..", dual_obj=".."); point: optional (x, X, y, Y) MpArrays."""
    nlimb = b.nlimb
    clusters = []
    for c in constraints:
        K = c.n_samples
        m_pairs = c.c.n // K
        m = int(round((np.sqrt(8 * m_pairs + 1) - 1) / 2))
        clusters.append(dict(m=m, K=K, L=c.L, delta=[int(c.V[l].shape[1]) if c.V[l].n else 0 for l in range(c.L)],
                             ranks=[[int(v) for v in c.ranks[l]] for l in range(c.L)]))
    sol = None
    if solution is not None or point is not None:
        sol = dict(solution or {})
        sol["has_point"] = point is not None
    hdr = json.dumps(dict(prec=32 * nlimb, n_y=int(b.n), b0=str(b0), clusters=clusters, solution=sol)).encode()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(hdr)))
        f.write(hdr)
        _put(f, b)
        for c in constraints:
            for l in range(c.L):
                _put(f, c.V[l].reshape(c.V[l].n))
            for l in range(c.L):
                _put(f, c.H[l])
            _put(f, c.B.reshape(c.B.n))
            _put(f, c.c)
        if point is not None:
            for a in point:
                _put(f, a.reshape(a.n))


def load_problem_file(path):
    """-> (constraints, b, meta): meta = the JSON header, with meta["point"] = (x, X, y, Y) when the file holds one."""
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError("not a CLRSDP1 problem file")
        (hl,) = struct.unpack("<Q", f.read(8))
        meta = json.loads(f.read(hl).decode())
        nlimb = meta["prec"] // 32
        n_y = meta["n_y"]
        b = _get(f, nlimb)
        if b.n != n_y:
            raise ValueError("objective length does not match the header")
        constraints = []
        for cl in meta["clusters"]:
            V, H, ranks = [], [], []
            for l in range(cl["L"]):
                rk = np.asarray(cl["ranks"][l], dtype=np.int32)
                v = _get(f, nlimb)
                nv = int(rk.sum())
                if v.n != nv * cl["delta"][l]:
                    raise ValueError("vector block does not match ranks x delta")
                V.append(v.reshape(nv, cl["delta"][l]))
                ranks.append(rk)
            for l in range(cl["L"]):
                H.append(_get(f, nlimb))
            dimS = cl["m"] * (cl["m"] + 1) // 2 * cl["K"]
            B = _get(f, nlimb)
            c = _get(f, nlimb)
            if B.n != dimS * n_y or c.n != dimS:
                raise ValueError("B / c do not match dim_S x n_y")
            constraints.append(Constraint(V=V, ranks=ranks, H=H, B=B.reshape(dimS, n_y), c=c))
        if meta.get("solution") and meta["solution"].get("has_point"):
            meta["point"] = tuple(_get(f, nlimb) for _ in range(4))
    return constraints, b, meta

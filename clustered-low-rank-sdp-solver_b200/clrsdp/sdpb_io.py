"""SDPB-format problem files: what the example's missing `WriteFilesSDPB.write_files(file_path,
constraints, blockinfo, b)` does (examples/SpherePacking.jl:6, :95-98) — the sampled constraint tuples `(A, B, c, H)` of
`prepareabc` written as an SDPB input directory, so that the same instance can be handed to SDPB itself for
cross-validation.

The reference solves "a more general version" of SDPB's problem (README.md:2) in SDPB's own conventions
(MPMP.jl:615, :642-657): maximise b0 + b.y subject to Tr(A_p Y) + (B y)_p = c_p, Y >= 0, with A_p built from the
sample-point vectors. A cluster is expressible in SDPB's format exactly when it is one of SDPB's blocks: at most two
polynomial weights (L <= 2: SDPB's "even" and "odd" bilinear bases), one vector per sample (rank 1) and positive signs
H (a positive H != 1 is folded into the vector as sqrt(H)); anything else - rank > 1 samples, negative H, L > 2, the
generality this solver adds over SDPB - raises ValueError naming the cluster.

Layout (the JSON `sdp` directory of SDPB 2.5+, one pair of files per block; written from the format's published
description, NOT validated against an SDPB build - there is none in this image):
  control.json       {"num_blocks": J, "command": "..."}
  objectives.json    {"constant": "<b0>", "b": ["<b_1>", ...]}
  block_info_<j>.json {"dim": m, "num_points": K}
  block_data_<j>.json {"c": [dim_S strings], "B": [[n_y strings] x dim_S],
                       "bilinear_bases_even": [[K strings] x delta_0],   (row d = basis element d at every sample)
                       "bilinear_bases_odd":  [[K strings] x delta_1]}   (absent weight: [])
Rows of c and B are ordered (r, s, k), s <= r, k fastest — the reference's order (MPMP.jl:387-400) — which is SDPB's.
Numbers are decimal strings carrying the full working precision (`prec` bits -> ceil(prec * log10(2)) + 3 digits).

`read_sdpb` reads such a directory back into `Constraint` tuples (round trip: tests/test_frontend.py).
"""
from __future__ import annotations

import json
import math
import os

import mpmath
import numpy as np

from .solver import Constraint
from .wire import MpArray


def _ctx(prec):
    mp = mpmath.mp.clone()
    mp.prec = prec + 32
    return mp


def _digits(prec):
    return int(math.ceil(prec * math.log10(2))) + 3


def _strs(mp, values, nd):
    return [mp.nstr(v, nd, strip_zeros=False, min_fixed=-1, max_fixed=-1) if v != 0 else "0" for v in values]


def write_sdpb(path, constraints, blockinfo, b: MpArray, b0="0", command="clrsdp.sdpb_io.write_sdpb"):
    """`write_files(file_path, constraints, blockinfo, b)` of the reference's example (ex:97). `blockinfo` is the
    `BlockInfo` of `get_block_info(constraints)`; it is checked against the constraints like the solver checks it."""
    nlimb = b.nlimb
    prec = 32 * nlimb
    mp, nd = _ctx(prec), _digits(prec)
    if blockinfo.J != len(constraints) or blockinfo.n_y != b.n:
        raise ValueError("blockinfo does not describe these constraints")
    os.makedirs(path, exist_ok=True)
    with mp.workprec(prec + 32):
        for j, c in enumerate(constraints):
            K, m = c.n_samples, blockinfo.m[j]
            if c.L > 2:
                raise ValueError(f"cluster {j}: {c.L} polynomial weights; SDPB's format has two bilinear bases")
            bases = []
            for l in range(c.L):
                if any(int(r) != 1 for r in c.ranks[l]):
                    raise ValueError(f"cluster {j}, block {l}: samples of rank != 1 are not expressible in SDPB's format")
                h = c.H[l].to_mpfs()
                if any(not v > 0 for v in h):
                    raise ValueError(f"cluster {j}, block {l}: SDPB's format needs positive signs H")
                delta = int(c.V[l].shape[1])
                v = c.V[l].reshape(c.V[l].n).to_mpfs()
                root = [mp.sqrt(mp.mpf(x)) for x in h]
                # row d of the basis matrix = element d of the vector of every sample, sqrt(H) folded in
                bases.append([_strs(mp, [mp.mpf(v[k * delta + d]) * root[k] for k in range(K)], nd) for d in range(delta)])
            while len(bases) < 2:
                bases.append([])
            dimS, n_y = c.B.shape
            Bv = c.B.reshape(c.B.n).to_mpfs()
            data = dict(c=_strs(mp, c.c.to_mpfs(), nd),
                        B=[_strs(mp, Bv[r * n_y:(r + 1) * n_y], nd) for r in range(dimS)],
                        bilinear_bases_even=bases[0], bilinear_bases_odd=bases[1])
            with open(os.path.join(path, f"block_info_{j}.json"), "w") as f:
                json.dump(dict(dim=int(m), num_points=int(K)), f)
            with open(os.path.join(path, f"block_data_{j}.json"), "w") as f:
                json.dump(data, f)
        with open(os.path.join(path, "objectives.json"), "w") as f:
            json.dump(dict(constant=str(b0), b=_strs(mp, b.to_mpfs(), nd)), f)
        with open(os.path.join(path, "control.json"), "w") as f:
            json.dump(dict(num_blocks=len(constraints), command=command, precision=prec), f)


def read_sdpb(path, prec=None):
    """-> (constraints, b, b0): the directory `write_sdpb` wrote (or one from SDPB's own tools in the same layout), as
    `Constraint` tuples with H = 1 and rank-1 samples, rounded to `prec` bits (default: the precision recorded by
    `write_sdpb`, else 256)."""
    with open(os.path.join(path, "control.json")) as f:
        ctl = json.load(f)
    prec = int(prec or ctl.get("precision", 256))
    nlimb = prec // 32
    mp = _ctx(prec)
    with mp.workprec(prec + 32):
        with open(os.path.join(path, "objectives.json")) as f:
            obj = json.load(f)
        b = MpArray.from_mpf([mp.mpf(s) for s in obj["b"]], nlimb)
        constraints = []
        for j in range(int(ctl["num_blocks"])):
            with open(os.path.join(path, f"block_info_{j}.json")) as f:
                info = json.load(f)
            with open(os.path.join(path, f"block_data_{j}.json")) as f:
                data = json.load(f)
            K, m = int(info["num_points"]), int(info["dim"])
            V, H, ranks = [], [], []
            for key in ("bilinear_bases_even", "bilinear_bases_odd"):
                rows = data.get(key) or []
                if not rows:
                    continue
                delta = len(rows)
                vals = [mp.mpf(rows[d][k]) for k in range(K) for d in range(delta)]
                V.append(MpArray.from_mpf(vals, nlimb).reshape(K, delta))
                H.append(MpArray.from_mpf([mp.mpf(1)] * K, nlimb))
                ranks.append(np.ones(K, dtype=np.int32))
            dimS = m * (m + 1) // 2 * K
            if len(data["c"]) != dimS or len(data["B"]) != dimS:
                raise ValueError(f"block {j}: c / B do not have dim * (dim + 1) / 2 * num_points rows")
            Bv = [mp.mpf(s) for row in data["B"] for s in row]
            constraints.append(Constraint(V=V, ranks=ranks, H=H, B=MpArray.from_mpf(Bv, nlimb).reshape(dimS, b.n),
                                          c=MpArray.from_mpf([mp.mpf(s) for s in data["c"]], nlimb)))
    return constraints, b, obj.get("constant", "0")

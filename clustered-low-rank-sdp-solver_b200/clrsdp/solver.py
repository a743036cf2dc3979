"""Host-side mirror of the reference's solver interface for the hot path.

Same names, argument meaning and error behaviour as MPMP.jl: `BlockInfo` (:467-513), `get_block_info`
(:516-560), `solverank1sdp` (:595-1025, keyword defaults :599-613, return tuple :1014-1024). The body
of the interior-point loop is NOT here: it runs on the GPU behind the C ABI (include/clrsdp.h); this
module only packs the problem into the wire format, drives `clrsdp_solve`/`clrsdp_iterate` and prints
the reference's log table. A Julia front end would do exactly the same through `ccall`
(INTEGRATION.md).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from fractions import Fraction

import numpy as np

from . import capi
from .wire import MpArray

_PRECISION = 256


def set_precision(bits: int):
    """Counterpart of `setprecision(BigFloat, bits)`; the reference takes its working precision from
    the global `precision(BigFloat)` (MPMP.jl:617 and every allocation)."""
    global _PRECISION
    if bits % 32 or not (128 <= bits <= 512):
        raise ValueError("precision must be a multiple of 32 in [128, 512]")
    _PRECISION = int(bits)


def precision() -> int:
    return _PRECISION


@dataclass
class Constraint:
    """The tuple (A, B, c, H) that `prepareabc` returns (MPMP.jl:385-406), stored densely.

    V[l]      MpArray (Nv_l, delta_l): the vectors A[l,k][rnk], k outer / rnk inner (hcat order, :1249-1254)
    ranks[l]  int array [K]: number of vectors at sample k (length of A[l,k])
    H[l]      MpArray (Nv_l,): A_sign[l,k][rnk] in the same order
    B         MpArray (dim_S, n_y), rows (r,s,k) with k fastest (:387-395)
    c         MpArray (dim_S,)
    """
    V: list
    ranks: list
    H: list
    B: MpArray
    c: MpArray

    @property
    def L(self):
        return len(self.V)

    @property
    def n_samples(self):
        return len(self.ranks[0])


@dataclass
class BlockInfo:
    """MPMP.jl:467-513 (field for field; indices 0-based)."""
    J: int
    n_y: int
    m: list
    L: list
    n_samples: list
    Y_blocksizes: list
    dim_S: list
    ranks: list
    x_indices: list = field(default_factory=list)
    rank_sums: list = field(default_factory=list)
    nz_k: list = field(default_factory=list)
    jl_pairs: list = field(default_factory=list)
    delta: list = field(default_factory=list)

    def __post_init__(self):
        J = self.J
        if not (len(self.m) == len(self.L) == len(self.n_samples) == len(self.dim_S) == J):
            raise ValueError("sizes of m,L,n_samples,dim_S must equal the number of constraints")
        if [len(r) for r in self.ranks] != list(self.L) or [len(y) for y in self.Y_blocksizes] != list(self.L):
            raise ValueError("Y[j] and ranks[j] must have length L[j]")
        self.x_indices = [int(sum(self.dim_S[:j])) for j in range(J + 1)]
        self.rank_sums = [[[0] + list(np.cumsum(self.ranks[j][l])) for l in range(self.L[j])] for j in range(J)]
        self.nz_k = [[next((k for k in range(self.n_samples[j]) if self.ranks[j][l][k] > 0), None)
                      for l in range(self.L[j])] for j in range(J)]
        # the reference reorders jl_pairs for its thread pool (:492-499); on the GPU all blocks of a
        # shape run as one batch, so the natural order is kept.
        self.jl_pairs = [(j, l) for j in range(J) for l in range(self.L[j])]
        if not self.delta:
            self.delta = [[self.Y_blocksizes[j][l] // self.m[j] for l in range(self.L[j])] for j in range(J)]


def get_block_info(constraints) -> BlockInfo:
    """MPMP.jl:516-560."""
    J = len(constraints)
    n_y = constraints[0].B.shape[1]
    L = [c.L for c in constraints]
    n_samples = [c.n_samples for c in constraints]
    m = []
    for j, c in enumerate(constraints):
        t = c.c.n // n_samples[j]
        mj = (-1 + int(np.sqrt(8 * t + 1) + 0.5)) // 2  # isqrt form of :531-534
        while mj * (mj + 1) // 2 > t:
            mj -= 1
        assert c.c.n == mj * (mj + 1) * n_samples[j] // 2
        m.append(mj)
    ranks = [[[int(r) for r in c.ranks[l]] for l in range(c.L)] for c in constraints]
    Y_blocksizes = [[m[j] * constraints[j].V[l].shape[1] for l in range(L[j])] for j in range(J)]
    dim_S = [m[j] * (m[j] + 1) // 2 * n_samples[j] for j in range(J)]
    return BlockInfo(J, n_y, m, L, n_samples, Y_blocksizes, dim_S, ranks)


def _to_mp_scalar(v, nlimb) -> MpArray:
    if isinstance(v, MpArray):
        return v
    try:
        import mpmath
        if isinstance(v, mpmath.mpf):
            return MpArray.from_mpf([v], nlimb)
    except ImportError:  # pragma: no cover
        pass
    if isinstance(v, float):
        return MpArray.from_double([v], nlimb)
    return MpArray.from_fraction([Fraction(v)], nlimb)


DEFAULTS = dict(  # MPMP.jl:599-613
    b0=0, maxiterations=500, beta_infeasible=Fraction(3, 10), beta_feasible=Fraction(1, 10),
    gamma=Fraction(7, 10), omega_p=Fraction(10) ** 10, omega_d=Fraction(10) ** 10,
    duality_gap_threshold=Fraction(1, 10 ** 15), primal_error_threshold=Fraction(1, 10 ** 30),
    dual_error_threshold=Fraction(1, 10 ** 30), need_primal_feasible=False, need_dual_feasible=False,
)


def real_params(nlimb, **kw) -> MpArray:
    p = dict(DEFAULTS)
    p.update(kw)
    order = ["beta_infeasible", "beta_feasible", "gamma", "omega_p", "omega_d", "duality_gap_threshold",
             "primal_error_threshold", "dual_error_threshold"]
    return MpArray.concat([_to_mp_scalar(p[k], nlimb) for k in order])


def flatten_blocks(M) -> MpArray:
    """A block-diagonal matrix given as list over j of lists over l of MpArray (nb, nb) -> the flat (j,l)-ordered,
    row-major layout the C ABI takes for X, Y and C."""
    if isinstance(M, MpArray):
        return M
    return MpArray.concat([blk.reshape(blk.n) for row in M for blk in row])


def load_problem(h: capi.Handle, constraints, b: MpArray, blockinfo: BlockInfo, b0=0, C=None):
    """set_structure + upload_cluster for every constraint + upload_objective (+ upload_C for C != 0)."""
    bi = blockinfo
    delta = [d for j in range(bi.J) for d in bi.delta[j]]
    ranks = [r for j in range(bi.J) for l in range(bi.L[j]) for r in bi.ranks[j][l]]
    h.set_structure(bi.n_y, bi.m, bi.L, bi.n_samples, delta, ranks)
    for j, c in enumerate(constraints):
        V = MpArray.concat(c.V) if len(c.V) > 1 else c.V[0].reshape(c.V[0].n)
        H = MpArray.concat(c.H) if len(c.H) > 1 else c.H[0]
        h.upload_cluster(j, V, H, c.B, c.c)
    h.upload_objective(b, _to_mp_scalar(b0, h.nlimb))
    if C is not None:
        h.upload_C(flatten_blocks(C))


def product_handle(prec=None, device=0) -> capi.Handle:
    """Handle on the CUDA library. Raises if libclrsdp.so is missing — there is no CPU fallback.
    `device`: a CUDA ordinal, or a list of ordinals for ONE handle that shards the clusters over several GPUs of the
    box inside this process (clrsdp_create_multi)."""
    return capi.Handle(capi.load_product_library(), "clrsdp_", prec or precision(), device)


def partition_clusters(blockinfo: BlockInfo, n_parts: int):
    """Cluster -> part assignment by the library's weighted partitioner (F16, MPMP.jl:425-465 on the cluster weights of
    SURVEY §8e): what clrsdp_create_multi does internally and what a one-process-per-GPU front end calls to pick each
    rank's clusters. Returns (owner[j], max part weight)."""
    bi = blockinfo
    w = [capi.cluster_weight(bi.m[j], bi.L[j], bi.n_samples[j], bi.delta[j], bi.n_y) for j in range(bi.J)]
    return capi.partition(w, n_parts)


HEADER = "%5s %8s %11s %11s %11s %10s %10s %10s %10s %10s %10s %10s" % (
    "iter", "time(s)", "mu", "P-obj", "D-obj", "gap", "P-error", "p-error", "d-error", "alpha_p", "alpha_d", "beta")


def format_row(r, t):
    """The reference's per-iteration row (MPMP.jl:923-937)."""
    return "%5d %8.1f %11.3e %11.3e %11.3e %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e" % (
        r.iter, t, r.mu, r.p_obj, r.d_obj, r.gap, r.P_err, r.p_err, r.d_err, r.alpha_p, r.alpha_d, r.beta_c)


def solverank1sdp(constraints, b, blockinfo: BlockInfo, *, C=0, b0=0, maxiterations=500,
                  beta_infeasible=DEFAULTS["beta_infeasible"], beta_feasible=DEFAULTS["beta_feasible"],
                  gamma=DEFAULTS["gamma"], omega_p=DEFAULTS["omega_p"], omega_d=DEFAULTS["omega_d"],
                  duality_gap_threshold=DEFAULTS["duality_gap_threshold"],
                  primal_error_threshold=DEFAULTS["primal_error_threshold"],
                  dual_error_threshold=DEFAULTS["dual_error_threshold"],
                  need_primal_feasible=False, need_dual_feasible=False, testing=True, initial_solutions=(),
                  verbose=True, handle=None, return_info=False):
    """Solve the SDP with low-rank constraint matrices (MPMP.jl:595-1025) on the B200 path.

    Returns (x, X, y, Y, P, p, d, dual_gap, primal_obj, dual_obj, time_total) like the reference
    (:1014-1024); X, Y, P are lists over j of lists over l of MpArray (nb, nb); the three scalars are
    mpmath mpf at working precision. `handle` lets tests pass a handle on the CPU oracle.
    """
    if isinstance(C, (int, float)) and C == 0:
        C = None                                     # the reference's AbsoluteZero (MPMP.jl:691-695)
    elif isinstance(C, (int, float)):
        raise ValueError("C must be 0 or a block-diagonal matrix with the structure of X (MPMP.jl:599)")
    own = handle is None
    h = handle or product_handle()
    nl = h.nlimb
    b = b if isinstance(b, MpArray) else MpArray.from_mpf(b, nl)
    load_problem(h, constraints, b, blockinfo, b0, C)
    h.set_params(real_params(nl, beta_infeasible=beta_infeasible, beta_feasible=beta_feasible, gamma=gamma,
                             omega_p=omega_p, omega_d=omega_d, duality_gap_threshold=duality_gap_threshold,
                             primal_error_threshold=primal_error_threshold,
                             dual_error_threshold=dual_error_threshold),
                 maxiterations, need_primal_feasible, need_dual_feasible, phase_timing=verbose)
    sizes = [bs for j in range(blockinfo.J) for bs in blockinfo.Y_blocksizes[j]]
    n_X = int(sum(s * s for s in sizes))
    if len(initial_solutions) == 4:  # warm start (:689)
        x0, X0, y0, Y0 = initial_solutions
        h.upload_point(x0, flatten_blocks(X0), y0, flatten_blocks(Y0))
    else:
        h.init_point()
    if verbose:
        print(HEADER)
    t0 = time.time()
    rows = []
    info = h.prepare()
    terminate = _terminate(info, need_primal_feasible, need_dual_feasible)
    it = 1
    while not terminate and it < maxiterations:  # (:742-753)
        r = h.iterate()
        rows.append(r)
        if verbose:
            print(format_row(r, time.time() - t0))
        terminate = r.terminate in (1, 2, 3)
        it += 1
    time_total = time.time() - t0
    if verbose:
        if rows and rows[-1].terminate in (1, 2, 3):
            print(capi.TERMINATE[rows[-1].terminate])
        print(HEADER)
        _print_timings(rows, time_total)
    x, Xf, y, Yf = h.download_point(int(sum(blockinfo.dim_S)), n_X, blockinfo.n_y)
    X, Y, P = [], [], []
    off = 0
    for j in range(blockinfo.J):
        X.append([]), Y.append([]), P.append([])
        for l in range(blockinfo.L[j]):
            s = blockinfo.Y_blocksizes[j][l]
            idx = np.arange(off, off + s * s)
            X[j].append(Xf.take(idx).reshape(s, s))
            Y[j].append(Yf.take(idx).reshape(s, s))
            P[j].append(h.fetch("P", j, l).reshape(s, s))
            off += s * s
    p, d = h.fetch("p"), h.fetch("d")
    # return values (:1021-1023): gap WITHOUT b0 (:1067-1074; <C,Y> is inside dual_obj), objectives with b0
    primal_obj, dual_obj = h.scalar("p_obj"), h.scalar("d_obj")
    import mpmath
    with mpmath.workprec(h.prec):
        b0m = _to_mp_scalar(b0, nl).to_mpf(0)
        po, do = primal_obj - b0m, dual_obj - b0m
        dual_gap = abs(po - do) / max(mpmath.mpf(1), abs(po + do))
    if own:
        h.close()
    out = (x, X, y, Y, P, p, d, dual_gap, primal_obj, dual_obj, time_total)
    return (out, rows) if return_info else out


def _terminate(info, need_p, need_d):
    return info.terminate in (1, 2, 3)


def _print_timings(rows, time_total):
    """The reference's timing report (MPMP.jl:972-1012); first two iterations excluded (:889)."""
    t = np.zeros(capi.T_COUNT)
    for r in rows[2:]:
        t += np.array(list(r.timings))
    print("\nTime spent: (the first two iterations are not included in the per-phase times)")
    print("%11s %11s %11s %11s %11s %11s %11s %11s" % ("total", "Decomp", "predict_dir", "correct_dir", "alpha", "Xinv", "R", "res"))
    print(("%11.5e " * 8) % (time_total, *t[:7]))
    print("Time inside decomp:")
    print("%11s %11s %11s %11s %11s" % ("schur", "chol_S", "comp CinvB", "comp Q", "chol_Q"))
    print(("%11.5e " * 5) % tuple(t[7:12]))
    print("Time inside search directions (both predictor & corrector step)")
    print("%11s %11s %11s %11s %11s" % ("calc Z", "calc rhs x", "solve system", "calc dX", "calc dY"))
    print(("%11.5e " * 5) % tuple(t[12:17]))
    print("(device time of the same iterations, CUDA events around each: %.5e s)" % sum(r.seconds for r in rows[2:]))


# ---- check-pointing the iterate (the reference's warm start: initial_solutions, MPMP.jl:613, :689) ----
def save_checkpoint(path, h: capi.Handle, blockinfo: BlockInfo, iteration=0):
    """Write the current iterate (x, X, y, Y) of handle `h` to `path` (.npz of the wire arrays: bit-exact, any
    precision). X and Y are stored flat in block order, as `clrsdp_download_point` delivers them."""
    sizes = [bs for j in range(blockinfo.J) for bs in blockinfo.Y_blocksizes[j]]
    n_X = int(sum(s * s for s in sizes))
    x, X, y, Y = h.download_point(int(sum(blockinfo.dim_S)), n_X, blockinfo.n_y)
    arrs = {}
    for name, a in (("x", x), ("X", X), ("y", y), ("Y", Y)):
        arrs[name + "_sign"], arrs[name + "_exp"], arrs[name + "_limb"] = a.sign, a.exp, a.limb
    np.savez_compressed(path, nlimb=h.nlimb, iteration=iteration, dim_S=np.asarray(blockinfo.dim_S),
                        blocks=np.asarray(sizes), n_y=blockinfo.n_y, **arrs)


def load_checkpoint(path, h: capi.Handle, blockinfo: BlockInfo):
    """Upload a saved iterate into `h` (the problem must already be loaded) and return the stored iteration number.
    The structure is checked; the precision must be the handle's (no silent re-rounding)."""
    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    sizes = [bs for j in range(blockinfo.J) for bs in blockinfo.Y_blocksizes[j]]
    if int(z["nlimb"]) != h.nlimb:
        raise ValueError(f"checkpoint is at {32 * int(z['nlimb'])} bits, the handle at {h.prec}")
    if list(z["dim_S"]) != list(blockinfo.dim_S) or list(z["blocks"]) != sizes or int(z["n_y"]) != blockinfo.n_y:
        raise ValueError("checkpoint does not match the block structure of this problem")
    pt = []
    for name in ("x", "X", "y", "Y"):
        a = MpArray(z[name + "_sign"].shape[0], h.nlimb)
        a.sign[:], a.exp[:], a.limb[:] = z[name + "_sign"], z[name + "_exp"], z[name + "_limb"]
        pt.append(a)
    h.upload_point(*pt)
    return int(z["iteration"])

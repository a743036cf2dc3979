"""ctypes binding of the C ABI declared in include/clrsdp.h.

`Handle(lib, prefix)` binds one set of entry points: the product library (`libclrsdp.so`, prefix
`clrsdp_`) or — in tests only — the CPU oracle (`oracle/libclrsdp_ref.so`, prefix `clrsdp_ref_`), which
exports the same functions. There is no fallback between the two: a missing product library raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .wire import MpArray, clrsdp_mp

T_COUNT = 17
TIMING_NAMES = [
    "decomp", "predict_dir", "correct_dir", "alpha", "Xinv", "R", "res",
    "schur", "chol_S", "comp_CinvB", "comp_Q", "chol_Q",
    "calc_Z", "calc_rhs_x", "solve_system", "calc_dX", "calc_dY",
]
P_COUNT = 8
SCALARS = ["mu", "p_obj", "d_obj", "gap", "primal_err", "dual_err", "alpha_p", "alpha_d", "beta_c",
           "mu_p", "mu_c", "lambda_x", "lambda_y"]
STATUS = {
    0: "ok", -1: "bad argument", -2: "CUDA error", -3: "NCCL error",
    -10: "X block not positive definite", -11: "Y block not positive definite",
    -12: "S could not be factorised", -13: "Q could not be factorised",
    -14: "step-length eigenvalue failed", -15: "call order violated",
    -16: "the iterate was lost (mu, a step length or an objective is zero, negative or not finite)",
}
TERMINATE = {0: "running", 1: "Primal feasible solution found", 2: "Dual feasible solution found",
             3: "Optimal solution found", 4: "maximum iterations reached"}


class IterInfo(ctypes.Structure):
    _fields_ = [
        ("iter", ctypes.c_int32), ("status", ctypes.c_int32), ("terminate", ctypes.c_int32),
        ("pd_feasible", ctypes.c_int32),
        ("mu", ctypes.c_double), ("p_obj", ctypes.c_double), ("d_obj", ctypes.c_double),
        ("gap", ctypes.c_double), ("P_err", ctypes.c_double), ("p_err", ctypes.c_double),
        ("d_err", ctypes.c_double), ("alpha_p", ctypes.c_double), ("alpha_d", ctypes.c_double),
        ("beta_c", ctypes.c_double),
        ("p_obj_new", ctypes.c_double), ("d_obj_new", ctypes.c_double), ("gap_new", ctypes.c_double),
        ("primal_err_new", ctypes.c_double), ("dual_err_new", ctypes.c_double),
        ("seconds", ctypes.c_double),
        ("timings", ctypes.c_double * T_COUNT),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "timings"}
        d["timings"] = dict(zip(TIMING_NAMES, list(self.timings)))
        return d


class IntParams(ctypes.Structure):
    _fields_ = [("maxiterations", ctypes.c_int32), ("need_primal_feasible", ctypes.c_int32),
                ("need_dual_feasible", ctypes.c_int32), ("phase_timing", ctypes.c_int32)]


class ClrsdpError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        msg = f"{where}: {STATUS.get(code, code)}"
        if code in (-10, -11, -12, -13, -14, -16):
            # the reference's error strings (MPMP.jl:793,1439,1503,1882)
            msg += " — try again with higher precision"
        if detail:
            msg += f" [{detail}]"
        super().__init__(msg)


_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(os.path.dirname(_HERE), "csrc", "libclrsdp.so")


def load_product_library():
    """Load libclrsdp.so (built in-tree by build.py / __graft_entry__.build). Fails loudly if absent."""
    if not os.path.exists(PRODUCT_LIB):
        raise FileNotFoundError(
            f"{PRODUCT_LIB} is missing: run `python __graft_entry__.py` (build()) first. "
            "There is no CPU fallback for the hot path.")
    return ctypes.CDLL(PRODUCT_LIB, mode=ctypes.RTLD_GLOBAL)


def partition(weights, parts):
    """clrsdp_partition (F16, the reference's distribute_weights_swapping, MPMP.jl:425-465): (set_of, max set weight)."""
    lib = load_product_library()
    f = lib.clrsdp_partition
    f.restype = ctypes.c_double
    f.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
    out = np.zeros(len(w), dtype=np.int32)
    mx = f(w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(w), int(parts), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    if mx < 0:
        raise ClrsdpError(-1, "clrsdp_partition")
    return out, mx


def cluster_weight(m, L, n_samples, delta, n_y):
    """clrsdp_cluster_weight: w_j of SURVEY §8e."""
    lib = load_product_library()
    f = lib.clrsdp_cluster_weight
    f.restype = ctypes.c_double
    f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    d = np.ascontiguousarray(np.asarray(delta, dtype=np.int32))
    return f(int(m), int(L), int(n_samples), d.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), int(n_y))


class Handle:
    """One solver handle behind the C ABI."""

    def __init__(self, lib, prefix: str, prec_bits: int, device_or_threads=0):
        """device_or_threads: CUDA ordinal (product library) / thread count (oracle); a LIST of CUDA ordinals creates a
        multi-device handle (clrsdp_create_multi: several GPUs, one process, the whole problem behind one handle)."""
        self.lib, self.prefix = lib, prefix
        self.prec = int(prec_bits)
        self.nlimb = self.prec // 32
        self._h = ctypes.c_void_p()
        if isinstance(device_or_threads, (list, tuple)):
            devs = np.ascontiguousarray(np.asarray(device_or_threads, dtype=np.int32))
            f = self._fn("create_multi")
            f.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
            st = f(ctypes.byref(self._h), self.prec, len(devs), devs.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
            if st != 0:
                raise ClrsdpError(st, prefix + "create_multi")
            return
        f = self._fn("create")
        f.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int]
        st = f(ctypes.byref(self._h), self.prec, int(device_or_threads))
        if st != 0:
            raise ClrsdpError(st, prefix + "create")

    def cluster_owner(self, J):
        f = self._fn("cluster_owner")
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        out = np.zeros(J, dtype=np.int32)
        self._check(f(self._h, int(J), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))), "cluster_owner")
        return out

    def _fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = ctypes.c_int
        return f

    def _check(self, st, where):
        if st != 0:
            detail = ""
            try:
                g = getattr(self.lib, self.prefix + "last_error")
                g.restype = ctypes.c_char_p
                g.argtypes = [ctypes.c_void_p]
                detail = (g(self._h) or b"").decode()
            except Exception:
                pass
            raise ClrsdpError(st, self.prefix + where, detail)

    def close(self):
        if self._h:
            f = self._fn("destroy")
            f.argtypes = [ctypes.c_void_p]
            f(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- problem ---------------------------------------------------------------------------------
    def set_structure(self, n_y, m, L, n_samples, delta, ranks):
        def arr(v):
            return np.ascontiguousarray(np.asarray(v, dtype=np.int32).reshape(-1))
        m, L, n_samples, delta, ranks = map(arr, (m, L, n_samples, delta, ranks))
        ip = ctypes.POINTER(ctypes.c_int)
        f = self._fn("set_structure")
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ip, ip, ip, ip, ip]
        self._check(f(self._h, len(m), int(n_y), *[a.ctypes.data_as(ip) for a in (m, L, n_samples, delta, ranks)]),
                    "set_structure")

    def upload_cluster(self, j, V: MpArray, H: MpArray, B: MpArray, c: MpArray):
        f = self._fn("upload_cluster")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, mp, mp, mp, mp]
        s = [a.c_struct() for a in (V, H, B, c)]
        self._check(f(self._h, int(j), *[ctypes.byref(x) for x in s]), "upload_cluster")

    def upload_C(self, C: MpArray | None):
        """Objective matrix C (kwarg of solverank1sdp, MPMP.jl:599): blocks in (j,l) order, each nb x nb row-major,
        concatenated; None restores C = 0."""
        f = self._fn("upload_C")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p, mp]
        if C is None:
            self._check(f(self._h, None), "upload_C")
        else:
            sc = C.c_struct()
            self._check(f(self._h, ctypes.byref(sc)), "upload_C")

    def upload_objective(self, b: MpArray, b0: MpArray):
        f = self._fn("upload_objective")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p, mp, mp]
        sb, s0 = b.c_struct(), b0.c_struct()
        self._check(f(self._h, ctypes.byref(sb), ctypes.byref(s0)), "upload_objective")

    def set_params(self, real_params: MpArray | None, maxiterations=500, need_primal_feasible=False,
                   need_dual_feasible=False, phase_timing=False):
        """phase_timing: fill the reference's 17 per-phase buckets (MPMP.jl:889-898) on every iteration, also when the
        iteration is replayed from a CUDA graph (event-record nodes inside the graph)."""
        f = self._fn("set_params")
        f.argtypes = [ctypes.c_void_p, ctypes.POINTER(clrsdp_mp), ctypes.POINTER(IntParams)]
        ip = IntParams(int(maxiterations), int(bool(need_primal_feasible)), int(bool(need_dual_feasible)), int(bool(phase_timing)))
        if real_params is not None:
            assert real_params.n == P_COUNT
            s = real_params.c_struct()
            st = f(self._h, ctypes.byref(s), ctypes.byref(ip))
        else:
            st = f(self._h, None, ctypes.byref(ip))
        self._check(st, "set_params")

    # ---- point -----------------------------------------------------------------------------------
    def init_point(self):
        f = self._fn("init_point")
        f.argtypes = [ctypes.c_void_p]
        self._check(f(self._h), "init_point")

    def upload_point(self, x, X, y, Y):
        f = self._fn("upload_point")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p, mp, mp, mp, mp]
        s = [a.c_struct() for a in (x, X, y, Y)]
        self._check(f(self._h, *[ctypes.byref(v) for v in s]), "upload_point")

    def download_point(self, n_x, n_X, n_y, out=None):
        """x, X, y, Y of the current iterate; `out` (a list returned by an earlier call) is overwritten in place"""
        if out is None:
            out = [MpArray(n, self.nlimb) for n in (n_x, n_X, n_y, n_X)]
        f = self._fn("download_point")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p, mp, mp, mp, mp]
        s = [a.c_struct() for a in out]
        self._check(f(self._h, *[ctypes.byref(v) for v in s]), "download_point")
        return out  # x, X, y, Y

    # ---- hot path --------------------------------------------------------------------------------
    def prepare(self) -> IterInfo:
        info = IterInfo()
        f = self._fn("prepare")
        f.argtypes = [ctypes.c_void_p, ctypes.POINTER(IterInfo)]
        self._check(f(self._h, ctypes.byref(info)), "prepare")
        return info

    def iterate(self) -> IterInfo:
        info = IterInfo()
        f = self._fn("iterate")
        f.argtypes = [ctypes.c_void_p, ctypes.POINTER(IterInfo)]
        self._check(f(self._h, ctypes.byref(info)), "iterate")
        return info

    def solve(self, max_rows=512):
        rows = (IterInfo * max_rows)()
        n = ctypes.c_int(0)
        f = self._fn("solve")
        f.argtypes = [ctypes.c_void_p, ctypes.POINTER(IterInfo), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        st = f(self._h, rows, max_rows, ctypes.byref(n))
        out = [rows[i] for i in range(min(n.value, max_rows))]
        self._check(st, "solve")
        return out

    def fetch(self, name: str, j: int = 0, l: int = 0) -> MpArray:
        f = getattr(self.lib, self.prefix + "fetch")
        f.restype = ctypes.c_int64
        f.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(clrsdp_mp)]
        n = f(self._h, name.encode(), int(j), int(l), None)
        if n < 0:
            raise ClrsdpError(int(n), f"{self.prefix}fetch({name})")
        out = MpArray(int(n), self.nlimb)
        if n:
            s = out.c_struct()
            n2 = f(self._h, name.encode(), int(j), int(l), ctypes.byref(s))
            if n2 < 0:
                raise ClrsdpError(int(n2), f"{self.prefix}fetch({name})")
        return out

    def scalar(self, name: str):
        return self.fetch("scalar", SCALARS.index(name)).to_mpf(0)

    # ---- phase-level ops ---------------------------------------------------------------------------
    def op_gemm(self, batch, M, N, K, A: MpArray, B: MpArray) -> MpArray:
        C = MpArray(batch * M * N, self.nlimb)
        f = self._fn("op_gemm")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [mp, mp, mp]
        sa, sb, sc = A.c_struct(), B.c_struct(), C.c_struct()
        self._check(f(self._h, batch, M, N, K, ctypes.byref(sa), ctypes.byref(sb), ctypes.byref(sc)), "op_gemm")
        return C.reshape(batch, M, N)

    def op_gemm_planes(self, batch, M, N, K, A: MpArray, B: MpArray, max_planes=80):
        planes = np.zeros((max_planes, batch, M, N), dtype=np.int32)
        npl = ctypes.c_int(max_planes)
        rexp = np.zeros((batch, M), dtype=np.int32)
        cexp = np.zeros((batch, N), dtype=np.int32)
        f = self._fn("op_gemm_planes")
        mp = ctypes.POINTER(clrsdp_mp)
        i32 = ctypes.POINTER(ctypes.c_int32)
        f.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [mp, mp, i32, ctypes.POINTER(ctypes.c_int), i32, i32]
        sa, sb = A.c_struct(), B.c_struct()
        self._check(f(self._h, batch, M, N, K, ctypes.byref(sa), ctypes.byref(sb), planes.ctypes.data_as(i32),
                      ctypes.byref(npl), rexp.ctypes.data_as(i32), cexp.ctypes.data_as(i32)), "op_gemm_planes")
        return planes[:npl.value], rexp, cexp

    def op_cholesky(self, batch, n, A: MpArray):
        L = MpArray(batch * n * n, self.nlimb)
        Li = MpArray(batch * n * n, self.nlimb)
        f = self._fn("op_cholesky")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, mp, mp, mp]
        sa, sl, si = A.c_struct(), L.c_struct(), Li.c_struct()
        self._check(f(self._h, batch, n, ctypes.byref(sa), ctypes.byref(sl), ctypes.byref(si)), "op_cholesky")
        return L.reshape(batch, n, n), Li.reshape(batch, n, n)

    def op_signed_factor(self, batch, n, A: MpArray):
        """(M, signs) with A^-1 = M^T diag(signs) M per matrix of the batch (include/clrsdp.h)."""
        M = MpArray(batch * n * n, self.nlimb)
        signs = np.zeros(batch * n, dtype=np.int32)
        f = self._fn("op_signed_factor")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, mp, mp, ctypes.POINTER(ctypes.c_int32)]
        sa, sm = A.c_struct(), M.c_struct()
        self._check(f(self._h, batch, n, ctypes.byref(sa), ctypes.byref(sm),
                      signs.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))), "op_signed_factor")
        return M.reshape(batch, n, n), signs.reshape(batch, n)

    def op_lambda_min(self, batch, n, A: MpArray) -> MpArray:
        lam = MpArray(batch, self.nlimb)
        f = self._fn("op_lambda_min")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, mp, mp]
        sa, sl = A.c_struct(), lam.c_struct()
        self._check(f(self._h, batch, n, ctypes.byref(sa), ctypes.byref(sl)), "op_lambda_min")
        return lam

    def op_elementwise(self, op: str, a: MpArray, b: MpArray | None) -> MpArray:
        c = MpArray(a.n, self.nlimb)
        f = self._fn("op_elementwise")
        mp = ctypes.POINTER(clrsdp_mp)
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, mp, mp, mp]
        sa, sc = a.c_struct(), c.c_struct()
        sb = b.c_struct() if b is not None else None
        self._check(f(self._h, ord(op), ctypes.byref(sa), ctypes.byref(sb) if sb is not None else None,
                      ctypes.byref(sc)), "op_elementwise")
        return c

    # ---- product-only helpers ----------------------------------------------------------------------
    def pin(self, *arrays):
        """register the numpy buffers of MpArrays (kept alive by the caller) for direct DMA (product library only)"""
        f = self._fn("pin_host")
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        for a in arrays:
            for buf in (a.sign, a.exp, a.limb):
                self._check(f(self._h, buf.ctypes.data, buf.nbytes), "pin_host")

    def measure_int8_peak(self) -> float:
        """int8 MAC/s of the tensor pipe, measured on this device (product library only)"""
        f = self._fn("measure_int8_peak")
        f.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
        v = ctypes.c_double(0.0)
        self._check(f(self._h, ctypes.byref(v)), "measure_int8_peak")
        return float(v.value)

    def launch_count(self) -> int:
        f = getattr(self.lib, self.prefix + "launch_count")
        f.restype = ctypes.c_int64
        f.argtypes = [ctypes.c_void_p]
        return int(f(self._h))

    def profile_reset(self, enable=True):
        f = self._fn("profile_reset")
        f.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self._check(f(self._h, int(enable)), "profile_reset")

    def profile_query(self, pattern: str):
        ms, n, work = ctypes.c_double(0), ctypes.c_int64(0), ctypes.c_double(0)
        f = self._fn("profile_query")
        f.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double),
                      ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double)]
        self._check(f(self._h, pattern.encode(), ctypes.byref(ms), ctypes.byref(n), ctypes.byref(work)),
                    "profile_query")
        return ms.value, n.value, work.value

    def profile_dump(self):
        f = self._fn("profile_dump")
        f.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        n = f(self._h, None, 0)
        buf = ctypes.create_string_buffer(max(n, 1))
        f(self._h, buf, n)
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, launches, work = line.split()
            out[name] = dict(ms=float(ms), launches=int(launches), work=float(work))
        return out

    def comm_init(self, n_ranks: int, rank: int, uid: bytes):
        f = self._fn("comm_init")
        f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
        self._check(f(self._h, n_ranks, rank, uid), "comm_init")

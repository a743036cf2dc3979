"""Wire format of multiprecision numbers crossing the C ABI (include/clrsdp.h, `clrsdp_mp`).

A number at precision p = 32*nlimb bits is  sign * (0.limbs)_2 * 2^exp  (MPFR / BigFloat convention,
the format the reference's ArbMatrix midpoints and `const T = BigFloat` values convert to without
rounding, MPMP.jl:17,617). Arrays are planar: sign[n], exp[n], limb[nlimb][n], limb 0 least significant.
"""
from __future__ import annotations

import ctypes
from fractions import Fraction

import numpy as np


class clrsdp_mp(ctypes.Structure):
    _fields_ = [
        ("sign", ctypes.POINTER(ctypes.c_int8)),
        ("exp", ctypes.POINTER(ctypes.c_int64)),
        ("limb", ctypes.POINTER(ctypes.c_uint32)),
        ("n", ctypes.c_int64),
    ]


class MpArray:
    """n numbers at nlimb*32 bits in wire layout (host, numpy-backed)."""

    __slots__ = ("sign", "exp", "limb", "nlimb", "shape")

    def __init__(self, n_or_shape, nlimb: int):
        shape = (int(n_or_shape),) if np.isscalar(n_or_shape) else tuple(int(s) for s in n_or_shape)
        n = int(np.prod(shape)) if shape else 1
        self.shape = shape
        self.nlimb = int(nlimb)
        self.sign = np.zeros(n, dtype=np.int8)
        self.exp = np.zeros(n, dtype=np.int64)
        self.limb = np.zeros((self.nlimb, n), dtype=np.uint32)

    # ---- basic container behaviour ---------------------------------------------------------------
    @property
    def n(self) -> int:
        return self.sign.shape[0]

    @property
    def prec(self) -> int:
        return 32 * self.nlimb

    def __len__(self):
        return self.n

    def reshape(self, *shape):
        out = self.view()
        shape = shape[0] if len(shape) == 1 and not np.isscalar(shape[0]) else shape
        assert int(np.prod(shape)) == self.n
        out.shape = tuple(shape)
        return out

    def view(self):
        out = MpArray.__new__(MpArray)
        out.sign, out.exp, out.limb, out.nlimb, out.shape = self.sign, self.exp, self.limb, self.nlimb, self.shape
        return out

    def take(self, idx) -> "MpArray":
        idx = np.asarray(idx).reshape(-1)
        out = MpArray(len(idx), self.nlimb)
        out.sign[:] = self.sign[idx]
        out.exp[:] = self.exp[idx]
        out.limb[:, :] = self.limb[:, idx]
        return out

    def widen(self, nlimb: int) -> "MpArray":
        """the same values at a higher precision (zero limbs appended at the low end; exact)."""
        assert nlimb >= self.nlimb
        out = MpArray(self.shape, nlimb)
        out.sign[:] = self.sign
        out.exp[:] = self.exp
        out.limb[nlimb - self.nlimb:, :] = self.limb
        return out

    def transpose2d(self) -> "MpArray":
        r, c = self.shape
        idx = np.arange(r * c).reshape(r, c).T.reshape(-1)
        return self.take(idx).reshape(c, r)

    @staticmethod
    def concat(parts) -> "MpArray":
        parts = list(parts)
        nl = parts[0].nlimb
        out = MpArray(sum(p.n for p in parts), nl)
        o = 0
        for p in parts:
            out.sign[o:o + p.n] = p.sign
            out.exp[o:o + p.n] = p.exp
            out.limb[:, o:o + p.n] = p.limb
            o += p.n
        return out

    def c_struct(self) -> clrsdp_mp:
        """ctypes view (works for both clrsdp_mp and clrsdp_mp_out, which share the layout)."""
        assert self.sign.flags.c_contiguous and self.exp.flags.c_contiguous and self.limb.flags.c_contiguous
        return clrsdp_mp(
            self.sign.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
            self.exp.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
            self.limb.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
            self.n,
        )

    # ---- conversions ------------------------------------------------------------------------------
    def set_int(self, i: int, mant: int, exp2: int):
        """element i = mant * 2^exp2 (mant any python int), rounded to nearest (ties to even) at p bits."""
        p = self.prec
        if mant == 0:
            self.sign[i] = 0
            self.exp[i] = 0
            self.limb[:, i] = 0
            return
        s = -1 if mant < 0 else 1
        a = -mant if mant < 0 else mant
        bl = a.bit_length()
        if bl > p:
            sh = bl - p
            q, rem = a >> sh, a & ((1 << sh) - 1)
            half = 1 << (sh - 1)
            if rem > half or (rem == half and (q & 1)):
                q += 1
                if q.bit_length() > p:
                    q >>= 1
                    bl += 1
            a = q
        else:
            a <<= p - bl
        self.sign[i] = s
        self.exp[i] = exp2 + bl
        for k in range(self.nlimb):
            self.limb[k, i] = (a >> (32 * k)) & 0xFFFFFFFF

    def get_int(self, i: int):
        """(mant, exp2) with value = mant * 2^exp2 exactly."""
        if self.sign[i] == 0:
            return 0, 0
        a = 0
        for k in range(self.nlimb):
            a |= int(self.limb[k, i]) << (32 * k)
        return int(self.sign[i]) * a, int(self.exp[i]) - self.prec

    def to_fraction(self, i: int) -> Fraction:
        m, e = self.get_int(i)
        return Fraction(m) * (Fraction(2) ** e)

    def to_fractions(self):
        return [self.to_fraction(i) for i in range(self.n)]

    def to_mpf(self, i: int):
        import mpmath
        m, e = self.get_int(i)
        return mpmath.mp.make_mpf(mpmath.libmp.from_man_exp(m, e))  # exact, independent of mp.prec

    def to_mpfs(self):
        return [self.to_mpf(i) for i in range(self.n)]

    def to_double(self) -> np.ndarray:
        hi = self.limb[self.nlimb - 1].astype(np.float64) * 2.0 ** -32
        if self.nlimb > 1:
            hi = hi + self.limb[self.nlimb - 2].astype(np.float64) * 2.0 ** -64
        with np.errstate(over="ignore", under="ignore"):
            e = np.clip(self.exp, -2000, 2000).astype(np.int32)
            out = np.ldexp(hi, e) * self.sign
        return out.reshape(self.shape) if self.shape else out

    @staticmethod
    def from_ints(mants, exp2, nlimb: int, shape=None) -> "MpArray":
        mants = list(mants)
        out = MpArray(len(mants), nlimb)
        if np.isscalar(exp2):
            for i, m in enumerate(mants):
                out.set_int(i, int(m), int(exp2))
        else:
            for i, (m, e) in enumerate(zip(mants, exp2)):
                out.set_int(i, int(m), int(e))
        return out.reshape(shape) if shape is not None else out

    @staticmethod
    def from_mpf(values, nlimb: int, shape=None) -> "MpArray":
        """from mpmath mpf (or anything mpmath.mpf() accepts); rounds to nearest at p bits."""
        import mpmath
        values = list(values)
        out = MpArray(len(values), nlimb)
        for i, v in enumerate(values):
            if not hasattr(v, "_mpf_"):  # never re-round an mpf of another context
                v = mpmath.mpf(v)
            sign, man, exp, _bc = v._mpf_
            out.set_int(i, -int(man) if sign else int(man), int(exp))
        return out.reshape(shape) if shape is not None else out

    @staticmethod
    def from_fraction(values, nlimb: int, shape=None) -> "MpArray":
        values = list(values)
        out = MpArray(len(values), nlimb)
        p = 32 * nlimb
        for i, v in enumerate(values):
            v = Fraction(v)
            if v == 0:
                continue
            # scale so that the quotient has p+2 significant bits, then round via set_int (sticky bit)
            num, den = v.numerator, v.denominator
            sh = p + 3 - (abs(num).bit_length() - den.bit_length())
            if sh > 0:
                q, r = divmod(abs(num) << sh, den)
            else:
                q, r = divmod(abs(num), den << (-sh))
            q = (q << 1) | (1 if r else 0)
            out.set_int(i, -q if num < 0 else q, -sh - 1)
        return out.reshape(shape) if shape is not None else out

    @staticmethod
    def from_scaled_int64(vals: np.ndarray, scale_exp: int, nlimb: int) -> "MpArray":
        """value = vals * 2^scale_exp for an int64 array with |vals| <= 2^53 (exact, vectorised)."""
        v = np.asarray(vals, dtype=np.int64)
        shape = v.shape
        v = v.reshape(-1)
        out = MpArray(v.size, nlimb)
        a = np.abs(v).astype(np.uint64)
        assert a.size == 0 or int(a.max()) <= (1 << 53)
        nz = a != 0
        bl = np.zeros(v.size, dtype=np.int64)
        an = a[nz]
        fl = np.frexp(an.astype(np.float64))[1].astype(np.int64)  # exact bit length below 2^53
        bl[nz] = fl
        top = np.zeros(v.size, dtype=np.uint64)
        top[nz] = an << (64 - fl).astype(np.uint64)
        out.sign[:] = np.sign(v).astype(np.int8)
        out.exp[nz] = bl[nz] + scale_exp
        out.limb[nlimb - 1] = (top >> np.uint64(32)).astype(np.uint32)
        if nlimb > 1:
            out.limb[nlimb - 2] = (top & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        return out.reshape(shape)

    @staticmethod
    def from_double(vals, nlimb: int) -> "MpArray":
        v = np.asarray(vals, dtype=np.float64)
        m, e = np.frexp(v.reshape(-1))
        mant = np.round(np.ldexp(m, 53)).astype(np.int64)
        out = MpArray.from_scaled_int64(mant, 0, nlimb)
        nz = mant != 0
        out.exp[nz] = e[nz]
        return out.reshape(v.shape)

    @staticmethod
    def zeros(n_or_shape, nlimb: int) -> "MpArray":
        return MpArray(n_or_shape, nlimb)


def rel_err_bits(a: MpArray, b: MpArray, scale: Fraction | None = None) -> float:
    """max_i |a_i - b_i| / scale, returned as -log2 (i.e. matching bits). scale defaults to max_i |b_i|."""
    import math
    assert a.n == b.n
    fa, fb = a.to_fractions(), b.to_fractions()
    if scale is None:
        scale = max((abs(v) for v in fb), default=Fraction(0))
    if scale == 0:
        scale = Fraction(1)
    worst = max((abs(x - y) for x, y in zip(fa, fb)), default=Fraction(0))
    if worst == 0:
        return float("inf")
    r = worst / scale
    return -(math.log2(r.numerator) - math.log2(r.denominator))

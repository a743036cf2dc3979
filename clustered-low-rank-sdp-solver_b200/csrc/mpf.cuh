// mpf.cuh — fixed-limb multiprecision floating point for device (and host) code.
//
// Replaces the Arb/BigFloat scalar substrate of the reference (MPMP.jl:5-17: every value is an Arb
// midpoint at `precision(BigFloat)` bits) by a register-resident format: NL 32-bit limbs of mantissa
// (normalised, top bit set), a 32-bit exponent and a sign. value = (-1)^neg * 0.m * 2^e  (MPFR
// convention, so the wire format converts without rounding). Every loop over limbs has compile-time
// bounds so the limbs live in registers; variable shifts use log-step limb moves + funnel shifts.
//
// Rounding: add/sub/mul/div/sqrt round to nearest up to a 2^-30 ulp slack (sticky bits below the guard
// limb are dropped), i.e. the result is within 0.5000001 ulp of the exact one — the same error model as
// the reference's `approx_*` routines, not bit-identical to MPFR.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MP_HD __host__ __device__ __forceinline__
#else
#define MP_HD inline
#endif

namespace mp {

constexpr int32_t EXP_ZERO = -(1 << 28);

template <int NL>
struct Num {
  uint32_t m[NL];  // m[0] least significant limb
  int32_t e;       // EXP_ZERO <=> the number is zero
  uint32_t neg;    // 0 / 1
};

// ---------------------------------------------------------------------------------------------------
MP_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
MP_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t r) {  // low 32 bits of (hi:lo) >> r, 0<=r<32
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, r);
#else
  return r ? ((lo >> r) | (hi << (32 - r))) : lo;
#endif
}
MP_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t r) {  // high 32 bits of (hi:lo) << r, 0<=r<32
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, r);
#else
  return r ? ((hi << r) | (lo >> (32 - r))) : hi;
#endif
}

template <int NL>
MP_HD bool is_zero(const Num<NL>& x) { return x.e == EXP_ZERO; }
template <int NL>
MP_HD void set_zero(Num<NL>& x) {
#pragma unroll
  for (int i = 0; i < NL; i++) x.m[i] = 0;
  x.e = EXP_ZERO;
  x.neg = 0;
}
template <int NL>
MP_HD Num<NL> zero() {
  Num<NL> x;
  set_zero(x);
  return x;
}
template <int NL>
MP_HD Num<NL> from_pow2(int k) {  // 2^k
  Num<NL> x;
#pragma unroll
  for (int i = 0; i < NL; i++) x.m[i] = 0;
  x.m[NL - 1] = 0x80000000u;
  x.e = k + 1;
  x.neg = 0;
  return x;
}
template <int NL>
MP_HD Num<NL> one() { return from_pow2<NL>(0); }
template <int NL>
MP_HD Num<NL> neg(Num<NL> x) {
  if (!is_zero(x)) x.neg ^= 1u;
  return x;
}
template <int NL>
MP_HD Num<NL> fabs(Num<NL> x) {
  x.neg = 0;
  return x;
}
template <int NL>
MP_HD Num<NL> mul_2exp(Num<NL> x, int k) {
  if (!is_zero(x)) x.e += k;
  return x;
}

// shift an N-limb array right/left by a whole number of limbs q (0 <= q < 2*N), zero filling
template <int N>
MP_HD void shr_limbs(uint32_t (&x)[N], uint32_t q) {
#pragma unroll
  for (int s = 1; s < N; s <<= 1)
    if (q & s) {
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = (i + s < N) ? x[(i + s < N) ? i + s : 0] : 0u;
    }
  if (q >= (uint32_t)N) {
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = 0;
  }
}
template <int N>
MP_HD void shl_limbs(uint32_t (&x)[N], uint32_t q) {
#pragma unroll
  for (int s = 1; s < N; s <<= 1)
    if (q & s) {
#pragma unroll
      for (int i = N - 1; i >= 0; i--) x[i] = (i - s >= 0) ? x[(i - s >= 0) ? i - s : 0] : 0u;
    }
  if (q >= (uint32_t)N) {
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = 0;
  }
}
template <int N>
MP_HD void shr_bits(uint32_t (&x)[N], uint32_t r) {  // 0 <= r < 32
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = funnel_r(x[i], (i + 1 < N) ? x[(i + 1 < N) ? i + 1 : 0] : 0u, r);
}
template <int N>
MP_HD void shl_bits(uint32_t (&x)[N], uint32_t r) {  // 0 <= r < 32
#pragma unroll
  for (int i = N - 1; i >= 0; i--) x[i] = funnel_l((i > 0) ? x[(i > 0) ? i - 1 : 0] : 0u, x[i], r);
}
// On the device the limb loops are PTX carry chains (add.cc / addc.cc / sub.cc / subc.cc / mad.lo.cc / madc.hi.cc): one
// instruction per limb instead of the compare-and-select sequences the compiler derives from 64-bit C arithmetic.
// Every chain starts with an instruction that sets the carry flag without reading it, and nothing the compiler emits
// between the (volatile) statements touches the flag.
// x += y, returns carry out
template <int N>
MP_HD uint32_t add_n(uint32_t (&x)[N], const uint32_t (&y)[N]) {
#if defined(__CUDA_ARCH__)
  asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(y[0]));
#pragma unroll
  for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
  uint32_t c;
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
  return c;
#else
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t s = (uint64_t)x[i] + y[i] + c;
    x[i] = (uint32_t)s;
    c = (uint32_t)(s >> 32);
  }
  return c;
#endif
}
// x -= y, returns borrow (1 if x < y: the result is then the two's complement of y - x)
template <int N>
MP_HD uint32_t sub_n(uint32_t (&x)[N], const uint32_t (&y)[N]) {
#if defined(__CUDA_ARCH__)
  asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(y[0]));
#pragma unroll
  for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
  uint32_t b;
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(b));  // 0 - 0 - borrow = 0 or 0xffffffff
  return b & 1u;
#else
  uint32_t b = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t s = (uint64_t)x[i] - y[i] - b;
    x[i] = (uint32_t)s;
    b = (uint32_t)(s >> 63);
  }
  return b;
#endif
}
// x = -x (two's complement)
template <int N>
MP_HD void neg_n(uint32_t (&x)[N]) {
#if defined(__CUDA_ARCH__)
  asm volatile("sub.cc.u32 %0, 0, %0;" : "+r"(x[0]));
#pragma unroll
  for (int i = 1; i < N; i++) asm volatile("subc.cc.u32 %0, 0, %0;" : "+r"(x[i]));
#else
  uint32_t c = 1;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t s = (uint64_t)(~x[i]) + c;
    x[i] = (uint32_t)s;
    c = (uint32_t)(s >> 32);
  }
#endif
}
// x += c (c = 0 or 1), returns carry out
template <int N>
MP_HD uint32_t inc_n(uint32_t (&x)[N], uint32_t c) {
#if defined(__CUDA_ARCH__)
  asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(c));
#pragma unroll
  for (int i = 1; i < N; i++) asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(x[i]));
  uint32_t co;
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(co));
  return co;
#else
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t s = (uint64_t)x[i] + c;
    x[i] = (uint32_t)s;
    c = (uint32_t)(s >> 32);
  }
  return c;
#endif
}
// x += y + cin (cin = 0 or 1), returns carry out
template <int N>
MP_HD uint32_t addc_n(uint32_t (&x)[N], const uint32_t (&y)[N], uint32_t cin) {
#if defined(__CUDA_ARCH__)
  uint32_t t;
  asm volatile("add.cc.u32 %0, %1, 0xffffffff;" : "=r"(t) : "r"(cin));  // carry flag = cin
#pragma unroll
  for (int i = 0; i < N; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
  uint32_t c;
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
  return c;
#else
  uint32_t c = cin;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t s = (uint64_t)x[i] + y[i] + c;
    x[i] = (uint32_t)s;
    c = (uint32_t)(s >> 32);
  }
  return c;
#endif
}
// The signed combination X <- | X +- Y | of two aligned magnitudes WITHOUT data-dependent branches (sub = 0: X + Y,
// sub = 1: X - Y). Returns through `flip` whether the result changed sign (|Y| > |X|) and through `ovf` whether the
// sum carried out of the top limb (the result has then been shifted right by one bit with the carry on top).
// On the GPU the lanes of a warp hold unrelated numbers: an if (same sign) add else subtract executes both sides for
// every warp, and the elimination / product kernels that live on these operations are instruction-bound.
template <int N>
MP_HD void combine_n(uint32_t (&X)[N], uint32_t (&Y)[N], uint32_t sub, uint32_t& flip, uint32_t& ovf) {
  const uint32_t m = 0u - sub;  // all ones when subtracting
#pragma unroll
  for (int i = 0; i < N; i++) Y[i] ^= m;
  const uint32_t c = addc_n<N>(X, Y, sub);  // X + Y, or X + ~Y + 1 = X - Y (carry = no borrow)
  flip = sub & (c ^ 1u);
  ovf = (sub ^ 1u) & c;
  const uint32_t nm = 0u - flip;            // negative difference: two's complement
#pragma unroll
  for (int i = 0; i < N; i++) X[i] ^= nm;
  inc_n<N>(X, flip);
  // carry out of a sum: one bit to the right, the carry becomes the top bit (shift count 0 or 1)
#pragma unroll
  for (int i = 0; i < N - 1; i++) X[i] = funnel_r(X[i], X[i + 1], ovf);
  X[N - 1] = (X[N - 1] >> ovf) | (ovf << 31);
}
// compare magnitudes of two N-limb arrays: -1, 0, 1
template <int N>
MP_HD int cmp_n(const uint32_t (&x)[N], const uint32_t (&y)[N]) {
  int r = 0;
#pragma unroll
  for (int i = 0; i < N; i++)  // low to high: the highest differing limb decides last
    r = (x[i] != y[i]) ? ((x[i] > y[i]) ? 1 : -1) : r;
  return r;
}
// normalise an N-limb array (shift left until the top bit is set); returns the shift, or -1 if zero
template <int N>
MP_HD int normalize_n(uint32_t (&x)[N]) {
  uint32_t q = 0;
  if (x[N - 1] == 0) {  // rare (deep cancellation): whole-limb shift
    bool found = false;
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
      if (!found) {
        if (x[i] == 0)
          q++;
        else
          found = true;
      }
    }
    if (!found) return -1;
    shl_limbs<N>(x, q);
  }
  uint32_t r = (uint32_t)clz32(x[N - 1]);
  if (r) shl_bits<N>(x, r);
  return (int)(32 * q + r);
}
// round an (NL+1)-limb normalised mantissa g|m (g = guard limb x[0]) to NL limbs; adjusts e on overflow
template <int NL>
MP_HD void round_guard(Num<NL>& out, const uint32_t (&x)[NL + 1], int32_t e, uint32_t negf) {
#pragma unroll
  for (int i = 0; i < NL; i++) out.m[i] = x[i + 1];
  uint32_t c = inc_n<NL>(out.m, x[0] >> 31);
  if (c) {  // mantissa was all ones: becomes 1.000.. = 0.1 * 2
    out.m[NL - 1] = 0x80000000u;
    e += 1;
  }
  out.e = e;
  out.neg = negf;
}

// ---------------------------------------------------------------------------------------------------
// (A branch-free variant of add - one path through combine_n, as in mul_sub_mul - was measured on the B200: the
// kernels that live on add (small products, traces, reductions) have warps whose lanes mostly agree in sign and whose
// shifts are short, and ran 1.3-1.7x slower with it; the branches stay.)
template <int NL>
MP_HD Num<NL> add(const Num<NL>& a, const Num<NL>& b) {
  if (is_zero(a)) return b;
  if (is_zero(b)) return a;
  bool sw = b.e > a.e;
  uint32_t X[NL + 1], Y[NL + 1];
  X[0] = 0;
  Y[0] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    X[i + 1] = sw ? b.m[i] : a.m[i];
    Y[i + 1] = sw ? a.m[i] : b.m[i];
  }
  int32_t ex = sw ? b.e : a.e, ey = sw ? a.e : b.e;
  uint32_t nx = sw ? b.neg : a.neg, ny = sw ? a.neg : b.neg;
  uint32_t d = (uint32_t)(ex - ey);
  Num<NL> out;
  if (d > 32u * NL + 31u) {  // y is below the guard limb
#pragma unroll
    for (int i = 0; i < NL; i++) out.m[i] = X[i + 1];
    out.e = ex;
    out.neg = nx;
    return out;
  }
  if (d >> 5) shr_limbs<NL + 1>(Y, d >> 5);
  if (d & 31u) shr_bits<NL + 1>(Y, d & 31u);
  if (nx == ny) {
    uint32_t c = add_n<NL + 1>(X, Y);
    if (c) {
      shr_bits<NL + 1>(X, 1);
      X[NL] |= 0x80000000u;
      ex += 1;
    }
    round_guard<NL>(out, X, ex, nx);
    return out;
  }
  // opposite signs: X - Y; a borrow (only possible when the exponents are equal) means |Y| > |X|: negate, flip the sign
  if (sub_n<NL + 1>(X, Y)) {
    neg_n<NL + 1>(X);
    nx = ny;
  }
  int sh = normalize_n<NL + 1>(X);
  if (sh < 0) return zero<NL>();
  round_guard<NL>(out, X, ex - sh, nx);
  return out;
}
template <int NL>
MP_HD Num<NL> sub(const Num<NL>& a, const Num<NL>& b) { return add(a, neg(b)); }

// top NL+2 limbs of the 2NL-limb product of the mantissas (truncated: columns NL-2 .. 2NL-1, error < NL units of
// the lowest kept limb). Not normalised: the value P / 2^(32(NL+2)) lies in [1/4, 1).
template <int NL>
MP_HD void mul_raw(const Num<NL>& a, const Num<NL>& b, uint32_t (&P)[NL + 2]) {
  constexpr int C0 = (NL >= 2) ? NL - 2 : 0;
#if defined(__CUDA_ARCH__)
  // column sums in a three-word accumulator t2:t1:t0; one product = mad.lo.cc + madc.hi.cc + addc
  uint32_t t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
  for (int c = C0; c <= 2 * NL - 2; c++) {
#pragma unroll
    for (int i = 0; i < NL; i++) {
      const int j = c - i;
      if (j >= 0 && j < NL)
        asm volatile(
            "mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
            "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32 %2, %2, 0;"
            : "+r"(t0), "+r"(t1), "+r"(t2)
            : "r"(a.m[i]), "r"(b.m[j]));
    }
    P[c - C0 + (NL >= 2 ? 0 : 1)] = t0;
    t0 = t1;
    t1 = t2;
    t2 = 0;
  }
  P[NL + 1] = t0;
  if (NL < 2) P[0] = 0;
#else
  uint64_t lo = 0;
  uint32_t hi = 0;
#pragma unroll
  for (int c = C0; c <= 2 * NL - 2; c++) {
#pragma unroll
    for (int i = 0; i < NL; i++) {
      int j = c - i;
      if (j >= 0 && j < NL) {
        uint64_t p = (uint64_t)a.m[i] * b.m[j];
        lo += p;
        hi += (lo < p) ? 1u : 0u;
      }
    }
    P[c - C0 + (NL >= 2 ? 0 : 1)] = (uint32_t)lo;
    lo = (lo >> 32) | ((uint64_t)hi << 32);
    hi = 0;
  }
  P[NL + 1] = (uint32_t)lo;
  if (NL < 2) P[0] = 0;
#endif
}
template <int NL>
MP_HD Num<NL> mul(const Num<NL>& a, const Num<NL>& b) {
  if (is_zero(a) || is_zero(b)) return zero<NL>();
  uint32_t P[NL + 2];
  mul_raw<NL>(a, b, P);
  // P[2..NL+1] = top NL limbs, P[1] = guard, P[0] = extra
  uint32_t X[NL + 1];
#pragma unroll
  for (int i = 0; i <= NL; i++) X[i] = P[i + 1];
  int32_t e = a.e + b.e;
  if (!(X[NL] & 0x80000000u)) {
#pragma unroll
    for (int i = NL; i >= 1; i--) X[i] = (X[i] << 1) | (X[i - 1] >> 31);
    X[0] = (X[0] << 1) | (P[0] >> 31);
    e -= 1;
  }
  Num<NL> out;
  round_guard<NL>(out, X, e, a.neg ^ b.neg);
  return out;
}
// a*b - c*d with ONE rounding: both products are kept to NL+1 accurate limbs (error < NL * 2^-32 ulp of the larger
// product), aligned, combined, normalised and rounded.
// The building block of the elimination steps and three-term recurrences on the latency-bound paths.
template <int NL>
MP_HD Num<NL> mul_sub_mul(const Num<NL>& a, const Num<NL>& b, const Num<NL>& c, const Num<NL>& d) {
  // One path for all operands (see combine_n): a product with a zero factor gets a zero mantissa and an exponent below
  // everything, so it is the operand that is shifted out, and the other product is normalised and rounded exactly as
  // mul() would (the rows of the elimination are full of zeros and of either sign: with branches a warp ran the
  // two-product, the one-product, the adding and the subtracting variants one after the other).
  const bool z1 = is_zero(a) || is_zero(b), z2 = is_zero(c) || is_zero(d);
  if (z1 && z2) return zero<NL>();  // (skips work only where a whole warp sits on structural zeros)
  uint32_t P1[NL + 2], P2[NL + 2], X[NL + 2], Y[NL + 2];
  mul_raw<NL>(a, b, P1);
  mul_raw<NL>(c, d, P2);
  constexpr int32_t E_NONE = -(1 << 30);
  const int32_t e1 = z1 ? E_NONE : a.e + b.e, e2 = z2 ? E_NONE : c.e + d.e;
  const uint32_t n1 = (a.neg ^ b.neg) & 1u, n2 = ((c.neg ^ d.neg) ^ 1u) & 1u;
  const bool sw = e2 > e1;  // X takes the product with the larger exponent
  const uint32_t k1 = z1 ? 0u : 0xffffffffu, k2 = z2 ? 0u : 0xffffffffu;
#pragma unroll
  for (int i = 0; i < NL + 2; i++) {
    X[i] = sw ? (P2[i] & k2) : (P1[i] & k1);
    Y[i] = sw ? (P1[i] & k1) : (P2[i] & k2);
  }
  const int32_t ex = sw ? e2 : e1, ey = sw ? e1 : e2;
  uint32_t nx = sw ? n2 : n1;
  const uint32_t ny = sw ? n1 : n2;
  const uint32_t dd = (uint32_t)(ex - ey);
  const uint32_t q = dd >> 5;
  shr_limbs<NL + 2>(Y, q > (uint32_t)(NL + 2) ? (uint32_t)(NL + 2) : q);
  shr_bits<NL + 2>(Y, dd & 31u);
  uint32_t flip, ovf;
  combine_n<NL + 2>(X, Y, nx ^ ny, flip, ovf);
  nx = flip ? ny : nx;
  const int sh = normalize_n<NL + 2>(X);
  if (sh < 0) return zero<NL>();
  uint32_t G[NL + 1];
#pragma unroll
  for (int i = 0; i <= NL; i++) G[i] = X[i + 1];
  Num<NL> out;
  round_guard<NL>(out, G, ex + (int32_t)ovf - sh, nx);
  return out;
}

template <int NL>
MP_HD int cmp(const Num<NL>& a, const Num<NL>& b) {  // -1, 0, 1
  bool za = is_zero(a), zb = is_zero(b);
  if (za && zb) return 0;
  if (za) return b.neg ? 1 : -1;
  if (zb) return a.neg ? -1 : 1;
  if (a.neg != b.neg) return a.neg ? -1 : 1;
  int r;
  if (a.e != b.e)
    r = (a.e > b.e) ? 1 : -1;
  else
    r = cmp_n<NL>(a.m, b.m);
  return a.neg ? -r : r;
}
template <int NL>
MP_HD int cmp_abs(const Num<NL>& a, const Num<NL>& b) { return cmp(fabs(a), fabs(b)); }

// ---- conversions -------------------------------------------------------------------------------------
template <int NL>
MP_HD double mant_to_double(const Num<NL>& x) {  // mantissa in [0.5,1) (0 for zero)
  if (is_zero(x)) return 0.0;
  double v = (double)x.m[NL - 1] * (1.0 / 4294967296.0);
  if (NL > 1) v += (double)x.m[NL - 2] * (1.0 / 18446744073709551616.0);
  return v;
}
template <int NL>
MP_HD double to_double(const Num<NL>& x) {
  if (is_zero(x)) return 0.0;
  int e = x.e;
  e = e > 2000 ? 2000 : (e < -2000 ? -2000 : e);
  double v = ldexp(mant_to_double(x), e);
  return x.neg ? -v : v;
}
template <int NL>
MP_HD Num<NL> from_double(double d) {  // finite d
  Num<NL> x;
  if (d == 0.0) return zero<NL>();
  int e;
  double f = frexp(d < 0 ? -d : d, &e);
  uint64_t mant = (uint64_t)ldexp(f, 53);  // exact, in [2^52, 2^53)
  uint64_t top = mant << 11;
#pragma unroll
  for (int i = 0; i < NL; i++) x.m[i] = 0;
  x.m[NL - 1] = (uint32_t)(top >> 32);
  if (NL > 1) x.m[NL - 2] = (uint32_t)top;
  x.e = e;
  x.neg = d < 0 ? 1u : 0u;
  return x;
}
template <int NL>
MP_HD Num<NL> from_int(int64_t v) {
  if (v == 0) return zero<NL>();
  uint32_t X[NL + 1];
#pragma unroll
  for (int i = 0; i <= NL; i++) X[i] = 0;
  uint64_t a = v < 0 ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
  X[NL] = (uint32_t)(a >> 32);
  X[NL - 1] = (uint32_t)a;
  int sh = normalize_n<NL + 1>(X);
  Num<NL> out;
  round_guard<NL>(out, X, 64 - sh, v < 0 ? 1u : 0u);
  return out;
}
template <int NL>
MP_HD Num<NL + 1> extend(const Num<NL>& x) {
  Num<NL + 1> y;
  y.m[0] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) y.m[i + 1] = x.m[i];
  y.e = x.e;
  y.neg = x.neg;
  return y;
}
template <int NL>
MP_HD Num<NL> narrow(const Num<NL + 1>& x) {
  if (is_zero(x)) return zero<NL>();
  Num<NL> out;
  round_guard<NL>(out, x.m, x.e, x.neg);
  return out;
}

// ---- reciprocal / division / square root ------------------------------------------------------------
// Newton iterations with precision doubling: the seed of a width-NL result is the width-(NL/2+1) result
// (recursively down to a double), so the total cost is ~1.4x the final iteration instead of log2(NL)
// full-width iterations. These sit on the critical path of the Cholesky pivots.
template <int NS, int NL>
MP_HD Num<NS> top_limbs(const Num<NL>& x) {  // truncate to the NS most significant limbs (NS <= NL)
  Num<NS> y;
#pragma unroll
  for (int i = 0; i < NS; i++) y.m[i] = x.m[NL - NS + i];
  y.e = x.e;
  y.neg = x.neg;
  return y;
}
template <int NL, int NS>
MP_HD Num<NL> widen(const Num<NS>& x) {  // zero-extend at the low end (NS <= NL)
  Num<NL> y;
#pragma unroll
  for (int i = 0; i < NL; i++) y.m[i] = (i >= NL - NS) ? x.m[(i >= NL - NS) ? i - (NL - NS) : 0] : 0u;
  y.e = x.e;
  y.neg = x.neg;
  return y;
}
// 1/bm for bm in [0.5,1): relative error ~2^-(32NL-6)
template <int NL>
MP_HD Num<NL> recip_core(const Num<NL>& bm) {
  Num<NL> r;
  if constexpr (NL <= 2) {
    r = from_double<NL>(1.0 / mant_to_double(bm));
    Num<NL> t = sub(one<NL>(), mul(bm, r));
    r = add(r, mul(r, t));
  } else {
    constexpr int NS = NL / 2 + 1;
    r = widen<NL, NS>(recip_core<NS>(top_limbs<NS, NL>(bm)));
    Num<NL> t = sub(one<NL>(), mul(bm, r));
    r = add(r, mul(r, t));
  }
  return r;
}
// t^-1/2 for t in [0.5,2)
template <int NL>
MP_HD Num<NL> rsqrt_core(const Num<NL>& t) {
  Num<NL> y;
  if constexpr (NL <= 2) {
    y = from_double<NL>(1.0 / ::sqrt(ldexp(mant_to_double(t), t.e)));
  } else {
    constexpr int NS = NL / 2 + 1;
    y = widen<NL, NS>(rsqrt_core<NS>(top_limbs<NS, NL>(t)));
  }
  Num<NL> h = sub(one<NL>(), mul(t, mul(y, y)));
  h = mul_2exp(h, -1);
  return add(y, mul(y, h));
}
template <int NL>
MP_HD Num<NL> recip(const Num<NL>& b) {  // b != 0; relative error a few ulp
  Num<NL> bm = b;
  bm.e = 0;
  bm.neg = 0;
  Num<NL> r = recip_core<NL>(bm);
  r.e -= b.e;
  r.neg = b.neg;
  return r;
}
template <int NL>
MP_HD Num<NL> div(const Num<NL>& a, const Num<NL>& b) {  // b != 0
  if (is_zero(a)) return zero<NL>();
  Num<NL + 1> q = mul(extend(a), recip(extend(b)));
  return narrow<NL>(q);
}
// sqrt(a) for a > 0; also returns 1/sqrt(a) (the Cholesky pivot needs both)
template <int NL>
MP_HD Num<NL> sqrt_rsqrt(const Num<NL>& a, Num<NL>& rinv) {
  Num<NL + 1> t = extend(a);
  int par = a.e & 1;
  int half = (a.e - par) >> 1;
  t.e = par;
  t.neg = 0;  // t in [0.5, 2), a = t * 4^half
  Num<NL + 1> y = rsqrt_core<NL + 1>(t);
  Num<NL + 1> s = mul(t, y);
  Num<NL> out = narrow<NL>(s);
  out.e += half;
  rinv = narrow<NL>(y);
  rinv.e -= half;
  return out;
}
template <int NL>
MP_HD Num<NL> sqrt(const Num<NL>& a) {  // a >= 0
  if (is_zero(a)) return zero<NL>();
  Num<NL> r;
  return sqrt_rsqrt(a, r);
}

// ---- planar (structure-of-arrays) storage in HBM ----------------------------------------------------
// A tensor of n numbers is (NL+1) planes of n 32-bit words: planes 0..NL-1 are the limbs (0 = least
// significant), plane NL is the header word (e << 1) | neg. A warp touching 32 consecutive numbers
// reads one fully coalesced 128-byte line per plane.
struct Tensor {
  uint32_t* w = nullptr;
  size_t n = 0;  // plane stride in words
};
MP_HD uint32_t pack_hdr(int32_t e, uint32_t negf) { return ((uint32_t)e << 1) | (negf & 1u); }
template <int NL>
MP_HD Num<NL> load(const uint32_t* __restrict__ w, size_t n, size_t i) {
  Num<NL> x;
#pragma unroll
  for (int k = 0; k < NL; k++) x.m[k] = w[(size_t)k * n + i];
  uint32_t h = w[(size_t)NL * n + i];
  x.e = ((int32_t)h) >> 1;
  x.neg = h & 1u;
  return x;
}
template <int NL>
MP_HD void store(uint32_t* __restrict__ w, size_t n, size_t i, const Num<NL>& x) {
#pragma unroll
  for (int k = 0; k < NL; k++) w[(size_t)k * n + i] = x.m[k];
  w[(size_t)NL * n + i] = pack_hdr(x.e, x.neg);
}
template <int NL>
MP_HD Num<NL> load(const Tensor& t, size_t i) { return load<NL>(t.w, t.n, i); }
template <int NL>
MP_HD void store(const Tensor& t, size_t i, const Num<NL>& x) { store<NL>(t.w, t.n, i, x); }

}  // namespace mp

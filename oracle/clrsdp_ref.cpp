// =====================================================================================================
// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or called by the product path.
//
// CPU restatement of the interior-point hot path of nanleij/Clustered-Low-Rank-SDP-solver
// (`solverank1sdp` and everything it calls, MPMP.jl:595-1898), phase by phase, at the same working
// precision, multithreaded over the same loops the reference parallelises with Threads.@threads.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path, and neither
// Julia nor Arb/FLINT exist in this image, so the reference itself cannot be run. Arb midpoint arithmetic
// (`approx_*` + `get_mid!`) is restated as round-to-nearest MPFR arithmetic at the same precision
// (libmpfr.so.6 4.2.1 through hand-declared prototypes). The restatement is pinned instead by
// tests/test_oracle_*.py: manufactured SDPs with known optimum, the sphere-packing bracket 0.793/0.813
// (examples/SpherePacking.jl:125-126) and an independent mpmath restatement on tiny instances.
//
// Deliberate deviations from the reference, all documented in DESIGN.md:
//   * step length (MPMP.jl:1857-1870): the reference takes all eigenvalues of the (symmetric) matrix
//     L^-1 dM L^-T with Arb's complex nonsymmetric QR; the oracle symmetrises, tridiagonalises
//     (Householder) and bisects with Sturm counts for the smallest one. Same number, cheaper (so the CPU
//     baseline timed from this oracle is an UNDER-estimate of the reference's cost).
//   * the BigFloat fallback chain (:774-798, :1874-1879) is replaced by status codes.
// Every function cites the reference lines it follows.
// =====================================================================================================
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../include/clrsdp.h"
#include "mpfr_decl.h"

namespace {

thread_local long g_prec = 256;

struct Real {
  __mpfr_struct v;
  Real() {
    mpfr_init2(&v, g_prec);
    mpfr_set_zero(&v, 1);
  }
  Real(const Real& o) {
    mpfr_init2(&v, g_prec);
    mpfr_set4(&v, &o.v, MPFR_RNDN, o.v._mpfr_sign);
  }
  Real(Real&& o) noexcept {
    v = o.v;
    o.v._mpfr_d = nullptr;
  }
  explicit Real(double d) {
    mpfr_init2(&v, g_prec);
    mpfr_set_d(&v, d, MPFR_RNDN);
  }
  Real& operator=(const Real& o) {
    if (this != &o) mpfr_set4(&v, &o.v, MPFR_RNDN, o.v._mpfr_sign);
    return *this;
  }
  Real& operator=(Real&& o) noexcept {
    if (this != &o) mpfr_swap(&v, &o.v);
    return *this;
  }
  ~Real() {
    if (v._mpfr_d) mpfr_clear(&v);
  }
  double d() const { return mpfr_get_d(&v, MPFR_RNDN); }
};

inline void r_add(Real& c, const Real& a, const Real& b) { mpfr_add(&c.v, &a.v, &b.v, MPFR_RNDN); }
inline void r_sub(Real& c, const Real& a, const Real& b) { mpfr_sub(&c.v, &a.v, &b.v, MPFR_RNDN); }
inline void r_mul(Real& c, const Real& a, const Real& b) { mpfr_mul(&c.v, &a.v, &b.v, MPFR_RNDN); }
inline void r_div(Real& c, const Real& a, const Real& b) { mpfr_div(&c.v, &a.v, &b.v, MPFR_RNDN); }
inline void r_fma(Real& acc, const Real& a, const Real& b) { mpfr_fma(&acc.v, &a.v, &b.v, &acc.v, MPFR_RNDN); }
// acc -= a*b
inline void r_fnma(Real& acc, const Real& a, const Real& b, Real& tmp) {
  mpfr_mul(&tmp.v, &a.v, &b.v, MPFR_RNDN);
  mpfr_sub(&acc.v, &acc.v, &tmp.v, MPFR_RNDN);
}
inline void r_neg(Real& c, const Real& a) { mpfr_neg(&c.v, &a.v, MPFR_RNDN); }
inline void r_abs(Real& c, const Real& a) { mpfr_set4(&c.v, &a.v, MPFR_RNDN, 1); }
inline int r_cmp(const Real& a, const Real& b) { return mpfr_cmp3(&a.v, &b.v, 1); }
inline void r_zero(Real& a) { mpfr_set_zero(&a.v, 1); }
inline bool r_is_zero(const Real& a) { return mpfr_zero_p(&a.v) != 0; }
inline void r_set_si(Real& a, long s) { mpfr_set_si(&a.v, s, MPFR_RNDN); }
inline void r_half(Real& a) { mpfr_div_2si(&a.v, &a.v, 1, MPFR_RNDN); }

struct Mat {
  int r = 0, c = 0;
  std::vector<Real> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_) {}
  Real& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  const Real& operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------------------
// dense kernels (the libarb entry points of SURVEY §2.3, restated)
// ---------------------------------------------------------------------------------------------------
// approx_mul!, two ways (selected per process by CLRSDP_REF_GEMM, read in clrsdp_ref_create):
//  * default: classical product, one rounding per fused multiply-add (what the parity tests check against);
//  * "fixed": block fixed point, the way libarb's approx_mul works: every row of A and every column of B is
//    aligned to its largest exponent and converted to integers of p/64 + 1 limbs, the dot products are exact integer
//    multiply-accumulates (mpn_mul_n / mpn_add_n, positive and negative terms apart) and each entry is rounded once.
//    It is the faster CPU product; bench.py times the CPU baseline with it and reports the fma-loop time beside it.
static int g_gemm_fixed = 0;
// ALGORITHM SWITCH (the default, "lu", is the reference's algorithm and what "the oracle" means everywhere):
// CLRSDP_REF_FACTOR (read per handle at create) replaces the reference's pivoted LU of S_j and Q (MPMP.jl:1436,:1501) by
// the factorisation of the GPU path so that its numerical behaviour can be studied on the CPU beside the LU, in MPFR
// arithmetic: "ldl" = equilibrated signed Cholesky S' = U^T Sigma U (pivots keep their sign), explicit inverse factors,
// W = L^-1 B, Q = W^T Sigma W (what csrc/ does since round 2); "chol" = the round-1 algorithm (pivots below
// 2^-(p-CLRSDP_REF_CLAMP) raised to that floor), kept because it reproduces round 1's loss of the iterate on sphere packing
// d = 8 at 256 bits. CLRSDP_REF_REFINE=<n> adds n steps of iterative refinement of the Schur solve (measured: no help).
static int g_clamp_bits = 16, g_refine = 0;
static int g_guard_sites = 0xff;        // CLRSDP_REF_GUARD_SITES: which products the reduced window applies to (1 = W, 2 = Q, 4 = rest)
static thread_local int g_site = 4;
static int g_fixed_guard = 64;  // CLRSDP_REF_FIXED_GUARD: guard bits of the fixed-point window below a row's largest entry (GPU: 16)

struct FixedRows {  // rows[r][k]: Lw limbs, sign, and the row exponent: value = +-limbs * 2^(E - 64 Lw)
  int Lw = 0, K = 0;
  std::vector<mp_limb_t> limbs;  // [rows][K][Lw]
  std::vector<signed char> sign; // [rows][K]
  std::vector<long> E;           // [rows], LONG_MIN for an all-zero row
};
// get(r, k) -> const Real&
template <class Get>
static void to_fixed(FixedRows& F, int rows, int K, long prec, Get get) {
  const int L = (int)((prec + 63) / 64), Lw = L + 1;
  F.Lw = Lw, F.K = K;
  F.limbs.assign((size_t)rows * K * Lw, 0);
  F.sign.assign((size_t)rows * K, 0);
  F.E.assign(rows, LONG_MIN);
  std::vector<mp_limb_t> tmp(Lw);
  for (int r = 0; r < rows; r++) {
    long E = LONG_MIN;
    for (int k = 0; k < K; k++) {
      const Real& x = get(r, k);
      if (!mpfr_zero_p(&x.v)) E = std::max(E, (long)x.v._mpfr_exp);
    }
    F.E[r] = E;
    if (E == LONG_MIN) continue;
    for (int k = 0; k < K; k++) {
      const Real& x = get(r, k);
      if (mpfr_zero_p(&x.v)) continue;
      const long sh = E - (long)x.v._mpfr_exp;
      if (sh >= 64L * Lw) continue;
      mp_limb_t* dst = &F.limbs[((size_t)r * K + k) * Lw];
      tmp[0] = 0;
      for (int i = 0; i < L; i++) tmp[i + 1] = x.v._mpfr_d[i];
      const int ws = (int)(sh / 64), bs = (int)(sh % 64);
      for (int i = 0; i + ws < Lw; i++) dst[i] = tmp[i + ws];
      if (bs) __gmpn_rshift(dst, dst, Lw, (unsigned)bs);
      if (g_fixed_guard < 64 && (g_guard_sites & g_site)) dst[0] &= ~(((mp_limb_t)1 << (64 - g_fixed_guard)) - 1);  // experiment: window of p + guard bits
      F.sign[(size_t)r * K + k] = x.v._mpfr_sign < 0 ? -1 : 1;
    }
  }
}
static void fixed_dot(Real& out, const FixedRows& A, int ra, const FixedRows& B, int rb, std::vector<mp_limb_t>& ws) {
  const int Lw = A.Lw, K = A.K, N = 2 * Lw + 1;
  if (A.E[ra] == LONG_MIN || B.E[rb] == LONG_MIN) {
    r_zero(out);
    return;
  }
  ws.assign((size_t)3 * N, 0);
  mp_limb_t *pos = ws.data(), *neg = pos + N, *prod = neg + N;
  const mp_limb_t* a = &A.limbs[(size_t)ra * K * Lw];
  const mp_limb_t* b = &B.limbs[(size_t)rb * K * Lw];
  for (int k = 0; k < K; k++) {
    const int sg = A.sign[(size_t)ra * K + k] * B.sign[(size_t)rb * K + k];
    if (!sg) continue;
    __gmpn_mul_n(prod, a + (size_t)k * Lw, b + (size_t)k * Lw, Lw);
    mp_limb_t* acc = sg > 0 ? pos : neg;
    acc[2 * Lw] += __gmpn_add_n(acc, acc, prod, 2 * Lw);
  }
  int sign = __gmpn_cmp(pos, neg, N);
  if (sign == 0) {
    r_zero(out);
    return;
  }
  if (sign > 0) __gmpn_sub_n(pos, pos, neg, N); else __gmpn_sub_n(pos, neg, pos, N);
  int top = N - 1;
  while (top >= 0 && pos[top] == 0) top--;
  const int lz = __builtin_clzl(pos[top]);
  const int n = top + 1;
  if (lz) __gmpn_lshift(pos, pos, n, (unsigned)lz);
  __mpfr_struct t;
  t._mpfr_prec = 64L * n;
  t._mpfr_sign = sign > 0 ? 1 : -1;
  t._mpfr_exp = A.E[ra] + B.E[rb] - 128L * Lw + 64L * n - lz;
  t._mpfr_d = pos;
  mpfr_set4(&out.v, &t, MPFR_RNDN, t._mpfr_sign);
}
static void gemm_fixed(Mat& out, const Mat& A, int c0, int nc, const Mat& B) {
  const long prec = A.r && A.c ? A(0, 0).v._mpfr_prec : 64;
  FixedRows FA, FB;
  to_fixed(FA, A.r, nc, prec, [&](int r, int k) -> const Real& { return A(r, c0 + k); });
  to_fixed(FB, B.c, nc, prec, [&](int r, int k) -> const Real& { return B(k, r); });
  std::vector<mp_limb_t> ws;
  for (int i = 0; i < A.r; i++)
    for (int j = 0; j < B.c; j++) fixed_dot(out(i, j), FA, i, FB, j, ws);
}
void gemm(Mat& C, const Mat& A, const Mat& B) {
  Mat out(A.r, B.c);
  if (g_gemm_fixed && A.c > 0) {
    gemm_fixed(out, A, 0, A.c, B);
  } else {
    for (int i = 0; i < A.r; i++)
      for (int j = 0; j < B.c; j++) {
        Real& acc = out(i, j);
        for (int k = 0; k < A.c; k++) r_fma(acc, A(i, k), B(k, j));
      }
  }
  C = std::move(out);
}
// C = A[:, c0:c0+nc] * B      (sub-column view of A, used by the pairings, MPMP.jl:1291-1296)
void gemm_cols(Mat& C, const Mat& A, int c0, int nc, const Mat& B) {
  Mat out(A.r, B.c);
  if (g_gemm_fixed && nc > 0) {
    gemm_fixed(out, A, c0, nc, B);
  } else {
    for (int i = 0; i < A.r; i++)
      for (int j = 0; j < B.c; j++) {
        Real& acc = out(i, j);
        for (int k = 0; k < nc; k++) r_fma(acc, A(i, c0 + k), B(k, j));
      }
  }
  C = std::move(out);
}
Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int i = 0; i < A.r; i++)
    for (int j = 0; j < A.c; j++) T(j, i) = A(i, j);
  return T;
}
// cho!: lower Cholesky factor; false when a pivot is not positive.
bool cholesky(Mat& L, const Mat& A) {
  int n = A.r;
  L = Mat(n, n);
  Real s, t;
  for (int j = 0; j < n; j++) {
    s = A(j, j);
    for (int k = 0; k < j; k++) r_fnma(s, L(j, k), L(j, k), t);
    if (mpfr_sgn(&s.v) <= 0) return false;
    mpfr_sqrt(&L(j, j).v, &s.v, MPFR_RNDN);
    for (int i = j + 1; i < n; i++) {
      s = A(i, j);
      for (int k = 0; k < j; k++) r_fnma(s, L(i, k), L(j, k), t);
      r_div(L(i, j), s, L(j, j));
    }
  }
  return true;
}
// approx_solve_tril!(X, L, B, unit): X = L^-1 B with L the lower triangle of Lm.
void solve_tril(Mat& X, const Mat& Lm, const Mat& B, bool unit) {
  int n = Lm.r, nc = B.c;
  Mat out(n, nc);
  Real s, t;
  if (g_gemm_fixed && n > 48 && nc > 0) {
    // blocked forward substitution (libarb's approx_solve_tril is recursive over block products too): the update of a
    // row block by the rows already solved is a block fixed-point product, the 32 x 32 diagonal blocks are solved
    // by substitution
    const int nb = 32;
    const long prec = B(0, 0).v._mpfr_prec;
    std::vector<mp_limb_t> ws;
    for (int r0 = 0; r0 < n; r0 += nb) {
      const int r1 = std::min(n, r0 + nb);
      FixedRows FA, FB;
      if (r0 > 0) {
        to_fixed(FA, r1 - r0, r0, prec, [&](int r, int k) -> const Real& { return Lm(r0 + r, k); });
        to_fixed(FB, nc, r0, prec, [&](int c, int k) -> const Real& { return out(k, c); });
      }
      for (int c = 0; c < nc; c++)
        for (int i = r0; i < r1; i++) {
          if (r0 > 0) {
            fixed_dot(t, FA, i - r0, FB, c, ws);
            r_sub(s, B(i, c), t);
          } else {
            s = B(i, c);
          }
          for (int k = r0; k < i; k++) r_fnma(s, Lm(i, k), out(k, c), t);
          if (unit)
            out(i, c) = s;
          else
            r_div(out(i, c), s, Lm(i, i));
        }
    }
    X = std::move(out);
    return;
  }
  for (int c = 0; c < nc; c++)
    for (int i = 0; i < n; i++) {
      s = B(i, c);
      for (int k = 0; k < i; k++) r_fnma(s, Lm(i, k), out(k, c), t);
      if (unit)
        out(i, c) = s;
      else
        r_div(out(i, c), s, Lm(i, i));
    }
  X = std::move(out);
}
// approx_solve_triu!(X, U, B, unit): X = U^-1 B with U the upper triangle of Um.
void solve_triu(Mat& X, const Mat& Um, const Mat& B, bool unit) {
  int n = Um.r, nc = B.c;
  Mat out(n, nc);
  Real s, t;
  for (int c = 0; c < nc; c++)
    for (int i = n - 1; i >= 0; i--) {
      s = B(i, c);
      for (int k = i + 1; k < n; k++) r_fnma(s, Um(i, k), out(k, c), t);
      if (unit)
        out(i, c) = s;
      else
        r_div(out(i, c), s, Um(i, i));
    }
  X = std::move(out);
}
// approx_lu!: in-place LU with partial pivoting (largest magnitude), perm 0-based; false if singular.
bool lu(std::vector<int>& perm, Mat& A) {
  int n = A.r;
  perm.resize(n);
  for (int i = 0; i < n; i++) perm[i] = i;
  Real t;
  for (int k = 0; k < n; k++) {
    int piv = -1;
    for (int i = k; i < n; i++) {
      if (r_is_zero(A(i, k))) continue;
      if (piv < 0 || mpfr_cmpabs(&A(i, k).v, &A(piv, k).v) > 0) piv = i;
    }
    if (piv < 0) return false;
    if (piv != k) {
      std::swap(perm[k], perm[piv]);
      for (int j = 0; j < n; j++) mpfr_swap(&A(k, j).v, &A(piv, j).v);
    }
    for (int i = k + 1; i < n; i++) {
      r_div(A(i, k), A(i, k), A(k, k));
      if (r_is_zero(A(i, k))) continue;
      for (int j = k + 1; j < n; j++) r_fnma(A(i, j), A(i, k), A(k, j), t);
    }
  }
  return true;
}
// spd_inv!: inverse of an SPD matrix through its Cholesky factor (X^-1 = L^-T L^-1).
bool spd_inverse(Mat& Inv, const Mat& A) {
  Mat L;
  if (!cholesky(L, A)) return false;
  int n = A.r;
  Mat I(n, n);
  for (int i = 0; i < n; i++) r_set_si(I(i, i), 1);
  Mat Li;
  solve_tril(Li, L, I, false);
  Mat out(n, n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      Real& acc = out(i, j);
      for (int k = i; k < n; k++) r_fma(acc, Li(k, i), Li(k, j));
      if (i != j) out(j, i) = acc;
    }
  Inv = std::move(out);
  return true;
}
// EXPERIMENT (CLRSDP_REF_FACTOR=chol): what the GPU path's chol_inverse computes. A' = D^-1 A D^-1 with
// D = diag(2^ceil(e_ii/2)); Cholesky of A' with pivots below 2^-(p-clamp) raised to that floor; Linv = L'^-1.
static int chol_inverse_equil(Mat& Linv, std::vector<long>& sc, std::vector<int>& sg, const Mat& A, long prec, int mode) {
  const int n = A.r;
  sc.assign(n, 0);
  for (int i = 0; i < n; i++)
    if (!mpfr_zero_p(&A(i, i).v)) {
      long e = (long)A(i, i).v._mpfr_exp;
      sc[i] = (e >= 0) ? (e + 1) / 2 : -((-e) / 2);
    }
  Mat Ap(n, n), L(n, n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) mpfr_mul_2si(&Ap(i, j).v, &A(i, j).v, -(sc[i] + sc[j]), MPFR_RNDN);
  Real thr, s, t;
  mpfr_set_ui_2exp(&thr.v, 1, -(prec - g_clamp_bits), MPFR_RNDN);
  int clamped = 0;
  sg.assign(n, 1);
  for (int j = 0; j < n; j++) {
    s = Ap(j, j);
    for (int k = 0; k < j; k++) {
      r_mul(t, L(j, k), L(j, k));
      if (sg[k] > 0) r_sub(s, s, t); else r_add(s, s, t);
    }
    if (mode == 2) {  // signed (LDL^T-like): keep the sign of the pivot, floor its magnitude
      if (mpfr_sgn(&s.v) < 0) {
        sg[j] = -1;
        r_neg(s, s);
        clamped++;
      }
      if (r_cmp(s, thr) < 0) s = thr;
    } else if (r_cmp(s, thr) < 0) {
      s = thr;
      clamped++;
    }
    mpfr_sqrt(&L(j, j).v, &s.v, MPFR_RNDN);
    for (int i = j + 1; i < n; i++) {
      s = Ap(i, j);
      for (int k = 0; k < j; k++) {
        r_mul(t, L(i, k), L(j, k));
        if (sg[k] > 0) r_sub(s, s, t); else r_add(s, s, t);
      }
      r_div(L(i, j), s, L(j, j));
      if (sg[j] < 0) r_neg(L(i, j), L(i, j));
    }
  }
  Mat I(n, n);
  for (int i = 0; i < n; i++) r_set_si(I(i, i), 1);
  solve_tril(Linv, L, I, false);
  return clamped;
}
// Smallest eigenvalue of a symmetric matrix: Householder tridiagonalisation + Sturm bisection.
// Stands in for approx_eig_qr! + min over real parts (MPMP.jl:1857-1870).
bool lambda_min_sym(Real& lam, Mat A) {
  int n = A.r;
  std::vector<Real> dg(n), e(n > 1 ? n - 1 : 0);
  Real t, s, alpha, beta, K;
  for (int k = 0; k + 2 < n; k++) {
    // x = A[k+1:, k]
    int len = n - k - 1;
    r_zero(s);
    for (int i = k + 1; i < n; i++) r_fma(s, A(i, k), A(i, k));
    Real tail;  // norm^2 without the first entry
    mpfr_fms(&tail.v, &A(k + 1, k).v, &A(k + 1, k).v, &s.v, MPFR_RNDN);
    r_neg(tail, tail);
    if (mpfr_sgn(&tail.v) <= 0) continue;  // already tridiagonal in this column
    mpfr_sqrt(&alpha.v, &s.v, MPFR_RNDN);
    if (mpfr_sgn(&A(k + 1, k).v) > 0) r_neg(alpha, alpha);
    std::vector<Real> v(len), p(len), q(len);
    for (int i = 0; i < len; i++) v[i] = A(k + 1 + i, k);
    r_sub(v[0], v[0], alpha);
    r_zero(beta);
    for (int i = 0; i < len; i++) r_fma(beta, v[i], v[i]);
    // H = I - 2 v v^T / beta ; p = 2 A v / beta ; K = v^T p / beta ; q = p - K v ; A -= v q^T + q v^T
    for (int i = 0; i < len; i++) {
      r_zero(p[i]);
      for (int j = 0; j < len; j++) r_fma(p[i], A(k + 1 + i, k + 1 + j), v[j]);
      mpfr_mul_2si(&p[i].v, &p[i].v, 1, MPFR_RNDN);
      r_div(p[i], p[i], beta);
    }
    r_zero(K);
    for (int i = 0; i < len; i++) r_fma(K, v[i], p[i]);
    r_div(K, K, beta);
    for (int i = 0; i < len; i++) {
      r_mul(t, K, v[i]);
      r_sub(q[i], p[i], t);
    }
    for (int i = 0; i < len; i++)
      for (int j = 0; j < len; j++) {
        r_fnma(A(k + 1 + i, k + 1 + j), v[i], q[j], t);
        r_fnma(A(k + 1 + i, k + 1 + j), q[i], v[j], t);
      }
    A(k + 1, k) = alpha;
    A(k, k + 1) = alpha;
    for (int i = k + 2; i < n; i++) {
      r_zero(A(i, k));
      r_zero(A(k, i));
    }
  }
  for (int i = 0; i < n; i++) dg[i] = A(i, i);
  for (int i = 0; i + 1 < n; i++) r_mul(e[i], A(i + 1, i), A(i + 1, i));  // squared off-diagonals
  // Gershgorin bounds
  Real lo, hi, rad, a1, a2;
  bool first = true;
  for (int i = 0; i < n; i++) {
    r_zero(rad);
    if (i > 0) {
      r_abs(t, A(i, i - 1));
      r_add(rad, rad, t);
    }
    if (i + 1 < n) {
      r_abs(t, A(i + 1, i));
      r_add(rad, rad, t);
    }
    r_sub(a1, dg[i], rad);
    r_add(a2, dg[i], rad);
    if (first || r_cmp(a1, lo) < 0) lo = a1;
    if (first || r_cmp(a2, hi) > 0) hi = a2;
    first = false;
  }
  if (n == 1) {
    lam = dg[0];
    return true;
  }
  // widen slightly so the count at lo is 0
  r_abs(t, lo);
  r_abs(s, hi);
  if (r_cmp(s, t) > 0) t = s;
  if (r_is_zero(t)) {
    r_zero(lam);
    return true;
  }
  mpfr_div_2si(&s.v, &t.v, 20, MPFR_RNDN);
  r_sub(lo, lo, s);
  r_add(hi, hi, s);
  auto count_below = [&](const Real& x) {  // number of eigenvalues < x
    int cnt = 0;
    Real qv, tt;
    r_sub(qv, dg[0], x);
    if (r_is_zero(qv)) mpfr_set_ui_2exp(&qv.v, 1, -16 * g_prec, MPFR_RNDN);
    if (mpfr_sgn(&qv.v) < 0) cnt++;
    for (int i = 1; i < n; i++) {
      r_div(tt, e[i - 1], qv);
      r_sub(qv, dg[i], x);
      r_sub(qv, qv, tt);
      if (r_is_zero(qv)) mpfr_set_ui_2exp(&qv.v, 1, -16 * g_prec, MPFR_RNDN);
      if (mpfr_sgn(&qv.v) < 0) cnt++;
    }
    return cnt;
  };
  Real mid, width, tol;
  for (int it = 0; it < 4 * (int)g_prec; it++) {
    r_add(mid, lo, hi);
    r_half(mid);
    if (r_cmp(mid, lo) <= 0 || r_cmp(mid, hi) >= 0) break;
    if (count_below(mid) >= 1)
      hi = mid;
    else
      lo = mid;
    r_sub(width, hi, lo);
    r_abs(t, lo);
    r_abs(s, hi);
    if (r_cmp(s, t) > 0) t = s;
    mpfr_div_2si(&tol.v, &t.v, g_prec + 2, MPFR_RNDN);
    if (r_cmp(width, tol) <= 0) break;
  }
  r_add(lam, lo, hi);
  r_half(lam);
  return true;
}

// ---------------------------------------------------------------------------------------------------
// problem / state
// ---------------------------------------------------------------------------------------------------
struct Block {  // one PSD block (j,l)
  int delta = 0, nb = 0, Nv = 0;
  std::vector<int> ranks, rank_sums;  // [K], [K+1]   (BlockInfo.ranks / rank_sums, MPMP.jl:476-477,488)
  Mat V, VT;                          // delta x Nv, Nv x delta  (hcat of A[l,k][rnk], MPMP.jl:1249-1260)
  std::vector<Real> H;                // [Nv]  A_sign[l,k][rnk]
};
struct Cluster {  // one constraint j
  int m = 0, L = 0, K = 0, dimS = 0;
  std::vector<Block> blk;
  Mat B;  // dimS x n_y
  std::vector<Real> c;
};
struct Decomp {  // (S, perms, LinvB, BTUinv, perm, Q) of MPMP.jl:1507
  std::vector<Mat> LU;
  std::vector<std::vector<int>> perms;
  std::vector<Mat> LinvB, BTUinv;
  std::vector<int> permQ;
  Mat QLU;
  // CLRSDP_REF_FACTOR=chol (experiment): L'_j^-1 of the equilibrated S_j, its scaling exponents, W_j = L_j^-1 B_j, Lq^-1
  std::vector<Mat> Linv, W;
  std::vector<std::vector<long>> sc;
  std::vector<std::vector<int>> sg;
  std::vector<int> sgQ;
  Mat LinvQ;
};
typedef std::vector<std::vector<Mat>> BlockDiag;  // [j][l]

}  // namespace

struct clrsdp_solver {
  int factor_mode = 0;  // experiment switch CLRSDP_REF_FACTOR, read at create: 0 = lu (the reference), 1 = chol, 2 = ldl
  long prec = 256;
  int nthreads = 1;
  int nlimb = 8;
  std::string err;
  // structure (BlockInfo, MPMP.jl:467-479)
  int J = 0, n_y = 0;
  std::vector<Cluster> cl;
  std::vector<int> x_idx;  // x_indices (0-based offsets, length J+1)
  int sumS = 0, ntot = 0;  // sum dim_S, size(X,1)
  std::vector<std::pair<int, int>> jl;
  // objective, params
  std::vector<Real> b;
  Real b0;
  Real beta_inf, beta_feas, gamma, omega_p, omega_d, gap_thr, perr_thr, derr_thr;
  clrsdp_int_params ip{500, 0, 0, 0};
  // state
  bool have_point = false, prepared = false;
  std::vector<Real> x, y;
  BlockDiag X, Y, R, Xinv, P, Z, dX, dY, XYsave;
  BlockDiag C;  // objective matrix (MPMP.jl:599); have_C = false is the reference's AbsoluteZero (:691-695)
  bool have_C = false;
  std::vector<Real> p, d, dx, dy;
  // intermediates kept for parity fetches
  std::vector<Mat> S_keep;
  Mat Q_keep;
  std::vector<std::vector<Mat>> Px_keep, Py_keep;
  Decomp dec;
  std::vector<std::vector<std::vector<std::vector<std::vector<Real>>>>> A_Y;  // [j][l][r][s][idx]
  BlockDiag dX_pred, dY_pred;
  std::vector<Real> dx_pred, dy_pred;
  // driver scalars
  int iter = 1;
  bool pd_feas = false;
  Real mu, p_obj, d_obj, dual_gap, primal_error, dual_error, alpha_p, alpha_d, beta_c, mu_p, mu_c;
  Real lam_x, lam_y;
  double t_schur = 0, t_cholS = 0, t_CinvB = 0, t_Q = 0, t_cholQ = 0;
  double t_dir[5] = {0, 0, 0, 0, 0};

  // -------------------------------------------------------------------------------------------------
  template <class F>
  void parallel_for(int n, F fn) {
    int nt = std::min(nthreads, n);
    if (nt <= 1) {
      g_prec = prec;
      for (int i = 0; i < n; i++) fn(i);
      return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
      th.emplace_back([&, t]() {
        g_prec = prec;
        for (int i = t; i < n; i += nt) fn(i);
      });
    for (auto& t : th) t.join();
  }
  BlockDiag like_blocks() const {
    BlockDiag M(J);
    for (int j = 0; j < J; j++)
      for (int l = 0; l < cl[j].L; l++) M[j].emplace_back(cl[j].blk[l].nb, cl[j].blk[l].nb);
    return M;
  }

  // dot(A,B) for nested BlockDiagonals (MPMP.jl:205-220)
  Real dot_blocks(const BlockDiag& A, const BlockDiag& B) {
    Real res;
    for (int j = 0; j < J; j++)
      for (int l = 0; l < cl[j].L; l++) {
        Real part;
        const Mat &a = A[j][l], &bb = B[j][l];
        for (size_t i = 0; i < a.a.size(); i++) r_fma(part, a.a[i], bb.a[i]);
        r_add(res, res, part);
      }
    return res;
  }
  // dot_c (MPMP.jl:1081-1092)
  Real dot_c(const std::vector<Real>& xv) {
    Real res;
    int xi = 0;
    for (int j = 0; j < J; j++)
      for (int i = 0; i < cl[j].dimS; i++) r_fma(res, cl[j].c[i], xv[xi++]);
    return res;
  }
  // compute_primal_objective / compute_dual_objective (MPMP.jl:1027-1034), C = AbsoluteZero
  Real primal_objective() {
    Real r = dot_c(x);
    r_add(r, r, b0);
    return r;
  }
  Real dot_CY() {  // dot(C, Y) (block_diag_dot, MPMP.jl:205-220); AbsoluteZero -> 0 (:590)
    Real r;
    if (!have_C) return r;
    for (size_t q = 0; q < jl.size(); q++) {
      int j = jl[q].first, l = jl[q].second;
      for (size_t i = 0; i < Y[j][l].a.size(); i++) r_fma(r, C[j][l].a[i], Y[j][l].a[i]);
    }
    return r;
  }
  Real dual_objective() {  // <C,Y> + <b,y> + b0 (:1032-1034)
    Real r = dot_CY();
    for (int i = 0; i < n_y; i++) r_fma(r, b[i], y[i]);
    r_add(r, r, b0);
    return r;
  }
  // compute_duality_gap(p,d) (MPMP.jl:1075-1078)
  Real duality_gap(const Real& po, const Real& dobj) {
    Real num, den, one, g;
    r_sub(num, po, dobj);
    r_abs(num, num);
    r_add(den, po, dobj);
    r_abs(den, den);
    r_set_si(one, 1);
    if (r_cmp(one, den) > 0) den = one;
    r_div(g, num, den);
    return g;
  }
  // compute_duality_gap(constraints,x,y,Y,C,b) (MPMP.jl:1067-1074): objectives WITHOUT b0
  Real duality_gap_point() {
    Real po = dot_c(x), dobj = dot_CY();
    for (int i = 0; i < n_y; i++) r_fma(dobj, b[i], y[i]);
    return duality_gap(po, dobj);
  }
  // compute_error (MPMP.jl:1037-1055)
  Real max_abs(const std::vector<Real>& v) {
    Real mx, t;
    for (auto& e : v) {
      r_abs(t, e);
      if (r_cmp(t, mx) > 0) mx = t;
    }
    return mx;
  }
  Real max_abs(const BlockDiag& M) {
    Real mx;
    for (auto& bj : M)
      for (auto& bl : bj) {
        Real t = max_abs(bl.a);
        if (r_cmp(t, mx) > 0) mx = t;
      }
    return mx;
  }

  // compute_weighted_A! (MPMP.jl:1621-1678): M = sum_i a_i A_i
  void weighted_A(BlockDiag& M, const std::vector<Real>& a) {
    parallel_for((int)jl.size(), [&](int q) {
      int j = jl[q].first, l = jl[q].second;
      const Cluster& c = cl[j];
      const Block& bk = c.blk[l];
      int dl = bk.delta, K = c.K;
      Mat& out = M[j][l];
      for (auto& e : out.a) r_zero(e);
      Mat vs_scaled(dl, bk.Nv), Qp;
      Real w;
      for (int r = 0; r < c.m; r++)
        for (int s = 0; s <= r; s++) {
          int v_off = (s + r * (r + 1) / 2) * K + x_idx[j];
          for (int k = 0; k < K; k++)
            for (int rk = 0; rk < bk.ranks[k]; rk++) {
              int col = bk.rank_sums[k] + rk;
              r_mul(w, a[v_off + k], bk.H[col]);  // ref(a,k+v_offset)*H  (:1654)
              for (int i = 0; i < dl; i++) r_mul(vs_scaled(i, col), w, bk.V(i, col));
            }
          g_site = 128; gemm(Qp, vs_scaled, bk.VT); g_site = 4;  // VD * V^T (:1659)
          if (r != s)
            for (auto& e : Qp.a) r_half(e);  // factor 1/2 of E_rs (:1661-1663)
          for (int i = 0; i < dl; i++)
            for (int i2 = 0; i2 < dl; i2++) out(s * dl + i, r * dl + i2) = Qp(i, i2);  // (:1664-1667)
        }
      if (c.m != 1)  // Symmetric(): mirror the upper triangle (:1671-1674)
        for (int i = 0; i < out.r; i++)
          for (int i2 = 0; i2 < i; i2++) out(i, i2) = out(i2, i);
    });
  }

  // trace_A(constraints, Z::BlockDiagonal, blockinfo) (MPMP.jl:1517-1584)
  std::vector<Real> trace_A_general(const BlockDiag& Zm) {
    std::vector<Real> result(sumS);
    parallel_for(J, [&](int j) {
      const Cluster& c = cl[j];
      Real res, t;
      for (int l = 0; l < c.L; l++) {
        const Block& bk = c.blk[l];
        int dl = bk.delta;
        for (int r = 0; r < c.m; r++)
          for (int s = 0; s <= r; s++) {
            // VZ = V^T Z[r,s];  (VZ o V^T) row sums
            std::vector<Real> part(bk.Nv);
            Real acc;
            for (int v = 0; v < bk.Nv; v++) {
              Real& rs = part[v];
              for (int i2 = 0; i2 < dl; i2++) {
                r_zero(acc);
                for (int i = 0; i < dl; i++) r_fma(acc, bk.VT(v, i), Zm[j][l](r * dl + i, s * dl + i2));
                r_mul(t, acc, bk.VT(v, i2));   // mul_entrywise! (:1559)
                r_add(rs, rs, t);              // * ones (:1560)
              }
            }
            int off = x_idx[j] + (s + r * (r + 1) / 2) * c.K;
            int idx = 0;
            for (int k = 0; k < c.K; k++) {
              r_zero(res);
              for (int rk = 0; rk < bk.ranks[k]; rk++) r_fma(res, bk.H[idx], part[idx]), idx++;
              r_add(result[off + k], res, result[off + k]);
            }
          }
      }
    });
    return result;
  }
  // trace_A(constraints, A_Y, blockinfo) (MPMP.jl:1585-1618)
  std::vector<Real> trace_A_from_AY() {
    std::vector<Real> result(sumS);
    parallel_for(J, [&](int j) {
      const Cluster& c = cl[j];
      Real res;
      for (int k = 0; k < c.K; k++)
        for (int l = 0; l < c.L; l++) {
          const Block& bk = c.blk[l];
          for (int rk = 0; rk < bk.ranks[k]; rk++)
            for (int r = 0; r < c.m; r++)
              for (int s = 0; s <= r; s++) {
                int tup = k + c.K * (s + r * (r + 1) / 2) + x_idx[j];
                int idx = rk + bk.rank_sums[k];
                r_mul(res, bk.H[idx], A_Y[j][l][r][s][idx]);
                r_add(result[tup], res, result[tup]);
              }
        }
    });
    return result;
  }

  // compute_residuals (MPMP.jl:1107-1144) with calculate_res_d (:1095-1102)
  void compute_residuals(bool use_AY) {
    weighted_A(P, x);
    parallel_for((int)jl.size(), [&](int q) {
      int j = jl[q].first, l = jl[q].second;
      for (size_t i = 0; i < P[j][l].a.size(); i++) r_sub(P[j][l].a[i], P[j][l].a[i], X[j][l].a[i]);
      if (have_C)  // P -= C (:1116-1118)
        for (size_t i = 0; i < P[j][l].a.size(); i++) r_sub(P[j][l].a[i], P[j][l].a[i], C[j][l].a[i]);
    });
    // d = c - B y - Tr(A_* Y)
    d.assign(sumS, Real());
    parallel_for(J, [&](int j) {
      const Cluster& c = cl[j];
      Real By;
      for (int i = 0; i < c.dimS; i++) {
        r_zero(By);
        for (int k = 0; k < n_y; k++) r_fma(By, c.B(i, k), y[k]);
        r_sub(d[x_idx[j] + i], c.c[i], By);
      }
    });
    std::vector<Real> tr = use_AY ? trace_A_from_AY() : trace_A_general(Y);
    for (int i = 0; i < sumS; i++) r_sub(d[i], d[i], tr[i]);
    // p = b - sum_j B_j^T x_j
    std::vector<std::vector<Real>> padd(J);
    parallel_for(J, [&](int j) {
      const Cluster& c = cl[j];
      padd[j].assign(n_y, Real());
      for (int k = 0; k < n_y; k++)
        for (int i = 0; i < c.dimS; i++) r_fma(padd[j][k], c.B(i, k), x[x_idx[j] + i]);
    });
    p.assign(n_y, Real());
    for (int j = 0; j < J; j++)
      for (int k = 0; k < n_y; k++) r_sub(p[k], p[k], padd[j][k]);
    for (int k = 0; k < n_y; k++) r_add(p[k], p[k], b[k]);
  }

  // compute_residual_R! (MPMP.jl:1189-1215)
  void residual_R(const Real& muv, bool second_order) {
    parallel_for((int)jl.size(), [&](int q) {
      int j = jl[q].first, l = jl[q].second;
      Mat tmp;
      Mat& Rb = R[j][l];
      for (auto& e : Rb.a) r_zero(e);
      for (int i = 0; i < Rb.r; i++) Rb(i, i) = muv;
      gemm(tmp, X[j][l], Y[j][l]);
      XYsave[j][l] = tmp;
      for (size_t i = 0; i < Rb.a.size(); i++) r_sub(Rb.a[i], Rb.a[i], tmp.a[i]);
      if (second_order) {
        gemm(tmp, dX[j][l], dY[j][l]);
        for (size_t i = 0; i < Rb.a.size(); i++) r_sub(Rb.a[i], Rb.a[i], tmp.a[i]);
      }
    });
  }

  // compute_S_integrated (MPMP.jl:1218-1414)
  void compute_S(std::vector<Mat>& S) {
    S.clear();
    for (int j = 0; j < J; j++) S.emplace_back(cl[j].dimS, cl[j].dimS);
    A_Y.assign(J, {});
    Px_keep.assign(J, {});
    Py_keep.assign(J, {});
    for (int j = 0; j < J; j++) {
      A_Y[j].resize(cl[j].L);
      Px_keep[j].resize(cl[j].L);
      Py_keep[j].resize(cl[j].L);
    }
    // pairings for all (j,l) (the reference threads over column chunks inside a serial j,l loop; the
    // result is the same matrix, so the oracle threads over (j,l) instead)
    parallel_for((int)jl.size(), [&](int q) {
      int j = jl[q].first, l = jl[q].second;
      const Cluster& c = cl[j];
      const Block& bk = c.blk[l];
      int dl = bk.delta, bs = bk.Nv, m = c.m;
      Mat Px(m * bs, m * bs), Py(m * bs, m * bs), partX, partY, Xp, Yp;
      for (int s = 0; s < m; s++) {
        g_site = 16; gemm_cols(partX, Xinv[j][l], s * dl, dl, bk.V); g_site = 4;  // X_inv[:, s-block] * V (:1291-1293)
        g_site = 16; gemm_cols(partY, Y[j][l], s * dl, dl, bk.V); g_site = 4;     // (:1294-1296)
        for (int r = 0; r < m; r++) {
          Mat subX(dl, bs), subY(dl, bs);
          for (int i = 0; i < dl; i++)
            for (int v = 0; v < bs; v++) {
              subX(i, v) = partX(r * dl + i, v);
              subY(i, v) = partY(r * dl + i, v);
            }
          g_site = 16; gemm(Xp, bk.VT, subX); g_site = 4;  // V^T (X^-1 V) (:1300-1306)
          g_site = 16; gemm(Yp, bk.VT, subY); g_site = 4;  // (:1308-1315)
          for (int a = 0; a < bs; a++)
            for (int v = 0; v < bs; v++) {
              Px(r * bs + a, s * bs + v) = Xp(a, v);
              Py(r * bs + a, s * bs + v) = Yp(a, v);
            }
        }
      }
      // A_Y (:1320-1330)
      A_Y[j][l].assign(m, std::vector<std::vector<Real>>(m));
      for (int r = 0; r < m; r++)
        for (int s = 0; s < m; s++) {
          A_Y[j][l][r][s].resize(bs);
          for (int k = 0; k < bs; k++) A_Y[j][l][r][s][k] = Py(r * bs + k, s * bs + k);
        }
      Px_keep[j][l] = std::move(Px);
      Py_keep[j][l] = std::move(Py);
    });
    // S accumulation (:1335-1406), threaded over (j,k1): different k1 write different columns
    std::vector<std::pair<int, int>> jk;
    for (int j = 0; j < J; j++)
      for (int k1 = 0; k1 < cl[j].K; k1++) jk.emplace_back(j, k1);
    parallel_for((int)jk.size(), [&](int q) {
      int j = jk[q].first, k1 = jk[q].second;
      const Cluster& c = cl[j];
      int K = c.K, m = c.m;
      Real tot;
      for (int l = 0; l < c.L; l++) {
        const Block& bk = c.blk[l];
        const Mat &Px = Px_keep[j][l], &Py = Py_keep[j][l];
        int bs = bk.Nv;
        for (int r1 = 0; r1 < m; r1++)
          for (int s1 = 0; s1 <= r1; s1++) {
            int hor = k1 + (s1 + r1 * (r1 + 1) / 2) * K;
            for (int r2 = 0; r2 < m; r2++)
              for (int s2 = 0; s2 <= r2; s2++)
                for (int k2 = 0; k2 < K; k2++) {
                  int ver = k2 + (s2 + r2 * (r2 + 1) / 2) * K;
                  if (ver > hor) continue;  // upper triangular part (:1349)
                  for (int rk1 = 0; rk1 < bk.ranks[k1]; rk1++)
                    for (int rk2 = 0; rk2 < bk.ranks[k2]; rk2++) {
                      int r1s = rk1 + bk.rank_sums[k1] + bs * r1;
                      int r2s = rk2 + bk.rank_sums[k2] + bs * r2;
                      int s1s = rk1 + bk.rank_sums[k1] + bs * s1;
                      int s2s = rk2 + bk.rank_sums[k2] + bs * s2;
                      r_mul(tot, Px(s1s, r2s), Py(s2s, r1s));  // (:1373-1377)
                      r_fma(tot, Px(r1s, r2s), Py(s2s, s1s));  // (:1378-1382)
                      r_fma(tot, Px(s1s, s2s), Py(r2s, r1s));  // (:1383-1387)
                      r_fma(tot, Px(r1s, s2s), Py(r2s, s1s));  // (:1388-1392)
                      r_mul(tot, bk.H[bk.rank_sums[k1] + rk1], tot);
                      r_mul(tot, bk.H[bk.rank_sums[k2] + rk2], tot);
                      mpfr_div_ui(&tot.v, &tot.v, 4, MPFR_RNDN);
                      r_add(S[j](ver, hor), tot, S[j](ver, hor));
                    }
                }
          }
      }
    });
    parallel_for(J, [&](int j) {  // Symmetric(S[j]) (:1409)
      for (int i = 0; i < S[j].r; i++)
        for (int i2 = 0; i2 < i; i2++) S[j](i, i2) = S[j](i2, i);
    });
  }

  // compute_T_decomposition (MPMP.jl:1417-1514)
  int T_decomposition() {
    double t0 = now_s();
    std::vector<Mat> S;
    compute_S(S);
    S_keep = S;
    t_schur = now_s() - t0;
    t0 = now_s();
    if (factor_mode) return T_decomposition_chol();
    dec.LU = std::move(S);
    dec.perms.assign(J, {});
    std::vector<int> ok(J, 1);
    parallel_for(J, [&](int j) { ok[j] = lu(dec.perms[j], dec.LU[j]) ? 1 : 0; });  // (:1435-1441)
    for (int j = 0; j < J; j++)
      if (!ok[j]) return CLRSDP_ERR_SINGULAR_S;
    t_cholS = now_s() - t0;
    t0 = now_s();
    dec.LinvB.assign(J, Mat());
    dec.BTUinv.assign(J, Mat());
    parallel_for(J, [&](int j) {  // (:1454-1464)
      Mat Ct = transpose(dec.LU[j]), tmp;
      solve_tril(tmp, Ct, cl[j].B, false);  // U^-T B
      dec.BTUinv[j] = transpose(tmp);
      Mat PB(cl[j].dimS, n_y);
      for (int i = 0; i < cl[j].dimS; i++)
        for (int k = 0; k < n_y; k++) PB(i, k) = cl[j].B(dec.perms[j][i], k);
      solve_tril(dec.LinvB[j], dec.LU[j], PB, true);  // L^-1 P B
    });
    t_CinvB = now_s() - t0;
    t0 = now_s();
    // Q = sum over row chunks of hcat(BTUinv) * vcat(LinvB) (:1467-1495)
    int nchunk = std::max(1, nthreads);
    std::vector<int> rowj, rowi;
    for (int j = 0; j < J; j++)
      for (int i = 0; i < cl[j].dimS; i++) rowj.push_back(j), rowi.push_back(i);
    int tot = (int)rowj.size();
    std::vector<Mat> Qp(nchunk);
    parallel_for(nchunk, [&](int t) {
      int min_size = tot / nchunk, n_min = (min_size + 1) * nchunk - tot;  // (:1472-1478)
      int lo = 0;
      for (int q = 0; q < t; q++) lo += (q < nchunk - n_min) ? min_size + 1 : min_size;
      int hi = lo + ((t < nchunk - n_min) ? min_size + 1 : min_size);
      Mat out(n_y, n_y);
      if (g_gemm_fixed && hi > lo) {
        FixedRows FA, FB;
        const long prec = out(0, 0).v._mpfr_prec;
        to_fixed(FA, n_y, hi - lo, prec, [&](int a, int k) -> const Real& { return dec.BTUinv[rowj[lo + k]](a, rowi[lo + k]); });
        to_fixed(FB, n_y, hi - lo, prec, [&](int bq, int k) -> const Real& { return dec.LinvB[rowj[lo + k]](rowi[lo + k], bq); });
        std::vector<mp_limb_t> ws;
        for (int a = 0; a < n_y; a++)
          for (int bq = 0; bq < n_y; bq++) fixed_dot(out(a, bq), FA, a, FB, bq, ws);
      } else {
        for (int a = 0; a < n_y; a++)
          for (int bq = 0; bq < n_y; bq++) {
            Real& acc = out(a, bq);
            for (int rr = lo; rr < hi; rr++) r_fma(acc, dec.BTUinv[rowj[rr]](a, rowi[rr]), dec.LinvB[rowj[rr]](rowi[rr], bq));
          }
      }
      Qp[t] = std::move(out);
    });
    Mat Q(n_y, n_y);
    for (int t = 0; t < nchunk; t++)
      for (size_t i = 0; i < Q.a.size(); i++) r_add(Q.a[i], Q.a[i], Qp[t].a[i]);
    Q_keep = Q;
    t_Q = now_s() - t0;
    t0 = now_s();
    dec.QLU = std::move(Q);
    if (!lu(dec.permQ, dec.QLU)) return CLRSDP_ERR_SINGULAR_Q;  // (:1499-1505)
    t_cholQ = now_s() - t0;
    return 0;
  }

  // EXPERIMENT (CLRSDP_REF_FACTOR=chol): the GPU path's decomposition, see the switch at the top of the file
  int n_clamped_S = 0, n_clamped_Q = 0;
  int T_decomposition_chol() {
    dec.Linv.assign(J, Mat());
    dec.W.assign(J, Mat());
    dec.sc.assign(J, {});
    dec.sg.assign(J, {});
    std::vector<int> cl_(J, 0);
    parallel_for(J, [&](int j) {
      cl_[j] = chol_inverse_equil(dec.Linv[j], dec.sc[j], dec.sg[j], S_keep[j], prec, factor_mode);
      Mat DB(cl[j].dimS, n_y);
      for (int i = 0; i < cl[j].dimS; i++)
        for (int k = 0; k < n_y; k++) mpfr_mul_2si(&DB(i, k).v, &cl[j].B(i, k).v, -dec.sc[j][i], MPFR_RNDN);
      g_site = 1;
      gemm(dec.W[j], dec.Linv[j], DB);
      g_site = 4;
    });
    n_clamped_S = 0;
    for (int j = 0; j < J; j++) n_clamped_S += cl_[j];
    Mat Q(n_y, n_y);
    for (int j = 0; j < J; j++) {
      Mat Wt = transpose(dec.W[j]), Qj;
      for (int i = 0; i < cl[j].dimS; i++)
        if (dec.sg[j][i] < 0)
          for (int k = 0; k < n_y; k++) r_neg(Wt(k, i), Wt(k, i));
      g_site = 2;
      gemm(Qj, Wt, dec.W[j]);
      g_site = 4;
      for (size_t i = 0; i < Q.a.size(); i++) r_add(Q.a[i], Q.a[i], Qj.a[i]);
    }
    Q_keep = Q;
    std::vector<long> scq;
    n_clamped_Q = chol_inverse_equil(dec.LinvQ, scq, dec.sgQ, Q, prec, factor_mode);
    for (int i = 0; i < n_y; i++)
      for (int k = 0; k < n_y; k++) mpfr_mul_2si(&dec.LinvQ(i, k).v, &dec.LinvQ(i, k).v, -scq[k], MPFR_RNDN);
    if (getenv("CLRSDP_REF_VERBOSE")) fprintf(stderr, "[chol] iter %d clamped S %d Q %d\n", iter, n_clamped_S, n_clamped_Q);
    return 0;
  }
  // one solve of the Schur system with the factors above: S dx - B dy = rx, B^T dx = ry
  void schur_solve_chol(std::vector<Real>& sdx, std::vector<Real>& sdy, const std::vector<Real>& rx, const std::vector<Real>& ry) {
    std::vector<Mat> tv(J);
    std::vector<Mat> ty(J);
    parallel_for(J, [&](int j) {
      int n = cl[j].dimS;
      Mat rhs(n, 1);
      for (int i = 0; i < n; i++) mpfr_mul_2si(&rhs(i, 0).v, &rx[x_idx[j] + i].v, -dec.sc[j][i], MPFR_RNDN);
      g_site = 8; gemm(tv[j], dec.Linv[j], rhs); g_site = 4;
      for (int i = 0; i < n; i++)
        if (dec.sg[j][i] < 0) r_neg(tv[j](i, 0), tv[j](i, 0));  // tv = Sigma L^-1 D^-1 rx
      Mat Wt = transpose(dec.W[j]);
      g_site = 8; gemm(ty[j], Wt, tv[j]); g_site = 4;
    });
    Mat dyr(n_y, 1), z, dyv;
    for (int k = 0; k < n_y; k++) {
      Real sum;
      for (int j = 0; j < J; j++) r_add(sum, sum, ty[j](k, 0));
      r_sub(dyr(k, 0), ry[k], sum);
    }
    g_site = 8; gemm(z, dec.LinvQ, dyr); g_site = 4;
    for (int k = 0; k < n_y; k++)
      if (dec.sgQ[k] < 0) r_neg(z(k, 0), z(k, 0));
    Mat LqT = transpose(dec.LinvQ);
    g_site = 8; gemm(dyv, LqT, z); g_site = 4;
    sdy.assign(n_y, Real());
    for (int k = 0; k < n_y; k++) sdy[k] = dyv(k, 0);
    sdx.assign(sumS, Real());
    parallel_for(J, [&](int j) {
      Mat u, sol;
      g_site = 8; gemm(u, dec.W[j], dyv); g_site = 4;
      for (int i = 0; i < cl[j].dimS; i++) {
        if (dec.sg[j][i] < 0) r_neg(u(i, 0), u(i, 0));  // Sigma (W dy) + Sigma t
        r_add(u(i, 0), u(i, 0), tv[j](i, 0));
      }
      Mat LiT = transpose(dec.Linv[j]);
      g_site = 8; gemm(sol, LiT, u); g_site = 4;
      for (int i = 0; i < cl[j].dimS; i++) mpfr_mul_2si(&sdx[x_idx[j] + i].v, &sol(i, 0).v, -dec.sc[j][i], MPFR_RNDN);
    });
  }
  void search_system_chol(const std::vector<Real>& rhs_x) {
    schur_solve_chol(dx, dy, rhs_x, p);
    for (int it = 0; it < g_refine; it++) {
      // r1 = rhs_x - (S dx - B dy), r2 = p - B^T dx
      std::vector<Real> r1(sumS), r2(n_y), cx, cy;
      Real t;
      for (int j = 0; j < J; j++) {
        int n = cl[j].dimS;
        for (int i = 0; i < n; i++) {
          Real acc = rhs_x[x_idx[j] + i];
          for (int k = 0; k < n; k++) r_fnma(acc, S_keep[j](i, k), dx[x_idx[j] + k], t);
          for (int k = 0; k < n_y; k++) r_fma(acc, cl[j].B(i, k), dy[k]);
          r1[x_idx[j] + i] = acc;
        }
      }
      for (int k = 0; k < n_y; k++) {
        Real acc = p[k];
        for (int j = 0; j < J; j++)
          for (int i = 0; i < cl[j].dimS; i++) r_fnma(acc, cl[j].B(i, k), dx[x_idx[j] + i], t);
        r2[k] = acc;
      }
      if (getenv("CLRSDP_REF_VERBOSE")) {
        Real m1 = max_abs(r1), m2 = max_abs(r2), mx = max_abs(dx), my = max_abs(dy);
        fprintf(stderr, "[chol] iter %d refine %d: |r1| %.2e |r2| %.2e |dx| %.2e |dy| %.2e\n", iter, it, m1.d(), m2.d(), mx.d(), my.d());
      }
      schur_solve_chol(cx, cy, r1, r2);
      for (int i = 0; i < sumS; i++) r_add(dx[i], dx[i], cx[i]);
      for (int k = 0; k < n_y; k++) r_add(dy[k], dy[k], cy[k]);
    }
  }

  // compute_search_direction (MPMP.jl:1682-1824)
  void search_direction() {
    double t0 = now_s();
    parallel_for((int)jl.size(), [&](int q) {  // Z = sym(X^-1 (P Y - R)) (:1700-1729)
      int j = jl[q].first, l = jl[q].second;
      Mat T;
      g_site = 32; gemm(T, P[j][l], Y[j][l]); g_site = 4;
      for (size_t i = 0; i < T.a.size(); i++) r_sub(T.a[i], T.a[i], R[j][l].a[i]);
      g_site = 32; gemm(T, Xinv[j][l], T); g_site = 4;
      Mat& Zb = Z[j][l];
      for (int i = 0; i < T.r; i++)
        for (int i2 = 0; i2 < T.c; i2++) {
          r_add(Zb(i, i2), T(i, i2), T(i2, i));
          r_half(Zb(i, i2));
        }
    });
    t_dir[0] += now_s() - t0;
    t0 = now_s();
    std::vector<Real> rhs_x(sumS), tr = trace_A_general(Z);  // rhs_x = -d - Tr(A_* Z) (:1735-1738)
    for (int i = 0; i < sumS; i++) {
      r_neg(rhs_x[i], d[i]);
      r_sub(rhs_x[i], rhs_x[i], tr[i]);
    }
    t_dir[1] += now_s() - t0;
    t0 = now_s();
    if (factor_mode) {
      search_system_chol(rhs_x);
      goto system_done;
    }
    {
    std::vector<Mat> temp_x(J), temp_y(J);
    parallel_for(J, [&](int j) {  // (:1751-1759)
      int n = cl[j].dimS;
      Mat rhs(n, 1);
      for (int i = 0; i < n; i++) rhs(i, 0) = rhs_x[x_idx[j] + dec.perms[j][i]];
      solve_tril(temp_x[j], dec.LU[j], rhs, true);
      gemm(temp_y[j], dec.BTUinv[j], temp_x[j]);
    });
    Mat dyv(n_y, 1);
    {
      std::vector<Real> sum_y(n_y);
      for (int j = 0; j < J; j++)
        for (int k = 0; k < n_y; k++) r_add(sum_y[k], sum_y[k], temp_y[j](k, 0));
      for (int k = 0; k < n_y; k++) r_sub(dyv(k, 0), p[k], sum_y[k]);  // dy = rhs_y - sum (:1761)
    }
    {  // approx_solve_lu_precomp! (:1764)
      Mat pb(n_y, 1), t1;
      for (int k = 0; k < n_y; k++) pb(k, 0) = dyv(dec.permQ[k], 0);
      solve_tril(t1, dec.QLU, pb, true);
      solve_triu(dyv, dec.QLU, t1, false);
    }
    dy.assign(n_y, Real());
    for (int k = 0; k < n_y; k++) dy[k] = dyv(k, 0);
    dx.assign(sumS, Real());
    parallel_for(J, [&](int j) {  // U dx = t + LinvB dy (:1771-1773)
      Mat t2, sol;
      gemm(t2, dec.LinvB[j], dyv);
      for (int i = 0; i < cl[j].dimS; i++) r_add(t2(i, 0), temp_x[j](i, 0), t2(i, 0));
      solve_triu(sol, dec.LU[j], t2, false);
      for (int i = 0; i < cl[j].dimS; i++) dx[x_idx[j] + i] = sol(i, 0);
    });
    }
  system_done:
    t_dir[2] += now_s() - t0;
    t0 = now_s();
    weighted_A(dX, dx);  // dX = sum dx_i A_i + P (:1780-1786)
    parallel_for((int)jl.size(), [&](int q) {
      int j = jl[q].first, l = jl[q].second;
      for (size_t i = 0; i < dX[j][l].a.size(); i++) r_add(dX[j][l].a[i], dX[j][l].a[i], P[j][l].a[i]);
    });
    t_dir[3] += now_s() - t0;
    t0 = now_s();
    parallel_for((int)jl.size(), [&](int q) {  // dY = sym(X^-1 (R - dX Y)) (:1791-1820)
      int j = jl[q].first, l = jl[q].second;
      Mat T;
      g_site = 64; gemm(T, dX[j][l], Y[j][l]); g_site = 4;
      for (size_t i = 0; i < T.a.size(); i++) r_sub(T.a[i], R[j][l].a[i], T.a[i]);
      g_site = 64; gemm(T, Xinv[j][l], T); g_site = 4;
      Mat& D = dY[j][l];
      for (int i = 0; i < T.r; i++)
        for (int i2 = 0; i2 < T.c; i2++) {
          r_add(D(i, i2), T(i, i2), T(i2, i));
          r_half(D(i, i2));
        }
    });
    t_dir[4] += now_s() - t0;
  }

  // compute_step_length (MPMP.jl:1829-1898)
  int step_length(Real& alpha, Real& lam_out, const BlockDiag& M, const BlockDiag& dM, bool isX) {
    std::vector<Real> mins(jl.size());
    std::vector<int> st(jl.size(), 0);
    parallel_for((int)jl.size(), [&](int q) {
      int j = jl[q].first, l = jl[q].second;
      Mat L, W;
      if (!cholesky(L, M[j][l])) {  // cho! (:1846)
        st[q] = 1;
        return;
      }
      solve_tril(W, L, dM[j][l], false);  // (:1853)
      W = transpose(W);                   // (:1854)
      solve_tril(W, L, W, false);         // (:1856)
      int n = W.r;
      Mat Ws(n, n);
      for (int i = 0; i < n; i++)
        for (int i2 = 0; i2 < n; i2++) {
          r_add(Ws(i, i2), W(i, i2), W(i2, i));
          r_half(Ws(i, i2));
        }
      if (!lambda_min_sym(mins[q], Ws)) st[q] = 2;
    });
    for (size_t q = 0; q < jl.size(); q++) {
      if (st[q] == 1) return isX ? CLRSDP_ERR_NOT_PD_X : CLRSDP_ERR_NOT_PD_Y;
      if (st[q] == 2) return CLRSDP_ERR_EIG;
    }
    Real mn = mins[0];
    for (size_t q = 1; q < jl.size(); q++)
      if (r_cmp(mins[q], mn) < 0) mn = mins[q];
    lam_out = mn;
    Real ng;
    r_neg(ng, gamma);
    if (r_cmp(mn, ng) > 0)
      r_set_si(alpha, 1);  // (:1893-1894)
    else
      r_div(alpha, ng, mn);  // -gamma/min_eig (:1896)
    return 0;
  }

  // terminate (MPMP.jl:1147-1173)
  int terminate_reason() {
    bool gap_opt = r_cmp(dual_gap, gap_thr) < 0;
    bool pf = r_cmp(primal_error, perr_thr) < 0;
    bool df = r_cmp(dual_error, derr_thr) < 0;
    if (ip.need_primal_feasible && pf) return CLRSDP_PRIMAL_FEASIBLE;
    if (ip.need_dual_feasible && df) return CLRSDP_DUAL_FEASIBLE;
    if (pf && df && gap_opt) return CLRSDP_OPTIMAL;
    return CLRSDP_RUNNING;
  }
  bool check_pd_feasibility() {  // (:1176-1185)
    return r_cmp(primal_error, perr_thr) < 0 && r_cmp(dual_error, derr_thr) < 0;
  }
  Real primal_err() {  // compute_primal_error (:1058-1062)
    Real a = max_abs(p), bq = max_abs(P);
    return r_cmp(a, bq) > 0 ? a : bq;
  }

  // loop initialisation (MPMP.jl:716-736)
  int prepare(clrsdp_iter_info* info) {
    g_prec = prec;
    if (!have_point) return CLRSDP_ERR_STATE;
    R = like_blocks();
    Xinv = like_blocks();
    P = like_blocks();
    Z = like_blocks();
    dX = like_blocks();
    dY = like_blocks();
    XYsave = like_blocks();
    iter = 1;
    r_zero(alpha_p);
    r_zero(alpha_d);
    Real nn;
    r_set_si(nn, ntot);
    mu = dot_blocks(X, Y);
    r_div(mu, mu, nn);
    p_obj = primal_objective();
    d_obj = dual_objective();
    dual_gap = duality_gap_point();
    compute_residuals(false);
    primal_error = primal_err();
    dual_error = max_abs(d);
    pd_feas = check_pd_feasibility();
    prepared = true;
    if (info) fill_info(info, 0, 0.0);
    return 0;
  }

  void fill_info(clrsdp_iter_info* info, int status, double secs) {
    memset(info, 0, sizeof(*info));
    info->iter = iter;
    info->status = status;
    info->pd_feasible = pd_feas;
    info->seconds = secs;
    info->terminate = terminate_reason();
    info->mu = mu.d();
    info->p_obj = info->p_obj_new = p_obj.d();
    info->d_obj = info->d_obj_new = d_obj.d();
    info->gap = info->gap_new = dual_gap.d();
    info->P_err = max_abs(P).d();
    info->p_err = max_abs(p).d();
    info->d_err = max_abs(d).d();
    info->primal_err_new = primal_error.d();
    info->dual_err_new = dual_error.d();
  }

  // one pass of the while body (MPMP.jl:754-953)
  int iterate(clrsdp_iter_info* info) {
    g_prec = prec;
    if (!prepared) return CLRSDP_ERR_STATE;
    double t_begin = now_s(), t0;
    clrsdp_iter_info row;
    memset(&row, 0, sizeof(row));
    Real nn;
    r_set_si(nn, ntot);
    mu = dot_blocks(X, Y);  // step 3 (:755-756)
    r_div(mu, mu, nn);
    if (pd_feas)
      r_zero(mu_p);
    else
      r_mul(mu_p, beta_inf, mu);
    row.iter = iter;
    row.mu = mu.d();
    row.p_obj = p_obj.d();
    row.d_obj = d_obj.d();
    row.gap = dual_gap.d();
    t0 = now_s();
    residual_R(mu_p, false);  // (:760)
    double time_R = now_s() - t0;
    t0 = now_s();
    {  // X inverse (:764-801)
      std::vector<int> ok(jl.size(), 1);
      parallel_for((int)jl.size(), [&](int q) {
        int j = jl[q].first, l = jl[q].second;
        ok[q] = spd_inverse(Xinv[j][l], X[j][l]) ? 1 : 0;
      });
      for (int o : ok)
        if (!o) return fail(info, CLRSDP_ERR_NOT_PD_X);
    }
    row.timings[CLRSDP_T_XINV] = now_s() - t0;
    t0 = now_s();
    int st = T_decomposition();  // (:806)
    if (st) return fail(info, st);
    row.timings[CLRSDP_T_DECOMP] = now_s() - t0;
    row.timings[CLRSDP_T_SCHUR] = t_schur;
    row.timings[CLRSDP_T_CHOL_S] = t_cholS;
    row.timings[CLRSDP_T_CINVB] = t_CinvB;
    row.timings[CLRSDP_T_Q] = t_Q;
    row.timings[CLRSDP_T_CHOL_Q] = t_cholQ;
    t0 = now_s();
    compute_residuals(true);  // (:812)
    row.timings[CLRSDP_T_RES] = now_s() - t0;
    row.P_err = max_abs(P).d();
    row.p_err = max_abs(p).d();
    row.d_err = max_abs(d).d();
    for (double& t : t_dir) t = 0;
    t0 = now_s();
    search_direction();  // predictor (:818)
    row.timings[CLRSDP_T_PREDICTOR] = now_s() - t0;
    dX_pred = dX;
    dY_pred = dY;
    dx_pred = dx;
    dy_pred = dy;
    // step 5 (:832-837)
    Real r, beta, one;
    r_set_si(one, 1);
    {
      BlockDiag XdX = X, YdY = Y;
      for (size_t q = 0; q < jl.size(); q++) {
        int j = jl[q].first, l = jl[q].second;
        for (size_t i = 0; i < XdX[j][l].a.size(); i++) {
          r_add(XdX[j][l].a[i], X[j][l].a[i], dX[j][l].a[i]);
          r_add(YdY[j][l].a[i], Y[j][l].a[i], dY[j][l].a[i]);
        }
      }
      r = dot_blocks(XdX, YdY);
      Real den;
      r_mul(den, mu, nn);
      r_div(r, r, den);
    }
    if (r_cmp(r, one) < 0)
      r_mul(beta, r, r);
    else
      beta = r;
    if (pd_feas) {
      beta_c = r_cmp(beta_feas, beta) > 0 ? beta_feas : beta;
      if (r_cmp(beta_c, one) > 0) beta_c = one;
    } else {
      beta_c = r_cmp(beta_inf, beta) > 0 ? beta_inf : beta;
    }
    r_mul(mu_c, beta_c, mu);
    t0 = now_s();
    residual_R(mu_c, true);  // step 6 (:841)
    time_R += now_s() - t0;
    row.timings[CLRSDP_T_R] = time_R;
    t0 = now_s();
    search_direction();  // corrector (:846)
    row.timings[CLRSDP_T_CORRECTOR] = now_s() - t0;
    for (int i = 0; i < 5; i++) row.timings[CLRSDP_T_Z + i] = t_dir[i];
    t0 = now_s();
    st = step_length(alpha_p, lam_x, X, dX, true);  // step 7 (:863-866)
    if (st) return fail(info, st);
    st = step_length(alpha_d, lam_y, Y, dY, false);
    if (st) return fail(info, st);
    row.timings[CLRSDP_T_ALPHA] = now_s() - t0;
    if (pd_feas) {  // (:871-874)
      if (r_cmp(alpha_d, alpha_p) < 0) alpha_p = alpha_d;
      alpha_d = alpha_p;
    }
    // step 8 (:877-887)
    for (int i = 0; i < sumS; i++) r_fma(x[i], dx[i], alpha_p);
    for (int i = 0; i < n_y; i++) r_fma(y[i], dy[i], alpha_d);
    parallel_for((int)jl.size(), [&](int q) {
      int j = jl[q].first, l = jl[q].second;
      for (size_t i = 0; i < X[j][l].a.size(); i++) {
        r_fma(X[j][l].a[i], dX[j][l].a[i], alpha_p);
        r_fma(Y[j][l].a[i], dY[j][l].a[i], alpha_d);
      }
    });
    row.alpha_p = alpha_p.d();
    row.alpha_d = alpha_d.d();
    row.beta_c = beta_c.d();
    // new objectives, STALE errors (:940-944)
    p_obj = primal_objective();
    d_obj = dual_objective();
    dual_gap = duality_gap(p_obj, d_obj);
    primal_error = primal_err();
    dual_error = max_abs(d);
    iter += 1;
    pd_feas = check_pd_feasibility();
    row.p_obj_new = p_obj.d();
    row.d_obj_new = d_obj.d();
    row.gap_new = dual_gap.d();
    row.primal_err_new = primal_error.d();
    row.dual_err_new = dual_error.d();
    row.pd_feasible = pd_feas;
    row.terminate = terminate_reason();
    if (row.terminate == CLRSDP_RUNNING && iter >= ip.maxiterations) row.terminate = CLRSDP_MAXITER;
    row.seconds = now_s() - t_begin;
    if (info) *info = row;
    return 0;
  }
  int fail(clrsdp_iter_info* info, int st) {
    if (info) {
      memset(info, 0, sizeof(*info));
      info->iter = iter;
      info->status = st;
    }
    return st;
  }
};

// ---------------------------------------------------------------------------------------------------
// wire format <-> MPFR
// ---------------------------------------------------------------------------------------------------
namespace {
void from_wire(Real& out, const clrsdp_mp* a, int64_t i, int nlimb) {
  if (a->sign[i] == 0) {
    r_zero(out);
    return;
  }
  int n64 = (int)((out.v._mpfr_prec + 63) / 64);
  int total32 = 2 * n64;
  for (int q = 0; q < n64; q++) out.v._mpfr_d[q] = 0;
  for (int k = 0; k < nlimb; k++) {
    int pos = total32 - nlimb + k;  // top-aligned
    uint64_t w = a->limb[(size_t)k * a->n + i];
    out.v._mpfr_d[pos / 2] |= (pos & 1) ? (w << 32) : w;
  }
  out.v._mpfr_exp = a->exp[i];
  out.v._mpfr_sign = a->sign[i] < 0 ? -1 : 1;
}
void to_wire(clrsdp_mp_out* o, int64_t i, const Real& x, int nlimb) {
  if (mpfr_zero_p(&x.v)) {
    o->sign[i] = 0;
    o->exp[i] = 0;
    for (int k = 0; k < nlimb; k++) o->limb[(size_t)k * o->n + i] = 0;
    return;
  }
  int n64 = (int)((x.v._mpfr_prec + 63) / 64);
  int total32 = 2 * n64;
  for (int k = 0; k < nlimb; k++) {
    int pos = total32 - nlimb + k;
    uint64_t w = x.v._mpfr_d[pos / 2];
    o->limb[(size_t)k * o->n + i] = (uint32_t)((pos & 1) ? (w >> 32) : w);
  }
  o->exp[i] = x.v._mpfr_exp;
  o->sign[i] = x.v._mpfr_sign < 0 ? -1 : 1;
}
}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API int clrsdp_ref_create(clrsdp_handle* h, int prec_bits, int nthreads) {
  {
    const char* gm = getenv("CLRSDP_REF_GEMM");
    g_gemm_fixed = (gm && std::string(gm) == "fixed") ? 1 : 0;
    if (const char* cb = getenv("CLRSDP_REF_CLAMP")) g_clamp_bits = atoi(cb);
    if (const char* fg = getenv("CLRSDP_REF_FIXED_GUARD")) g_fixed_guard = std::min(64, std::max(1, atoi(fg)));
    if (const char* gs = getenv("CLRSDP_REF_GUARD_SITES")) g_guard_sites = atoi(gs);
    if (const char* rf = getenv("CLRSDP_REF_REFINE")) g_refine = atoi(rf);
  }
  if (!h || prec_bits < 64 || prec_bits % 32) return CLRSDP_ERR_BAD_ARG;
  g_prec = prec_bits;
  clrsdp_solver* s = new clrsdp_solver();
  {
    const char* fm = getenv("CLRSDP_REF_FACTOR");
    s->factor_mode = (fm && std::string(fm) == "chol") ? 1 : ((fm && std::string(fm) == "ldl") ? 2 : 0);
  }
  s->prec = prec_bits;
  s->nlimb = prec_bits / 32;
  s->nthreads = nthreads > 0 ? nthreads : (int)std::max(1u, std::thread::hardware_concurrency());
  // defaults of MPMP.jl:602-609 (exact decimal -> p-bit conversions happen in set_params; these are
  // the same values rounded from doubles only if the caller never sets them)
  r_set_si(s->beta_inf, 3);
  mpfr_div_ui(&s->beta_inf.v, &s->beta_inf.v, 10, MPFR_RNDN);
  r_set_si(s->beta_feas, 1);
  mpfr_div_ui(&s->beta_feas.v, &s->beta_feas.v, 10, MPFR_RNDN);
  r_set_si(s->gamma, 7);
  mpfr_div_ui(&s->gamma.v, &s->gamma.v, 10, MPFR_RNDN);
  mpfr_set_d(&s->omega_p.v, 1e10, MPFR_RNDN);
  mpfr_set_d(&s->omega_d.v, 1e10, MPFR_RNDN);
  auto pow10neg = [](Real& r, int e) {
    r_set_si(r, 1);
    for (int i = 0; i < e; i++) mpfr_div_ui(&r.v, &r.v, 10, MPFR_RNDN);
  };
  pow10neg(s->gap_thr, 15);
  pow10neg(s->perr_thr, 30);
  pow10neg(s->derr_thr, 30);
  *h = s;
  return 0;
}
REF_API int clrsdp_ref_destroy(clrsdp_handle h) {
  if (h) {
    g_prec = h->prec;
    delete h;
  }
  return 0;
}
REF_API const char* clrsdp_ref_last_error(clrsdp_handle h) { return h ? h->err.c_str() : "null handle"; }

REF_API int clrsdp_ref_set_structure(clrsdp_handle h, int J, int n_y, const int* m, const int* L,
                                     const int* n_samples, const int* delta, const int* ranks) {
  if (!h || J <= 0 || n_y <= 0) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  h->J = J;
  h->n_y = n_y;
  h->cl.assign(J, Cluster());
  h->x_idx.assign(J + 1, 0);
  h->jl.clear();
  h->ntot = 0;
  int di = 0, ri = 0;
  for (int j = 0; j < J; j++) {
    Cluster& c = h->cl[j];
    c.m = m[j];
    c.L = L[j];
    c.K = n_samples[j];
    c.dimS = c.m * (c.m + 1) / 2 * c.K;
    c.blk.assign(c.L, Block());
    for (int l = 0; l < c.L; l++) {
      Block& bk = c.blk[l];
      bk.delta = delta[di++];
      bk.nb = c.m * bk.delta;
      bk.ranks.assign(ranks + ri, ranks + ri + c.K);
      ri += c.K;
      bk.rank_sums.assign(c.K + 1, 0);
      for (int k = 0; k < c.K; k++) bk.rank_sums[k + 1] = bk.rank_sums[k] + bk.ranks[k];
      bk.Nv = bk.rank_sums[c.K];
      h->jl.emplace_back(j, l);
      h->ntot += bk.nb;
    }
    h->x_idx[j + 1] = h->x_idx[j] + c.dimS;
  }
  h->sumS = h->x_idx[J];
  h->b.assign(n_y, Real());
  h->have_point = h->prepared = false;
  return 0;
}

REF_API int clrsdp_ref_upload_cluster(clrsdp_handle h, int j, const clrsdp_mp* V, const clrsdp_mp* H,
                                      const clrsdp_mp* B, const clrsdp_mp* c) {
  if (!h || j < 0 || j >= h->J) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  Cluster& cq = h->cl[j];
  int64_t vi = 0, hi = 0;
  for (int l = 0; l < cq.L; l++) {
    Block& bk = cq.blk[l];
    bk.V = Mat(bk.delta, bk.Nv);
    bk.H.assign(bk.Nv, Real());
    for (int v = 0; v < bk.Nv; v++) {
      for (int i = 0; i < bk.delta; i++) from_wire(bk.V(i, v), V, vi++, h->nlimb);
      from_wire(bk.H[v], H, hi++, h->nlimb);
    }
    bk.VT = transpose(bk.V);
  }
  if (vi != V->n || hi != H->n || B->n != (int64_t)cq.dimS * h->n_y || c->n != cq.dimS) return CLRSDP_ERR_BAD_ARG;
  cq.B = Mat(cq.dimS, h->n_y);
  for (int64_t i = 0; i < B->n; i++) from_wire(cq.B.a[i], B, i, h->nlimb);
  cq.c.assign(cq.dimS, Real());
  for (int i = 0; i < cq.dimS; i++) from_wire(cq.c[i], c, i, h->nlimb);
  return 0;
}

REF_API int clrsdp_ref_upload_objective(clrsdp_handle h, const clrsdp_mp* b, const clrsdp_mp* b0) {
  if (!h || b->n != h->n_y) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  for (int i = 0; i < h->n_y; i++) from_wire(h->b[i], b, i, h->nlimb);
  if (b0 && b0->n >= 1)
    from_wire(h->b0, b0, 0, h->nlimb);
  else
    r_zero(h->b0);
  return 0;
}

// C: blocks in (j,l) order, each nb x nb row-major, concatenated (like X); NULL or n == 0 restores C = 0
REF_API int clrsdp_ref_upload_C(clrsdp_handle h, const clrsdp_mp* Cw) {
  if (!h) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  if (!Cw || Cw->n == 0) {
    h->have_C = false;
    return 0;
  }
  int64_t tot = 0;
  for (auto& c : h->cl)
    for (auto& bk : c.blk) tot += (int64_t)bk.nb * bk.nb;
  if (Cw->n != tot) return CLRSDP_ERR_BAD_ARG;
  h->C = h->like_blocks();
  int64_t off = 0;
  for (size_t j = 0; j < h->cl.size(); j++)
    for (size_t l = 0; l < h->cl[j].blk.size(); l++)
      for (auto& e : h->C[j][l].a) from_wire(e, Cw, off++, h->nlimb);
  h->have_C = true;
  h->prepared = false;
  return 0;
}

REF_API int clrsdp_ref_set_params(clrsdp_handle h, const clrsdp_mp* rp, const clrsdp_int_params* ip) {
  if (!h) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  if (rp) {
    if (rp->n != CLRSDP_P_COUNT) return CLRSDP_ERR_BAD_ARG;
    Real* dst[CLRSDP_P_COUNT] = {&h->beta_inf, &h->beta_feas, &h->gamma,    &h->omega_p,
                                 &h->omega_d,  &h->gap_thr,   &h->perr_thr, &h->derr_thr};
    for (int i = 0; i < CLRSDP_P_COUNT; i++) from_wire(*dst[i], rp, i, h->nlimb);
  }
  if (ip) h->ip = *ip;
  return 0;
}

REF_API int clrsdp_ref_init_point(clrsdp_handle h) {  // MPMP.jl:660-686
  if (!h || h->J == 0) return CLRSDP_ERR_STATE;
  g_prec = h->prec;
  h->x.assign(h->sumS, Real());
  h->y.assign(h->n_y, Real());
  h->X = h->like_blocks();
  h->Y = h->like_blocks();
  for (auto& jl : h->jl) {
    Mat &Xb = h->X[jl.first][jl.second], &Yb = h->Y[jl.first][jl.second];
    for (int i = 0; i < Xb.r; i++) {
      Xb(i, i) = h->omega_p;
      Yb(i, i) = h->omega_d;
    }
  }
  h->have_point = true;
  h->prepared = false;
  return 0;
}

REF_API int clrsdp_ref_upload_point(clrsdp_handle h, const clrsdp_mp* x, const clrsdp_mp* X,
                                    const clrsdp_mp* y, const clrsdp_mp* Y) {  // MPMP.jl:689
  if (!h || h->J == 0) return CLRSDP_ERR_STATE;
  g_prec = h->prec;
  if (x->n != h->sumS || y->n != h->n_y) return CLRSDP_ERR_BAD_ARG;
  h->x.assign(h->sumS, Real());
  h->y.assign(h->n_y, Real());
  for (int i = 0; i < h->sumS; i++) from_wire(h->x[i], x, i, h->nlimb);
  for (int i = 0; i < h->n_y; i++) from_wire(h->y[i], y, i, h->nlimb);
  h->X = h->like_blocks();
  h->Y = h->like_blocks();
  int64_t off = 0;
  for (auto& jl : h->jl) {
    Mat &Xb = h->X[jl.first][jl.second], &Yb = h->Y[jl.first][jl.second];
    for (size_t i = 0; i < Xb.a.size(); i++) {
      from_wire(Xb.a[i], X, off + i, h->nlimb);
      from_wire(Yb.a[i], Y, off + i, h->nlimb);
    }
    off += Xb.a.size();
  }
  if (off != X->n || off != Y->n) return CLRSDP_ERR_BAD_ARG;
  h->have_point = true;
  h->prepared = false;
  return 0;
}

REF_API int clrsdp_ref_download_point(clrsdp_handle h, clrsdp_mp_out* x, clrsdp_mp_out* X,
                                      clrsdp_mp_out* y, clrsdp_mp_out* Y) {
  if (!h || !h->have_point) return CLRSDP_ERR_STATE;
  g_prec = h->prec;
  if (x)
    for (int i = 0; i < h->sumS; i++) to_wire(x, i, h->x[i], h->nlimb);
  if (y)
    for (int i = 0; i < h->n_y; i++) to_wire(y, i, h->y[i], h->nlimb);
  int64_t off = 0;
  for (auto& jl : h->jl) {
    Mat &Xb = h->X[jl.first][jl.second], &Yb = h->Y[jl.first][jl.second];
    for (size_t i = 0; i < Xb.a.size(); i++) {
      if (X) to_wire(X, off + i, Xb.a[i], h->nlimb);
      if (Y) to_wire(Y, off + i, Yb.a[i], h->nlimb);
    }
    off += Xb.a.size();
  }
  return 0;
}

REF_API int clrsdp_ref_prepare(clrsdp_handle h, clrsdp_iter_info* info) { return h ? h->prepare(info) : CLRSDP_ERR_BAD_ARG; }
REF_API int clrsdp_ref_iterate(clrsdp_handle h, clrsdp_iter_info* info) { return h ? h->iterate(info) : CLRSDP_ERR_BAD_ARG; }

REF_API int clrsdp_ref_solve(clrsdp_handle h, clrsdp_iter_info* rows, int max_rows, int* n_rows) {
  if (!h) return CLRSDP_ERR_BAD_ARG;
  int st = h->prepare(nullptr);
  if (st) return st;
  int n = 0;
  // while !terminate(...) && iter < maxiterations (MPMP.jl:742-753)
  while (h->terminate_reason() == CLRSDP_RUNNING && h->iter < h->ip.maxiterations) {
    clrsdp_iter_info row;
    st = h->iterate(&row);
    if (rows && n < max_rows) rows[n] = row;
    n++;
    if (st) break;
  }
  if (n_rows) *n_rows = n;
  return st;
}

namespace {
int64_t put_vec(clrsdp_mp_out* out, const std::vector<Real>& v, int nlimb) {
  if (out) {
    if (out->n < (int64_t)v.size()) return CLRSDP_ERR_BAD_ARG;
    for (size_t i = 0; i < v.size(); i++) to_wire(out, i, v[i], nlimb);
  }
  return (int64_t)v.size();
}
}  // namespace

REF_API int64_t clrsdp_ref_fetch(clrsdp_handle h, const char* name, int j, int l, clrsdp_mp_out* out) {
  if (!h || !name) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  std::string nm(name);
  auto blockmat = [&](const BlockDiag& M) -> int64_t {
    if (j < 0 || j >= (int)M.size() || l < 0 || l >= (int)M[j].size()) return CLRSDP_ERR_BAD_ARG;
    return put_vec(out, M[j][l].a, h->nlimb);
  };
  if (nm == "x") return put_vec(out, h->x, h->nlimb);
  if (nm == "y") return put_vec(out, h->y, h->nlimb);
  if (nm == "dx") return put_vec(out, h->dx, h->nlimb);
  if (nm == "dy") return put_vec(out, h->dy, h->nlimb);
  if (nm == "dx_pred") return put_vec(out, h->dx_pred, h->nlimb);
  if (nm == "dy_pred") return put_vec(out, h->dy_pred, h->nlimb);
  if (nm == "p") return put_vec(out, h->p, h->nlimb);
  if (nm == "d") return put_vec(out, h->d, h->nlimb);
  if (nm == "b") return put_vec(out, h->b, h->nlimb);
  if (nm == "c") {
    std::vector<Real> cc;
    for (auto& c : h->cl) cc.insert(cc.end(), c.c.begin(), c.c.end());
    return put_vec(out, cc, h->nlimb);
  }
  if (nm == "X") return blockmat(h->X);
  if (nm == "Y") return blockmat(h->Y);
  if (nm == "Xinv") return blockmat(h->Xinv);
  if (nm == "R") return blockmat(h->R);
  if (nm == "P") return blockmat(h->P);
  if (nm == "Z") return blockmat(h->Z);
  if (nm == "dX") return blockmat(h->dX);
  if (nm == "dY") return blockmat(h->dY);
  if (nm == "dX_pred") return blockmat(h->dX_pred);
  if (nm == "dY_pred") return blockmat(h->dY_pred);
  if (nm == "XY") return blockmat(h->XYsave);
  if (nm == "Px") return blockmat(h->Px_keep);
  if (nm == "Py") return blockmat(h->Py_keep);
  if (nm == "S") {
    if (j < 0 || j >= (int)h->S_keep.size()) return CLRSDP_ERR_BAD_ARG;
    return put_vec(out, h->S_keep[j].a, h->nlimb);
  }
  if (nm == "Q") return put_vec(out, h->Q_keep.a, h->nlimb);
  if (nm == "scalar") {
    const Real* tab[CLRSDP_S_COUNT] = {&h->mu,      &h->p_obj,   &h->d_obj,  &h->dual_gap, &h->primal_error,
                                       &h->dual_error, &h->alpha_p, &h->alpha_d, &h->beta_c, &h->mu_p,
                                       &h->mu_c,    &h->lam_x,   &h->lam_y};
    if (j < 0 || j >= CLRSDP_S_COUNT) return CLRSDP_ERR_BAD_ARG;
    if (out) {
      if (out->n < 1) return CLRSDP_ERR_BAD_ARG;
      to_wire(out, 0, *tab[j], h->nlimb);
    }
    return 1;
  }
  return CLRSDP_ERR_BAD_ARG;
}

// ---- phase-level ops (the same dense kernels, exposed for parity tests and the CPU baseline) --------
REF_API int clrsdp_ref_op_gemm(clrsdp_handle h, int batch, int M, int N, int K, const clrsdp_mp* A,
                               const clrsdp_mp* B, clrsdp_mp_out* C) {
  if (!h) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  h->parallel_for(batch, [&](int bq) {
    Mat a(M, K), b(K, N), c;
    for (int i = 0; i < M * K; i++) from_wire(a.a[i], A, (int64_t)bq * M * K + i, h->nlimb);
    for (int i = 0; i < K * N; i++) from_wire(b.a[i], B, (int64_t)bq * K * N + i, h->nlimb);
    gemm(c, a, b);
    for (int i = 0; i < M * N; i++) to_wire(C, (int64_t)bq * M * N + i, c.a[i], h->nlimb);
  });
  return 0;
}
REF_API int clrsdp_ref_op_cholesky(clrsdp_handle h, int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* L,
                                   clrsdp_mp_out* Linv) {
  if (!h) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  std::vector<int> ok(batch, 1);
  h->parallel_for(batch, [&](int bq) {
    Mat a(n, n), Lm, Li;
    for (int i = 0; i < n * n; i++) from_wire(a.a[i], A, (int64_t)bq * n * n + i, h->nlimb);
    if (!cholesky(Lm, a)) {
      ok[bq] = 0;
      return;
    }
    if (L)
      for (int i = 0; i < n * n; i++) to_wire(L, (int64_t)bq * n * n + i, Lm.a[i], h->nlimb);
    if (Linv) {
      Mat I(n, n);
      for (int i = 0; i < n; i++) r_set_si(I(i, i), 1);
      solve_tril(Li, Lm, I, false);
      for (int i = 0; i < n * n; i++) to_wire(Linv, (int64_t)bq * n * n + i, Li.a[i], h->nlimb);
    }
  });
  for (int o : ok)
    if (!o) return CLRSDP_ERR_NOT_PD_X;
  return 0;
}
REF_API int clrsdp_ref_op_lambda_min(clrsdp_handle h, int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* lam) {
  if (!h) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  h->parallel_for(batch, [&](int bq) {
    Mat a(n, n);
    for (int i = 0; i < n * n; i++) from_wire(a.a[i], A, (int64_t)bq * n * n + i, h->nlimb);
    Real lm;
    lambda_min_sym(lm, a);
    to_wire(lam, bq, lm, h->nlimb);
  });
  return 0;
}
REF_API int clrsdp_ref_op_elementwise(clrsdp_handle h, int op, const clrsdp_mp* a, const clrsdp_mp* b,
                                      clrsdp_mp_out* c) {
  if (!h) return CLRSDP_ERR_BAD_ARG;
  g_prec = h->prec;
  Real x, y, z;
  for (int64_t i = 0; i < a->n; i++) {
    from_wire(x, a, i, h->nlimb);
    if (b) from_wire(y, b, i, h->nlimb);
    switch (op) {
      case '+': r_add(z, x, y); break;
      case '-': r_sub(z, x, y); break;
      case '*': r_mul(z, x, y); break;
      case '/': r_div(z, x, y); break;
      case 's': mpfr_sqrt(&z.v, &x.v, MPFR_RNDN); break;
      default: return CLRSDP_ERR_BAD_ARG;
    }
    to_wire(c, i, z, h->nlimb);
  }
  return 0;
}

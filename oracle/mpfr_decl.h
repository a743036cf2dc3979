// TEST INFRASTRUCTURE ONLY (oracle/). Hand-declared prototypes for the header-less libmpfr.so.6
// (MPFR 4.2.1) / libgmp.so.10 installed in this image. Layout of __mpfr_struct is MPFR's public ABI
// (also Julia's BigFloat). Only what the oracle uses is declared.
#pragma once
#include <limits.h>
#include <stdint.h>

extern "C" {
typedef long mpfr_prec_t;
typedef long mpfr_exp_t;
typedef unsigned long mp_limb_t;
typedef struct {
  mpfr_prec_t _mpfr_prec;
  int _mpfr_sign;
  mpfr_exp_t _mpfr_exp;
  mp_limb_t* _mpfr_d;
} __mpfr_struct;
typedef __mpfr_struct* mpfr_ptr;
typedef const __mpfr_struct* mpfr_srcptr;
typedef int mpfr_rnd_t;
#define MPFR_RNDN 0
#define MPFR_EXP_ZERO_ (LONG_MIN + 1)
#define MPFR_EXP_NAN_ (LONG_MIN + 2)
#define MPFR_EXP_INF_ (LONG_MIN + 3)

void mpfr_init2(mpfr_ptr, mpfr_prec_t);
void mpfr_clear(mpfr_ptr);
int mpfr_set4(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t, int);
int mpfr_set_d(mpfr_ptr, double, mpfr_rnd_t);
int mpfr_set_si(mpfr_ptr, long, mpfr_rnd_t);
int mpfr_set_ui_2exp(mpfr_ptr, unsigned long, mpfr_exp_t, mpfr_rnd_t);
void mpfr_set_zero(mpfr_ptr, int);
int mpfr_add(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sub(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_mul(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sqr(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_div(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sqrt(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_fma(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_fms(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_neg(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_cmp3(mpfr_srcptr, mpfr_srcptr, int);
int mpfr_cmpabs(mpfr_srcptr, mpfr_srcptr);
int mpfr_cmp_si_2exp(mpfr_srcptr, long, mpfr_exp_t);
int mpfr_sgn(mpfr_srcptr);
int mpfr_zero_p(mpfr_srcptr);
double mpfr_get_d(mpfr_srcptr, mpfr_rnd_t);
int mpfr_mul_2si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_div_2si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_div_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_mul_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
void mpfr_swap(mpfr_ptr, mpfr_ptr);
// libgmp's mpn layer (public ABI names): used by the block fixed-point product (the way libarb's approx_mul works)
typedef long mp_size_t;
void __gmpn_mul_n(mp_limb_t*, const mp_limb_t*, const mp_limb_t*, mp_size_t);
mp_limb_t __gmpn_add_n(mp_limb_t*, const mp_limb_t*, const mp_limb_t*, mp_size_t);
mp_limb_t __gmpn_sub_n(mp_limb_t*, const mp_limb_t*, const mp_limb_t*, mp_size_t);
mp_limb_t __gmpn_rshift(mp_limb_t*, const mp_limb_t*, mp_size_t, unsigned int);
mp_limb_t __gmpn_lshift(mp_limb_t*, const mp_limb_t*, mp_size_t, unsigned int);
int __gmpn_cmp(const mp_limb_t*, const mp_limb_t*, mp_size_t);
}

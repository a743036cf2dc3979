// Host build of csrc/mpf.cuh (the same source the device kernels compile) for CPU unit tests.
#include "../../include/clrsdp.h"
#include "../../clustered-low-rank-sdp-solver_b200/csrc/mpf.cuh"

template <int NL>
static mp::Num<NL> rd(const clrsdp_mp* a, int64_t i) {
  mp::Num<NL> x;
  if (a->sign[i] == 0) return mp::zero<NL>();
  for (int k = 0; k < NL; k++) x.m[k] = a->limb[(size_t)k * a->n + i];
  x.e = (int32_t)a->exp[i];
  x.neg = a->sign[i] < 0;
  return x;
}
template <int NL>
static void wr(clrsdp_mp_out* o, int64_t i, const mp::Num<NL>& x) {
  if (mp::is_zero(x)) {
    o->sign[i] = 0; o->exp[i] = 0;
    for (int k = 0; k < NL; k++) o->limb[(size_t)k * o->n + i] = 0;
    return;
  }
  for (int k = 0; k < NL; k++) o->limb[(size_t)k * o->n + i] = x.m[k];
  o->exp[i] = x.e;
  o->sign[i] = x.neg ? -1 : 1;
}
template <int NL>
static int run(int op, const clrsdp_mp* a, const clrsdp_mp* b, clrsdp_mp_out* c) {
  for (int64_t i = 0; i < a->n; i++) {
    mp::Num<NL> x = rd<NL>(a, i), y = b ? rd<NL>(b, i) : mp::zero<NL>(), z;
    switch (op) {
      case '+': z = mp::add(x, y); break;
      case '-': z = mp::sub(x, y); break;
      case '*': z = mp::mul(x, y); break;
      case '/': z = mp::div(x, y); break;
      case 's': z = mp::sqrt(x); break;
      case 'r': { mp::Num<NL> r; mp::sqrt_rsqrt(x, r); z = r; } break;
      case 'm': { int64_t j = (i + 1) % a->n; z = mp::mul_sub_mul(x, y, rd<NL>(a, j), rd<NL>(b, j)); } break;
      case 'd': z = mp::from_double<NL>(mp::to_double(x)); break;
      case 'c': z = mp::from_int<NL>(mp::cmp(x, y)); break;
      case 'i': z = mp::from_int<NL>((int64_t)a->exp[i] * (a->sign[i] < 0 ? -1 : 1)); break;
      default: return -1;
    }
    wr<NL>(c, i, z);
  }
  return 0;
}
extern "C" int mpf_host_op(int nlimb, int op, const clrsdp_mp* a, const clrsdp_mp* b, clrsdp_mp_out* c) {
  switch (nlimb) {
    case 4: return run<4>(op, a, b, c);
    case 8: return run<8>(op, a, b, c);
    case 12: return run<12>(op, a, b, c);
    case 16: return run<16>(op, a, b, c);
  }
  return -1;
}

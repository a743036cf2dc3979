// Measuring aid (not a test): latency of the scalar multiprecision operations in a dependent chain, one warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../clustered-low-rank-sdp-solver_b200/csrc mpf_latency.cu -o mpf_latency
#include <cstdio>
#include <cuda_runtime.h>
#include "mpf.cuh"

template <int NL, int OP>
__global__ void chain(uint32_t* out, long long* cyc, int iters, int nwarps_active, int divergent) {
  mp::Num<NL> a, b, c, d;
  const int t = threadIdx.x;
  for (int i = 0; i < NL; i++) {
    a.m[i] = 0x9e3779b9u * (t + 1 + i) | 0x80000000u * (i == NL - 1);
    b.m[i] = 0x7f4a7c15u * (t + 3 + i) | 0x80000000u * (i == NL - 1);
    c.m[i] = 0x85ebca6bu * (t + 5 + i) | 0x80000000u * (i == NL - 1);
    d.m[i] = 0xc2b2ae35u * (t + 7 + i) | 0x80000000u * (i == NL - 1);
  }
  a.e = 0, b.e = 0, c.e = -1, d.e = 0;
  a.neg = b.neg = c.neg = d.neg = 0;
  if (divergent) {  // lanes with either sign, every fifth with a zero factor, every seventh far apart in exponent
    c.neg = t & 1;
    if (t % 5 == 0) d = mp::zero<NL>();
    if (t % 7 == 0) c.e = -40;
  }
  if ((t >> 5) >= nwarps_active) return;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (OP == 0) a = mp::mul_sub_mul(b, a, c, d);       // elimination step: mu*x - gam*y
    if (OP == 1) { a = mp::mul(a, b); a.e = 0; }
    if (OP == 2) { a = mp::add(a, c); a.e = 0; }
    if (OP == 3) { a = mp::add(a, mp::mul(b, c)); a.e = 0; }
    if (OP == 0) a.e = 0;
  }
  long long t1 = clock64();
  if (t == 0) cyc[0] = t1 - t0;
  uint32_t s = 0;
  for (int i = 0; i < NL; i++) s ^= a.m[i];
  out[t] = s + a.e + a.neg;
}
template <int NL>
void run(const char* name) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 4096 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 200;
  for (int dv = 0; dv < 2; dv++)
  for (int nw : {1, 4, 16}) {
    long long h[4];
    for (int op = 0; op < 4; op++) {
      if (op == 0) chain<NL, 0><<<1, 512>>>(out, cyc, iters, nw, dv);
      if (op == 1) chain<NL, 1><<<1, 512>>>(out, cyc, iters, nw, dv);
      if (op == 2) chain<NL, 2><<<1, 512>>>(out, cyc, iters, nw, dv);
      if (op == 3) chain<NL, 3><<<1, 512>>>(out, cyc, iters, nw, dv);
      cudaDeviceSynchronize();
      cudaMemcpy(&h[op], cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("%s %s warps=%2d  cycles per op: mul_sub_mul %6lld  mul %6lld  add %6lld  add(mul) %6lld\n", name, dv ? "divergent" : "uniform  ", nw, h[0] / iters, h[1] / iters,
           h[2] / iters, h[3] / iters);
  }
}
int main() {
  run<8>("256-bit");
  run<16>("512-bit");
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}

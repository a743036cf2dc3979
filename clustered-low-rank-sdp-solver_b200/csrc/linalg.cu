// linalg.cu — CUDA-core multiprecision kernels (see linalg.cuh). One thread owns one number in registers;
// tensors are planar in HBM so that a warp's 32 numbers are one coalesced line per limb plane.
#include "linalg.cuh"
#include "../../include/clrsdp.h"
#include "comm.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <atomic>
#include <cstring>

namespace clr {
namespace cg = cooperative_groups;
using mp::Num;
constexpr int MAX_DEVICES = 64;  // device ordinals of one box (per-device one-off settings)

#define DISPATCH_NL(nl, ...)                                  \
  switch (nl) {                                               \
    case 4: { constexpr int NL = 4; __VA_ARGS__; } break;     \
    case 8: { constexpr int NL = 8; __VA_ARGS__; } break;     \
    case 12: { constexpr int NL = 12; __VA_ARGS__; } break;   \
    case 16: { constexpr int NL = 16; __VA_ARGS__; } break;   \
    default: throw SolverError(-1, "unsupported precision");  \
  }

// Out-of-line wrappers: the latency-bound kernels below have dozens of call sites of the (fully unrolled)
// limb loops; keeping each operation a real function keeps the code (and compile time) bounded. The
// HBM-bound elementwise kernels keep the inlined versions.
template <int NL> __device__ __noinline__ Num<NL> nadd(const Num<NL>& a, const Num<NL>& b) { return mp::add(a, b); }
template <int NL> __device__ __noinline__ Num<NL> nsub(const Num<NL>& a, const Num<NL>& b) { return mp::sub(a, b); }
template <int NL> __device__ __noinline__ Num<NL> nmul(const Num<NL>& a, const Num<NL>& b) { return mp::mul(a, b); }
template <int NL> __device__ __noinline__ Num<NL> ndiv(const Num<NL>& a, const Num<NL>& b) { return mp::div(a, b); }
template <int NL> __device__ __noinline__ Num<NL> nrecip(const Num<NL>& a) { return mp::recip(a); }
template <int NL> __device__ __noinline__ Num<NL> nsqrt(const Num<NL>& a) { return mp::sqrt(a); }
template <int NL> __device__ __noinline__ Num<NL> nsqrt_rsqrt(const Num<NL>& a, Num<NL>& r) { return mp::sqrt_rsqrt(a, r); }
template <int NL> __device__ __noinline__ Num<NL> nmsm(const Num<NL>& a, const Num<NL>& b, const Num<NL>& c, const Num<NL>& d) { return mp::mul_sub_mul(a, b, c, d); }
template <int NL> __device__ __noinline__ int ncmp(const Num<NL>& a, const Num<NL>& b) { return mp::cmp(a, b); }

// =========================================================================================================
// shared-memory / shuffle helpers for mp numbers
// =========================================================================================================
template <int NL>
__device__ __forceinline__ void smem_put(uint32_t* s, int slot, const Num<NL>& x) {
  uint32_t* p = s + (size_t)slot * (NL + 2);
#pragma unroll
  for (int k = 0; k < NL; k++) p[k] = x.m[k];
  p[NL] = (uint32_t)x.e;
  p[NL + 1] = x.neg;
}
template <int NL>
__device__ __forceinline__ Num<NL> smem_get(const uint32_t* s, int slot) {
  const uint32_t* p = s + (size_t)slot * (NL + 2);
  Num<NL> x;
#pragma unroll
  for (int k = 0; k < NL; k++) x.m[k] = p[k];
  x.e = (int32_t)p[NL];
  x.neg = p[NL + 1];
  return x;
}
template <int NL>
__device__ __forceinline__ Num<NL> shfl_xor_num(const Num<NL>& v, int o) {
  Num<NL> r;
#pragma unroll
  for (int k = 0; k < NL; k++) r.m[k] = __shfl_xor_sync(0xffffffffu, v.m[k], o);
  r.e = __shfl_xor_sync(0xffffffffu, v.e, o);
  r.neg = __shfl_xor_sync(0xffffffffu, v.neg, o);
  return r;
}
enum { RED_ADD = 0, RED_MAX = 1, RED_MIN = 2 };
template <int NL, int OP>
__device__ __forceinline__ Num<NL> red_op(const Num<NL>& a, const Num<NL>& b) {
  if (OP == RED_ADD) return nadd(a, b);
  if (OP == RED_MAX) return ncmp(a, b) >= 0 ? a : b;
  return ncmp(a, b) <= 0 ? a : b;
}
// block-wide reduction; the result is returned to every thread. scratch: 33*(NL+2) words.
// Every thread of the block must call it (it contains __syncthreads).
template <int NL, int OP>
__device__ Num<NL> block_reduce(Num<NL> v, uint32_t* scratch) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  const int lane = tid & 31, warp = tid >> 5, nwarp = (nthr + 31) >> 5;
#pragma unroll 1
  for (int o = 16; o; o >>= 1) {
    Num<NL> t = shfl_xor_num(v, o);
    v = red_op<NL, OP>(v, t);
  }
  __syncthreads();
  if (lane == 0) smem_put<NL>(scratch, warp, v);
  __syncthreads();
  if (warp == 0) {
    // lanes beyond nwarp replicate slot 0 for max/min (idempotent) and contribute zero for add
    Num<NL> z = smem_get<NL>(scratch, lane < nwarp ? lane : 0);
    if (OP == RED_ADD && lane >= nwarp) z = mp::zero<NL>();
#pragma unroll 1
    for (int o = 16; o; o >>= 1) {
      Num<NL> t = shfl_xor_num(z, o);
      z = red_op<NL, OP>(z, t);
    }
    if (lane == 0) smem_put<NL>(scratch, 32, z);
  }
  __syncthreads();
  return smem_get<NL>(scratch, 32);
}
template <int NL>
__device__ __forceinline__ Num<NL> ldm(const mp::Tensor& t, int64_t i) { return mp::load<NL>(t.w, t.n, (size_t)i); }
template <int NL>
__device__ __forceinline__ void stm(const mp::Tensor& t, int64_t i, const Num<NL>& x) { mp::store<NL>(t.w, t.n, (size_t)i, x); }

// =========================================================================================================
// Cholesky (upper factor), one CTA per matrix, threads (column c, part)
// =========================================================================================================
constexpr int TRI_THREADS = 512;
template <int NL>
__global__ void __launch_bounds__(TRI_THREADS)
chol_kernel(mp::Tensor A, const int64_t* __restrict__ offA, int64_t shiftA, int ldA, mp::Tensor U,
            const int64_t* __restrict__ offU, int64_t shiftU, int ld, mp::Tensor rdiag, int n, int* __restrict__ status) {
  extern __shared__ uint32_t sm[];
  __shared__ int bad;
  const int b = blockIdx.x, tx = threadIdx.x, part = threadIdx.y, CX = blockDim.x, P = blockDim.y;
  const int tid = part * CX + tx, nthr = CX * P;
  const int64_t oa = offA[b] + shiftA, ou = offU[b] + shiftU;
  if (tid == 0) bad = status[b];  // sticky across the panels of a blocked factorisation
  // U <- upper triangle of A, zeros below
  for (int r = part; r < n; r += P)
    for (int c = tx; c < n; c += CX)
      stm<NL>(U, ou + (int64_t)r * ld + c, (r <= c) ? ldm<NL>(A, oa + (int64_t)r * ldA + c) : mp::zero<NL>());
  __syncthreads();
  for (int k = 0; k < n; k++) {
    if (tid == 0) {
      Num<NL> a = ldm<NL>(U, ou + (int64_t)k * ld + k);
      if (mp::is_zero(a) || a.neg) {
        bad = 1;
      } else {
        Num<NL> rinv;
        Num<NL> d = nsqrt_rsqrt(a, rinv);
        stm<NL>(U, ou + (int64_t)k * ld + k, d);
        stm<NL>(rdiag, (int64_t)b * n + k, rinv);
        smem_put<NL>(sm, 0, rinv);
      }
    }
    __syncthreads();
    if (bad) break;
    {
      Num<NL> rinv = smem_get<NL>(sm, 0);
      for (int c = k + 1 + tid; c < n; c += nthr)
        stm<NL>(U, ou + (int64_t)k * ld + c, nmul(ldm<NL>(U, ou + (int64_t)k * ld + c), rinv));
    }
    __syncthreads();
    for (int c = k + 1 + tx; c < n; c += CX) {
      Num<NL> ukc = ldm<NL>(U, ou + (int64_t)k * ld + c);
      for (int r = k + 1 + part; r <= c; r += P) {
        Num<NL> ukr = ldm<NL>(U, ou + (int64_t)k * ld + r);
        int64_t at = ou + (int64_t)r * ld + c;
        stm<NL>(U, at, nsub(ldm<NL>(U, at), nmul(ukr, ukc)));
      }
    }
    __syncthreads();
  }
  if (tid == 0) status[b] = bad;
}

// V = U^-1 (upper) by right-looking back substitution; Linv = V^T written alongside.
template <int NL>
__global__ void __launch_bounds__(TRI_THREADS)
trinv_kernel(mp::Tensor U, const int64_t* __restrict__ offU, int64_t shiftU, mp::Tensor rdiag, mp::Tensor V,
             const int64_t* __restrict__ offV, int64_t shiftV, mp::Tensor Linv, const int64_t* __restrict__ offL,
             int64_t shiftL, int ld, int n) {
  const int b = blockIdx.x, tx = threadIdx.x, part = threadIdx.y, CX = blockDim.x, P = blockDim.y;
  const int tid = part * CX + tx, nthr = CX * P;
  const int64_t ou = offU[b] + shiftU, ov = offV[b] + shiftV, ol = offL ? offL[b] + shiftL : 0;
  for (int r = part; r < n; r += P)
    for (int c = tx; c < n; c += CX) {
      stm<NL>(V, ov + (int64_t)r * ld + c, mp::zero<NL>());
      if (offL && c > r) stm<NL>(Linv, ol + (int64_t)r * ld + c, mp::zero<NL>());
    }
  __syncthreads();
  for (int k = n - 1; k >= 0; k--) {
    {
      Num<NL> rk = ldm<NL>(rdiag, (int64_t)b * n + k);
      for (int c = k + tid; c < n; c += nthr) {
        Num<NL> acc = ldm<NL>(V, ov + (int64_t)k * ld + c);
        Num<NL> v = (c == k) ? nsub(mp::one<NL>(), acc) : mp::neg(acc);
        v = nmul(v, rk);
        stm<NL>(V, ov + (int64_t)k * ld + c, v);
        if (offL) stm<NL>(Linv, ol + (int64_t)c * ld + k, v);
      }
    }
    __syncthreads();
    for (int c = k + tx; c < n; c += CX) {
      Num<NL> vkc = ldm<NL>(V, ov + (int64_t)k * ld + c);
      for (int i = part; i < k; i += P) {
        Num<NL> uik = ldm<NL>(U, ou + (int64_t)i * ld + k);
        int64_t at = ov + (int64_t)i * ld + c;
        stm<NL>(V, at, nadd(ldm<NL>(V, at), nmul(uik, vkc)));
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fused panel factorisation: U (upper Cholesky factor) and G = L^-1 = U^-T of one w x w SPD block, w <= 32,
// entirely in shared memory (packed triangles). Gaussian elimination on the augmented matrix [A | I]:
// the row operations that turn A into U turn I into L^-1, so the inverse costs no extra dependency chain.
//
// The pivot chain is the latency floor of the whole solver (n dependent steps per n x n matrix), so the
// loop is DIVISION- AND SQUARE-ROOT-FREE: step k replaces every later row by
//        row_r <- mu_k * row_r - (R_k[r] * 2^-e_k) * row_k ,   mu_k = d'_k * 2^-e_k in [1/2, 1),
// (d'_k = current pivot), i.e. all unpivoted rows carry the common scale tau_k = prod_{i<k} mu_i. The
// true factors are recovered at the end, for all rows in parallel, by one rsqrt per row:
//        U_k = R'_k * f_k ,  G_k = G'_k * f_k ,  f_k = (tau_k * d'_k)^-1/2 .
// Warp 0 runs the chain one step ahead (row k+1, one element per lane: two multiplications and one
// subtraction per step) while the other warps apply step k to the rows below: ONE __syncthreads per step.
// ---------------------------------------------------------------------------------------------------------
constexpr int PANEL_THREADS = 544;  // warp 0: pivot row one step ahead; warps 1..15: elimination; warp 16: the row scale tau
constexpr int PANEL_ELIM = 480;     // threads of warps 1..15
constexpr int PANEL_W = 32;
__host__ __device__ inline int pk_u(int w, int r, int c) { return r * w - (r * (r - 1)) / 2 + (c - r); }  // r <= c
__host__ __device__ inline int pk_g(int r, int j) { return (r * (r + 1)) / 2 + j; }                      // j <= r
template <int NL>
__device__ __forceinline__ void panel_elim(uint32_t* Us, uint32_t* Gs, int w, int k, int r, int qp, const Num<NL>& mu,
                                           int ek) {
  // qp enumerates the active columns of row r: G part q = qp <= k, U part q = qp + (r - k - 1) >= r.
  // ONE instance of the fused operation for both parts (addresses by selection): the lanes of a warp straddle the two
  // parts in almost every step, and with a branch per part the warp ran the ~700-instruction a*b - c*d twice.
  const Num<NL> gam = mp::mul_2exp(smem_get<NL>(Us, pk_u(w, k, r)), -ek);
  const bool gp = qp <= k;
  const int q = qp + (r - k - 1);
  uint32_t* const base = gp ? Gs : Us;
  const int at = gp ? pk_g(r, qp) : pk_u(w, r, q);
  const int pt = gp ? pk_g(k, qp) : pk_u(w, k, q);
  smem_put<NL>(base, at, mp::mul_sub_mul(mu, smem_get<NL>(base, at), gam, smem_get<NL>(base, pt)));
}
// CLUSTER version (CL > 1 CTAs per matrix, launched as one thread-block cluster): the elimination is throughput-bound
// on the integer pipes of ONE SM when a CTA owns the whole panel (w^3/2 fused a*b - c*d operations of ~800 instructions
// each), and the factorisations with few matrices in the batch (Q: one matrix; the S_j: one per cluster) leave most SMs
// idle. Rows are owned cyclically (row r by CTA r % CL): a CTA eliminates its own rows in its own shared memory, and the
// owner of row k+1 - which its warp 0 finishes one step ahead, as before - keeps it in place; after the cluster barrier
// of the step the other CTAs copy it out of the owner's shared memory through DSMEM. Every CTA therefore ends with all
// final rows, computes the row scales f_k itself and stores its share of the two outputs.
template <int NL>
__global__ void __launch_bounds__(PANEL_THREADS)
panel_factor_kernel(mp::Tensor A, const int64_t* __restrict__ offA, int64_t shiftA, int ldA, mp::Tensor Lm,
                    const int64_t* __restrict__ offL, int64_t shiftL, int ldL, int w, int write_u,
                    int* __restrict__ status, int* __restrict__ sig, int sig_ld, int CL, int dbg) {
  extern __shared__ uint32_t sm[];
  __shared__ long long tdbg[4][36];
  const int ntri = w * (w + 1) / 2;
  uint32_t* Us = sm;                                  // packed upper triangle of the scaled work matrix R'
  uint32_t* Gs = sm + (size_t)ntri * (NL + 2);        // packed lower triangle of G' (diagonal slots: tau)
  uint32_t* Fs = sm + (size_t)2 * ntri * (NL + 2);    // f_k per row
  __shared__ int bad;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = CL > 1 ? (int)cluster.block_rank() : 0;
  const int b = blockIdx.x / CL, tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int64_t oa = offA[b] + shiftA, ol = offL[b] + shiftL;
  if (tid == 0) bad = status[b];
  for (int idx = tid; idx < w * w; idx += nthr) {
    int r = idx / w, c = idx % w;
    if (r <= c) smem_put<NL>(Us, pk_u(w, r, c), ldm<NL>(A, oa + (int64_t)r * ldA + c));
    if (c <= r) smem_put<NL>(Gs, pk_g(r, c), (r == 0 && c == 0) ? mp::one<NL>() : mp::zero<NL>());
  }
  __syncthreads();
  // SIGNED mode (sig != nullptr: the Schur complements S_j and Q, which the reference factors by pivoted LU and never
  // tests for definiteness, MPMP.jl:1436,1501): the factorisation is A = U^T Sigma U with Sigma = diag(+-1), i.e. an
  // LDL^T whose pivots keep their sign. Near the optimum S_j is singular to working precision and its computed trailing
  // pivots come out with either sign; a Cholesky factorisation has to perturb them (by far more than their rounding
  // error) to go on, which destroys the search direction, whereas the signed factorisation is backward stable at the
  // rounding level like the LU. The elimination below never divides, so a negative pivot costs nothing: the common row
  // scale tau just carries its sign. Only a pivot that is exactly zero or below 2^-(p+8) times the scale of its row is
  // replaced (by that bound, sign kept) - a perturbation below the rounding errors already in the matrix.
  constexpr int FLOOR_BITS = 32 * NL + 8;
  if (tid == 0) {  // (every CTA of a cluster does this on its own copy: same data, same result)
    Num<NL> a = smem_get<NL>(Us, pk_u(w, 0, 0));
    if (sig) {
      if (mp::is_zero(a) || a.e <= -FLOOR_BITS) {
        Num<NL> f = mp::from_pow2<NL>(-FLOOR_BITS);
        f.neg = mp::is_zero(a) ? 0u : a.neg;
        smem_put<NL>(Us, pk_u(w, 0, 0), f);
      }
    } else if (mp::is_zero(a) || a.neg) {
      bad = 1;
    }
  }
  if (CL > 1) cluster.sync(); else __syncthreads();   // (cluster: nobody pushes into a CTA that is still loading)
  for (int k = 0; k + 1 < w && !bad; k++) {
    // rows 0..k are final (scaled); apply elimination step k to the rows below
    if (dbg && lane == 0 && (warp == 0 || warp == 1 || warp == 15)) tdbg[warp == 15 ? 2 : warp][k] = clock64();
    Num<NL> mu = smem_get<NL>(Us, pk_u(w, k, k));
    const int ek = mu.e;
    mu.e = 0;
    if (warp == 0) {
      const int r = k + 1;
      if (r % CL == crank) {
        if (lane < w) panel_elim<NL>(Us, Gs, w, k, r, lane, mu, ek);  // exactly w active columns: one per lane
        __syncwarp();
        if (lane == 0) {
          Num<NL> a = smem_get<NL>(Us, pk_u(w, r, r));
          if (sig) {  // the unpivoted rows carry the scale tau_{k+1} = mu_k tau_k: threshold by exponents only (no
                      // multiplication on the pivot chain): |a| < 2^te  <=>  a.e <= te
            const int te = smem_get<NL>(Gs, pk_g(k, k)).e - FLOOR_BITS;
            if (mp::is_zero(a) || a.e <= te) {
              Num<NL> f = mp::from_pow2<NL>(te);
              f.neg = mp::is_zero(a) ? 0u : a.neg;
              smem_put<NL>(Us, pk_u(w, r, r), f);
            }
          } else if (mp::is_zero(a) || a.neg) {
            bad = 1;
          }
        }
      }
    } else if (warp == 16) {
      // the common diagonal of the unpivoted rows of G': tau_{k+1} = mu_k tau_k (only row k+1's slot is kept). A warp of
      // its own: on an elimination warp this multiplication (1500 cycles with the call) was added to the longest path
      // of every step.
      if (lane == 0) smem_put<NL>(Gs, pk_g(k + 1, k + 1), nmul(mu, smem_get<NL>(Gs, pk_g(k, k))));
    } else {
      // rows r = k+2+i, i in [0,R): row i has cnt_i = w-1-i active columns
      const int R = w - (k + 2);
      if (CL == 1) {
        // rows i and R-1-i are paired so that every pair has the same number of elements
        const int C = 2 * (w - 1) - (R - 1);  // cnt_i + cnt_{R-1-i}
        const int npair = R >> 1;
        const int total = npair * C + ((R & 1) ? (w - 1 - (R >> 1)) : 0);
        for (int idx = tid - 32; idx < total; idx += PANEL_ELIM) {
          int i, qp;
          if (idx < npair * C) {
            int pr = idx / C, j = idx % C, c0 = w - 1 - pr;
            if (j < c0) i = pr, qp = j; else i = R - 1 - pr, qp = j - c0;
          } else {
            i = R >> 1, qp = idx - npair * C;
          }
          panel_elim<NL>(Us, Gs, w, k, k + 2 + i, qp, mu, ek);
        }
      } else {
        // own rows only: i = i0 + j CL; w - 1 slots per row (at most one round of the 480 threads for CL >= 2, w = 32)
        const int i0 = ((crank - (k + 2)) % CL + CL) % CL;
        const int Rm = R > i0 ? (R - i0 + CL - 1) / CL : 0;
        for (int idx = tid - 32; idx < Rm * (w - 1); idx += PANEL_ELIM) {
          const int j = idx / (w - 1), qp = idx % (w - 1), i = i0 + j * CL;
          if (qp < w - 1 - i) panel_elim<NL>(Us, Gs, w, k, k + 2 + i, qp, mu, ek);
        }
      }
    }
    if (dbg && lane == 0 && warp == 1) tdbg[3][k] = clock64();
    if (CL > 1) {
      // every CTA but the owner PULLS the finished row k+1 (G part: columns 0..k, U part: columns k+1..w-1; the diagonal
      // slot of G - tau - is computed by every CTA itself) and the failure flag out of the owner's shared memory: one
      // 8-byte DSMEM load per thread (a push by the owner's warp 0 - 72 remote stores per lane to 7 destinations on the
      // pivot chain - measured 3.6 us per step). The owner never touches the row again, so no second cluster barrier.
      cluster.sync();
      const int r = k + 1, owner = r % CL;
      if (crank != owner) {
        const int g0 = pk_g(r, 0) * (NL + 2), gw = (k + 1) * (NL + 2);
        const int u0 = pk_u(w, r, r) * (NL + 2), uw = (w - r) * (NL + 2);
        const uint32_t* rG = cluster.map_shared_rank(Gs, owner);
        const uint32_t* rU = cluster.map_shared_rank(Us, owner);
        // (NL + 2 is even and every number starts at a multiple of NL + 2 words: 8-byte alignment)
        for (int i = 2 * tid; i < gw; i += 2 * nthr)
          *reinterpret_cast<uint2*>(Gs + g0 + i) = *reinterpret_cast<const uint2*>(rG + g0 + i);
        for (int i = 2 * tid; i < uw; i += 2 * nthr)
          *reinterpret_cast<uint2*>(Us + u0 + i) = *reinterpret_cast<const uint2*>(rU + u0 + i);
        if (tid == nthr - 1 && *cluster.map_shared_rank(&bad, owner)) bad = 1;
      }
    }
    __syncthreads();
  }
  if (dbg && tid == 0) tdbg[0][w - 1] = clock64();
  if (CL > 1) cluster.sync();  // nobody leaves while its rows may still be read
  // f_k = rsqrt(|tau_k d'_k|), one row per lane. Signed mode: the true pivot d_k = d'_k / tau_k has the sign
  // sigma_k = sign(tau_k d'_k); row k of L^-1 (L = U^T) is G'_k * sign(tau_k) f_k, row k of U is R'_k * sign(d'_k) f_k.
  __shared__ uint32_t sgn_tau[PANEL_W], sgn_d[PANEL_W];
  if (warp == 0 && !bad && lane < w) {
    Num<NL> tau = smem_get<NL>(Gs, pk_g(lane, lane)), dk = smem_get<NL>(Us, pk_u(w, lane, lane));
    sgn_tau[lane] = sig ? tau.neg : 0u;
    sgn_d[lane] = sig ? dk.neg : 0u;
    Num<NL> prod = nmul(tau, dk);
    if (sig) {
      if (crank == 0) sig[(int64_t)b * sig_ld + lane] = (int)prod.neg;
      prod.neg = 0u;
    }
    Num<NL> f, root = nsqrt_rsqrt(prod, f);
    (void)root;
    smem_put<NL>(Fs, lane, f);
  }
  __syncthreads();
  if (!bad) {
    for (int idx = tid + crank * nthr; idx < w * w; idx += nthr * CL) {
      int r = idx / w, c = idx % w;
      Num<NL> f = smem_get<NL>(Fs, r);
      if (write_u) {
        Num<NL> u = r <= c ? nmul(smem_get<NL>(Us, pk_u(w, r, c)), f) : mp::zero<NL>();
        if (sgn_d[r] && !mp::is_zero(u)) u.neg ^= 1u;
        stm<NL>(A, oa + (int64_t)r * ldA + c, u);
      }
      Num<NL> g = c <= r ? nmul(smem_get<NL>(Gs, pk_g(r, c)), f) : mp::zero<NL>();
      if (sgn_tau[r] && !mp::is_zero(g)) g.neg ^= 1u;
      stm<NL>(Lm, ol + (int64_t)r * ldL + c, g);
    }
  }
  if (tid == 0 && crank == 0) status[b] = bad;
  if (dbg && blockIdx.x == 0 && tid == 0) {
    long long tend = clock64();
    printf("panel dbg (w=%d CL=%d): step: warp0 start->next | warp1 work | warp15 start\n", w, CL);
    for (int k = 0; k + 1 < w; k++)
      printf("  k=%2d  step %6lld  w1 busy %6lld  w15 lag %6lld\n", k, tdbg[0][k + 1] - tdbg[0][k], tdbg[3][k] - tdbg[1][k],
             tdbg[2][k] - tdbg[0][k]);
    printf("  tail %lld\n", tend - tdbg[0][w - 1]);
  }
}
int panel_width(int nl) { (void)nl; return PANEL_W; }
// CLRSDP_PANEL_CLUSTER (measuring aid, read per launch so that tests can switch it): unset = one CTA per matrix,
// 0 = as many CTAs per matrix as keep the grid within the SMs, n > 0 = clusters of n CTAs.
static int panel_cluster_env() {
  const char* e = getenv("CLRSDP_PANEL_CLUSTER");
  return e ? atoi(e) : -1;
}
void panel_factor(Ctx& ctx, int nl, const MatBatch& A, const MatBatch& Linv, bool write_u, int* d_status, int* d_sig,
                  int sig_ld) {
  if (A.n > panel_width(nl)) throw SolverError(-1, "panel_factor: block larger than the panel width");
  // CTAs per matrix: as many as keep the grid within the SMs (a power of two up to the portable cluster size 8)
  int CL = 1;
  if (panel_cluster_env() == 0)
    while (CL < 8 && (int64_t)A.batch * CL * 2 <= ctx.sm_count) CL *= 2;
  if (A.n < 8) CL = 1;
  if (panel_cluster_env() > 0) CL = panel_cluster_env();
  static int panel_dbg = getenv("CLRSDP_PANEL_DEBUG") ? atoi(getenv("CLRSDP_PANEL_DEBUG")) : 0;
  DISPATCH_NL(nl, {
    size_t words = ((size_t)A.n * (A.n + 1) + A.n) * (NL + 2);
    // (function attributes belong to the DEVICE that is current: a process-wide "done" flag left every device but the first
    // of a multi-device handle without the opt-in, and its first panel above 48 KB failed to launch)
    static std::atomic<bool> attr[MAX_DEVICES][17];
    if (!attr[ctx.device % MAX_DEVICES][NL].load()) {
      CLR_CUDA(cudaFuncSetAttribute(panel_factor_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      attr[ctx.device % MAX_DEVICES][NL].store(true);
    }
    std::string nm = "panel_factor_n" + std::to_string(A.n);
    int tk = ctx.begin(nm.c_str());
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(A.batch * CL)), cfg.blockDim = dim3(PANEL_THREADS);
    cfg.dynamicSmemBytes = words * sizeof(uint32_t), cfg.stream = ctx.stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)CL, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    CLR_CUDA(cudaLaunchKernelEx(&cfg, panel_factor_kernel<NL>, A.t, A.d_off, A.shift, A.stride(), Linv.t, Linv.d_off,
                                Linv.shift, Linv.stride(), A.n, write_u ? 1 : 0, d_status, d_sig, sig_ld, CL, panel_dbg));
    ctx.end(tk);
  });
}

static dim3 tri_block(int n) {
  int cx = std::min(((n + 31) / 32) * 32, 256);
  int p = std::max(1, TRI_THREADS / cx);
  return dim3(cx, p, 1);
}

void chol_upper(Ctx& ctx, int nl, const MatBatch& A, const MatBatch& U, mp::Tensor rdiag, int* d_status) {
  dim3 blk = tri_block(A.n);
  DISPATCH_NL(nl, {
    std::string nm = "chol_n" + std::to_string(A.n);
    int tk = ctx.begin(nm.c_str());
    if (U.stride() != U.n && A.t.w != U.t.w) throw SolverError(-1, "chol_upper: sub-blocks must be factored in place");
    chol_kernel<NL><<<A.batch, blk, (NL + 2) * sizeof(uint32_t), ctx.stream>>>(
        A.t, A.d_off, A.shift, A.stride(), U.t, U.d_off, U.shift, U.stride(), rdiag, A.n, d_status);
    ctx.end(tk);
  });
}
void tri_inverse(Ctx& ctx, int nl, const MatBatch& U, mp::Tensor rdiag, const MatBatch& V, const MatBatch* Linv) {
  dim3 blk = tri_block(U.n);
  DISPATCH_NL(nl, {
    std::string nm = "trinv_n" + std::to_string(U.n);
    int tk = ctx.begin(nm.c_str());
    if (V.stride() != U.stride() || (Linv && Linv->stride() != U.stride()))
      throw SolverError(-1, "tri_inverse: operands must share the leading dimension");
    trinv_kernel<NL><<<U.batch, blk, 0, ctx.stream>>>(U.t, U.d_off, U.shift, rdiag, V.t, V.d_off, V.shift,
                                                      Linv ? Linv->t : V.t, Linv ? Linv->d_off : nullptr,
                                                      Linv ? Linv->shift : 0, U.stride(), U.n);
    ctx.end(tk);
  });
}

// =========================================================================================================
// smallest eigenvalue of a symmetric matrix: Householder tridiagonalisation, then multisection with Sturm
// counts + Newton polish on the characteristic polynomial
// =========================================================================================================
// number of sign changes in the Sturm sequence p_0..p_n of T - x I, stopped at the first one
// (returns true iff at least one eigenvalue is < x)
template <int NL>
__device__ bool any_eig_below(const uint32_t* dsm, const uint32_t* e2sm, int n, const Num<NL>& x) {
  Num<NL> p2 = mp::one<NL>();
  Num<NL> p1 = nsub(smem_get<NL>(dsm, 0), x);
  if (mp::is_zero(p1) || p1.neg) return true;
  for (int i = 1; i < n; i++) {
    Num<NL> a = nsub(smem_get<NL>(dsm, i), x);
    Num<NL> p = nsub(nmul(a, p1), nmul(smem_get<NL>(e2sm, i - 1), p2));
    if (mp::is_zero(p) || p.neg) return true;  // all previous p_i > 0, so this is the first sign change
    p2 = p1;
    p1 = p;
    // rescale to keep exponents bounded
    if (p1.e > (1 << 20) || p1.e < -(1 << 20)) {
      p2.e -= p1.e;
      p1.e = 0;
    }
  }
  return false;
}

template <int NL>
__global__ void __launch_bounds__(512)
lambda_min_kernel(mp::Tensor W, const int64_t* __restrict__ offW, int n, mp::Tensor out, const int* __restrict__ out_index,
                  const int* __restrict__ only_if) {
  extern __shared__ uint32_t sm[];
  // smem layout (in Num slots of NL+2 words): v[n], q[n], d[n], e2[n], eabs[n], red[33], misc[8], partial[P*CX]
  const int CX = blockDim.x, P = blockDim.y;
  uint32_t* vsm = sm;
  uint32_t* qsm = vsm + (size_t)n * (NL + 2);
  uint32_t* dsm = qsm + (size_t)n * (NL + 2);
  uint32_t* e2sm = dsm + (size_t)n * (NL + 2);
  uint32_t* easm = e2sm + (size_t)n * (NL + 2);
  uint32_t* red = easm + (size_t)n * (NL + 2);
  uint32_t* misc = red + 33 * (NL + 2);
  uint32_t* part_sm = misc + 8 * (NL + 2);
  __shared__ int flag, first_t;
  const int b = blockIdx.x, c = threadIdx.x, part = threadIdx.y;
  const int tid = part * CX + c, nthr = CX * P;
  if (only_if && !only_if[b]) return;  // already solved by the mixed-precision kernel
  const int64_t ow = offW[b];
  const int64_t out_at = out_index ? out_index[b] : b;
  auto Wat = [&](int r, int cc) { return ow + (int64_t)r * n + cc; };

  if (n == 1) {
    if (tid == 0) stm<NL>(out, out_at, ldm<NL>(W, ow));
    return;
  }
  // ---- Householder tridiagonalisation (symmetric full storage, both triangles kept up to date) ----
  for (int k = 0; k + 2 < n; k++) {
    Num<NL> xc = mp::zero<NL>();
    if (part == 0 && c > k && c < n) xc = ldm<NL>(W, Wat(k, c));
    Num<NL> sq = (part == 0 && c > k + 1 && c < n) ? nmul(xc, xc) : mp::zero<NL>();
    Num<NL> tail = block_reduce<NL, RED_ADD>(sq, red);
    if (tid == 0) {
      Num<NL> x1 = ldm<NL>(W, Wat(k, k + 1));
      if (mp::is_zero(tail)) {
        flag = 1;  // column already tridiagonal
        smem_put<NL>(e2sm, k, nmul(x1, x1));
        smem_put<NL>(easm, k, mp::fabs(x1));
      } else {
        flag = 0;
        Num<NL> sigma = nadd(tail, nmul(x1, x1));
        Num<NL> alpha = nsqrt(sigma);
        if (!x1.neg) alpha = mp::neg(alpha);
        Num<NL> beta = mp::mul_2exp(nsub(sigma, nmul(alpha, x1)), 1);  // v^T v
        Num<NL> binv = nrecip(beta);
        smem_put<NL>(misc, 0, alpha);
        smem_put<NL>(misc, 1, binv);
        smem_put<NL>(e2sm, k, sigma);  // alpha^2
        smem_put<NL>(easm, k, mp::fabs(alpha));
      }
    }
    __syncthreads();
    if (flag) continue;
    if (part == 0 && c > k && c < n) {
      Num<NL> v = xc;
      if (c == k + 1) v = nsub(v, smem_get<NL>(misc, 0));
      smem_put<NL>(vsm, c, v);
    }
    __syncthreads();
    // p = 2/beta * A v  (restricted to the trailing block), partial sums over row parts
    if (c > k && c < n) {
      Num<NL> acc = mp::zero<NL>();
      for (int j = k + 1 + part; j < n; j += P) acc = nadd(acc, nmul(ldm<NL>(W, Wat(j, c)), smem_get<NL>(vsm, j)));
      smem_put<NL>(part_sm, part * CX + c, acc);
    }
    __syncthreads();
    Num<NL> pc = mp::zero<NL>(), vc = mp::zero<NL>();
    if (part == 0 && c > k && c < n) {
      for (int q = 0; q < P; q++) pc = nadd(pc, smem_get<NL>(part_sm, q * CX + c));
      pc = mp::mul_2exp(nmul(pc, smem_get<NL>(misc, 1)), 1);
      vc = smem_get<NL>(vsm, c);
    }
    Num<NL> vp = block_reduce<NL, RED_ADD>(nmul(vc, pc), red);
    if (part == 0 && c > k && c < n) {
      Num<NL> Kc = nmul(vp, smem_get<NL>(misc, 1));  // v^T p / beta
      smem_put<NL>(qsm, c, nsub(pc, nmul(Kc, vc)));
    }
    __syncthreads();
    if (c > k && c < n) {
      Num<NL> qc = smem_get<NL>(qsm, c), vcc = smem_get<NL>(vsm, c);
      for (int j = k + 1 + part; j < n; j += P) {
        Num<NL> upd = nadd(nmul(smem_get<NL>(vsm, j), qc), nmul(smem_get<NL>(qsm, j), vcc));
        int64_t at = Wat(j, c);
        stm<NL>(W, at, nsub(ldm<NL>(W, at), upd));
      }
    }
    __syncthreads();
  }
  // diagonal and squared sub-diagonal of the tridiagonal matrix
  for (int i = tid; i < n; i += nthr) smem_put<NL>(dsm, i, ldm<NL>(W, Wat(i, i)));
  if (tid == 0) {
    Num<NL> x1 = ldm<NL>(W, Wat(n - 2, n - 1));
    smem_put<NL>(e2sm, n - 2, nmul(x1, x1));
    smem_put<NL>(easm, n - 2, mp::fabs(x1));
  }
  __syncthreads();
  // ---- Gershgorin interval (thread 0) ----
  if (tid == 0) {
    Num<NL> lo = mp::zero<NL>(), hi = mp::zero<NL>(), nrm = mp::zero<NL>();
    for (int i = 0; i < n; i++) {
      Num<NL> rad = mp::zero<NL>();
      if (i > 0) rad = nadd(rad, smem_get<NL>(easm, i - 1));
      if (i + 1 < n) rad = nadd(rad, smem_get<NL>(easm, i));
      Num<NL> di = smem_get<NL>(dsm, i);
      Num<NL> a1 = nsub(di, rad), a2 = nadd(di, rad);
      if (i == 0 || ncmp(a1, lo) < 0) lo = a1;
      if (i == 0 || ncmp(a2, hi) > 0) hi = a2;
    }
    nrm = mp::cmp_abs(lo, hi) > 0 ? mp::fabs(lo) : mp::fabs(hi);
    Num<NL> pad = mp::mul_2exp(nrm, -20);
    if (mp::is_zero(nrm)) pad = mp::from_pow2<NL>(-1000);
    smem_put<NL>(misc, 2, nsub(lo, pad));
    smem_put<NL>(misc, 3, nadd(hi, pad));
    smem_put<NL>(misc, 4, nrm);
    flag = 0;
  }
  __syncthreads();
  int Tp = 1, lg = 0;
  while (Tp * 2 <= nthr) Tp *= 2, lg++;
  bool newton_allowed = true;
  // invariant: no eigenvalue below lo, at least one below hi
  for (int round = 0; round < 200; round++) {
    Num<NL> lo = smem_get<NL>(misc, 2), hi = smem_get<NL>(misc, 3);
    Num<NL> width = nsub(hi, lo);
    Num<NL> scale = mp::cmp_abs(lo, hi) > 0 ? mp::fabs(lo) : mp::fabs(hi);
    // converged when the interval is below 2^-(p-3) relative (or cannot shrink any more)
    if (mp::is_zero(width) || width.neg || width.e <= scale.e - (32 * NL - 3)) break;
    bool narrow_enough = width.e <= scale.e - 40;
    if (narrow_enough && newton_allowed) {
      // ---- Newton from the left end on p_n(x) = det(T - xI): monotone, quadratic for a simple root ----
      if (tid == 0) {
        Num<NL> x = lo;
        int ok = 0;
        for (int it = 0; it < 10; it++) {
          Num<NL> p2 = mp::one<NL>(), dp2 = mp::zero<NL>();
          Num<NL> p1 = nsub(smem_get<NL>(dsm, 0), x), dp1 = mp::neg(mp::one<NL>());
          for (int i = 1; i < n; i++) {
            Num<NL> a = nsub(smem_get<NL>(dsm, i), x), e2 = smem_get<NL>(e2sm, i - 1);
            Num<NL> p = nsub(nmul(a, p1), nmul(e2, p2));
            Num<NL> dp = nsub(nsub(nmul(a, dp1), p1), nmul(e2, dp2));
            p2 = p1, dp2 = dp1, p1 = p, dp1 = dp;
            if (p1.e > (1 << 20) || p1.e < -(1 << 20)) {
              int s = p1.e;
              p1.e -= s;
              if (!mp::is_zero(p2)) p2.e -= s;
              if (!mp::is_zero(dp1)) dp1.e -= s;
              if (!mp::is_zero(dp2)) dp2.e -= s;
            }
          }
          if (mp::is_zero(p1)) { ok = 1; break; }
          if (mp::is_zero(dp1)) break;
          Num<NL> step = ndiv(p1, dp1);  // x_new = x - p/p'
          Num<NL> xn = nsub(x, step);
          if (ncmp(xn, x) < 0 || ncmp(xn, hi) > 0) break;  // left the bracket: not the simple-root regime
          bool tiny = mp::is_zero(step) || step.e <= xn.e - (32 * NL - 6);
          x = xn;
          if (tiny) { ok = 1; break; }
        }
        if (ok) {
          // verify: there must be an eigenvalue below x(1 + 2^-(p-8))
          Num<NL> eps = mp::mul_2exp(mp::fabs(x), -(32 * NL - 8));
          if (mp::is_zero(eps)) eps = mp::mul_2exp(smem_get<NL>(misc, 4), -(32 * NL));
          Num<NL> xu = nadd(x, eps), xl = nsub(x, eps);
          if (any_eig_below<NL>(dsm, e2sm, n, xu) && !any_eig_below<NL>(dsm, e2sm, n, xl)) {
            smem_put<NL>(misc, 2, x);
            smem_put<NL>(misc, 3, x);
            flag = 2;
          } else {
            flag = 1;
          }
        } else {
          flag = 1;
        }
      }
      __syncthreads();
      int f = flag;
      __syncthreads();
      if (f == 2) break;
      newton_allowed = false;
      continue;
    }
    // ---- one multisection round: Tp probes, keep the first sub-interval that contains an eigenvalue ----
    if (tid == 0) first_t = Tp - 1;
    __syncthreads();
    Num<NL> w = mp::mul_2exp(width, -lg);
    Num<NL> xt = hi;
    if (tid < Tp - 1) {
      xt = nadd(lo, nmul(w, mp::from_int<NL>(tid + 1)));
      if (any_eig_below<NL>(dsm, e2sm, n, xt)) atomicMin(&first_t, tid);
    }
    __syncthreads();
    int ft = first_t;
    if (tid == ft) smem_put<NL>(misc, 3, xt);
    if (ft > 0 && tid == ft - 1) smem_put<NL>(misc, 2, xt);
    __syncthreads();
  }
  if (tid == 0) {
    Num<NL> lo = smem_get<NL>(misc, 2), hi = smem_get<NL>(misc, 3);
    stm<NL>(out, out_at, mp::mul_2exp(nadd(lo, hi), -1));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Mixed-precision smallest eigenvalue (the production path of compute_step_length, MPMP.jl:1857-1870).
//
// The tridiagonalisation above costs 4/3 n^3 multiprecision operations on a dependency chain. Here the n^3
// work is done in FP64 and only O(n^2) work per refinement step in multiprecision:
//   1. Wd = fp64(W * 2^-emax); Householder tridiagonalisation of Wd in shared memory, Sturm multisection
//      for a lower bound lo of lambda_min(Wd) (absolute accuracy ~2^-45 |W|);
//   2. Cholesky (FP64) of Wd - sigma I, sigma = lo - 2^-36 |W|  (positive definite by construction), three
//      inverse iterations for an FP64 eigenvector v;
//   3. multiprecision refinement: r = W v - rho v with rho = v'Wv / v'v at full precision; the correction
//      (W - sigma)^-1 r is computed in FP64 with the Cholesky factor (r scaled by its exponent), projected
//      against v and subtracted. The error contracts by ~2^-36 |W| / gap per step; iterate until the
//      residual is at the rounding level of the working precision, so rho is an eigenvalue to 2^-(p-12) |W|
//      (and quadratically better in the residual while the gap is resolved).
// A matrix whose refinement does not contract (clustered smallest eigenvalues closer than ~2^-30 |W|), or whose
// FP64 factorisation breaks down, is flagged and handled by lambda_min_kernel (flags[b] = 1).
// ---------------------------------------------------------------------------------------------------------
constexpr int LMX_THREADS = 256;
// -DCLRSDP_LMX_TIMING: per-phase clock64 deltas of block 0 (measuring aid)
#ifdef CLRSDP_LMX_TIMING
#define LMX_T(i) lmx_t[i] = clock64()
#define LMX_COUNT_ITER lmx_iters++
#else
#define LMX_T(i)
#define LMX_COUNT_ITER
#endif
constexpr int LMX_NT = 5;  // vector slots per lane of warp 0: n <= 160
constexpr int LMX_MAX_ITERS = 64;

// x <- (Wd - sigma)^-1 x for a vector distributed over the lanes of one warp (element i = lane + 32 t).
// A holds the Cholesky factor L in both triangles (A[k][i] = A[i][k] = L[max][min]); rd[k] = 1/L[k][k].
__device__ __forceinline__ void lmx_solve(const double* __restrict__ A, const double* __restrict__ rd, int n, int LD,
                                          int lane, double (&x)[LMX_NT]) {
#pragma unroll
  for (int t0 = 0; t0 < LMX_NT; t0++) {  // L y = x
    if (32 * t0 >= n) break;
    for (int kl = 0; kl < 32; kl++) {
      const int k = 32 * t0 + kl;
      if (k >= n) break;
      const double yk = __shfl_sync(0xffffffffu, x[t0], kl) * rd[k];
      if (lane == kl) x[t0] = yk;
      const double* row = A + (size_t)k * LD;
#pragma unroll
      for (int t = t0; t < LMX_NT; t++) {
        const int i = lane + 32 * t;
        if (i > k && i < n) x[t] -= row[i] * yk;
      }
    }
  }
#pragma unroll
  for (int t0 = LMX_NT - 1; t0 >= 0; t0--) {  // L^T z = y
    if (32 * t0 >= n) continue;
    for (int kl = 31; kl >= 0; kl--) {
      const int k = 32 * t0 + kl;
      if (k >= n) continue;
      const double zk = __shfl_sync(0xffffffffu, x[t0], kl) * rd[k];
      if (lane == kl) x[t0] = zk;
      const double* row = A + (size_t)k * LD;
#pragma unroll
      for (int t = 0; t <= t0; t++) {
        const int i = lane + 32 * t;
        if (i < k) x[t] -= row[i] * zk;
      }
    }
  }
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int NL>
__global__ void __launch_bounds__(LMX_THREADS)
lambda_min_mixed_kernel(mp::Tensor W, const int64_t* __restrict__ offW, int n, mp::Tensor out,
                        const int* __restrict__ out_index, int* __restrict__ flags) {
  extern __shared__ double smd[];
  const int LD = n | 1;
  double* A = smd;
  double* dT = A + (size_t)n * LD;
  double* eT = dT + n;
  double* vv = eT + n;
  double* ww = vv + n;
  double* part = ww + n;
  double* sc = part + LMX_THREADS;  // 16 scalars
  uint32_t* vsm = reinterpret_cast<uint32_t*>(sc + 16);        // mp vector v
  uint32_t* psm = vsm + (size_t)n * (NL + 2);                  // mp partial sums [LMX_THREADS]
  uint32_t* red = psm + (size_t)LMX_THREADS * (NL + 2);        // block_reduce scratch [33]
  __shared__ int s_int[LMX_THREADS / 32 + 2];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ow = offW[b];
  const int64_t out_at = out_index ? out_index[b] : b;
  const uint32_t* hdr = W.w + (size_t)NL * W.n;

#ifdef CLRSDP_LMX_TIMING
  long long lmx_t[6] = {0, 0, 0, 0, 0, 0};
  int lmx_iters = 0;
#endif
  if (n == 1) {
    if (tid == 0) {
      stm<NL>(out, out_at, ldm<NL>(W, ow));
      flags[b] = 0;
    }
    return;
  }
  // ---- common exponent ----
  int emax = mp::EXP_ZERO;
  for (int idx = tid; idx < n * n; idx += LMX_THREADS) emax = max(emax, ((int32_t)hdr[ow + idx]) >> 1);
  emax = __reduce_max_sync(0xffffffffu, emax);
  if (lane == 0) s_int[warp] = emax;
  __syncthreads();
  emax = s_int[0];
#pragma unroll
  for (int q = 1; q < LMX_THREADS / 32; q++) emax = max(emax, s_int[q]);
  __syncthreads();
  if (emax == mp::EXP_ZERO) {  // the zero matrix
    if (tid == 0) {
      stm<NL>(out, out_at, mp::zero<NL>());
      flags[b] = 0;
    }
    return;
  }
  auto load_fp64 = [&]() {
    for (int idx = tid; idx < n * n; idx += LMX_THREADS) {
      const int r = idx / n, c = idx - r * n;
      Num<NL> x = ldm<NL>(W, ow + idx);
      double v = 0.0;
      if (!mp::is_zero(x) && x.e - emax > -1000) {
        v = ldexp(mp::mant_to_double(x), x.e - emax);
        if (x.neg) v = -v;
      }
      A[(size_t)r * LD + c] = v;
    }
  };
  load_fp64();
  __syncthreads();
  // symmetrise the FP64 copy (the callers pass symmetric matrices; this makes the copy exactly symmetric)
  for (int idx = tid; idx < n * n; idx += LMX_THREADS) {
    const int r = idx / n, c = idx - r * n;
    if (r < c) {
      double s = 0.5 * (A[(size_t)r * LD + c] + A[(size_t)c * LD + r]);
      A[(size_t)r * LD + c] = s;
      A[(size_t)c * LD + r] = s;
    }
  }
  __syncthreads();

  LMX_T(0);
  // ---- 1. Householder tridiagonalisation in FP64 (full symmetric storage) ----
  const int R = ((n + 31) / 32) * 32;
  const int parts = LMX_THREADS / R > 0 ? LMX_THREADS / R : 1;
  const int row = tid % R, prt = tid / R;
  for (int k = 0; k + 2 < n; k++) {
    if (warp == 0) {
      const double* rk = A + (size_t)k * LD;
      double sg = 0.0;
      for (int i = k + 2 + lane; i < n; i += 32) sg += rk[i] * rk[i];
      sg = warp_sum(sg);
      const double x1 = rk[k + 1];
      if (sg == 0.0) {
        if (lane == 0) {
          eT[k] = x1;
          dT[k] = rk[k];
          sc[2 + (k & 1)] = 1.0;  // skip flag
        }
      } else {
        const double alpha = -copysign(sqrt(x1 * x1 + sg), x1);
        const double v1 = x1 - alpha;
        const double binv = 2.0 / (sg + v1 * v1);
        for (int i = k + 2 + lane; i < n; i += 32) vv[i] = rk[i];
        if (lane == 0) {
          vv[k + 1] = v1;
          eT[k] = alpha;
          dT[k] = rk[k];
          sc[0] = binv;
          sc[2 + (k & 1)] = 0.0;
        }
      }
    }
    __syncthreads();
    if (sc[2 + (k & 1)] != 0.0) continue;
    {
      double acc = 0.0;
      if (row > k && row < n && prt < parts)
        for (int j = k + 1 + prt; j < n; j += parts) acc += A[(size_t)j * LD + row] * vv[j];
      part[tid] = acc;
    }
    __syncthreads();
    if (warp == 0) {
      const double binv = sc[0];
      double pl[LMX_NT], vp = 0.0;
#pragma unroll
      for (int t = 0; t < LMX_NT; t++) {
        const int i = lane + 32 * t;
        pl[t] = 0.0;
        if (i > k && i < n) {
          double s = 0.0;
          for (int q = 0; q < parts; q++) s += part[q * R + i];
          pl[t] = s * binv;
          vp += pl[t] * vv[i];
        }
      }
      vp = warp_sum(vp);
      const double Kc = 0.5 * vp * binv;
#pragma unroll
      for (int t = 0; t < LMX_NT; t++) {
        const int i = lane + 32 * t;
        if (i > k && i < n) ww[i] = pl[t] - Kc * vv[i];
      }
    }
    __syncthreads();
    {
      const int m = n - k - 1;
      for (int idx = tid; idx < m * m; idx += LMX_THREADS) {
        const int i = k + 1 + idx / m, j = k + 1 + idx % m;
        A[(size_t)i * LD + j] -= vv[i] * ww[j] + ww[i] * vv[j];
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    dT[n - 2] = A[(size_t)(n - 2) * LD + (n - 2)];
    dT[n - 1] = A[(size_t)(n - 1) * LD + (n - 1)];
    eT[n - 2] = A[(size_t)(n - 2) * LD + (n - 1)];
    eT[n - 1] = 0.0;
  }
  __syncthreads();
  LMX_T(1);
  // ---- Gershgorin bounds of T ----
  if (warp == 0) {
    double lo = 1e300, hi = -1e300;
    for (int i = lane; i < n; i += 32) {
      double rad = (i > 0 ? fabs(eT[i - 1]) : 0.0) + (i + 1 < n ? fabs(eT[i]) : 0.0);
      lo = fmin(lo, dT[i] - rad);
      hi = fmax(hi, dT[i] + rad);
    }
    lo = -warp_max(-lo);
    hi = warp_max(hi);
    if (lane == 0) {
      double nrm = fmax(fabs(lo), fabs(hi));
      sc[4] = lo - 1e-9 * nrm - 1e-300;
      sc[5] = hi + 1e-9 * nrm + 1e-300;
      sc[6] = nrm;
    }
  }
  __syncthreads();
  const double nrm = sc[6];
  // ---- multisection with Sturm counts: invariant: no eigenvalue below lo, at least one below hi ----
  for (int round = 0; round < 9; round++) {
    const double lo = sc[4], hi = sc[5];
    if (!(hi - lo > nrm * 0x1p-50)) break;
    if (tid == 0) s_int[0] = LMX_THREADS - 1;
    __syncthreads();
    const double wdt = (hi - lo) * (1.0 / LMX_THREADS);
    double xt = hi;
    if (tid < LMX_THREADS - 1) {
      xt = lo + wdt * (double)(tid + 1);
      const double pivmin = 1e-290;
      double q = dT[0] - xt;
      bool below = q < 0.0;
      for (int i = 1; i < n && !below; i++) {
        if (fabs(q) < pivmin) q = pivmin;  // q >= 0 here
        q = dT[i] - xt - eT[i - 1] * eT[i - 1] / q;
        below = q < 0.0;
      }
      if (below) atomicMin(&s_int[0], tid);
    }
    __syncthreads();
    const int ft = s_int[0];
    __syncthreads();
    if (tid == ft) sc[5] = xt;
    if (ft > 0 && tid == ft - 1) sc[4] = xt;
    __syncthreads();
  }
  const double lam0 = sc[4];
  const double sigma = lam0 - nrm * 0x1p-36;
  __syncthreads();
  LMX_T(2);
  // ---- 2. Cholesky of Wd - sigma I (factor kept in both triangles) ----
  load_fp64();
  __syncthreads();
  for (int i = tid; i < n; i += LMX_THREADS) A[(size_t)i * LD + i] -= sigma;
  if (tid == 0) s_int[1] = 0;
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += LMX_THREADS) {  // exact symmetry (see above)
    const int r = idx / n, c = idx - r * n;
    if (r < c) {
      double s = 0.5 * (A[(size_t)r * LD + c] + A[(size_t)c * LD + r]);
      A[(size_t)r * LD + c] = s;
      A[(size_t)c * LD + r] = s;
    }
  }
  __syncthreads();
  for (int k = 0; k < n; k++) {
    const double piv = A[(size_t)k * LD + k];
    if (!(piv > 0.0)) {  // uniform: every thread reads the same value
      if (tid == 0) s_int[1] = 1;
      break;
    }
    const double rs = rsqrt(piv);
    __syncthreads();  // everyone has read the pivot
    for (int i = k + tid; i < n; i += LMX_THREADS) {
      if (i == k) {
        A[(size_t)k * LD + k] = piv * rs;
        dT[k] = rs;
      } else {
        const double l = A[(size_t)k * LD + i] * rs;
        A[(size_t)k * LD + i] = l;
        A[(size_t)i * LD + k] = l;
      }
    }
    __syncthreads();
    const int m = n - k - 1;
    const double* rk = A + (size_t)k * LD;
    for (int idx = tid; idx < m * m; idx += LMX_THREADS) {
      const int i = k + 1 + idx / m, j = k + 1 + idx % m;
      A[(size_t)i * LD + j] -= rk[i] * rk[j];
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_int[1]) {
    if (tid == 0) flags[b] = 1;
    return;
  }
  LMX_T(3);
  // ---- FP64 inverse iteration (warp 0) ----
  if (warp == 0) {
    double x[LMX_NT];
#pragma unroll
    for (int t = 0; t < LMX_NT; t++) {
      const uint32_t i = (uint32_t)(lane + 32 * t);
      x[t] = (int)i < n ? (double)((i * 2654435761u >> 8) & 0xFFFFu) * (1.0 / 65536.0) - 0.5 : 0.0;
    }
    for (int it = 0; it < 4; it++) {
      lmx_solve(A, dT, n, LD, lane, x);
      double mx = 0.0;
#pragma unroll
      for (int t = 0; t < LMX_NT; t++) mx = fmax(mx, fabs(x[t]));
      mx = warp_max(mx);
      double s2 = 0.0;
#pragma unroll
      for (int t = 0; t < LMX_NT; t++) {
        x[t] = mx > 0.0 ? x[t] / mx : 0.0;
        s2 += x[t] * x[t];
      }
      s2 = warp_sum(s2);
      const double rn = s2 > 0.0 ? rsqrt(s2) : 0.0;
#pragma unroll
      for (int t = 0; t < LMX_NT; t++) x[t] *= rn;
    }
#pragma unroll
    for (int t = 0; t < LMX_NT; t++) {
      const int i = lane + 32 * t;
      if (i < n) smem_put<NL>(vsm, i, mp::from_double<NL>(x[t]));
    }
  }
  __syncthreads();
  LMX_T(4);
  // ---- 3. multiprecision refinement ----
  // exponent of |W| as an mp quantity: nrm * 2^emax
  int nrm_e;
  (void)frexp(nrm, &nrm_e);
  nrm_e += emax;
  const int target = nrm_e - (32 * NL - 12);
  int prev_rexp = 1 << 30, stalls = 0;
  bool done = false, failed = false;
  Num<NL> svv = mp::zero<NL>(), svw = mp::zero<NL>();
  const int nwv = (n + 31) >> 5;
  for (int iter = 0; iter < LMX_MAX_ITERS; iter++) {
    LMX_COUNT_ITER;
    // w = W v (W symmetric: column `row` is read as row-major W[j][row], coalesced over the threads)
    {
      Num<NL> acc = mp::zero<NL>();
      if (row < n && prt < parts)
        for (int j = prt; j < n; j += parts)
          acc = mp::add(acc, mp::mul(ldm<NL>(W, ow + (int64_t)j * n + row), smem_get<NL>(vsm, j)));
      smem_put<NL>(psm, tid, acc);
    }
    __syncthreads();
    Num<NL> wi = mp::zero<NL>(), vi = mp::zero<NL>();
    if (tid < n) {
      for (int q = 0; q < parts; q++) wi = nadd(wi, smem_get<NL>(psm, q * R + tid));
      vi = smem_get<NL>(vsm, tid);
    }
    // v'v and v'w: only the first ceil(n/32) warps hold entries; both sums go through the shuffles together
    {
      Num<NL> sa = mp::zero<NL>(), sb = mp::zero<NL>();
      if (tid < n) {
        sa = nmul(vi, vi);
        sb = nmul(vi, wi);
      }
      if (warp < nwv) {
#pragma unroll 1
        for (int o = 16; o; o >>= 1) {
          Num<NL> ta = shfl_xor_num(sa, o), tb = shfl_xor_num(sb, o);
          sa = nadd(sa, ta);
          sb = nadd(sb, tb);
        }
        if (lane == 0) {
          smem_put<NL>(red, 2 * warp, sa);
          smem_put<NL>(red, 2 * warp + 1, sb);
        }
      }
    }
    __syncthreads();
    svv = smem_get<NL>(red, 0);
    svw = smem_get<NL>(red, 1);
    for (int q = 1; q < nwv; q++) {
      svv = nadd(svv, smem_get<NL>(red, 2 * q));
      svw = nadd(svw, smem_get<NL>(red, 2 * q + 1));
    }
    if (mp::is_zero(svv)) {
      failed = true;
      break;
    }
    // scaled residual r' = (v'v) w - (v'w) v = (v'v) (W v - rho v): no division inside the loop (v'v = 1 + O(2^-52))
    Num<NL> ri = mp::zero<NL>();
    if (tid < n) ri = nmsm(svv, wi, svw, vi);
    int rexp = __reduce_max_sync(0xffffffffu, ri.e);
    __syncthreads();
    if (lane == 0) s_int[warp] = rexp;
    __syncthreads();
    rexp = s_int[0];
#pragma unroll
    for (int q = 1; q < LMX_THREADS / 32; q++) rexp = max(rexp, s_int[q]);
    if (rexp <= target) {
      done = true;
      break;
    }
    if (rexp > prev_rexp - 8) {
      if (++stalls >= 3) {
        failed = true;
        break;
      }
    }
    prev_rexp = rexp;
    // FP64 correction of the scaled residual
    if (tid < n) {
      double rd = 0.0;
      if (!mp::is_zero(ri) && ri.e - rexp > -1000) {
        rd = ldexp(mp::mant_to_double(ri), ri.e - rexp);
        if (ri.neg) rd = -rd;
      }
      ww[tid] = rd;
      vv[tid] = mp::to_double(vi);
    }
    __syncthreads();
    if (warp == 0) {
      double x[LMX_NT], vd[LMX_NT];
#pragma unroll
      for (int t = 0; t < LMX_NT; t++) {
        const int i = lane + 32 * t;
        x[t] = i < n ? ww[i] : 0.0;
        vd[t] = i < n ? vv[i] : 0.0;
      }
      lmx_solve(A, dT, n, LD, lane, x);
      double xv = 0.0, v2 = 0.0;
#pragma unroll
      for (int t = 0; t < LMX_NT; t++) xv += x[t] * vd[t], v2 += vd[t] * vd[t];
      xv = warp_sum(xv);
      v2 = warp_sum(v2);
      const double cf = v2 > 0.0 ? xv / v2 : 0.0;
#pragma unroll
      for (int t = 0; t < LMX_NT; t++) {
        const int i = lane + 32 * t;
        if (i < n) ww[i] = x[t] - cf * vd[t];
      }
    }
    __syncthreads();
    if (tid < n) {
      const double dl = ww[tid];
      if (dl != 0.0 && isfinite(dl)) {
        Num<NL> corr = mp::mul_2exp(mp::from_double<NL>(dl), rexp - emax);
        smem_put<NL>(vsm, tid, nsub(vi, corr));
      }
    }
    __syncthreads();
  }
  LMX_T(5);
#ifdef CLRSDP_LMX_TIMING
  if (tid == 0 && b == 0)
    printf("[lmx] n=%d cycles: tridiag %lld bisect %lld chol %lld invit %lld refine %lld (%d iterations)\n", n, lmx_t[1] - lmx_t[0],
           lmx_t[2] - lmx_t[1], lmx_t[3] - lmx_t[2], lmx_t[4] - lmx_t[3], lmx_t[5] - lmx_t[4], lmx_iters);
#endif
  if (tid == 0) {
    if (done && !failed) {
      stm<NL>(out, out_at, ndiv(svw, svv));  // Rayleigh quotient of the converged vector
      flags[b] = 0;
    } else {
      flags[b] = 1;
    }
  }
}

static size_t lmx_smem_bytes(int n, int NL) {
  const int LD = n | 1;
  return sizeof(double) * ((size_t)n * LD + 4 * (size_t)n + LMX_THREADS + 16) +
         sizeof(uint32_t) * ((size_t)n + LMX_THREADS + 33) * (NL + 2);
}
static bool lmx_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CLRSDP_LAMBDA_MIXED");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

void lambda_min(Ctx& ctx, int nl, const MatBatch& W, mp::Tensor out, const int* d_out_index, int* d_flags) {
  if (W.n > 512) throw SolverError(-1, "lambda_min: n > 512 not supported");
  int cx = std::min(1024, ((W.n + 31) / 32) * 32);
  int P = std::max(1, 512 / cx);
  dim3 blk(cx, P, 1);
  DISPATCH_NL(nl, {
    static std::atomic<bool> attr[MAX_DEVICES][17];  // per device, see panel_factor
    if (!attr[ctx.device % MAX_DEVICES][NL].load()) {
      CLR_CUDA(cudaFuncSetAttribute(lambda_min_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      CLR_CUDA(cudaFuncSetAttribute(lambda_min_mixed_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      attr[ctx.device % MAX_DEVICES][NL].store(true);
    }
    // production path: FP64 factorisation + multiprecision refinement; matrices it flags (and sizes that do not fit
    // its shared-memory working set) go through the all-multiprecision kernel
    const size_t mixed_bytes = lmx_smem_bytes(W.n, NL);
    const bool mixed = d_flags && lmx_enabled() && W.n <= 32 * LMX_NT && mixed_bytes <= 226 * 1024;
    if (mixed) {
      std::string nm = "lambda_min_mixed_n" + std::to_string(W.n);
      int tk = ctx.begin(nm.c_str());
      lambda_min_mixed_kernel<NL><<<W.batch, LMX_THREADS, mixed_bytes, ctx.stream>>>(W.t, W.d_off, W.n, out, d_out_index, d_flags);
      ctx.end(tk);
    }
    size_t words = (size_t)(5 * W.n + 33 + 8 + P * cx) * (NL + 2);
    std::string nm = (mixed ? "lambda_min_fallback_n" : "lambda_min_n") + std::to_string(W.n);
    int tk = ctx.begin(nm.c_str());
    lambda_min_kernel<NL><<<W.batch, blk, words * sizeof(uint32_t), ctx.stream>>>(W.t, W.d_off, W.n, out, d_out_index,
                                                                                   mixed ? d_flags : nullptr);
    ctx.end(tk);
  });
}

// =========================================================================================================
// elementwise kernels
// =========================================================================================================
template <int NL>
__global__ void lincomb_kernel(mp::Tensor out, int64_t oo, mp::Tensor a, int64_t ao, int sa, mp::Tensor b, int64_t bo,
                               int sb, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Num<NL> r = mp::zero<NL>();
    if (sa) {
      r = ldm<NL>(a, ao + i);
      if (sa < 0) r = mp::neg(r);
    }
    if (sb) {
      Num<NL> y = ldm<NL>(b, bo + i);
      r = sb > 0 ? mp::add(r, y) : mp::sub(r, y);
    }
    stm<NL>(out, oo + i, r);
  }
}
template <int NL>
__global__ void axpy_kernel(mp::Tensor y, int64_t yo, mp::Tensor x, int64_t xo, mp::Tensor scal, int slot, int64_t n) {
  Num<NL> s = ldm<NL>(scal, slot);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    stm<NL>(y, yo + i, mp::add(ldm<NL>(y, yo + i), mp::mul(s, ldm<NL>(x, xo + i))));
}
// per-block kernels: thread per element of a batch of n x n blocks
template <int NL>
__global__ void residual_R_kernel(mp::Tensor R, const int64_t* __restrict__ off, int batch, int n, mp::Tensor scal,
                                  int slot, mp::Tensor T1, mp::Tensor T2, int has2) {
  Num<NL> s = ldm<NL>(scal, slot);
  int64_t total = (int64_t)batch * n * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(idx / ((int64_t)n * n));
    int rem = (int)(idx % ((int64_t)n * n));
    int i = rem / n, j = rem % n;
    int64_t at = off[b] + rem;
    Num<NL> r = (i == j) ? s : mp::zero<NL>();
    r = mp::sub(r, ldm<NL>(T1, at));
    if (has2) r = mp::sub(r, ldm<NL>(T2, at));
    stm<NL>(R, at, r);
  }
}
template <int NL>
__global__ void symmetrize_kernel(mp::Tensor out, const int64_t* __restrict__ off, int batch, int n, mp::Tensor in) {
  int64_t total = (int64_t)batch * n * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(idx / ((int64_t)n * n));
    int rem = (int)(idx % ((int64_t)n * n));
    int i = rem / n, j = rem % n;
    if (i > j) continue;
    int64_t o = off[b];
    Num<NL> r = mp::mul_2exp(mp::add(ldm<NL>(in, o + (int64_t)i * n + j), ldm<NL>(in, o + (int64_t)j * n + i)), -1);
    stm<NL>(out, o + (int64_t)i * n + j, r);
    if (i != j) stm<NL>(out, o + (int64_t)j * n + i, r);
  }
}
template <int NL>
__global__ void set_identity_kernel(mp::Tensor M, const int64_t* __restrict__ off, int batch, int n, mp::Tensor scal,
                                    int slot) {
  Num<NL> s = ldm<NL>(scal, slot);
  int64_t total = (int64_t)batch * n * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(idx / ((int64_t)n * n));
    int rem = (int)(idx % ((int64_t)n * n));
    stm<NL>(M, off[b] + rem, (rem / n == rem % n) ? s : mp::zero<NL>());
  }
}
static int ew_grid(Ctx& ctx, int64_t n, int threads = 128) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, threads), (int64_t)ctx.sm_count * 32));
}
void ew_lincomb(Ctx& ctx, int nl, mp::Tensor out, int64_t oo, mp::Tensor a, int64_t ao, int sa, mp::Tensor b,
                int64_t bo, int sb, int64_t n) {
  if (n <= 0) return;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("ew_lincomb", (double)n * 4.0 * (NL + 1) * (1 + (sa != 0) + (sb != 0)));
    lincomb_kernel<NL><<<ew_grid(ctx, n), 128, 0, ctx.stream>>>(out, oo, a, ao, sa, b, bo, sb, n);
    ctx.end(tk);
  });
}
template <int NL>
__global__ void binary_kernel(int op, mp::Tensor c, mp::Tensor a, mp::Tensor b, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Num<NL> x = ldm<NL>(a, i), y = ldm<NL>(b, i), z;
    switch (op) {
      case '+': z = mp::add(x, y); break;
      case '-': z = mp::sub(x, y); break;
      case '*': z = mp::mul(x, y); break;
      case '/': z = mp::div(x, y); break;
      default: z = mp::sqrt(x); break;
    }
    stm<NL>(c, i, z);
  }
}
void ew_binary(Ctx& ctx, int nl, int op, mp::Tensor c, mp::Tensor a, mp::Tensor b, int64_t n) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("ew_binary");
    binary_kernel<NL><<<ew_grid(ctx, n), 128, 0, ctx.stream>>>(op, c, a, b, n);
    ctx.end(tk);
  });
}
template <int NL>
__global__ void mat_copy_kernel(mp::Tensor D, const int64_t* __restrict__ offD, int64_t shD, int ldD, mp::Tensor S,
                                const int64_t* __restrict__ offS, int64_t shS, int ldS, int batch, int n, int mode) {
  int64_t total = (int64_t)batch * n * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(idx / ((int64_t)n * n));
    int rem = (int)(idx % ((int64_t)n * n));
    int r = rem / n, c = rem % n;
    Num<NL> v = mp::zero<NL>();
    if (mode == 0 || (mode == 1 && r <= c)) v = ldm<NL>(S, offS[b] + shS + (int64_t)r * ldS + c);
    stm<NL>(D, offD[b] + shD + (int64_t)r * ldD + c, v);
  }
}
void mat_copy(Ctx& ctx, int nl, const MatBatch& dst, const MatBatch& src, bool upper_only) {
  int64_t total = (int64_t)dst.batch * dst.n * dst.n;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("mat_copy", (double)total * 4.0 * (NL + 1) * 2);
    mat_copy_kernel<NL><<<ew_grid(ctx, total), 128, 0, ctx.stream>>>(dst.t, dst.d_off, dst.shift, dst.stride(), src.t,
                                                                     src.d_off, src.shift, src.stride(), dst.batch,
                                                                     dst.n, upper_only ? 1 : 0);
    ctx.end(tk);
  });
}
// ---- symmetric equilibration by powers of two (exact) --------------------------------------------------------
// s[b*n + i] = ceil(exponent(A_ii) / 2): A' = D^-1 A D^-1 with D = diag(2^s) has its diagonal in [1/4, 1) (0 if A_ii <= 0:
// the factorisation then reports the matrix as not positive definite).
template <int NL>
__global__ void equil_exponents_kernel(mp::Tensor A, const int64_t* __restrict__ offA, int64_t shA, int ldA, int batch, int n,
                                       int* __restrict__ s) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * n) return;
  int b = idx / n, i = idx - b * n;
  uint32_t h = A.w[(size_t)NL * A.n + offA[b] + shA + (int64_t)i * ldA + i];
  int32_t e = ((int32_t)h) >> 1;
  s[idx] = (e == mp::EXP_ZERO) ? 0 : ((e + 1) >> 1);  // arithmetic shift: ceil(e / 2) for either sign
}
// D = upper triangle (mode 1) or all (mode 0) of S, entry (r, c) scaled by 2^-(s_r + s_c)
template <int NL>
__global__ void mat_copy_scaled_kernel(mp::Tensor D, const int64_t* __restrict__ offD, int64_t shD, int ldD, mp::Tensor S,
                                       const int64_t* __restrict__ offS, int64_t shS, int ldS, int batch, int n, int mode,
                                       const int* __restrict__ sc) {
  int64_t total = (int64_t)batch * n * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(idx / ((int64_t)n * n));
    int rem = (int)(idx % ((int64_t)n * n));
    int r = rem / n, c = rem % n;
    Num<NL> v = mp::zero<NL>();
    if (mode == 0 || (mode == 1 && r <= c)) {
      v = ldm<NL>(S, offS[b] + shS + (int64_t)r * ldS + c);
      v = mp::mul_2exp(v, -(sc[b * n + r] + sc[b * n + c]));
    }
    stm<NL>(D, offD[b] + shD + (int64_t)r * ldD + c, v);
  }
}
// M[r][c] *= 2^(sign * s_c): the column scaling that turns the factors of the equilibrated matrix into those of A
template <int NL>
__global__ void col_scale_kernel(mp::Tensor M, const int64_t* __restrict__ offM, int64_t shM, int ldM, int batch, int n, int sign,
                                 const int* __restrict__ sc) {
  int64_t total = (int64_t)batch * n * n;
  uint32_t* hdr = M.w + (size_t)NL * M.n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(idx / ((int64_t)n * n));
    int rem = (int)(idx % ((int64_t)n * n));
    int r = rem / n, c = rem % n;
    const int64_t at = offM[b] + shM + (int64_t)r * ldM + c;
    const uint32_t h = hdr[at];
    const int32_t e = ((int32_t)h) >> 1;
    if (e != mp::EXP_ZERO) hdr[at] = mp::pack_hdr(e + sign * sc[b * n + c], h & 1u);
  }
}
void equil_exponents(Ctx& ctx, int nl, const MatBatch& A, int* d_scale) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("equil_exponents");
    equil_exponents_kernel<NL><<<ceil_div((int64_t)A.batch * A.n, 128), 128, 0, ctx.stream>>>(A.t, A.d_off, A.shift, A.stride(),
                                                                                               A.batch, A.n, d_scale);
    ctx.end(tk);
  });
}
void mat_copy_scaled(Ctx& ctx, int nl, const MatBatch& dst, const MatBatch& src, bool upper_only, const int* d_scale) {
  int64_t total = (int64_t)dst.batch * dst.n * dst.n;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("mat_copy", (double)total * 4.0 * (NL + 1) * 2);
    mat_copy_scaled_kernel<NL><<<ew_grid(ctx, total), 128, 0, ctx.stream>>>(dst.t, dst.d_off, dst.shift, dst.stride(), src.t,
                                                                            src.d_off, src.shift, src.stride(), dst.batch,
                                                                            dst.n, upper_only ? 1 : 0, d_scale);
    ctx.end(tk);
  });
}
void col_scale(Ctx& ctx, int nl, const MatBatch& M, int sign, const int* d_scale) {
  int64_t total = (int64_t)M.batch * M.n * M.n;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("col_scale", (double)total * 8.0);
    col_scale_kernel<NL><<<ew_grid(ctx, total), 128, 0, ctx.stream>>>(M.t, M.d_off, M.shift, M.stride(), M.batch, M.n, sign, d_scale);
    ctx.end(tk);
  });
}
// v[off + i] *= 2^(sign * sc[i]) (headers only); scatter of the per-batch exponents into a vector indexed like x
__global__ void vec_scale_kernel(uint32_t* __restrict__ hdr, int64_t n, int sign, const int* __restrict__ sc) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t h = hdr[i];
    const int32_t e = ((int32_t)h) >> 1;
    if (e != mp::EXP_ZERO) hdr[i] = mp::pack_hdr(e + sign * sc[i], h & 1u);
  }
}
__global__ void scatter_scale_kernel(const int* __restrict__ src, int batch, int n, const int64_t* __restrict__ off,
                                     int* __restrict__ dst) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * n) return;
  int b = idx / n, i = idx - b * n;
  dst[off[b] + i] = src[idx];
}
// v[off + i] = -v[off + i] where sg[i] != 0 (headers only)
__global__ void vec_flip_kernel(uint32_t* __restrict__ hdr, int64_t n, const int* __restrict__ sg) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t h = hdr[i];
    if (sg[i] && (((int32_t)h) >> 1) != mp::EXP_ZERO) hdr[i] = h ^ 1u;
  }
}
void vec_flip(Ctx& ctx, int nl, mp::Tensor v, int64_t off, int64_t n, const int* d_sig) {
  if (n <= 0) return;
  int tk = ctx.begin("vec_flip", (double)n * 12.0);
  vec_flip_kernel<<<ew_grid(ctx, n), 128, 0, ctx.stream>>>(v.w + (size_t)nl * v.n + off, n, d_sig);
  ctx.end(tk);
}
void vec_scale(Ctx& ctx, int nl, mp::Tensor v, int64_t off, int64_t n, int sign, const int* d_scale) {
  if (n <= 0) return;
  int tk = ctx.begin("vec_scale", (double)n * 12.0);
  vec_scale_kernel<<<ew_grid(ctx, n), 128, 0, ctx.stream>>>(v.w + (size_t)nl * v.n + off, n, sign, d_scale);
  ctx.end(tk);
}
void scatter_scale(Ctx& ctx, const int* d_src, int batch, int n, const int64_t* d_off, int* d_dst) {
  int tk = ctx.begin("scatter_scale", (double)batch * n * 8.0);
  scatter_scale_kernel<<<ceil_div((int64_t)batch * n, 128), 128, 0, ctx.stream>>>(d_src, batch, n, d_off, d_dst);
  ctx.end(tk);
}
void mat_zero(Ctx& ctx, int nl, const MatBatch& dst) {
  int64_t total = (int64_t)dst.batch * dst.n * dst.n;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("mat_zero", (double)total * 4.0 * (NL + 1));
    mat_copy_kernel<NL><<<ew_grid(ctx, total), 128, 0, ctx.stream>>>(dst.t, dst.d_off, dst.shift, dst.stride(), dst.t,
                                                                     dst.d_off, dst.shift, dst.stride(), dst.batch,
                                                                     dst.n, 2);
    ctx.end(tk);
  });
}
void ew_zero(Ctx& ctx, int nl, mp::Tensor t, int64_t off, int64_t n) { ew_lincomb(ctx, nl, t, off, t, off, 0, t, off, 0, n); }
void ew_axpy(Ctx& ctx, int nl, mp::Tensor y, int64_t yo, mp::Tensor x, int64_t xo, mp::Tensor scal, int slot,
             int64_t n) {
  if (n <= 0) return;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("ew_axpy", (double)n * 4.0 * (NL + 1) * 3);
    axpy_kernel<NL><<<ew_grid(ctx, n), 128, 0, ctx.stream>>>(y, yo, x, xo, scal, slot, n);
    ctx.end(tk);
  });
}
void ew_residual_R(Ctx& ctx, int nl, const MatBatch& R, mp::Tensor scal, int slot, mp::Tensor T1, const mp::Tensor* T2) {
  int64_t total = (int64_t)R.batch * R.n * R.n;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("ew_residual_R", (double)total * 4.0 * (NL + 1) * (T2 ? 3 : 2));
    residual_R_kernel<NL><<<ew_grid(ctx, total), 128, 0, ctx.stream>>>(R.t, R.d_off, R.batch, R.n, scal, slot, T1,
                                                                       T2 ? *T2 : T1, T2 ? 1 : 0);
    ctx.end(tk);
  });
}
void ew_symmetrize(Ctx& ctx, int nl, const MatBatch& out, mp::Tensor in) {
  int64_t total = (int64_t)out.batch * out.n * out.n;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("ew_symmetrize", (double)total * 4.0 * (NL + 1) * 2);
    symmetrize_kernel<NL><<<ew_grid(ctx, total), 128, 0, ctx.stream>>>(out.t, out.d_off, out.batch, out.n, in);
    ctx.end(tk);
  });
}
void ew_set_identity(Ctx& ctx, int nl, const MatBatch& M, mp::Tensor scal, int slot) {
  int64_t total = (int64_t)M.batch * M.n * M.n;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("ew_set_identity");
    set_identity_kernel<NL><<<ew_grid(ctx, total), 128, 0, ctx.stream>>>(M.t, M.d_off, M.batch, M.n, scal, slot);
    ctx.end(tk);
  });
}

// =========================================================================================================
// reductions: grid-stride partials -> one partial per CTA -> final CTA
// =========================================================================================================
constexpr int RED_BLOCKS = 592;  // 4 CTAs per SM on 148 SMs
size_t reduce_work_elems() { return RED_BLOCKS; }
enum { RK_DOT = 0, RK_DOT_SUM = 1, RK_MAXABS = 2, RK_MIN = 3, RK_SUM = 4 };

template <int NL, int KIND>
__global__ void reduce_stage1(mp::Tensor a, int64_t ao, mp::Tensor da, mp::Tensor b, int64_t bo, mp::Tensor db,
                              int64_t n, mp::Tensor work) {
  extern __shared__ uint32_t sm[];
  constexpr int OP = (KIND == RK_MAXABS) ? RED_MAX : (KIND == RK_MIN ? RED_MIN : RED_ADD);
  Num<NL> acc = mp::zero<NL>();
  bool have = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Num<NL> v;
    if (KIND == RK_DOT) v = nmul(ldm<NL>(a, ao + i), ldm<NL>(b, bo + i));
    if (KIND == RK_DOT_SUM)
      v = nmul(nadd(ldm<NL>(a, ao + i), ldm<NL>(da, ao + i)), nadd(ldm<NL>(b, bo + i), ldm<NL>(db, bo + i)));
    if (KIND == RK_MAXABS) v = mp::fabs(ldm<NL>(a, ao + i));
    if (KIND == RK_MIN || KIND == RK_SUM) v = ldm<NL>(a, ao + i);
    if (KIND == RK_MIN)
      acc = have ? red_op<NL, OP>(acc, v) : v;
    else
      acc = red_op<NL, OP>(acc, v);
    have = true;
  }
  if (KIND == RK_MIN) {
    // threads without data must not inject a zero: borrow the value of element 0 (idempotent for min)
    if (!have) acc = ldm<NL>(a, ao);
  }
  Num<NL> r = block_reduce<NL, OP>(acc, sm);
  if (threadIdx.x == 0) stm<NL>(work, blockIdx.x, r);
}
template <int NL, int OP>
__global__ void reduce_stage2(mp::Tensor work, int nparts, mp::Tensor scal, int slot) {
  extern __shared__ uint32_t sm[];
  Num<NL> acc = ldm<NL>(work, threadIdx.x < nparts ? threadIdx.x : 0);
  if (OP == RED_ADD && threadIdx.x >= nparts) acc = mp::zero<NL>();
  for (int i = threadIdx.x + blockDim.x; i < nparts; i += blockDim.x) acc = red_op<NL, OP>(acc, ldm<NL>(work, i));
  Num<NL> r = block_reduce<NL, OP>(acc, sm);
  if (threadIdx.x == 0) stm<NL>(scal, slot, r);
}
template <int NL, int KIND>
static void reduce_launch(Ctx& ctx, mp::Tensor a, int64_t ao, mp::Tensor da, mp::Tensor b, int64_t bo, mp::Tensor db,
                          int64_t n, mp::Tensor scal, int slot, mp::Tensor work, const char* name, int nin) {
  constexpr int OP = (KIND == RK_MAXABS) ? RED_MAX : (KIND == RK_MIN ? RED_MIN : RED_ADD);
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(RED_BLOCKS, ceil_div(n, 128)));
  size_t smem = 33 * (NL + 2) * sizeof(uint32_t);
  int tk = ctx.begin(name, (double)n * 4.0 * (NL + 1) * nin);
  reduce_stage1<NL, KIND><<<grid, 128, smem, ctx.stream>>>(a, ao, da, b, bo, db, n, work);
  ctx.end(tk);
  tk = ctx.begin("reduce_final");
  reduce_stage2<NL, OP><<<1, 256, smem, ctx.stream>>>(work, grid, scal, slot);
  ctx.end(tk);
}
void reduce_dot(Ctx& ctx, int nl, mp::Tensor a, int64_t ao, mp::Tensor b, int64_t bo, int64_t n, mp::Tensor scal,
                int slot, mp::Tensor work) {
  DISPATCH_NL(nl, (reduce_launch<NL, RK_DOT>(ctx, a, ao, a, b, bo, b, n, scal, slot, work, "reduce_dot", 2)));
}
void reduce_dot_sum(Ctx& ctx, int nl, mp::Tensor a, mp::Tensor da, mp::Tensor b, mp::Tensor db, int64_t n,
                    mp::Tensor scal, int slot, mp::Tensor work) {
  DISPATCH_NL(nl, (reduce_launch<NL, RK_DOT_SUM>(ctx, a, 0, da, b, 0, db, n, scal, slot, work, "reduce_dot_sum", 4)));
}
void reduce_maxabs(Ctx& ctx, int nl, mp::Tensor a, int64_t ao, int64_t n, mp::Tensor scal, int slot, mp::Tensor work) {
  DISPATCH_NL(nl, (reduce_launch<NL, RK_MAXABS>(ctx, a, ao, a, a, 0, a, n, scal, slot, work, "reduce_maxabs", 1)));
}
void reduce_min(Ctx& ctx, int nl, mp::Tensor a, int64_t ao, int64_t n, mp::Tensor scal, int slot, mp::Tensor work) {
  DISPATCH_NL(nl, (reduce_launch<NL, RK_MIN>(ctx, a, ao, a, a, 0, a, n, scal, slot, work, "reduce_min", 1)));
}

// =========================================================================================================
// GEMV: one warp per (row, K-part); lanes stride K; partials summed by a second kernel when K is split
// =========================================================================================================
struct GemvDev {
  mp::Tensor A, x, out, work;
  int64_t a0, x0, oo, rs, ks;
  int rows, K, nparts;
  const int* row_item;
  const int64_t* aoff;
  const int64_t* xoff;
  const int* row0;
  const int* itemK;
  int item_trans;
  int lpr;  // lanes per row of the row-major kernel (a power of two <= 32)
  // fused header-word operations (see GemvArgs)
  const int* x_flip;
  const int* x_scale;
  int x_scale_sign;
  const int* o_flip;
  const int* o_scale;
  int o_scale_sign;
  mp::Tensor e;
  int64_t e0;
  int e_mode;
};
template <int NL>
__device__ __forceinline__ Num<NL> gemv_x(const GemvDev& g, int64_t at) {
  Num<NL> v = ldm<NL>(g.x, at);
  if (!mp::is_zero(v)) {
    if (g.x_scale) v.e += g.x_scale_sign * g.x_scale[at - g.x0];
    if (g.x_flip) v.neg ^= (uint32_t)(g.x_flip[at - g.x0] & 1);
  }
  return v;
}
// result of row r: sign flip, power-of-two scale, combination with e (in this order), store
template <int NL>
__device__ __forceinline__ void gemv_finish(const GemvDev& g, int r, Num<NL> acc) {
  if (!mp::is_zero(acc)) {
    if (g.o_flip) acc.neg ^= (uint32_t)(g.o_flip[r] & 1);
    if (g.o_scale) acc.e += g.o_scale_sign * g.o_scale[r];
  }
  if (g.e_mode == 1)
    acc = mp::add(ldm<NL>(g.e, g.e0 + r), acc);
  else if (g.e_mode == 2)
    acc = mp::sub(ldm<NL>(g.e, g.e0 + r), acc);
  stm<NL>(g.out, g.oo + r, acc);
}
template <int NL>
__global__ void gemv_kernel(GemvDev g) {
  // a group of `lpr` adjacent lanes per (row, K-part): short rows (the K = dim_S items) would otherwise spend as much on the
  // shuffle reduction (5 multiprecision additions) as on their 4 multiply-adds per lane
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lpr = g.lpr, sub = (int)(gt % lpr);
  const int64_t grp = gt / lpr;
  const bool live = grp < (int64_t)g.rows * g.nparts;
  int r = live ? (int)(grp / g.nparts) : 0, part = live ? (int)(grp % g.nparts) : 0;
  int64_t ab = g.a0, xb = g.x0;
  int rl = r, K = g.K;
  int64_t rs = g.rs, ks = g.ks;
  if (g.row_item) {
    int it = g.row_item[r];
    ab = g.aoff[it];
    xb = g.xoff[it];
    rl = r - g.row0[it];
    K = g.itemK[it];
    rs = g.item_trans ? 1 : K;
    ks = g.item_trans ? K : 1;
  }
  int chunk = (K + g.nparts - 1) / g.nparts;
  int k0 = part * chunk, k1 = live ? min(K, k0 + chunk) : 0;
  Num<NL> acc = mp::zero<NL>();
  for (int k = k0 + sub; k < k1; k += lpr)
    acc = mp::add(acc, mp::mul(ldm<NL>(g.A, ab + (int64_t)rl * rs + (int64_t)k * ks), gemv_x<NL>(g, xb + k)));
#pragma unroll 1
  for (int o = lpr >> 1; o; o >>= 1) acc = nadd(acc, shfl_xor_num(acc, o));
  if (live && sub == 0) {
    if (g.nparts == 1)
      gemv_finish<NL>(g, r, acc);
    else
      stm<NL>(g.work, (int64_t)part * g.rows + r, acc);
  }
}
// Transposed access (element (r, k) at k*ld + r: B^T x, W dy, L^-T u): a warp owns 32 CONSECUTIVE rows, one per lane, so
// every load of the warp is one coalesced 128-byte line per limb plane (the row-per-warp kernel above would touch 32
// sectors for 32 words); each lane runs its own multiply-add chain over its K-part, the parts are summed by
// gemv_sum_kernel. With items, lanes look their item up individually (a warp may straddle two items).
template <int NL>
__global__ void gemv_t_kernel(GemvDev g) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int ngrp = (g.rows + 31) >> 5;
  if (warp >= ngrp * g.nparts) return;
  const int grp = warp / g.nparts, part = warp % g.nparts;
  const int r = grp * 32 + lane;
  if (r >= g.rows) return;
  int64_t ab = g.a0, xb = g.x0, ld = g.ks;
  int rl = r, K = g.K;
  if (g.row_item) {
    const int it = g.row_item[r];
    ab = g.aoff[it];
    xb = g.xoff[it];
    rl = r - g.row0[it];
    K = g.itemK[it];
    ld = K;
  }
  const int chunk = (K + g.nparts - 1) / g.nparts;
  const int k0 = part * chunk, k1 = min(K, k0 + chunk);
  Num<NL> acc = mp::zero<NL>();
  for (int k = k0; k < k1; k++) acc = mp::add(acc, mp::mul(ldm<NL>(g.A, ab + (int64_t)k * ld + rl), gemv_x<NL>(g, xb + k)));
  if (g.nparts == 1)
    gemv_finish<NL>(g, r, acc);
  else
    stm<NL>(g.work, (int64_t)part * g.rows + r, acc);
}
template <int NL>
__global__ void gemv_sum_kernel(GemvDev g) {
  // One group of 2^k >= min(nparts, 32) adjacent lanes per row: the parts are loaded in parallel and meet in a shuffle
  // reduction (one thread per row summed up to 16 parts in a serial chain of dependent loads: 16 us per call, 19 calls on
  // the critical path of an iteration).
  const int lpr = g.lpr;
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int sub = (int)(gt % lpr);
  const int64_t r = gt / lpr;
  const bool live = r < g.rows;
  Num<NL> acc = mp::zero<NL>();
  if (live)
    for (int p = sub; p < g.nparts; p += lpr) acc = nadd(acc, ldm<NL>(g.work, (int64_t)p * g.rows + r));
#pragma unroll 1
  for (int o = lpr >> 1; o; o >>= 1) acc = nadd(acc, shfl_xor_num(acc, o));
  if (live && sub == 0) gemv_finish<NL>(g, (int)r, acc);
}
static int gemv_parts(int rows, int K) {
  if ((int64_t)rows >= 148 * 8 || K <= 256) return 1;
  int p = std::min(ceil_div(K, 128), ceil_div(148 * 16, std::max(rows, 1)));
  return std::max(1, p);
}
static int gemv_parts_t(int rows, int K) {  // transposed kernel: about 16 warps per SM, at least 8 terms per part
  const int ngrp = (rows + 31) / 32;
  return std::max(1, std::min(ceil_div(K, 8), (148 * 16) / std::max(ngrp, 1)));
}
size_t gemv_work_elems(int rows, int K) { return (size_t)rows * std::max(gemv_parts(rows, K), gemv_parts_t(rows, K)); }
void gemv(Ctx& ctx, int nl, const GemvArgs& a, mp::Tensor work) {
  if (a.rows <= 0) return;
  GemvDev g{a.A, a.x, a.out, work, a.a0, a.x0, a.oo, a.rs, a.ks, a.rows, a.K, 1, a.d_row_item, a.d_aoff, a.d_xoff,
            a.d_row0, a.d_K, a.item_trans, 32, a.x_flip, a.x_scale, a.x_scale_sign, a.o_flip, a.o_scale, a.o_scale_sign,
            a.e, a.e0, a.e_mode};
  const bool transposed = a.d_row_item ? a.item_trans != 0 : (a.rs == 1 && a.ks != 1);
  // (items: every item is K_item x K_item here, so a.rows / number of items bounds K; the work buffer is sized by the caller
  // with gemv_work_elems(rows, rows))
  g.nparts = transposed ? gemv_parts_t(a.rows, a.d_row_item ? std::max(1, a.K_hint) : a.K) : (a.d_row_item ? 1 : gemv_parts(a.rows, a.K));
  {  // at least ~16 terms per lane
    const int kk = (a.d_row_item ? std::max(1, a.K_hint) : a.K) / g.nparts;
    static int lpr_env = getenv("CLRSDP_GEMV_LPR") ? atoi(getenv("CLRSDP_GEMV_LPR")) : 0;  // measuring aid
    while (g.lpr > 4 && kk < 16 * g.lpr) g.lpr >>= 1;
    if (lpr_env) g.lpr = lpr_env;
  }
  DISPATCH_NL(nl, {
    int tk = ctx.begin(transposed ? "gemv_t" : "gemv", (double)a.rows * a.K * 4.0 * (NL + 1));
    if (transposed) {
      int64_t warps = (int64_t)((a.rows + 31) / 32) * g.nparts;
      gemv_t_kernel<NL><<<ceil_div(warps * 32, 128), 128, 0, ctx.stream>>>(g);
    } else {
      int64_t thr = (int64_t)a.rows * g.nparts * g.lpr;
      gemv_kernel<NL><<<ceil_div(thr, 128), 128, 0, ctx.stream>>>(g);
    }
    ctx.end(tk);
    if (g.nparts > 1) {
      tk = ctx.begin("gemv_sum");
      g.lpr = 2;
      while (g.lpr < 32 && g.lpr < g.nparts) g.lpr <<= 1;
      gemv_sum_kernel<NL><<<ceil_div((int64_t)a.rows * g.lpr, 128), 128, 0, ctx.stream>>>(g);
      ctx.end(tk);
    }
  });
}

// =========================================================================================================
// wire format <-> header word (the pinned transfer path of solver.cu)
// =========================================================================================================
template <int NL>
__global__ void wire_pack_kernel(mp::Tensor t, int64_t off, int64_t n, const int8_t* __restrict__ sign,
                                 const int64_t* __restrict__ exp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int sg = sign[i];
    uint32_t* hdr = t.w + (size_t)NL * t.n + off + i;
    if (sg == 0) {
      *hdr = mp::pack_hdr(mp::EXP_ZERO, 0);
#pragma unroll
      for (int k = 0; k < NL; k++) t.w[(size_t)k * t.n + off + i] = 0;
    } else {
      *hdr = mp::pack_hdr((int32_t)exp[i], sg < 0 ? 1u : 0u);
    }
  }
}
template <int NL>
__global__ void wire_unpack_kernel(mp::Tensor t, int64_t off, int64_t n, int8_t* __restrict__ sign, int64_t* __restrict__ exp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t h = t.w[(size_t)NL * t.n + off + i];
    const int32_t e = ((int32_t)h) >> 1;
    const bool z = e == mp::EXP_ZERO;
    sign[i] = z ? 0 : ((h & 1u) ? -1 : 1);
    exp[i] = z ? 0 : e;
  }
}
void wire_pack(Ctx& ctx, int nl, mp::Tensor t, int64_t off, int64_t n, const int8_t* d_sign, const int64_t* d_exp) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("wire_pack", (double)n * 13.0);
    wire_pack_kernel<NL><<<(int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx.sm_count * 8), 256, 0, ctx.stream>>>(t, off, n, d_sign, d_exp);
    ctx.end(tk);
  });
}
void wire_unpack(Ctx& ctx, int nl, mp::Tensor t, int64_t off, int64_t n, int8_t* d_sign, int64_t* d_exp) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("wire_unpack", (double)n * 13.0);
    wire_unpack_kernel<NL><<<(int)std::min<int64_t>(ceil_div(n, 256), (int64_t)ctx.sm_count * 8), 256, 0, ctx.stream>>>(t, off, n, d_sign, d_exp);
    ctx.end(tk);
  });
}

// =========================================================================================================
// small batched products on the CUDA cores
// =========================================================================================================
// C[b][i][j] (epi) sum_k A(b,i,k) * B(b,j,k) for products too small to amortise the tensor-core pipeline (slice,
// 45 pipeline stages per tile, carry: ~90 us of latency whatever the size). KS adjacent lanes share one output
// entry (K split KS ways, shuffle reduction); plain multiprecision multiply-add, the oracle's own error model.
struct SmallGemmDev {
  mp::Tensor A, B, C, E;
  const int64_t* offA;
  const int64_t* offB;
  const int64_t* offC;
  int64_t a0, abs_, ars, aks, b0, bbs, brs, bks, c0, cbs, crs, ccs;
  int batch, M, N, K, ks_log, epi;
  const int* ksign;  // optional [batch][ksign_ld]: term k of item b enters with a minus sign where ksign != 0
  int ksign_ld;
};
template <int NL>
__global__ void __launch_bounds__(256) small_gemm_kernel(SmallGemmDev g) {
  const int KS = 1 << g.ks_log;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t out = gid >> g.ks_log;
  const int part = (int)(gid & (KS - 1));
  const int64_t total = (int64_t)g.batch * g.M * g.N;
  const bool live = out < total;
  Num<NL> acc = mp::zero<NL>();
  int b = 0, i = 0, j = 0;
  if (live) {
    j = (int)(out % g.N);
    i = (int)((out / g.N) % g.M);
    b = (int)(out / ((int64_t)g.N * g.M));
    const int64_t ab = g.a0 + (g.offA ? g.offA[b] : (int64_t)b * g.abs_) + (int64_t)i * g.ars;
    const int64_t bb = g.b0 + (g.offB ? g.offB[b] : (int64_t)b * g.bbs) + (int64_t)j * g.brs;
    const int* ksg = g.ksign ? g.ksign + (int64_t)b * g.ksign_ld : nullptr;
    for (int k = part; k < g.K; k += KS) {
      Num<NL> t = mp::mul(ldm<NL>(g.A, ab + (int64_t)k * g.aks), ldm<NL>(g.B, bb + (int64_t)k * g.bks));
      if (ksg && ksg[k] && !mp::is_zero(t)) t.neg ^= 1u;
      acc = mp::add(acc, t);
    }
  }
#pragma unroll 1
  for (int o = KS >> 1; o; o >>= 1) acc = nadd(acc, shfl_xor_num(acc, o));
  if (live && part == 0) {
    const int64_t at = g.c0 + (g.offC ? g.offC[b] : (int64_t)b * g.cbs) + (int64_t)i * g.crs + (int64_t)j * g.ccs;
    if (g.epi == 4) {  // EPI_NEG
      acc = mp::neg(acc);
    } else if (g.epi != 0) {
      Num<NL> e = ldm<NL>(g.E, at);
      if (g.epi == 1)  // EPI_SUB_FROM: E - A*B
        acc = nsub(e, acc);
      else if (g.epi == 2)  // EPI_MINUS_SUB: A*B - E
        acc = nsub(acc, e);
      else
        acc = nadd(acc, e);
    }
    stm<NL>(g.C, at, acc);
  }
}
// C(b,i,j) = S(b,i,j) for rectangular strided blocks (C addressing as in SmallGemmArgs; S through the a* fields)
template <int NL>
__global__ void rect_copy_kernel(SmallGemmDev g) {
  const int64_t total = (int64_t)g.batch * g.M * g.N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % g.N), i = (int)((idx / g.N) % g.M), b = (int)(idx / ((int64_t)g.N * g.M));
    const int64_t from = g.a0 + (g.offA ? g.offA[b] : (int64_t)b * g.abs_) + (int64_t)i * g.ars + (int64_t)j * g.aks;
    const int64_t to = g.c0 + (g.offC ? g.offC[b] : (int64_t)b * g.cbs) + (int64_t)i * g.crs + (int64_t)j * g.ccs;
    stm<NL>(g.C, to, ldm<NL>(g.A, from));
  }
}
void rect_copy(Ctx& ctx, int nl, const SmallGemmArgs& a) {
  SmallGemmDev g;
  memset(&g, 0, sizeof(g));
  g.A = a.A, g.C = a.C;
  g.offA = a.offA, g.offC = a.offC;
  g.a0 = a.a0, g.abs_ = a.abs_, g.ars = a.ars, g.aks = a.aks;
  g.c0 = a.c0, g.cbs = a.cbs, g.crs = a.crs, g.ccs = a.ccs;
  g.batch = a.batch, g.M = a.M, g.N = a.N;
  const int64_t total = (int64_t)a.batch * a.M * a.N;
  if (total <= 0) return;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("rect_copy", (double)total * 8.0 * (NL + 1));
    rect_copy_kernel<NL><<<(int)std::min<int64_t>(ceil_div(total, 128), (int64_t)ctx.sm_count * 16), 128, 0, ctx.stream>>>(g);
    ctx.end(tk);
  });
}
void small_gemm(Ctx& ctx, int nl, const SmallGemmArgs& a) {
  SmallGemmDev g;
  g.A = a.A, g.B = a.B, g.C = a.C, g.E = a.E.w ? a.E : a.C;
  g.offA = a.offA, g.offB = a.offB, g.offC = a.offC;
  g.a0 = a.a0, g.abs_ = a.abs_, g.ars = a.ars, g.aks = a.aks;
  g.b0 = a.b0, g.bbs = a.bbs, g.brs = a.brs, g.bks = a.bks;
  g.c0 = a.c0, g.cbs = a.cbs, g.crs = a.crs, g.ccs = a.ccs;
  g.batch = a.batch, g.M = a.M, g.N = a.N, g.K = a.K, g.epi = a.epi;
  g.ksign = a.ksign, g.ksign_ld = a.ksign_ld;
  const int64_t outs = (int64_t)a.batch * a.M * a.N;
  if (outs <= 0) return;
  // split K over adjacent lanes until the grid fills the machine (about 4 x 256 threads per SM)
  int ks_log = 0;
  while (ks_log < 5 && (2 << ks_log) <= a.K && (outs << ks_log) < (int64_t)ctx.sm_count * 1024) ks_log++;
  g.ks_log = ks_log;
  DISPATCH_NL(nl, {
    int tk = ctx.begin("small_gemm", (double)outs * a.K);
    small_gemm_kernel<NL><<<ceil_div(outs << ks_log, 256), 256, 0, ctx.stream>>>(g);
    ctx.end(tk);
  });
}

// =========================================================================================================
// structure-aware kernels
// =========================================================================================================
__device__ __forceinline__ void tri_decode(int pr, int& r, int& s) {  // pr = s + r(r+1)/2, s <= r
  r = 0;
  while ((r + 1) * (r + 2) / 2 <= pr) r++;
  s = pr - r * (r + 1) / 2;
}

// thread per (cluster, ver, hor) with ver <= hor
template <int NL>
__global__ void schur_kernel(StructTables st, mp::Tensor Px, mp::Tensor Py, mp::Tensor H, mp::Tensor S, int64_t total,
                             const int* __restrict__ task_cluster_start) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    // find the cluster by binary search over the prefix of dimS^2 (c_Soff)
    int lo = 0, hi = st.J - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (st.c_Soff[mid] <= idx) lo = mid; else hi = mid - 1;
    }
    const int j = lo;
    const int dimS = st.c_dimS[j], K = st.c_K[j], m = st.c_m[j];
    int64_t rem = idx - st.c_Soff[j];
    int ver = (int)(rem / dimS), hor = (int)(rem % dimS);
    if (ver > hor) continue;
    int r1, s1, r2, s2;
    tri_decode(hor / K, r1, s1);
    tri_decode(ver / K, r2, s2);
    int k1 = hor % K, k2 = ver % K;
    Num<NL> acc = mp::zero<NL>();
    for (int l = 0; l < st.c_L[j]; l++) {
      int bk = st.c_blk0[j] + l;
      int Nv = st.b_Nv[bk];
      const int* rsum = st.rank_sums + st.b_rs0[bk];
      int64_t po = st.b_Poff[bk], ho = st.b_Hoff[bk];
      int ld = m * Nv;
      for (int a1 = rsum[k1]; a1 < rsum[k1 + 1]; a1++)
        for (int a2 = rsum[k2]; a2 < rsum[k2 + 1]; a2++) {
          int r1s = a1 + Nv * r1, r2s = a2 + Nv * r2, s1s = a1 + Nv * s1, s2s = a2 + Nv * s2;
          Num<NL> tot;
          if (m == 1) {
            // all four terms coincide (MPMP.jl:1373-1392 with r=s): tot = 4 Px[a1][a2] Py[a2][a1]
            tot = mp::mul_2exp(nmul(ldm<NL>(Px, po + (int64_t)a1 * ld + a2), ldm<NL>(Py, po + (int64_t)a2 * ld + a1)), 2);
          } else {
            tot = nmul(ldm<NL>(Px, po + (int64_t)s1s * ld + r2s), ldm<NL>(Py, po + (int64_t)s2s * ld + r1s));
            tot = nadd(tot, nmul(ldm<NL>(Px, po + (int64_t)r1s * ld + r2s), ldm<NL>(Py, po + (int64_t)s2s * ld + s1s)));
            tot = nadd(tot, nmul(ldm<NL>(Px, po + (int64_t)s1s * ld + s2s), ldm<NL>(Py, po + (int64_t)r2s * ld + r1s)));
            tot = nadd(tot, nmul(ldm<NL>(Px, po + (int64_t)r1s * ld + s2s), ldm<NL>(Py, po + (int64_t)r2s * ld + s1s)));
          }
          tot = nmul(ldm<NL>(H, ho + a1), tot);
          tot = nmul(ldm<NL>(H, ho + a2), tot);
          acc = nadd(acc, mp::mul_2exp(tot, -2));
        }
    }
    int64_t so = st.c_Soff[j];
    stm<NL>(S, so + (int64_t)ver * dimS + hor, acc);
    if (ver != hor) stm<NL>(S, so + (int64_t)hor * dimS + ver, acc);
  }
  (void)task_cluster_start;
}
void schur_assemble(Ctx& ctx, int nl, const StructTables& st, mp::Tensor Px, mp::Tensor Py, mp::Tensor H, mp::Tensor S,
                    int64_t S_total) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("schur_assemble", (double)S_total * 4.0 * (NL + 1) * 2.0);
    schur_kernel<NL><<<ew_grid(ctx, S_total), 128, 0, ctx.stream>>>(st, Px, Py, H, S, S_total, nullptr);
    ctx.end(tk);
  });
}

// thread per x entry (j, r, s, k)
template <int NL>
__global__ void trace_kernel(StructTables st, mp::Tensor A, mp::Tensor H, mp::Tensor out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.sumS) return;
  int j = st.x_cluster[i];
  int K = st.c_K[j], m = st.c_m[j];
  int loc = i - st.c_xoff[j];
  int r, s;
  tri_decode(loc / K, r, s);
  int k = loc % K;
  Num<NL> acc = mp::zero<NL>();
  for (int l = 0; l < st.c_L[j]; l++) {
    int bk = st.c_blk0[j] + l;
    int Nv = st.b_Nv[bk];
    const int* rsum = st.rank_sums + st.b_rs0[bk];
    int64_t ho = st.b_Hoff[bk];
    for (int a = rsum[k]; a < rsum[k + 1]; a++) {
      // from the pairings: Py[(r,a),(s,a)]   (the general method on Z V is trace_zv_kernel below)
      const int ld = m * Nv;
      const Num<NL> val = ldm<NL>(A, st.b_Poff[bk] + (int64_t)(r * Nv + a) * ld + (s * Nv + a));
      acc = nadd(acc, nmul(ldm<NL>(H, ho + a), val));
    }
  }
  stm<NL>(out, i, acc);
}
void trace_from_pairings(Ctx& ctx, int nl, const StructTables& st, mp::Tensor Py, mp::Tensor H, mp::Tensor out) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("trace_pairings");
    trace_kernel<NL><<<ceil_div(st.sumS, 128), 128, 0, ctx.stream>>>(st, Py, H, out);
    ctx.end(tk);
  });
}
// trace_A, general method (MPMP.jl:1517-1618): out[i] = sum_l sum_a H[a] * <Vt[a], Tt[(r,s), a]> with one WARP per constraint
// index i: the lanes stride the vector (coalesced lines of both operands) and the partial sums meet in a shuffle
// reduction. (One thread per index - trace_kernel<NL, 1> - ran serial dot products of length delta with a stride of
// delta words between the lanes of a warp, on 256 warps for the whole GPU: 270 us per call at BASELINE config 3.)
template <int NL>
__global__ void trace_zv_kernel(StructTables st, mp::Tensor Vt, mp::Tensor Tt, mp::Tensor H, mp::Tensor out) {
  const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= st.sumS) return;
  const int j = st.x_cluster[i];
  const int K = st.c_K[j], m = st.c_m[j];
  const int loc = i - st.c_xoff[j];
  int r, s;
  tri_decode(loc / K, r, s);
  const int k = loc % K;
  Num<NL> acc = mp::zero<NL>();
  for (int l = 0; l < st.c_L[j]; l++) {
    const int bk = st.c_blk0[j] + l;
    const int Nv = st.b_Nv[bk], dl = st.b_delta[bk];
    const int* rsum = st.rank_sums + st.b_rs0[bk];
    const int64_t ho = st.b_Hoff[bk];
    for (int a = rsum[k]; a < rsum[k + 1]; a++) {
      const int64_t vo = st.b_Voff[bk] + (int64_t)a * dl, to = st.b_Toff[bk] + ((int64_t)(r * m + s) * Nv + a) * dl;
      Num<NL> val = mp::zero<NL>();
      for (int q = lane; q < dl; q += 32) val = nadd(val, nmul(ldm<NL>(Vt, vo + q), ldm<NL>(Tt, to + q)));
      acc = nadd(acc, nmul(ldm<NL>(H, ho + a), val));
    }
  }
#pragma unroll 1
  for (int o = 16; o; o >>= 1) acc = nadd(acc, shfl_xor_num(acc, o));
  if (lane == 0) stm<NL>(out, i, acc);
}
void trace_from_ZV(Ctx& ctx, int nl, const StructTables& st, mp::Tensor Vt, mp::Tensor Tt, mp::Tensor H, mp::Tensor out) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("trace_ZV");
    trace_zv_kernel<NL><<<ceil_div((int64_t)st.sumS * 32, 128), 128, 0, ctx.stream>>>(st, Vt, Tt, H, out);
    ctx.end(tk);
  });
}

// thread per VD element: VD[blk][pair][i][v]
template <int NL>
__global__ void scale_vectors_kernel(StructTables st, mp::Tensor Vt, mp::Tensor H, mp::Tensor a, mp::Tensor VD,
                                     int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = st.n_blocks - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (st.b_VDoff[mid] <= idx) lo = mid; else hi = mid - 1;
    }
    int bk = lo, j = st.b_cluster[bk];
    int Nv = st.b_Nv[bk], dl = st.b_delta[bk], K = st.c_K[j];
    int64_t rem = idx - st.b_VDoff[bk];
    int v = (int)(rem % Nv);
    int i = (int)((rem / Nv) % dl);
    int pair = (int)(rem / ((int64_t)Nv * dl));
    int k = st.samp[st.b_Hoff[bk] + v];
    Num<NL> w = nmul(ldm<NL>(a, st.c_xoff[j] + pair * K + k), ldm<NL>(H, st.b_Hoff[bk] + v));
    stm<NL>(VD, idx, nmul(w, ldm<NL>(Vt, st.b_Voff[bk] + (int64_t)v * dl + i)));
  }
}
void scale_vectors(Ctx& ctx, int nl, const StructTables& st, mp::Tensor Vt, mp::Tensor H, mp::Tensor a, mp::Tensor VD,
                   int64_t VD_total) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("scale_vectors", (double)VD_total * 4.0 * (NL + 1) * 2.0);
    scale_vectors_kernel<NL><<<ew_grid(ctx, VD_total), 128, 0, ctx.stream>>>(st, Vt, H, a, VD, VD_total);
    ctx.end(tk);
  });
}

// thread per block element (I <= J): gathers the (r,s) sub-block products, halves off-diagonal blocks,
// mirrors, and adds sign*E
template <int NL>
__global__ void assemble_kernel(StructTables st, mp::Tensor QP, mp::Tensor out, mp::Tensor E, int sign, int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = st.n_blocks - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (st.b_off[mid] <= idx) lo = mid; else hi = mid - 1;
    }
    int bk = lo, j = st.b_cluster[bk];
    int dl = st.b_delta[bk], nb = st.c_m[j] * dl;
    int64_t rem = idx - st.b_off[bk];
    int I = (int)(rem / nb), Jc = (int)(rem % nb);
    if (I > Jc) continue;
    int s = I / dl, i = I % dl, r = Jc / dl, i2 = Jc % dl;
    int pair = s + r * (r + 1) / 2;
    Num<NL> v = ldm<NL>(QP, st.b_QPoff[bk] + (int64_t)pair * dl * dl + (int64_t)i * dl + i2);
    if (r != s) v = mp::mul_2exp(v, -1);
    int64_t a1 = st.b_off[bk] + (int64_t)I * nb + Jc, a2 = st.b_off[bk] + (int64_t)Jc * nb + I;
    Num<NL> e1 = ldm<NL>(E, a1);
    stm<NL>(out, a1, sign > 0 ? nadd(v, e1) : nsub(v, e1));
    if (I != Jc) {
      Num<NL> e2 = ldm<NL>(E, a2);
      stm<NL>(out, a2, sign > 0 ? nadd(v, e2) : nsub(v, e2));
    }
  }
}
void assemble_weighted(Ctx& ctx, int nl, const StructTables& st, mp::Tensor QP, mp::Tensor out, mp::Tensor E, int sign,
                       int64_t blk_total, const int*) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("assemble_weighted", (double)blk_total * 4.0 * (NL + 1) * 3.0);
    assemble_kernel<NL><<<ew_grid(ctx, blk_total), 128, 0, ctx.stream>>>(st, QP, out, E, sign, blk_total);
    ctx.end(tk);
  });
}

// =========================================================================================================
// combination of the per-rank partial tensors after ncclAllGather (rank order => identical on all ranks)
// =========================================================================================================
template <int NL>
__global__ void combine_ranks_kernel(const uint32_t* __restrict__ g, int nranks, int64_t n, mp::Tensor out, int64_t off,
                                     int op) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Num<NL> acc = mp::load<NL>(g, (size_t)n, (size_t)i);
    for (int r = 1; r < nranks; r++) {
      Num<NL> v = mp::load<NL>(g + (size_t)r * (NL + 1) * n, (size_t)n, (size_t)i);
      if (op == COMB_SUM)
        acc = mp::add(acc, v);
      else if (op == COMB_MAX)
        acc = mp::cmp(acc, v) >= 0 ? acc : v;
      else
        acc = mp::cmp(acc, v) <= 0 ? acc : v;
    }
    stm<NL>(out, off + i, acc);
  }
}
void combine_ranks(Ctx& ctx, int nl, const uint32_t* gathered, int nranks, int64_t n, mp::Tensor out, int64_t off, int op) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("combine_ranks", (double)n * 4.0 * (NL + 1) * (nranks + 1));
    combine_ranks_kernel<NL><<<ew_grid(ctx, n), 128, 0, ctx.stream>>>(gathered, nranks, n, out, off, op);
    ctx.end(tk);
  });
}

// =========================================================================================================
// driver scalars (single thread; the reference does these in Arb on the host, MPMP.jl:755-756 etc.)
// =========================================================================================================
template <int NL>
__global__ void scalar_kernel(int prog, mp::Tensor sc, int* flags, double* dout, const int* __restrict__ status,
                              int n_status) {
  // a failed factorisation (X or Y not positive definite) must not reach the update: the step lengths become zero, so
  // the iterate of the last good iteration stays on the device for a resume at higher precision (the reference raises
  // before its update, MPMP.jl:793 precedes :877)
  int failed = 0;
  if (status) {
    for (int i = threadIdx.x; i < n_status; i += 32) failed |= status[i];
    failed = __any_sync(0xffffffffu, failed != 0);
  }
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  auto G = [&](int s) { return ldm<NL>(sc, s); };
  auto Pp = [&](int s, const Num<NL>& v) { stm<NL>(sc, s, v); };
  const Num<NL> one = mp::one<NL>();
  if (prog == SP_MU) {  // mu = dot(X,Y)/size(X,1); mu_p = pd_feas ? 0 : beta_inf*mu   (:755-756)
    Num<NL> mu = ndiv(G(SL_DOT_XY), G(SL_NTOT));
    Pp(SL_MU, mu);
    Pp(SL_MU_P, flags[0] ? mp::zero<NL>() : nmul(G(SL_BETA_INF), mu));
  } else if (prog == SP_BETA) {  // (:832-837)
    Num<NL> r = ndiv(G(SL_DOT_SUM), nmul(G(SL_MU), G(SL_NTOT)));
    Num<NL> beta = ncmp(r, one) < 0 ? nmul(r, r) : r;
    Num<NL> bc;
    if (flags[0]) {
      bc = ncmp(G(SL_BETA_FEAS), beta) > 0 ? G(SL_BETA_FEAS) : beta;
      if (ncmp(bc, one) > 0) bc = one;
    } else {
      bc = ncmp(G(SL_BETA_INF), beta) > 0 ? G(SL_BETA_INF) : beta;
    }
    Pp(SL_BETA_C, bc);
    Pp(SL_MU_C, nmul(bc, G(SL_MU)));
  } else if (prog == SP_ALPHA) {  // (:1893-1897) for X and Y, then (:871-874)
    Num<NL> ng = mp::neg(G(SL_GAMMA));
    Num<NL> ap = ncmp(G(SL_LAM_X), ng) > 0 ? one : ndiv(ng, G(SL_LAM_X));
    Num<NL> ad = ncmp(G(SL_LAM_Y), ng) > 0 ? one : ndiv(ng, G(SL_LAM_Y));
    if (flags[0]) {
      if (ncmp(ad, ap) < 0) ap = ad;
      ad = ap;
    }
    if (failed) ap = ad = mp::zero<NL>();
    Pp(SL_ALPHA_P, ap);
    Pp(SL_ALPHA_D, ad);
  } else if (prog == SP_OBJECTIVES || prog == SP_OBJECTIVES_INIT) {  // objectives and gap (:1027-1034, :1067-1078)
    const Num<NL> by = nadd(G(SL_CY), G(SL_BY));  // <C,Y> + <b,y> (:1033); SL_CY stays 0 while C = 0
    Num<NL> po = nadd(G(SL_CX), G(SL_B0)), dobj = nadd(by, G(SL_B0));
    Pp(SL_P_OBJ, po);
    Pp(SL_D_OBJ, dobj);
    if (prog == SP_OBJECTIVES_INIT) po = G(SL_CX), dobj = by;  // gap of the initial point: no b0 (:1067-1074)
    Num<NL> num = mp::fabs(nsub(po, dobj)), den = mp::fabs(nadd(po, dobj));
    if (ncmp(one, den) > 0) den = one;
    Pp(SL_GAP, ndiv(num, den));
  } else if (prog == SP_ERRORS) {  // errors, feasibility, termination (:1058-1064, :1147-1185)
    Num<NL> pe = ncmp(G(SL_PERR_P), G(SL_PERR_p)) > 0 ? G(SL_PERR_P) : G(SL_PERR_p);
    Pp(SL_PRIMAL_ERR, pe);
    Pp(SL_DUAL_ERR, G(SL_DERR));
    bool pf = ncmp(pe, G(SL_PERR_THR)) < 0, df = ncmp(G(SL_DERR), G(SL_DERR_THR)) < 0;
    bool gap_opt = ncmp(G(SL_GAP), G(SL_GAP_THR)) < 0;
    flags[0] = (pf && df) ? 1 : 0;
    int term = 0;
    if (flags[2] && pf)
      term = 1;
    else if (flags[3] && df)
      term = 2;
    else if (pf && df && gap_opt)
      term = 3;
    flags[1] = term;
  }
  if (dout)
    for (int s = 0; s < SL_COUNT; s++) dout[s] = mp::to_double(G(s));
}
void scalar_program(Ctx& ctx, int nl, int prog, mp::Tensor scal, int* d_flags, double* d_out, const int* d_status,
                    int n_status) {
  DISPATCH_NL(nl, {
    int tk = ctx.begin("scalar_program");
    scalar_kernel<NL><<<1, 32, 0, ctx.stream>>>(prog, scal, d_flags, d_out, d_status, n_status);
    ctx.end(tk);
  });
}

}  // namespace clr

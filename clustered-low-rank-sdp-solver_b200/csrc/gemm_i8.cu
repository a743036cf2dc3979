// gemm_i8.cu — sliced int8 GEMM on the sm_100a tensor cores (tcgen05 kind::i8, TMA, TMEM).
//
// GPU replacement of `Arblib.approx_mul!` (18 call sites in MPMP.jl). Like libarb's own
// block algorithm it works in block fixed point: every row of either operand is scaled to a common
// exponent and cut into S = p/8 + GUARD balanced radix-256 digits (int8); the product is then a sum of
// exact int8 x int8 -> int32 matrix products
//        plane_t = sum_{a+b=t} A_a * B_b^T ,        C = 2^(eA_i+eB_j) * sum_t plane_t * 256^-t
// truncated to the T = S most significant planes, followed by carry propagation and renormalisation.
#include "gemm_i8.cuh"

#include <algorithm>
#include <cstring>

namespace clr {

// =====================================================================================================
// 1. slicing: mp rows -> exponent per row + balanced int8 digits
// =====================================================================================================
struct GatherArgs {
  const uint32_t* w;
  size_t n;
  const int64_t* d_off;
  int64_t off0, bstride, rs, ks;
  int batch, rows, K, Kp, rows_total;
  const int* kshift;  // optional [batch][K]: entry (b, r, k) is taken as x * 2^-kshift[b*K + k] (exact)
  const int* ksign;   // optional [batch][ksign_ld]: entry (b, r, k) is negated where ksign[b*ksign_ld + k] != 0
  int ksign_ld;
};

__device__ __forceinline__ int64_t item_off(const GatherArgs& g, int b) {
  return g.off0 + (g.d_off ? g.d_off[b] : (int64_t)b * g.bstride);
}

// Fixed-point conversion of one entry relative to its row exponent and balanced radix-256 digits:
//   F = trunc( x * 2^(8S-2-rexp) ),  |F| < 2^(8S-2);   F = sum_a d_a 256^(S-1-a),  d_a in [-128,127]
// The balanced digits need no digit-serial carry loop: with M = 0x80 in each of the S byte positions,
//   byte_i(F + M) = (b_i + c_i + 128) mod 256,   c_{i+1} = [b_i + c_i >= 128]   (the carry of the addition IS the
// balancing carry), so d_i = byte_i(F + M) - 128 = byte_i((F + M) xor M) read as int8: one multiword add and one xor.
template <int NL, int S>
__device__ __forceinline__ void fixed_point_digits(const GatherArgs& g, int b, int64_t at, int k, bool in_range, int32_t rexp,
                                                   uint32_t (&W)[NL + 2]) {
  constexpr int NLW = NL + 2;
  static_assert(8 * S + 2 <= 32 * NLW, "digit window too small");
#pragma unroll
  for (int i = 0; i < NLW; i++) W[i] = 0;
  bool negf = false;
  if (in_range) {
    mp::Num<NL> x = mp::load<NL>(g.w, g.n, (size_t)at);
    if (!mp::is_zero(x)) {
      if (g.kshift) x.e -= g.kshift[b * g.K + k];
      uint32_t d = (uint32_t)(rexp - x.e);
      uint32_t sr = 32u * NLW - 8u * S + 2u + d;
      if (sr < 32u * NLW) {
#pragma unroll
        for (int i = 0; i < NL; i++) W[i + 2] = x.m[i];
        if (sr >> 5) mp::shr_limbs<NLW>(W, sr >> 5);
        if (sr & 31u) mp::shr_bits<NLW>(W, sr & 31u);
        negf = (x.neg != 0) != (g.ksign != nullptr && g.ksign[(int64_t)b * g.ksign_ld + k] != 0);
      }
    }
  }
  if (negf) mp::neg_n<NLW>(W);  // two's complement
  uint32_t M[NLW];
#pragma unroll
  for (int i = 0; i < NLW; i++) M[i] = (4 * i + 4 <= S) ? 0x80808080u : ((4 * i < S) ? (0x80808080u >> (8 * (4 - (S - 4 * i)))) : 0u);
  mp::add_n<NLW>(W, M);
#pragma unroll
  for (int i = 0; i < NLW; i++) W[i] ^= M[i];
}

// Four consecutive entries k4 .. k4+3 of one row: their digits are packed into one 32-bit word per digit plane (a 4 x 4
// byte transpose of the four entries' words by PRMT), so a warp writes 128 contiguous bytes per plane and store
// instruction instead of 32.
template <int NL, int S>
__device__ __forceinline__ void slice_entry4(const GatherArgs& g, int row, int k4, int32_t rexp, int8_t* __restrict__ digits) {
  constexpr int NLW = NL + 2;
  uint32_t W0[NLW], W1[NLW], W2[NLW], W3[NLW];
  const int b = row / g.rows, r = row % g.rows;
  const int64_t base = item_off(g, b) + (int64_t)r * g.rs;
  fixed_point_digits<NL, S>(g, b, base + (int64_t)(k4 + 0) * g.ks, k4 + 0, k4 + 0 < g.K, rexp, W0);
  fixed_point_digits<NL, S>(g, b, base + (int64_t)(k4 + 1) * g.ks, k4 + 1, k4 + 1 < g.K, rexp, W1);
  fixed_point_digits<NL, S>(g, b, base + (int64_t)(k4 + 2) * g.ks, k4 + 2, k4 + 2 < g.K, rexp, W2);
  fixed_point_digits<NL, S>(g, b, base + (int64_t)(k4 + 3) * g.ks, k4 + 3, k4 + 3 < g.K, rexp, W3);
  uint32_t* o = reinterpret_cast<uint32_t*>(digits + ((size_t)(S - 1) * g.rows_total + row) * g.Kp + k4);
  const size_t pstride4 = ((size_t)g.rows_total * g.Kp) >> 2;
#pragma unroll
  for (int wi = 0; 4 * wi < S; wi++) {
    const uint32_t lo01 = __byte_perm(W0[wi], W1[wi], 0x5140), hi01 = __byte_perm(W0[wi], W1[wi], 0x7362);
    const uint32_t lo23 = __byte_perm(W2[wi], W3[wi], 0x5140), hi23 = __byte_perm(W2[wi], W3[wi], 0x7362);
    *o = __byte_perm(lo01, lo23, 0x5410);
    o -= pstride4;
    if (4 * wi + 1 < S) { *o = __byte_perm(lo01, lo23, 0x7632); o -= pstride4; }
    if (4 * wi + 2 < S) { *o = __byte_perm(hi01, hi23, 0x5410); o -= pstride4; }
    if (4 * wi + 3 < S) { *o = __byte_perm(hi01, hi23, 0x7632); o -= pstride4; }
  }
}

// Row exponent + slicing in one launch: pass 1 takes the maximum exponent over the row's non-zero entries (EXP_ZERO for an
// all-zero row), pass 2 converts the row. Every thread converts FOUR consecutive entries of one row and stores their
// digits as one packed 32-bit word per digit plane (slice_entry4): one-entry-per-thread slicing spent a third of its
// ~700 instructions per entry on 34 single-byte stores and their address arithmetic.
//   wpr >= 1: a block of 8 warps owns 8/wpr rows (wpr warps per row; long rows, or few rows);
//   wpr == 0: SHORT rows (Kp <= 128): lpr = Kp/4 lanes per row, 32/lpr rows per warp, the exponent maximum by a segmented
//             shuffle - no shared memory, no block barrier.
constexpr int SLICE_THREADS = 256;
template <int NL, int S>
__global__ void __launch_bounds__(SLICE_THREADS, (NL <= 8 ? 4 : 2)) slice_rows_kernel(GatherArgs g, int wpr, int32_t* __restrict__ exps,
                                                                    int8_t* __restrict__ digits) {
  __shared__ int32_t smx[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t* hdr = g.w + (size_t)NL * g.n;
  if (wpr == 0) {
    const int lpr = g.Kp >> 2, rpw = 32 / lpr;           // Kp in {32, 64, 128}: 8, 16 or 32 lanes per row
    const int sub = lane & (lpr - 1), k4 = 4 * sub;
    const int nwarps = (g.rows_total + rpw - 1) / rpw;
    for (int wg = blockIdx.x * 8 + warp; wg < nwarps; wg += gridDim.x * 8) {
      const int row = wg * rpw + lane / lpr;
      const bool live = row < g.rows_total;
      int32_t mx = mp::EXP_ZERO;
      if (live) {
        const int b = row / g.rows, r = row - b * g.rows;
        const int64_t base = item_off(g, b) + (int64_t)r * g.rs;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int k = k4 + i;
          if (k < g.K) {
            int32_t e = ((int32_t)hdr[base + (int64_t)k * g.ks]) >> 1;
            if (e != mp::EXP_ZERO && g.kshift) e -= g.kshift[b * g.K + k];
            mx = max(mx, e);
          }
        }
      }
      for (int o = lpr >> 1; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (live) {
        if (sub == 0) exps[row] = mx;
        slice_entry4<NL, S>(g, row, k4, mx, digits);
      }
    }
    return;
  }
  const int rpb = 8 / wpr, sub = warp % wpr;
  const int ngroups = (g.rows_total + rpb - 1) / rpb;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int row = grp * rpb + warp / wpr;
    const bool live = row < g.rows_total;
    int32_t mx = mp::EXP_ZERO;
    int64_t base = 0;
    int b = 0;
    if (live) {
      b = row / g.rows;
      const int r = row % g.rows;
      base = item_off(g, b) + (int64_t)r * g.rs;
      if (g.kshift) {
        for (int k = sub * 32 + lane; k < g.K; k += 32 * wpr) {
          int32_t e = ((int32_t)hdr[base + (int64_t)k * g.ks]) >> 1;
          if (e != mp::EXP_ZERO) mx = max(mx, e - g.kshift[b * g.K + k]);
        }
      } else {
        for (int k = sub * 32 + lane; k < g.K; k += 32 * wpr) mx = max(mx, ((int32_t)hdr[base + (int64_t)k * g.ks]) >> 1);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (wpr > 1) {
      __syncthreads();
      if (lane == 0) smx[warp] = mx;
      __syncthreads();
      for (int q = 0; q < wpr; q++) mx = max(mx, smx[(warp / wpr) * wpr + q]);
    }
    if (live) {
      if (sub == 0 && lane == 0) exps[row] = mx;
      for (int k4 = 4 * (sub * 32 + lane); k4 < g.Kp; k4 += 128 * wpr) slice_entry4<NL, S>(g, row, k4, mx, digits);
    }
  }
}

// =====================================================================================================
// 2. the tensor-core kernel
// =====================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// TMEM <-> registers for CPS = 4, 8 or 16 consecutive columns of this thread's lane
template <int CPS>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[CPS]);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
template <int CPS>
__device__ __forceinline__ void tmem_st_zero(uint32_t taddr);
template <>
__device__ __forceinline__ void tmem_st_zero<16>(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_st_zero<8>(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
}
template <>
__device__ __forceinline__ void tmem_st_zero<4>(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
}

// one lane of a converged warp; everything around the elected instruction stays warp-uniform, so descriptors and
// barrier addresses live in uniform registers (issuing from `if (lane == 0)` makes the compiler wrap every tcgen05/TMA
// instruction in a lane-serialising R2UR loop, ~250 cycles per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u)
      : "memory");
}

struct MmaParams {
  int T, M, N, batch, nsplit, Kp, Kc, BK, BN, m_tiles, n_tiles, item0, stages;
  int stack;  // digits of A stacked along the 128 MMA rows (1, 2 or 4): item tiles of 128/stack rows
  int Mpad;   // rows per item in the block buffer (m_tiles * 128)
  const int* rowA;
  const int* rowB;
  const int2* tile_map;  // symmetric product (A == B): the (mt, nt) tiles that touch the upper triangle; nullptr = all tiles
  int n_tile_pairs;
  int32_t* blocks;  // [T][nsplit][batch][ceil(N/4)][Mpad][4]: block q, lane group r holds the partial plane q + r
  uint32_t idesc, sbo16, layout_type;
  uint32_t idesc2;  // instruction descriptor of the N = 2 BN MMA over a pair of adjacent B digit tiles (0 = do not pair)
  int debug;  // CLRSDP_MMA_DEBUG (measuring aid): 1 = skip the block stores, 2 = skip the TMEM loads, 32 = cycle counters
  long long* dbg;
  // FUSED epilogue (single K-split): the epilogue warps propagate the carries themselves while the planes arrive (least
  // significant first) and finish the numbers (normalise, round, epilogue operation, store): no int32 planes in HBM, no
  // carry launch. `blocks` then holds the raw two's-complement words [NW][batch][ceil(N/4)][Mpad][4] of every lane group.
  const int32_t* expA;  // per row of A / B (global row index)
  const int32_t* expB;
  uint32_t* cw;         // destination tensor (planar mp) and the extra operand of the epilogue operation
  size_t cn;
  const uint32_t* ew;
  size_t en;
  const int64_t* c_off;
  int64_t c_off0, c_bstride, c_rs, c_cs;
  int epi;
};

constexpr int EPI_SETS = 4;       // epilogue warp sets: block q is drained by set q % EPI_SETS
constexpr int NISSUE = 2;         // MMA-issuing warps: stage s is issued by warp 1 + s % NISSUE
constexpr int FIRST_EPI = 1 + NISSUE;
constexpr int MMA_THREADS = 32 * (FIRST_EPI + 4 * EPI_SETS);  // warp 0: TMA producer, warps 1..NISSUE: MMA issuers, then the epilogue sets
constexpr int MAX_STAGES = 8;
constexpr int DG = 4;             // digits per operand group: one pipeline stage feeds a DG x DG square of digit pairs
constexpr int ACC_SLOTS = 8;      // accumulator slots in TMEM (block q lives in slot q % ACC_SLOTS)

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t sbo16, uint32_t layout_type) {
  // K-major canonical layout (cute::UMMA::SmemDescriptor): start>>4 | LBO[16,30) | SBO[32,46) | version=1 at
  // [46,48) | layout_type [61,64)
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)sbo16 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// Issuer `which` of NISSUE: the warps share the stages round-robin, so the barrier wait, the descriptor set-up and
// the commit of one stage (~600 cycles on this single-lane instruction stream) run while the tensor pipe works through
// the MMAs of the other warp's stage. Every MMA accumulates (the epilogue hands the accumulators back zeroed), so the
// order in which the two warps' MMAs reach the pipe does not matter: the sums are integers.
template <int STACK>
__device__ __forceinline__ void mma_issue(const MmaParams& p, uint64_t* bars, uint8_t* smem, uint32_t stage_bytes,
                                          uint32_t a_tile, uint32_t b_tile, uint32_t a_bytes, uint32_t tmem_base,
                                          int kblocks, int which) {
  const int T = p.T, Dmax = (T - 1) / DG;
  constexpr int QSPAN = (DG - STACK) + (DG - 1);
  const uint32_t a_step = a_tile >> 4, b_step = b_tile >> 4;  // descriptor address units (16 bytes)
  const bool two_k = p.BK > 32;
  int stage = 0, turn = 0;
  uint32_t phase = 0;
  long long t_full = 0, t_slot = 0, t_start = clock64();
  mbar_wait(smem_u32(&bars[32]), 0u);  // accumulators zeroed
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int D = Dmax; D >= 0; D--) {
    const int sb = (DG * D) & (ACC_SLOTS - 1);  // slot of block 4D
    const int cmax = T - 1 - DG * D;            // blocks 4D + c with c > cmax do not exist
    // The new blocks 4D .. 4D+3 of this anti-diagonal need their slots back from the epilogue; those of 4D .. 4D+2 were
    // handed over for draining only at the end of the previous anti-diagonal. A warp's first stage of D therefore
    // issues its pairs into the older slots first and waits for the fresh ones in the middle of the stage. Both warps
    // wait, also a warp without a stage in this D: its commits below must not overtake the other warp's commits of the
    // slot's previous block.
    bool need_wait = true;
    auto wait_slots = [&]() {
      const long long cs0 = p.dbg ? clock64() : 0;
      for (int c = 0; c <= min(DG - 1, cmax); c++) {
        const uint32_t u = (uint32_t)(T - 1 - (DG * D + c)) >> 3;  // earlier blocks in the same slot
        if (u) mbar_wait(smem_u32(&bars[24 + ((sb + c) & (ACC_SLOTS - 1))]), (u & 1u) ^ 1u);
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (p.dbg) t_slot += clock64() - cs0;
    };
    for (int I = 0; I <= D; I++) {
      const int J = D - I;
      if (DG * I >= T || DG * J >= T) continue;
      const int na = min(DG / STACK, (T - DG * I + STACK - 1) / STACK), nb = min(DG, T - DG * J);
      for (int kb = 0; kb < kblocks; kb++) {
        if (turn == which) {
          const long long c0 = p.dbg ? clock64() : 0;
          mbar_wait(smem_u32(&bars[stage]), phase);
          if (p.dbg) t_full += clock64() - c0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc0 = make_smem_desc(sa, p.sbo16, p.layout_type);
          const uint64_t bdesc0 = make_smem_desc(sa + a_bytes, p.sbo16, p.layout_type);
          const bool full = na == DG / STACK && nb == DG && cmax >= QSPAN && !need_wait;
          if (full && p.idesc2) {
            // interior stage, B digits taken in PAIRS along N: the tiles of digits jb and jb + 1 are adjacent in shared
            // memory (one TMA box, digit-major) and their blocks c, c + 1 adjacent in TMEM, so one MMA of N = 2 BN does
            // both. An SS-mode kind::i8 MMA reads 128 x 32 bytes of A and N x 32 bytes of B from shared memory at 128 B/clk:
            // at N = 64 that is 48 cycles for 32 cycles of arithmetic, at N = 128 the two are balanced (64 / 64). Only
            // the pair that would straddle the end of the accumulator ring (slot 7 | slot 0) is issued as two halves.
            if (elect_one()) {
#pragma unroll
              for (int ia = 0; ia < DG / STACK; ia++) {
#pragma unroll
                for (int jb = 0; jb < DG; jb += 2) {
                  const int s0 = (sb + ia * STACK + jb) & (ACC_SLOTS - 1);
                  const uint32_t d_tmem = tmem_base + (uint32_t)(s0 * p.BN);
                  const uint64_t ad = adesc0 + (uint64_t)(ia * a_step), bd = bdesc0 + (uint64_t)(jb * b_step);
                  if (s0 != ACC_SLOTS - 1) {
                    umma_i8(d_tmem, ad, bd, p.idesc2, 1u);
                    if (two_k) umma_i8(d_tmem, ad + 2, bd + 2, p.idesc2, 1u);
                  } else {
                    umma_i8(d_tmem, ad, bd, p.idesc, 1u);
                    if (two_k) umma_i8(d_tmem, ad + 2, bd + 2, p.idesc, 1u);
                    umma_i8(tmem_base, ad, bd + b_step, p.idesc, 1u);
                    if (two_k) umma_i8(tmem_base, ad + 2, bd + b_step + 2, p.idesc, 1u);
                  }
                }
              }
              umma_commit(smem_u32(&bars[8 + stage]));
            }
          } else if (full) {  // interior stage (the common case): no predicates
            if (elect_one()) {
#pragma unroll
              for (int ia = 0; ia < DG / STACK; ia++) {
#pragma unroll
                for (int jb = 0; jb < DG; jb++) {
                  const uint32_t d_tmem = tmem_base + (uint32_t)(((sb + ia * STACK + jb) & (ACC_SLOTS - 1)) * p.BN);
                  const uint64_t ad = adesc0 + (uint64_t)(ia * a_step), bd = bdesc0 + (uint64_t)(jb * b_step);
                  umma_i8(d_tmem, ad, bd, p.idesc, 1u);
                  if (two_k) umma_i8(d_tmem, ad + 2, bd + 2, p.idesc, 1u);
                }
              }
              umma_commit(smem_u32(&bars[8 + stage]));
            }
          } else {
            // my first stage of this D: first the pairs into block 4D+3 and above (a slot that has been free for a
            // whole anti-diagonal, or blocks carried over), then - after the slot wait - those into 4D .. 4D+2
            const bool split = need_wait;
            if (split) {
              const uint32_t u3 = (uint32_t)(T - 1 - (DG * D + DG - 1)) >> 3;
              if (DG - 1 <= cmax && u3) mbar_wait(smem_u32(&bars[24 + ((sb + DG - 1) & (ACC_SLOTS - 1))]), (u3 & 1u) ^ 1u);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (elect_one()) {
#pragma unroll
              for (int ia = 0; ia < DG / STACK; ia++) {
#pragma unroll
                for (int jb = 0; jb < DG; jb++) {
                  const int c = ia * STACK + jb;  // block 4D + c, lane group r holds plane 4D + c + r
                  if (ia < na && jb < nb && c <= cmax && (!split || c >= DG - 1)) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(((sb + c) & (ACC_SLOTS - 1)) * p.BN);
                    const uint64_t ad = adesc0 + (uint64_t)(ia * a_step), bd = bdesc0 + (uint64_t)(jb * b_step);
                    umma_i8(d_tmem, ad, bd, p.idesc, 1u);
                    if (two_k) umma_i8(d_tmem, ad + 2, bd + 2, p.idesc, 1u);
                  }
                }
              }
            }
            __syncwarp();
            if (split) {
              wait_slots();
              need_wait = false;
              if (elect_one()) {
#pragma unroll
                for (int ia = 0; ia < DG / STACK; ia++) {
#pragma unroll
                  for (int jb = 0; jb < DG; jb++) {
                    const int c = ia * STACK + jb;
                    if (ia < na && jb < nb && c <= cmax && c < DG - 1) {
                      const uint32_t d_tmem = tmem_base + (uint32_t)(((sb + c) & (ACC_SLOTS - 1)) * p.BN);
                      const uint64_t ad = adesc0 + (uint64_t)(ia * a_step), bd = bdesc0 + (uint64_t)(jb * b_step);
                      umma_i8(d_tmem, ad, bd, p.idesc, 1u);
                      if (two_k) umma_i8(d_tmem, ad + 2, bd + 2, p.idesc, 1u);
                    }
                  }
                }
              }
              __syncwarp();
            }
            if (elect_one()) umma_commit(smem_u32(&bars[8 + stage]));  // frees the smem slot when these MMAs retire
          }
          __syncwarp();
        }
        if (++turn == NISSUE) turn = 0;
        if (++stage == p.stages) stage = 0, phase ^= 1u;
      }
    }
    if (need_wait) wait_slots();  // a warp without a stage in this D (see above)
    // blocks that no later anti-diagonal touches are complete once BOTH warps' MMAs have retired: each warp commits
    const int q_hi = min(T - 1, DG * D + QSPAN);
    const int q_lo = (D == 0) ? 0 : DG * D + QSPAN - (DG - 1);
    if (elect_one())
      for (int q = q_hi; q >= q_lo; q--) umma_commit(smem_u32(&bars[16 + (q & (ACC_SLOTS - 1))]));
    __syncwarp();
  }
  if (p.dbg && blockIdx.x == 0 && which == 0 && (threadIdx.x & 31) == 0) {
    p.dbg[0] = clock64() - t_start;
    p.dbg[1] = t_full;
    p.dbg[6] = t_slot;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Epilogue of one warp: columns [eset*CPS, (eset+1)*CPS) of the tile, TMEM lanes [32*q4, 32*q4 + 32), all blocks.
//
// !FUSED: the int32 blocks go to HBM as they are (chunk-major [N/4][Mpad][4]); carry_kernel finishes them (K-splits,
//         symmetric products, the exact-plane test entry).
// FUSED:  per entry a running 64-bit carry and the 32-bit word under construction stay in registers: plane t (byte
//         position T-1-t from the least significant end) contributes byte (carry + v) & 0xFF, carry <- (carry + v) >> 8.
//         A word is final as soon as its four bytes are (carries only travel upwards) and is streamed to the raw-word
//         buffer (L2-resident: 4*NW bytes per entry and lane group instead of 4*T). After the last plane the carry is
//         spelled out into the words above (sign extension). Then the same threads finish the tile: sum of the lane groups
//         of an entry (M-stacking: group r holds the planes of digit a + r, processed one byte position later), two's
//         complement sign, normalisation, rounding to p bits, exponent, epilogue operation, store - what carry_kernel does,
//         without the round trip of the planes through HBM and without the extra launch.
// ---------------------------------------------------------------------------------------------------------------------
template <int NL, int CPS, bool FUSED>
__device__ __forceinline__ void epilogue_warp(const MmaParams& p, uint64_t* bars, uint32_t tmem_base, int q4, int eset, int lane,
                                              int mt, int nt, int item, int split) {
  constexpr int TT = 4 * NL + GUARD_DIGITS;     // digit planes of this precision (= p.T)
  constexpr int NW = (TT + 3) / 4 + 2;          // words of the raw two's-complement result
  const int T = p.T, stack = p.stack;
  const int qspan = (DG - stack) + (DG - 1);
  const int Dmax = (T - 1) / DG;
  const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
  {  // hand every accumulator to the issuers zeroed: each warp clears its own columns of all slots
    for (int sl = 0; sl < ACC_SLOTS; sl++) tmem_st_zero<CPS>(tmem_base + lane_addr + (uint32_t)(sl * p.BN + eset * CPS));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars[32]));
  }
  const int lrow = q4 * 32 + lane;
  const int mip = 128 / stack;
  const int grp = lrow / mip;           // lane group: plane q + grp
  const bool row_ok = (stack == 1) ? (mt * 128 + lrow < p.M) : ((lrow - grp * mip) < p.M);
  const int row = mt * 128 + lrow;
  const int n4 = (p.N + 3) >> 2;
  const int chunk0 = (nt * p.BN + eset * CPS) >> 2;   // first 4-column chunk of this thread
  // acc[c] = carry + sum of the planes of the word under construction, plane at byte k weighted 2^(8k): one IMAD.WIDE per
  // plane and entry (|v| < 2^31, so |acc| < 2^56); when the word's four bytes are in, its low half is final
  int64_t acc[CPS];
#pragma unroll
  for (int c = 0; c < CPS; c++) acc[c] = 0;
  // raw words of (word w, this item): int4 index ((w * batch + item) * n4 + chunk) * Mpad + row
  int4* const wbase = reinterpret_cast<int4*>(p.blocks) + (size_t)item * n4 * p.Mpad + row;
  const size_t wstride = (size_t)p.batch * n4 * p.Mpad;
  auto flush = [&](int w) {
    if (row_ok && !(p.debug & 1)) {
#pragma unroll
      for (int i = 0; i < CPS / 4; i++)
        if (chunk0 + i < n4)
          wbase[(size_t)w * wstride + (size_t)(chunk0 + i) * p.Mpad] =
              make_int4((int)(uint32_t)acc[4 * i], (int)(uint32_t)acc[4 * i + 1], (int)(uint32_t)acc[4 * i + 2], (int)(uint32_t)acc[4 * i + 3]);
    }
#pragma unroll
    for (int c = 0; c < CPS; c++) acc[c] >>= 32;  // arithmetic: the carry into the next word
  };
  long long dbg_wait = 0;
  const long long dbg_t0 = p.dbg ? clock64() : 0;
  for (int D = Dmax; D >= 0; D--) {
    const int q_hi = min(T - 1, DG * D + qspan);
    const int q_lo = (D == 0) ? 0 : DG * D + qspan - (DG - 1);
    for (int q = q_hi; q >= q_lo; q--) {
      const int slot = q & (ACC_SLOTS - 1);
      const uint32_t u = (uint32_t)(T - 1 - q) >> 3;
      const long long cw0 = p.dbg ? clock64() : 0;
      mbar_wait(smem_u32(&bars[16 + slot]), u & 1u);
      if (p.dbg) dbg_wait += clock64() - cw0;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(slot * p.BN + eset * CPS);
      uint32_t v[CPS];
      if (!(p.debug & 2)) {
        tmem_ld<CPS>(taddr, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int c = 0; c < CPS; c++) v[c] = 0;
      }
      // the accumulator slot goes back as soon as it has been read and cleared
      tmem_st_zero<CPS>(taddr);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[24 + slot]));
      if (!FUSED) {
        // chunk-major block layout [N/4][Mpad][4]: the 32 rows of a warp are 512 contiguous bytes per 4-column chunk
        if (row_ok && (q + grp < T) && !(p.debug & 1)) {
          int4* out = reinterpret_cast<int4*>(p.blocks) + (((size_t)q * p.nsplit + split) * p.batch + item) * (size_t)n4 * p.Mpad + row;
#pragma unroll
          for (int i = 0; i < CPS / 4; i++)
            if (chunk0 + i < n4)
              out[(size_t)(chunk0 + i) * p.Mpad] = make_int4((int)v[4 * i], (int)v[4 * i + 1], (int)v[4 * i + 2], (int)v[4 * i + 3]);
        }
      } else {
        const int t = q + grp;  // the plane these rows of block q belong to
        if (t < T) {
          const int pos = T - 1 - t;
          const int32_t weight = 1 << (8 * (pos & 3));
#pragma unroll
          for (int c = 0; c < CPS; c++)  // one IMAD.WIDE each (nvcc turns the C++ product into a six-instruction 64-bit shift-add)
            asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"((int32_t)v[c]), "r"(weight));
          if ((pos & 3) == 3) flush(pos >> 2);
        }
      }
    }
  }
  if (p.dbg && blockIdx.x == 0 && eset == 0 && q4 == 0 && lane == 0) {
    p.dbg[8] = clock64() - dbg_t0;  // epilogue: all blocks drained
    p.dbg[9] = dbg_wait;            // of which waiting for complete blocks
  }
  if (!FUSED) return;
  // the remaining carry (signed) is spelled out in the words above the last plane (sign extension up to word NW - 1)
  // (the next byte position of this lane group is T - grp: the words below (T - grp) >> 2 are complete and flushed)
  for (int w = (T - grp) >> 2; w < NW; w++) flush(w);
  // every lane group's words of this tile are in memory: epilogue warps only (named barrier 1)
  asm volatile("bar.sync 1, %0;" ::"r"(32 * 4 * EPI_SETS) : "memory");
  if (p.debug & 1) return;
  // Finish: a warp takes 4 x 8 patches of the tile (its reads of the raw words are runs of 4 rows x 16 bytes in the
  // chunk-major layout, its result stores 4 runs of 8 consecutive entries per limb plane), one entry per lane.
  const int ewarp = 4 * eset + q4;                       // 0 .. 15
  const int rows_tile = (stack == 1) ? 128 : mip;        // output rows of this tile
  const int npc = p.BN >> 3, npatch = (rows_tile >> 2) * npc;
  const int gitem = p.item0 + item;
  const int rowA0 = (p.rowA ? p.rowA[gitem] : gitem * p.M), rowB0 = (p.rowB ? p.rowB[gitem] : gitem * p.N);
  const int64_t cbase = p.c_off0 + (p.c_off ? p.c_off[gitem] : (int64_t)gitem * p.c_bstride);
  const uint32_t* const raw = reinterpret_cast<const uint32_t*>(p.blocks) + (size_t)item * n4 * p.Mpad * 4;
  const size_t wstride4 = wstride * 4;
  // The raw words of the NEXT patch are requested before the current one is normalised and stored (two L2 round trips
  // in flight per thread: with 16 warps per SM the finish is bound by that latency, not by the arithmetic).
  auto entry_of = [&](int pt, int& il, int& i, int& j) {
    il = 4 * (pt / npc) + (lane >> 3);  // row inside the tile
    i = mt * 128 + il;
    j = nt * p.BN + 8 * (pt % npc) + (lane & 7);
    return pt < npatch && i < p.M && j < p.N;
  };
  auto request = [&](int pt, uint32_t (&R0)[NW], uint32_t (&R1)[NW]) {
    int il, i, j;
    if (!entry_of(pt, il, i, j)) return;
    const uint32_t* e0 = raw + ((size_t)(j >> 2) * p.Mpad + (mt * 128 + il)) * 4 + (j & 3);
#pragma unroll
    for (int w = 0; w < NW; w++) R0[w] = __ldcg(e0 + (size_t)w * wstride4);
    if (stack > 1) {
#pragma unroll
      for (int w = 0; w < NW; w++) R1[w] = __ldcg(e0 + (size_t)w * wstride4 + (size_t)mip * 4);
    }
  };
  uint32_t N0[NW], N1[NW];
  request(ewarp, N0, N1);
  for (int pt = ewarp; pt < npatch; pt += 4 * EPI_SETS) {
    int il, i, j;
    const bool live = entry_of(pt, il, i, j);
    uint32_t W[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) W[w] = N0[w];
    if (live && stack > 1) {  // lane groups are numbers on the same scale: multiword sum mod 2^(32 NW)
      mp::add_n<NW>(W, N1);
      const uint32_t* e0 = raw + ((size_t)(j >> 2) * p.Mpad + (mt * 128 + il)) * 4 + (j & 3);
      for (int r = 2; r < stack; r++) {
        uint32_t V[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) V[w] = __ldcg(e0 + (size_t)w * wstride4 + (size_t)r * mip * 4);
        mp::add_n<NW>(W, V);
      }
    }
    request(pt + 4 * EPI_SETS, N0, N1);
    if (!live) continue;
    const bool negf = (W[NW - 1] >> 31) != 0;
    if (negf) {
      uint32_t cc = 1;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const uint64_t sm = (uint64_t)(~W[w]) + cc;
        W[w] = (uint32_t)sm;
        cc = (uint32_t)(sm >> 32);
      }
    }
    const int32_t ea = p.expA[rowA0 + i], eb = p.expB[rowB0 + j];
    mp::Num<NL> r;
    const int shn = mp::normalize_n<NW>(W);
    if (shn < 0 || ea == mp::EXP_ZERO || eb == mp::EXP_ZERO) {
      r = mp::zero<NL>();
    } else {
      uint32_t X[NL + 1];
#pragma unroll
      for (int w = 0; w <= NL; w++) X[w] = W[NW - NL - 1 + w];
      // value = N_int * 2^(ea + eb - 4 - 8T), N_int = (W / 2^(32 NW)) * 2^(32 NW - sh)
      mp::round_guard<NL>(r, X, 32 * NW - shn + ea + eb - 4 - 8 * T, negf ? 1u : 0u);
    }
    const int64_t at = cbase + (int64_t)i * p.c_rs + (int64_t)j * p.c_cs;
    if (p.epi == EPI_NEG) {
      r = mp::neg(r);
    } else if (p.epi != EPI_STORE) {
      mp::Num<NL> e = mp::load<NL>(p.ew, p.en, (size_t)at);
      if (p.epi == EPI_SUB_FROM)
        r = mp::sub(e, r);
      else if (p.epi == EPI_MINUS_SUB)
        r = mp::sub(r, e);
      else
        r = mp::add(r, e);
    }
    mp::store<NL>(p.cw, p.cn, (size_t)at, r);
  }
  if (p.dbg && blockIdx.x == 0 && eset == 0 && q4 == 0 && lane == 0) p.dbg[10] = clock64() - dbg_t0;  // epilogue incl. finish
}

// One CTA = one 128 x BN output tile of one batch item (x one K-split) and ALL its digit planes.
//
// The digit pairs (a, b), a + b < T, are visited in DG x DG squares: a pipeline stage holds the tiles of DG digits of A
// and DG digits of B for one K block (two TMA boxes), and feeds DG*DG/stack MMAs, so every byte brought in from L2 is
// used DG/2 times more often than in a pair-at-a-time schedule (the L2 -> smem path, ~42 B/clk/SM, is what bounds a
// digit-pair product at 128 x 64 x 32). Squares are processed by anti-diagonals D = I + J from the least significant
// end; the pairs of one anti-diagonal touch the blocks q = a + b in [4D, 4D + 6], which live in a ring of 8 TMEM
// accumulators; when an anti-diagonal is finished its four lowest-order blocks are complete and are drained by the
// epilogue warps while the next anti-diagonal accumulates into the other slots.
//
// Items with at most 64 (32) rows stack 2 (4) consecutive digits of A along the 128 rows of the MMA: rows
// [r*128/stack, (r+1)*128/stack) of block q then hold the contribution of digit a + r, i.e. of plane q + r, and the
// carry kernel adds the `stack` lane groups. This keeps all 128 rows of the tensor core busy on the 64 x 64 blocks.
template <int NL, bool FUSED>
__global__ void __launch_bounds__(MMA_THREADS, 1)
mma_planes_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, MmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  if (p.dbg && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.dbg[blockIdx.x == 0 ? 2 : 4] = (long long)gt;
  }
  const uint32_t a_tile = 128u * p.BK, b_tile = (uint32_t)p.BN * p.BK;
  const uint32_t a_bytes = a_tile * (DG / p.stack), b_bytes = b_tile * DG;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  // bars[0..7] full, [8..15] empty, [16..23] slot_full, [24..31] slot_empty, [32] accumulators zeroed, then the TMEM
  // base pointer
  uint32_t* tmem_slot = (uint32_t*)(bars + 33);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // K-split slowest: the CTAs resident at the same time (consecutive block indices) then work on the SAME K range of
  // different tiles, i.e. they share the rows of A (same mt) and of B (same nt), and their combined working set per digit
  // group fits L2. With the split fastest the Q product of BASELINE config 5 (1024 x 1024 x 16384, 9 splits) read 160 GB
  // from DRAM per launch (5.3 TB/s, L2 hit rate 40 %): it was HBM-bound, not tensor-bound.
  int idx = blockIdx.x;
  const int tiles_all = (int)(gridDim.x / (unsigned)p.nsplit);
  const int split = idx / tiles_all;
  idx -= split * tiles_all;
  int nt, mt;
  if (p.tile_map) {
    const int2 t = p.tile_map[idx % p.n_tile_pairs];
    idx /= p.n_tile_pairs;
    mt = t.x, nt = t.y;
  } else {
    nt = idx % p.n_tiles;
    idx /= p.n_tiles;
    mt = idx % p.m_tiles;
    idx /= p.m_tiles;
  }
  const int item = idx;  // local item in this launch
  const int gitem = p.item0 + item;
  const int rowA0 = (p.rowA ? p.rowA[gitem] : gitem * p.M) + mt * 128;
  const int rowB0 = (p.rowB ? p.rowB[gitem] : gitem * p.N) + nt * p.BN;
  const int k_begin = split * p.Kc;
  const int kblocks = (min(p.Kp, k_begin + p.Kc) - k_begin) / p.BK;
  const int T = p.T, stack = p.stack;
  const int qspan = (DG - stack) + (DG - 1);  // blocks touched by anti-diagonal D: [4D, 4D + qspan]
  const int Dmax = (T - 1) / DG;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(ACC_SLOTS * p.BN)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; s++) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[8 + s]), 1);
    }
    for (int s = 0; s < ACC_SLOTS; s++) {
      mbar_init(smem_u32(&bars[16 + s]), NISSUE);
      mbar_init(smem_u32(&bars[24 + s]), 4 * EPI_SETS);
    }
    mbar_init(smem_u32(&bars[32]), 4 * EPI_SETS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int D = Dmax; D >= 0; D--)
        for (int I = 0; I <= D; I++) {
          const int J = D - I;
          if (DG * I >= T || DG * J >= T) continue;
          for (int kb = 0; kb < kblocks; kb++) {
            mbar_wait(smem_u32(&bars[8 + stage]), phase ^ 1u);
            uint32_t full = smem_u32(&bars[stage]);
            uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            int k0 = k_begin + kb * p.BK;
            if (elect_one()) {
              mbar_expect_tx(full, a_bytes + b_bytes);
              tma_load_3d(sa, &tmA, full, k0, rowA0, DG * I);            // box (BK, 128/stack rows, DG digits)
              tma_load_3d(sa + a_bytes, &tmB, full, k0, rowB0, DG * J);  // box (BK, BN rows, DG digits)
            }
            __syncwarp();
            if (++stage == p.stages) stage = 0, phase ^= 1u;
          }
        }
    }
  } else if (warp < FIRST_EPI) {
    // ------------------------------ MMA issuers -------------------------------
    // Each warp runs the (warp-uniform) schedule and one elected lane issues its stages. A tcgen05.mma of
    // 128 x 64 x 32 retires every 48 cycles and the issuing thread is throttled to that rate, so every other
    // instruction on this path would be idle tensor pipe if a single warp issued everything.
    if (stack == 1)
      mma_issue<1>(p, bars, smem, stage_bytes, a_tile, b_tile, a_bytes, tmem_base, kblocks, warp - 1);
    else if (stack == 2)
      mma_issue<2>(p, bars, smem, stage_bytes, a_tile, b_tile, a_bytes, tmem_base, kblocks, warp - 1);
    else
      mma_issue<4>(p, bars, smem, stage_bytes, a_tile, b_tile, a_bytes, tmem_base, kblocks, warp - 1);
  } else {
    // ------------------------------ epilogue: TMEM -> registers -> HBM --------------------------------
    // Four warp sets (one warp per TMEM lane quarter in each); set e owns the columns [e*BN/4, (e+1)*BN/4) of EVERY block,
    // so a thread sees all planes of its entries in the order they complete (least significant first).
    const int q4 = warp & 3;                      // TMEM lane quarter this warp may access
    const int eset = (warp - FIRST_EPI) >> 2;     // column group
    if (p.BN == 64)
      epilogue_warp<NL, 16, FUSED>(p, bars, tmem_base, q4, eset, lane, mt, nt, item, split);
    else if (p.BN == 32)
      epilogue_warp<NL, 8, FUSED>(p, bars, tmem_base, q4, eset, lane, mt, nt, item, split);
    else
      epilogue_warp<NL, 4, FUSED>(p, bars, tmem_base, q4, eset, lane, mt, nt, item, split);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
  if (p.dbg && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    p.dbg[blockIdx.x == 0 ? 3 : 5] = (long long)gt;
  }
}

// =====================================================================================================
// 3. carry propagation + renormalisation:  planes -> mp numbers
// =====================================================================================================
struct CarryArgs {
  const int32_t* planes;  // blocks [T][nsplit][batch][ceil(N/4)][Mpad][4]; plane p = sum_r block[p - r] rows [r*mip, r*mip + M)
  int stack, mip, Mpad;
  int sym_bn;  // > 0: symmetric product computed on the tiles touching the upper triangle only (tile width sym_bn)
  const int32_t* expA;    // per row of A (global row index)
  const int32_t* expB;
  const int* rowA;
  const int* rowB;
  int T, nsplit, batch, M, N, item0;
  uint32_t* cw;
  size_t cn;
  const uint32_t* ew;
  size_t en;
  const int64_t* d_off;
  int64_t off0, bstride, rs, cs;
  int epi;
};

// VAR 0: 17 planes in flight per load chunk (9 with four lane groups), registers as the compiler likes (128);
// VAR 1 (default up to 256 bits; CLRSDP_CARRY_VAR overrides): 9 planes per chunk under __launch_bounds__(128, 5): 20 warps per
// SM instead of 16.
template <int NL, int T, int STACK, int VAR>
__global__ void __launch_bounds__(128, (VAR && STACK != 4) ? 5 : 4) carry_kernel(CarryArgs c) {
  constexpr int NW = (T + 3) / 4 + 2;
  // a warp owns a 4 x 8 patch of one item: its plane reads are two runs of 4 rows x 16 bytes in the chunk-major block
  // layout, its result stores 4 runs of 8 consecutive entries
  const int pm = (c.M + 3) >> 2, pn = (c.N + 7) >> 3, n4 = (c.N + 3) >> 2;
  const int64_t total = (int64_t)c.batch * pm * pn * 32;
  const size_t sstride = (size_t)c.batch * n4 * c.Mpad * 4;  // one K-split of one block
  const size_t pstride = (size_t)c.nsplit * sstride;          // one block
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int ln = (int)(idx & 31);
    int64_t patch = idx >> 5;
    const int pj = (int)(patch % pn);
    patch /= pn;
    const int pi = (int)(patch % pm);
    const int b = (int)(patch / pm);
    const int i = 4 * pi + (ln >> 3), j = 8 * pj + (ln & 7);
    if (i >= c.M || j >= c.N) continue;
    int gb = c.item0 + b;
    // symmetric product: an entry whose tile was skipped is read from its mirror image (same planes, same exponents)
    int si = i, sj = j;
    if (c.sym_bn > 0 && (j / c.sym_bn) * c.sym_bn + c.sym_bn - 1 < (i >> 7) * 128) si = j, sj = i;
    const size_t e0 = (((size_t)b * n4 + (sj >> 2)) * c.Mpad + si) * 4 + (sj & 3);  // entry (b, si, sj) in lane group 0
    const size_t gstride = (size_t)c.mip * 4;                                      // next lane group
    uint32_t W[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) W[q] = 0;
    int64_t carry = 0;
    // planes from the least significant (t = T-1) upwards, CH at a time: the CH (x nsplit) loads of a chunk are
    // independent and all in flight before the carry chain consumes them
    constexpr int CH = (STACK == 4 || VAR) ? 9 : 17;  // (STACK = 4 with 17: 254 registers, 8 warps per SM)
#pragma unroll
    for (int c0 = T - 1; c0 >= 0; c0 -= CH) {
      int64_t acc[CH];
#pragma unroll
      for (int u = 0; u < CH; u++) acc[u] = 0;
      for (int s = 0; s < c.nsplit; s++)
#pragma unroll
        for (int r = 0; r < STACK; r++) {  // plane t = sum over lane groups r of block t - r
          const int32_t* ps = c.planes + (size_t)s * sstride + e0 + (size_t)r * gstride;
          int32_t v32[CH];
#pragma unroll
          for (int u = 0; u < CH; u++) v32[u] = (c0 - u - r >= 0) ? ps[(size_t)(c0 - u - r >= 0 ? c0 - u - r : 0) * pstride] : 0;
#pragma unroll
          for (int u = 0; u < CH; u++) acc[u] += v32[u];
        }
#pragma unroll
      for (int u = 0; u < CH; u++) {
        const int t = c0 - u;
        if (t >= 0) {
          int64_t v = carry + acc[u];
          const int pos = T - 1 - t;  // byte position from the least significant end
          W[pos >> 2] |= (uint32_t)(v & 0xFF) << (8 * (pos & 3));
          carry = v >> 8;
        }
      }
    }
    // remaining carry (signed) goes above byte T-1
    {
      constexpr int pos = T;
      uint64_t cu = (uint64_t)carry;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        int pp = pos + q;
        if ((pp >> 2) < NW) W[pp >> 2] |= (uint32_t)((cu >> (8 * q)) & 0xFF) << (8 * (pp & 3));
      }
      // sign-extend into any words above
      if (carry < 0) {
#pragma unroll
        for (int pp = pos + 8; pp < 4 * NW; pp++) W[pp >> 2] |= 0xFFu << (8 * (pp & 3));
      }
    }
    bool negf = carry < 0;
    if (negf) {
      uint32_t cc = 1;
#pragma unroll
      for (int q = 0; q < NW; q++) {
        uint64_t s = (uint64_t)(~W[q]) + cc;
        W[q] = (uint32_t)s;
        cc = (uint32_t)(s >> 32);
      }
    }
    int ra = (c.rowA ? c.rowA[gb] : gb * c.M) + si, rb = (c.rowB ? c.rowB[gb] : gb * c.N) + sj;
    int32_t ea = c.expA[ra], eb = c.expB[rb];
    mp::Num<NL> r;
    int sh = mp::normalize_n<NW>(W);
    if (sh < 0 || ea == mp::EXP_ZERO || eb == mp::EXP_ZERO) {
      r = mp::zero<NL>();
    } else {
      uint32_t X[NL + 1];
#pragma unroll
      for (int q = 0; q <= NL; q++) X[q] = W[NW - NL - 1 + q];
      // value = N_int * 2^(ea + eb - 4 - 8T), N_int = (W / 2^(32 NW)) * 2^(32 NW - sh)
      mp::round_guard<NL>(r, X, 32 * NW - sh + ea + eb - 4 - 8 * T, negf ? 1u : 0u);
    }
    int64_t at = c.off0 + (c.d_off ? c.d_off[gb] : (int64_t)gb * c.bstride) + (int64_t)i * c.rs + (int64_t)j * c.cs;
    if (c.epi == EPI_NEG) {
      r = mp::neg(r);
    } else if (c.epi != EPI_STORE) {
      mp::Num<NL> e = mp::load<NL>(c.ew, c.en, (size_t)at);
      if (c.epi == EPI_SUB_FROM)
        r = mp::sub(e, r);
      else if (c.epi == EPI_MINUS_SUB)
        r = mp::sub(r, e);
      else
        r = mp::add(r, e);
    }
    mp::store<NL>(c.cw, c.cn, (size_t)at, r);
  }
}

// =====================================================================================================
// host side
// =====================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CLR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw CudaError("cuTensorMapEncodeTiled not available");
    fn = (EncodeTiledFn)p;
  }
  return fn;
}
static CUtensorMap make_map(const Slice& s, int box_rows, int BK, int box_digits) {
  CUtensorMap tm;
  // rows past rows_total (tiles of the last item) are out of bounds => zero-filled by TMA
  cuuint64_t dims[3] = {(cuuint64_t)s.Kp, (cuuint64_t)s.rows_total, (cuuint64_t)s.S};
  cuuint64_t strides[2] = {(cuuint64_t)s.Kp, (cuuint64_t)s.Kp * s.rows_total};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, (cuuint32_t)box_digits};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = BK == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (BK == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, s.digits.p, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return tm;
}

typedef void (*MmaKernel)(const CUtensorMap, const CUtensorMap, MmaParams);
static MmaKernel mma_kernel_of(int nl, bool fused) {
  switch (nl) {
    case 4: return fused ? mma_planes_kernel<4, true> : mma_planes_kernel<4, false>;
    case 8: return fused ? mma_planes_kernel<8, true> : mma_planes_kernel<8, false>;
    case 12: return fused ? mma_planes_kernel<12, true> : mma_planes_kernel<12, false>;
    case 16: return fused ? mma_planes_kernel<16, true> : mma_planes_kernel<16, false>;
    default: throw SolverError(-1, "unsupported precision");
  }
}
GemmEngine::GemmEngine(Ctx& c, int nl) : ctx_(c), nl_(nl), S_(num_digits(nl)) {
  for (int f = 0; f < 2; f++)
    CLR_CUDA(cudaFuncSetAttribute(mma_kernel_of(nl, f != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (const char* e = getenv("CLRSDP_FUSED_CARRY")) fused_ok_ = atoi(e) != 0;
}

template <int NL>
static void slice_impl(Ctx& ctx, const OperandDesc& op, Slice& out) {
  constexpr int S = 4 * NL + GUARD_DIGITS;
  int K = op.K;
  int Kp = K <= 32 ? 32 : (K <= 64 ? 64 : ((K + 127) / 128) * 128);
  out.S = S;
  out.K = K;
  out.Kp = Kp;
  out.rows_item = op.rows;
  out.batch = op.batch;
  out.rows_total = op.rows * op.batch;
  size_t bytes = (size_t)S * out.rows_total * Kp;
  out.digits.ensure(bytes);
  out.exps.ensure(sizeof(int32_t) * (size_t)std::max(out.rows_total, 1));
  GatherArgs g{op.src.w, op.src.n, op.d_off, op.off0, op.bstride, op.rs, op.ks, op.batch, op.rows, K, Kp, out.rows_total, op.d_kshift, op.d_ksign, op.ksign_ld};
  int64_t total = (int64_t)out.rows_total * Kp;
  // long rows (or few of them) get a whole block per row, short rows one warp
  int wpr = (Kp > 256 || (int64_t)out.rows_total * 32 < (int64_t)ctx.sm_count * 256) ? 8 : 1;
  int grid;
  if (Kp <= 128) {   // short rows: Kp/4 lanes per row, no block barrier
    wpr = 0;
    const int rpw = 32 / (Kp / 4);
    grid = (int)std::min<int64_t>(ceil_div(ceil_div(out.rows_total, rpw), 8), (int64_t)ctx.sm_count * 8);
  } else {
    grid = (int)std::min<int64_t>(ceil_div(out.rows_total, 8 / wpr), (int64_t)ctx.sm_count * 8);
  }
  // algorithmic bytes: read (p/8+4) per entry, write S digit bytes
  int tk = ctx.begin("slice", (double)total * (4.0 * (NL + 1) + S));
  slice_rows_kernel<NL, S><<<grid, SLICE_THREADS, 0, ctx.stream>>>(g, wpr, out.exps.as<int32_t>(), out.digits.as<int8_t>());
  ctx.end(tk);
}

void GemmEngine::slice(const OperandDesc& op, Slice& out) {
  switch (nl_) {
    case 4: slice_impl<4>(ctx_, op, out); break;
    case 8: slice_impl<8>(ctx_, op, out); break;
    case 12: slice_impl<12>(ctx_, op, out); break;
    case 16: slice_impl<16>(ctx_, op, out); break;
    default: throw SolverError(-1, "unsupported precision");
  }
}

static int bk_cap() {
  static int v = -1;
  if (v < 0) v = 64;
  return v;
}
static int stack_of(int M) { return M <= 32 ? 4 : (M <= 64 ? 2 : 1); }
static int bn_of(int N) { return N <= 16 ? 16 : (N <= 32 ? 32 : 64); }
// Tile width of a product. A batch of small items that fills at most half of the SMs with 64-wide tiles is cut into
// 32-wide ones: the time of such a CTA is the drain of its T accumulator blocks (TMEM -> registers -> global), which
// halves with the tile width, and twice as many SMs work (cfg3: 64 items of 64 x 64 x 64 -> 128 CTAs on 148 SMs).
static int bn_for(const GemmPlan& plan, int sm_count, bool symmetric) {
  int bn = bn_of(plan.N);
  static int split = -1;
  if (split < 0) split = getenv("CLRSDP_BN_SPLIT") ? atoi(getenv("CLRSDP_BN_SPLIT")) : 1;
  if (split && bn == 64 && !symmetric) {
    int64_t tiles = (int64_t)plan.batch * ceil_div(plan.M, 128) * ceil_div(plan.N, 64);
    if (2 * tiles <= sm_count) bn = 32;
  }
  return bn;
}

void GemmEngine::run_mma(const Slice& A, const Slice& B, const GemmPlan& plan, int item0, int nitems, int nsplit,
                         int Kc, bool symmetric, const FusedOut* fo) {
  MmaParams p;
  memset(&p, 0, sizeof(p));
  if (fo) {
    if (nsplit != 1 || symmetric) throw SolverError(-1, "fused epilogue: single K-split, all tiles");
    p.expA = A.exps.as<int32_t>(), p.expB = B.exps.as<int32_t>();
    p.cw = fo->C.dst.w, p.cn = fo->C.dst.n;
    p.ew = fo->extra ? fo->extra->w : fo->C.dst.w, p.en = fo->extra ? fo->extra->n : fo->C.dst.n;
    p.c_off = fo->C.d_off, p.c_off0 = fo->C.off0, p.c_bstride = fo->C.bstride, p.c_rs = fo->C.rs, p.c_cs = fo->C.cs;
    p.epi = fo->epi;
  }
  p.T = S_;
  p.M = plan.M;
  p.N = plan.N;
  p.batch = nitems;
  p.nsplit = nsplit;
  p.Kp = A.Kp;
  p.Kc = Kc;
  p.BK = std::min(A.Kp, bk_cap());
  p.BN = bn_for(plan, ctx_.sm_count, symmetric);
  p.stack = stack_of(plan.M);
  p.m_tiles = ceil_div(plan.M, 128);
  p.Mpad = p.m_tiles * 128;
  p.n_tiles = ceil_div(plan.N, p.BN);
  p.item0 = item0;
  p.rowA = plan.d_rowA;
  p.rowB = plan.d_rowB;
  p.blocks = planes_.as<int32_t>();
  {
    static int dbg = -1;
    if (dbg < 0) dbg = getenv("CLRSDP_MMA_DEBUG") ? atoi(getenv("CLRSDP_MMA_DEBUG")) : 0;
    p.debug = dbg;
  }
  // instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) at [4,6), a/b format INT8 (1) at
  // [7,10)/[10,13), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
  p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  {
    static int pair = -1;
    if (pair < 0) pair = getenv("CLRSDP_PAIR_N") ? atoi(getenv("CLRSDP_PAIR_N")) : 1;
    p.idesc2 = pair ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * p.BN) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)) : 0u;
  }
  p.sbo16 = (8u * p.BK) >> 4;
  p.layout_type = p.BK == 128 ? 2u : (p.BK == 64 ? 4u : 6u);
  size_t a_bytes = (size_t)128 * p.BK * (DG / p.stack), b_bytes = (size_t)p.BN * p.BK * DG;
  size_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~(size_t)1023);
  // dynamic smem: [1 KB alignment slack][stages][1 KB of barriers]
  p.stages = (int)std::min<size_t>(MAX_STAGES, (196 * 1024) / stage_bytes);
  size_t smem = 1024 + (size_t)p.stages * stage_bytes + 1024;
  CUtensorMap tmA = make_map(A, 128 / p.stack, p.BK, DG), tmB = make_map(B, p.BN, p.BK, DG);
  int64_t tiles_per_item = (int64_t)p.m_tiles * p.n_tiles;
  if (symmetric) {
    std::vector<int2> map;
    for (int mt = 0; mt < p.m_tiles; mt++)
      for (int nt = 0; nt < p.n_tiles; nt++)
        if (nt * p.BN + p.BN - 1 >= mt * 128) map.push_back(make_int2(mt, nt));
    if (sym_map_host_ != map.size() * 1000003 + (size_t)p.m_tiles * 1009 + p.n_tiles) {  // (re)upload when the shape changes
      sym_map_.ensure(map.size() * sizeof(int2));
      CLR_CUDA(cudaMemcpyAsync(sym_map_.p, map.data(), map.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx_.stream));
      CLR_CUDA(cudaStreamSynchronize(ctx_.stream));
      sym_map_host_ = map.size() * 1000003 + (size_t)p.m_tiles * 1009 + p.n_tiles;
    }
    p.tile_map = sym_map_.as<int2>();
    p.n_tile_pairs = (int)map.size();
    tiles_per_item = (int64_t)map.size();
  }
  int64_t grid = (int64_t)nitems * tiles_per_item * nsplit;
  double macs = (double)nitems * p.m_tiles * 128.0 * p.n_tiles * p.BN * (double)A.Kp * (p.T * (p.T + 1) / 2.0) / p.stack;
  last_int8_macs += macs;
  // algorithmic MACs: M*N*K * s(s+1)/2 with s = p/8, no guard digits, no tile padding
  double s_alg = 4.0 * nl_;
  double alg = (double)nitems * (symmetric ? 0.5 * plan.M * (plan.M + 1.0) : (double)plan.M * plan.N) * (double)A.K * (s_alg * (s_alg + 1) / 2.0);
  std::string nm = std::string(fo ? "mma_planes_fused_M" : "mma_planes_M") + std::to_string(plan.M) + "_N" + std::to_string(plan.N) + "_K" + std::to_string(A.K) + "_b" + std::to_string(nitems);
  int tk = ctx_.begin(nm.c_str(), alg);
  static long long* d_dbg = nullptr;
  if (p.debug & 32) {
    if (!d_dbg) CLR_CUDA(cudaMalloc(&d_dbg, 128));
    CLR_CUDA(cudaMemsetAsync(d_dbg, 0, 128, ctx_.stream));
    p.dbg = d_dbg;
  }
  mma_kernel_of(nl_, fo != nullptr)<<<(unsigned)grid, MMA_THREADS, smem, ctx_.stream>>>(tmA, tmB, p);
  ctx_.end(tk);
  if (p.debug & 32) {
    long long h[16];
    CLR_CUDA(cudaStreamSynchronize(ctx_.stream));
    CLR_CUDA(cudaMemcpy(h, d_dbg, 128, cudaMemcpyDeviceToHost));
    fprintf(stderr, "[mma dbg] epilogue warp 0 of CTA 0: %lld cycles to drain all blocks (%lld waiting for complete blocks), %lld incl. the finish\n",
            h[8], h[9], h[10]);
    fprintf(stderr, "[mma dbg] %s grid=%lld stages=%d stack=%d BN=%d BK=%d: issuer cycles %lld, waiting for TMA %lld, for accumulator slots %lld; CTA0 life %lld ns, last CTA starts %+lld ns after CTA0 and lives %lld ns\n", nm.c_str(),
            (long long)grid, p.stages, p.stack, p.BN, p.BK, h[0], h[1], h[6], h[3] - h[2], h[4] - h[2], h[5] - h[4]);
  }
}

template <int NL>
static void carry_impl(Ctx& ctx, const CarryArgs& c) {
  constexpr int T = 4 * NL + GUARD_DIGITS;
  int64_t total = (int64_t)c.batch * c.M * c.N;
  int64_t threads = (int64_t)c.batch * ((c.M + 3) / 4) * ((c.N + 7) / 8) * 32;
  int grid = (int)std::min<int64_t>(ceil_div(threads, 128), (int64_t)ctx.sm_count * 32);
  // algorithmic bytes: read T int32 planes (x nsplit), write (p/8+4)
  int tk = ctx.begin("carry", (double)total * (4.0 * T * c.nsplit + 4.0 * (NL + 1)));
  // measured (r1g): VAR 1 is 5 % faster at 256 bits (cfg3: 1.52 -> 1.44 ms per iteration), 9 % slower at 512 bits (cfg5shard)
  static int var_env = -2;
  if (var_env == -2) var_env = getenv("CLRSDP_CARRY_VAR") ? atoi(getenv("CLRSDP_CARRY_VAR")) : -1;
  const int var = var_env >= 0 ? var_env : (NL <= 8 ? 1 : 0);
  if (var) {
    if (c.stack == 1)
      carry_kernel<NL, T, 1, 1><<<grid, 128, 0, ctx.stream>>>(c);
    else if (c.stack == 2)
      carry_kernel<NL, T, 2, 1><<<grid, 128, 0, ctx.stream>>>(c);
    else
      carry_kernel<NL, T, 4, 1><<<grid, 128, 0, ctx.stream>>>(c);
  } else if (c.stack == 1)
    carry_kernel<NL, T, 1, 0><<<grid, 128, 0, ctx.stream>>>(c);
  else if (c.stack == 2)
    carry_kernel<NL, T, 2, 0><<<grid, 128, 0, ctx.stream>>>(c);
  else
    carry_kernel<NL, T, 4, 0><<<grid, 128, 0, ctx.stream>>>(c);
  ctx.end(tk);
}

// K is cut into chunks for int32 exactness ((t+1) * Kc * 2^14 < 2^31  =>  T * Kc <= 131071) and, when there are fewer
// output tiles than SMs, for parallelism (the chunks of one tile run on different SMs and are summed by the carry kernel)
static void split_k(int sm_count, int T, int Kp, int BK, int64_t tiles, int& Kc, int& nsplit) {
  int kc_safe = (131071 / T) / BK * BK;
  Kc = std::min(Kp, kc_safe);
  if (tiles < sm_count && Kp >= 512) {
    // one wave: at most sm_count CTAs (one CTA per SM: 512 TMEM columns, ~200 KB smem), chunks of equal length
    int want = (int)std::min<int64_t>(sm_count / tiles, Kp / 256);
    if (want > 1) Kc = std::min(Kc, ceil_div(ceil_div(Kp, want), BK) * BK);
  }
  nsplit = ceil_div(Kp, Kc);
}

void GemmEngine::multiply(const Slice& A, const Slice& B, const GemmPlan& plan, const OutDesc& C, int epi,
                          const mp::Tensor* extra, bool symmetric) {
  if (symmetric && (A.rows_total != B.rows_total || plan.M != plan.N || plan.d_rowA || plan.d_rowB)) symmetric = false;
  if (A.Kp != B.Kp || A.S != S_ || B.S != S_) throw SolverError(-1, "gemm: operand mismatch");
  const int T = S_;
  const int BK = std::min(A.Kp, bk_cap()), BN = bn_for(plan, ctx_.sm_count, symmetric), stack = stack_of(plan.M);
  const int m_tiles = ceil_div(plan.M, 128), Mpad = m_tiles * 128;
  int64_t tiles = (int64_t)plan.batch * m_tiles * ceil_div(plan.N, BN);
  if (symmetric) {  // only the tiles that touch the upper triangle
    int cnt = 0;
    for (int mt = 0; mt < m_tiles; mt++)
      for (int nt = 0; nt < ceil_div(plan.N, BN); nt++) cnt += (nt * BN + BN - 1 >= mt * 128) ? 1 : 0;
    tiles = (int64_t)plan.batch * cnt;
  }
  int Kc, nsplit;
  split_k(ctx_.sm_count, T, A.Kp, BK, tiles, Kc, nsplit);
  // chunk the batch so that the block workspace stays bounded
  size_t per_item = (size_t)T * nsplit * Mpad * (((size_t)plan.N + 3) / 4 * 4) * sizeof(int32_t);
  size_t cap = (size_t)1536 << 20;
  int chunk = (int)std::max<size_t>(1, std::min<size_t>(plan.batch, cap / std::max<size_t>(per_item, 1)));
  planes_.ensure(per_item * chunk);
  // one K-split and no mirrored tiles: the epilogue of the tensor-core kernel finishes the numbers itself
  const bool fused = fused_ok_ && nsplit == 1 && !symmetric;
  for (int item0 = 0; item0 < plan.batch; item0 += chunk) {
    int n = std::min(chunk, plan.batch - item0);
    if (fused) {
      FusedOut fo{C, epi, extra};
      run_mma(A, B, plan, item0, n, nsplit, Kc, symmetric, &fo);
      continue;
    }
    run_mma(A, B, plan, item0, n, nsplit, Kc, symmetric);
    CarryArgs c;
    memset(&c, 0, sizeof(c));
    c.planes = planes_.as<int32_t>();
    c.sym_bn = symmetric ? BN : 0;
    c.stack = stack;
    c.mip = 128 / stack;
    c.Mpad = Mpad;
    c.expA = A.exps.as<int32_t>();
    c.expB = B.exps.as<int32_t>();
    c.rowA = plan.d_rowA;
    c.rowB = plan.d_rowB;
    c.T = T;
    c.nsplit = nsplit;
    c.batch = n;
    c.M = plan.M;
    c.N = plan.N;
    c.item0 = item0;
    c.cw = C.dst.w;
    c.cn = C.dst.n;
    c.ew = extra ? extra->w : C.dst.w;
    c.en = extra ? extra->n : C.dst.n;
    c.d_off = C.d_off;
    c.off0 = C.off0;
    c.bstride = C.bstride;
    c.rs = C.rs;
    c.cs = C.cs;
    c.epi = epi;
    switch (nl_) {
      case 4: carry_impl<4>(ctx_, c); break;
      case 8: carry_impl<8>(ctx_, c); break;
      case 12: carry_impl<12>(ctx_, c); break;
      case 16: carry_impl<16>(ctx_, c); break;
    }
  }
}

// ---- measured int8 tensor peak (the denominator of bench.py's roofline) ---------------------------------------------
// Every SM issues `iters` back-to-back tcgen05.mma kind::i8 128 x 256 x 32 on operands resident in shared memory (no
// loads, one accumulator, one commit at the end): the rate at which the tensor pipe retires int8 MACs, nothing else.
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x01010101u * (i & 1);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = make_smem_desc(smem_u32(smem), (8u * 128) >> 4, 2u);
    const uint64_t bdesc = make_smem_desc(smem_u32(smem) + 16384, (8u * 128) >> 4, 2u);
    for (int i = 0; i < iters; i++) {
      if (elect_one()) umma_i8(tmem_base, adesc, bdesc, idesc, 1u);
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}
double GemmEngine::measure_i8_peak() {
  const int iters = 4000;
  CLR_CUDA(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  cudaEvent_t a, b;
  CLR_CUDA(cudaEventCreate(&a));
  CLR_CUDA(cudaEventCreate(&b));
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    CLR_CUDA(cudaEventRecord(a, ctx_.stream));
    i8_peak_kernel<<<ctx_.sm_count, 128, 56 * 1024, ctx_.stream>>>(iters);
    CLR_CUDA(cudaEventRecord(b, ctx_.stream));
    CLR_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    CLR_CUDA(cudaEventElapsedTime(&ms, a, b));
    double macs = (double)ctx_.sm_count * iters * 128.0 * 256.0 * 32.0;
    if (rep > 0) best = std::max(best, macs / (ms * 1e-3));  // first launch = warm-up
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return best;
}

void GemmEngine::planes_only(const Slice& A, const Slice& B, const GemmPlan& plan, int32_t* h_planes, int* T_out) {
  const int T = S_;
  int BK = std::min(A.Kp, bk_cap());
  int kc_safe = (131071 / T) / BK * BK;
  if (A.Kp > kc_safe) throw SolverError(-1, "planes_only: K too large for a single split");
  const int stack = stack_of(plan.M), mip = 128 / stack, Mpad = ceil_div(plan.M, 128) * 128;
  const int n4 = (plan.N + 3) / 4;
  size_t bytes = (size_t)T * plan.batch * Mpad * n4 * 4 * sizeof(int32_t);
  planes_.ensure(bytes);
  run_mma(A, B, plan, 0, plan.batch, 1, A.Kp, false);
  std::vector<int32_t> blk(bytes / sizeof(int32_t));
  CLR_CUDA(cudaMemcpyAsync(blk.data(), planes_.p, bytes, cudaMemcpyDeviceToHost, ctx_.stream));
  ctx_.sync();
  // plane t = sum over the lane groups r of block t - r (exact integer sums)
  const size_t bstride = (size_t)plan.batch * n4 * Mpad * 4;
  for (int t = 0; t < T; t++)
    for (int b = 0; b < plan.batch; b++)
      for (int i = 0; i < plan.M; i++)
        for (int j = 0; j < plan.N; j++) {
          int32_t v = 0;
          for (int r = 0; r < stack && r <= t; r++)
            v += blk[(size_t)(t - r) * bstride + (((size_t)b * n4 + (j >> 2)) * Mpad + r * mip + i) * 4 + (j & 3)];
          h_planes[(((size_t)t * plan.batch + b) * plan.M + i) * plan.N + j] = v;
        }
  *T_out = T;
}

}  // namespace clr

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
};

__device__ __forceinline__ int64_t item_off(const GatherArgs& g, int b) {
  return g.off0 + (g.d_off ? g.d_off[b] : (int64_t)b * g.bstride);
}

// Fixed-point conversion of one entry relative to its row exponent and balanced radix-256 digits:
//   F = trunc( x * 2^(8S-2-rexp) ),  |F| < 2^(8S-2);   F = sum_a d_a 256^(S-1-a),  d_a in [-128,127]
template <int NL, int S>
__device__ __forceinline__ void slice_entry(const GatherArgs& g, int row, int k, int32_t rexp, int8_t* __restrict__ digits) {
  constexpr int NLW = NL + 2;
  static_assert(8 * S + 2 <= 32 * NLW, "digit window too small");
  uint32_t W[NLW];
#pragma unroll
  for (int i = 0; i < NLW; i++) W[i] = 0;
  bool negf = false;
  if (k < g.K) {
    int b = row / g.rows, r = row % g.rows;
    int64_t at = item_off(g, b) + (int64_t)r * g.rs + (int64_t)k * g.ks;
    mp::Num<NL> x = mp::load<NL>(g.w, g.n, (size_t)at);
    if (!mp::is_zero(x)) {
      uint32_t d = (uint32_t)(rexp - x.e);
      uint32_t sr = 32u * NLW - 8u * S + 2u + d;
      if (sr < 32u * NLW) {
#pragma unroll
        for (int i = 0; i < NL; i++) W[i + 2] = x.m[i];
        mp::shr_limbs<NLW>(W, sr >> 5);
        mp::shr_bits<NLW>(W, sr & 31u);
        negf = x.neg != 0;
      }
    }
  }
  if (negf) {  // two's complement
    uint32_t c = 1;
#pragma unroll
    for (int i = 0; i < NLW; i++) {
      uint64_t s = (uint64_t)(~W[i]) + c;
      W[i] = (uint32_t)s;
      c = (uint32_t)(s >> 32);
    }
  }
  int carry = 0;
  int8_t* out = digits + ((size_t)(S - 1) * g.rows_total + row) * g.Kp + k;
  const size_t pstride = (size_t)g.rows_total * g.Kp;
#pragma unroll
  for (int i = 0; i < S; i++) {
    int v = (int)((W[i >> 2] >> (8 * (i & 3))) & 0xFFu) + carry;
    carry = v >= 128 ? 1 : 0;
    v -= carry << 8;
    *out = (int8_t)v;
    out -= pstride;
  }
}

// Row exponent + slicing in one launch. A block of 8 warps owns 8/WPR rows (WPR warps per row): pass 1 takes the
// maximum exponent over the row's non-zero entries (EXP_ZERO for an all-zero row), pass 2 converts the row.
// The second read of the row hits L1/L2.
constexpr int SLICE_THREADS = 256;
template <int NL, int S>
__global__ void __launch_bounds__(SLICE_THREADS, 3) slice_rows_kernel(GatherArgs g, int wpr, int32_t* __restrict__ exps,
                                                                    int8_t* __restrict__ digits) {
  __shared__ int32_t smx[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rpb = 8 / wpr, sub = warp % wpr;
  const uint32_t* hdr = g.w + (size_t)NL * g.n;
  const int ngroups = (g.rows_total + rpb - 1) / rpb;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int row = grp * rpb + warp / wpr;
    const bool live = row < g.rows_total;
    int32_t mx = mp::EXP_ZERO;
    int64_t base = 0;
    if (live) {
      int b = row / g.rows, r = row % g.rows;
      base = item_off(g, b) + (int64_t)r * g.rs;
      for (int k = sub * 32 + lane; k < g.K; k += 32 * wpr) mx = max(mx, ((int32_t)hdr[base + (int64_t)k * g.ks]) >> 1);
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
      for (int k = sub * 32 + lane; k < g.Kp; k += 32 * wpr) slice_entry<NL, S>(g, row, k, mx, digits);
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

struct MmaParams {
  int T, M, N, batch, nsplit, Kp, Kc, BK, BN, m_tiles, n_tiles, npairs, item0, stages;
  const int* rowA;
  const int* rowB;
  int32_t* planes;  // [T][nsplit][batch][M][N]
  uint32_t idesc, sbo16, layout_type;
};

constexpr int MMA_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int MAX_STAGES = 8;

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

__global__ void __launch_bounds__(MMA_THREADS, 1)
mma_planes_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, MmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // every operand tile starts on a 1024-byte boundary (required by the 128B swizzle atom)
  const uint32_t a_bytes = 128u * p.BK, b_bytes = (uint32_t)p.BN * p.BK;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  uint64_t* bars = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  // bars[0..7] full, [8..15] empty, [16..17] tmem_full, [18..19] tmem_empty, then tmem base ptr
  uint32_t* tmem_slot = (uint32_t*)(bars + 20);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // decode work: pair fastest so that CTAs sharing operand tiles are co-scheduled (L2 reuse)
  int idx = blockIdx.x;
  const int pair = idx % p.npairs;
  idx /= p.npairs;
  const int split = idx % p.nsplit;
  idx /= p.nsplit;
  const int nt = idx % p.n_tiles;
  idx /= p.n_tiles;
  const int mt = idx % p.m_tiles;
  idx /= p.m_tiles;
  const int item = idx;  // local item in this launch
  const int gitem = p.item0 + item;
  const int rowA0 = (p.rowA ? p.rowA[gitem] : gitem * p.M) + mt * 128;
  const int rowB0 = (p.rowB ? p.rowB[gitem] : gitem * p.N) + nt * p.BN;
  const int k_begin = split * p.Kc;
  const int kblocks = (min(p.Kp, k_begin + p.Kc) - k_begin) / p.BK;
  int planes_t[2];
  planes_t[0] = p.T - 1 - pair;  // the long plane first
  planes_t[1] = pair;
  const int nplanes = (planes_t[0] == planes_t[1]) ? 1 : 2;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * p.BN) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; s++) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[8 + s]), 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(smem_u32(&bars[16 + s]), 1);
      mbar_init(smem_u32(&bars[18 + s]), 4);
    }
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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pl = 0; pl < nplanes; pl++) {
        const int t = planes_t[pl];
        for (int a = 0; a <= t; a++)
          for (int kb = 0; kb < kblocks; kb++) {
            mbar_wait(smem_u32(&bars[8 + stage]), phase ^ 1u);
            uint32_t full = smem_u32(&bars[stage]);
            mbar_expect_tx(full, a_bytes + b_bytes);
            uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            int k0 = k_begin + kb * p.BK;
            tma_load_3d(sa, &tmA, full, k0, rowA0, a);
            tma_load_3d(sa + a_bytes, &tmB, full, k0, rowB0, t - a);
            if (++stage == p.stages) stage = 0, phase ^= 1u;
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pl = 0; pl < nplanes; pl++) {
        const int t = planes_t[pl];
        // accumulator buffer pl: first use of each buffer needs no wait (fresh barrier parity trick)
        mbar_wait(smem_u32(&bars[18 + pl]), 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(pl * p.BN);
        const int nkb = (t + 1) * kblocks;
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(smem_u32(&bars[stage]), phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          uint64_t adesc = make_smem_desc(sa, p.sbo16, p.layout_type);
          uint64_t bdesc = make_smem_desc(sa + a_bytes, p.sbo16, p.layout_type);
          for (int kk = 0; kk < p.BK / 32; kk++) {
            umma_i8(d_tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), p.idesc,
                    (kb > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&bars[8 + stage]));  // frees the smem slot when these MMAs retire
          if (++stage == p.stages) stage = 0, phase ^= 1u;
        }
        umma_commit(smem_u32(&bars[16 + pl]));  // accumulator pl complete
      }
    }
  } else {
    // ------------------------------ epilogue: TMEM -> registers -> HBM planes --------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = mt * 128 + q * 32 + lane;
    for (int pl = 0; pl < nplanes; pl++) {
      const int t = planes_t[pl];
      mbar_wait(smem_u32(&bars[16 + pl]), 0u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      int32_t* out = p.planes + ((((size_t)t * p.nsplit + split) * p.batch + item) * p.M + row) * (size_t)p.N;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pl * p.BN + c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        int col = nt * p.BN + c0;
        if (row < p.M) {
          if (col + 16 <= p.N && (p.N & 3) == 0) {
            int4* o4 = reinterpret_cast<int4*>(out + col);
#pragma unroll
            for (int i = 0; i < 4; i++) o4[i] = make_int4((int)v[4 * i], (int)v[4 * i + 1], (int)v[4 * i + 2], (int)v[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; i++)
              if (col + i < p.N) out[col + i] = (int32_t)v[i];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[18 + pl]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// =====================================================================================================
// 3. carry propagation + renormalisation:  planes -> mp numbers
// =====================================================================================================
struct CarryArgs {
  const int32_t* planes;  // [T][nsplit][batch][M][N]
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

template <int NL, int T>
__global__ void carry_kernel(CarryArgs c) {
  constexpr int NW = (T + 3) / 4 + 2;
  int64_t total = (int64_t)c.batch * c.M * c.N;
  size_t pstride = (size_t)c.nsplit * total;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(idx % c.N);
    int i = (int)((idx / c.N) % c.M);
    int b = (int)(idx / ((int64_t)c.N * c.M));
    int gb = c.item0 + b;
    uint32_t W[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) W[q] = 0;
    int64_t carry = 0;
    // planes from the least significant (t = T-1) upwards, CH at a time: the CH (x nsplit) loads of a chunk are
    // independent and all in flight before the carry chain consumes them
    constexpr int CH = 17;
#pragma unroll
    for (int c0 = T - 1; c0 >= 0; c0 -= CH) {
      int64_t acc[CH];
#pragma unroll
      for (int u = 0; u < CH; u++) acc[u] = 0;
      for (int s = 0; s < c.nsplit; s++) {
        const int32_t* ps = c.planes + (size_t)s * total + idx;
        int32_t v32[CH];
#pragma unroll
        for (int u = 0; u < CH; u++) v32[u] = (c0 - u >= 0) ? ps[(size_t)(c0 - u >= 0 ? c0 - u : 0) * pstride] : 0;
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
    int ra = (c.rowA ? c.rowA[gb] : gb * c.M) + i, rb = (c.rowB ? c.rowB[gb] : gb * c.N) + j;
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
static CUtensorMap make_map(const Slice& s, int box_rows, int BK) {
  CUtensorMap tm;
  // rows past rows_total (tiles of the last item) are out of bounds => zero-filled by TMA
  cuuint64_t dims[3] = {(cuuint64_t)s.Kp, (cuuint64_t)s.rows_total, (cuuint64_t)s.S};
  cuuint64_t strides[2] = {(cuuint64_t)s.Kp, (cuuint64_t)s.Kp * s.rows_total};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = BK == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (BK == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, s.digits.p, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return tm;
}

GemmEngine::GemmEngine(Ctx& c, int nl) : ctx_(c), nl_(nl), S_(num_digits(nl)) {
  CLR_CUDA(cudaFuncSetAttribute(mma_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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
  GatherArgs g{op.src.w, op.src.n, op.d_off, op.off0, op.bstride, op.rs, op.ks, op.batch, op.rows, K, Kp, out.rows_total};
  int64_t total = (int64_t)out.rows_total * Kp;
  // long rows (or few of them) get a whole block per row, short rows one warp
  int wpr = (Kp > 256 || (int64_t)out.rows_total * 32 < (int64_t)ctx.sm_count * 256) ? 8 : 1;
  if (Kp <= 32) wpr = 1;
  int rpb = 8 / wpr;
  int grid = (int)std::min<int64_t>(ceil_div(out.rows_total, rpb), (int64_t)ctx.sm_count * 8);
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

void GemmEngine::run_mma(const Slice& A, const Slice& B, const GemmPlan& plan, int item0, int nitems, int nsplit,
                         int Kc) {
  MmaParams p;
  memset(&p, 0, sizeof(p));
  p.T = S_;
  p.M = plan.M;
  p.N = plan.N;
  p.batch = nitems;
  p.nsplit = nsplit;
  p.Kp = A.Kp;
  p.Kc = Kc;
  p.BK = std::min(A.Kp, 128);
  p.BN = std::min(256, ((plan.N + 15) / 16) * 16);
  p.m_tiles = ceil_div(plan.M, 128);
  p.n_tiles = ceil_div(plan.N, p.BN);
  p.npairs = (p.T + 1) / 2;
  p.item0 = item0;
  p.rowA = plan.d_rowA;
  p.rowB = plan.d_rowB;
  p.planes = planes_.as<int32_t>();
  // instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) at [4,6), a/b format INT8 (1) at
  // [7,10)/[10,13), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
  p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.sbo16 = (8u * p.BK) >> 4;
  p.layout_type = p.BK == 128 ? 2u : (p.BK == 64 ? 4u : 6u);
  size_t stage_bytes = (size_t)128 * p.BK + ((((size_t)p.BN * p.BK) + 1023) & ~(size_t)1023);
  p.stages = (int)std::min<size_t>(MAX_STAGES, (200 * 1024) / stage_bytes);
  size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
  CUtensorMap tmA = make_map(A, 128, p.BK), tmB = make_map(B, p.BN, p.BK);
  int64_t grid = (int64_t)nitems * p.m_tiles * p.n_tiles * nsplit * p.npairs;
  double macs = (double)nitems * p.m_tiles * 128.0 * p.n_tiles * p.BN * (double)A.Kp * (p.T * (p.T + 1) / 2.0);
  last_int8_macs += macs;
  // algorithmic MACs: M*N*K * s(s+1)/2 with s = p/8, no guard digits, no tile padding
  double s_alg = 4.0 * nl_;
  double alg = (double)nitems * plan.M * plan.N * (double)A.K * (s_alg * (s_alg + 1) / 2.0);
  std::string nm = "mma_planes_M" + std::to_string(plan.M) + "_N" + std::to_string(plan.N) + "_K" + std::to_string(A.K) + "_b" + std::to_string(nitems);
  int tk = ctx_.begin(nm.c_str(), alg);
  mma_planes_kernel<<<(unsigned)grid, MMA_THREADS, smem, ctx_.stream>>>(tmA, tmB, p);
  ctx_.end(tk);
}

template <int NL>
static void carry_impl(Ctx& ctx, const CarryArgs& c) {
  constexpr int T = 4 * NL + GUARD_DIGITS;
  int64_t total = (int64_t)c.batch * c.M * c.N;
  int grid = (int)std::min<int64_t>(ceil_div(total, 128), (int64_t)ctx.sm_count * 32);
  // algorithmic bytes: read T int32 planes (x nsplit), write (p/8+4)
  int tk = ctx.begin("carry", (double)total * (4.0 * T * c.nsplit + 4.0 * (NL + 1)));
  carry_kernel<NL, T><<<grid, 128, 0, ctx.stream>>>(c);
  ctx.end(tk);
}

void GemmEngine::multiply(const Slice& A, const Slice& B, const GemmPlan& plan, const OutDesc& C, int epi,
                          const mp::Tensor* extra) {
  if (A.Kp != B.Kp || A.S != S_ || B.S != S_) throw SolverError(-1, "gemm: operand mismatch");
  const int T = S_;
  int BK = std::min(A.Kp, 128);
  // int32 exactness: (t+1) * Kc * 2^14 < 2^31  =>  T * Kc <= 131071
  int kc_safe = (131071 / T) / BK * BK;
  int Kc = std::min(A.Kp, kc_safe);
  int BN = std::min(256, ((plan.N + 15) / 16) * 16);
  int64_t tiles = (int64_t)plan.batch * ceil_div(plan.M, 128) * ceil_div(plan.N, BN) * ((T + 1) / 2);
  if (tiles < 2 * ctx_.sm_count && A.Kp >= 2048) Kc = std::min(Kc, 1024);  // split-K for parallelism
  int nsplit = ceil_div(A.Kp, Kc);
  // chunk the batch so that the plane workspace stays bounded
  size_t per_item = (size_t)T * nsplit * plan.M * plan.N * sizeof(int32_t);
  size_t cap = (size_t)1536 << 20;
  int chunk = (int)std::max<size_t>(1, std::min<size_t>(plan.batch, cap / std::max<size_t>(per_item, 1)));
  planes_.ensure(per_item * chunk);
  for (int item0 = 0; item0 < plan.batch; item0 += chunk) {
    int n = std::min(chunk, plan.batch - item0);
    run_mma(A, B, plan, item0, n, nsplit, Kc);
    CarryArgs c;
    memset(&c, 0, sizeof(c));
    c.planes = planes_.as<int32_t>();
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

void GemmEngine::planes_only(const Slice& A, const Slice& B, const GemmPlan& plan, int32_t* h_planes, int* T_out) {
  const int T = S_;
  int BK = std::min(A.Kp, 128);
  int kc_safe = (131071 / T) / BK * BK;
  if (A.Kp > kc_safe) throw SolverError(-1, "planes_only: K too large for a single split");
  size_t bytes = (size_t)T * plan.batch * plan.M * plan.N * sizeof(int32_t);
  planes_.ensure(bytes);
  run_mma(A, B, plan, 0, plan.batch, 1, A.Kp);
  CLR_CUDA(cudaMemcpyAsync(h_planes, planes_.p, bytes, cudaMemcpyDeviceToHost, ctx_.stream));
  ctx_.sync();
  *T_out = T;
}

}  // namespace clr

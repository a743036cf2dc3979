// gemm_i8.cuh — interface of the sliced int8 tensor-core GEMM (the GPU stand-in for approx_mul!,
// 18 call sites in MPMP.jl).  C = A * B^T-operand form:
//     C[b][i][j] = sum_k  Arows[b][i][k] * Brows[b][j][k]
// where both operands are "row operands" (the contraction index k is the fast index of every row), so
// the same sliced object can serve as left or right factor.
//
//   slice      : mp rows -> per-row exponent + S balanced radix-256 digits (int8), planes [S][rows][Kp]
//   mma_planes : plane t = sum_{a+b=t} digits_a(A) * digits_b(B)^T   (tcgen05 kind::i8, int32 in TMEM,
//                operands by TMA), t = 0..T-1, exact
//   carry      : sum_t plane_t * 256^-t -> carry propagation, normalisation, rounding to p bits
#pragma once
#include "common.cuh"

namespace clr {

constexpr int GUARD_DIGITS = 2;  // digits beyond p/8 kept per operand and per product
inline int num_digits(int nl) { return 4 * nl + GUARD_DIGITS; }

// gather description of a logical operand [batch][rows][K] inside an mp tensor:
//   element (b, r, k) lives at  off(b) + r*rs + k*ks,   off(b) = off0 + (d_off ? d_off[b] : b*bstride)
struct OperandDesc {
  mp::Tensor src;
  const int64_t* d_off = nullptr;
  int64_t off0 = 0, bstride = 0;
  int64_t rs = 0, ks = 1;
  int batch = 1, rows = 0, K = 0;
  // optional [batch][K] exponents: the operand is sliced as x(b, r, k) * 2^-d_kshift[b*K + k] (exact scaling of the
  // contraction index, e.g. D^-1 B with the equilibration D of the matrix that B is multiplied into)
  const int* d_kshift = nullptr;
  // optional [batch][ksign_ld] flags: entry (b, r, k) is negated where d_ksign[b*ksign_ld + k] != 0 (the Sigma of a
  // signed factorisation A = U^T Sigma U applied to the contraction index)
  const int* d_ksign = nullptr;
  int ksign_ld = 0;
};
// destination of a product: element (b, i, j) at off(b) + i*rs + j*cs, off(b) as above
struct OutDesc {
  mp::Tensor dst;
  const int64_t* d_off = nullptr;
  int64_t off0 = 0, bstride = 0;
  int64_t rs = 0, cs = 1;
};
// epilogue of the carry kernel
enum : int {
  EPI_STORE = 0,      // C = A*B
  EPI_SUB_FROM = 1,   // C = E - A*B     (E = extra operand, same addressing as C)
  EPI_MINUS_SUB = 2,  // C = A*B - E
  EPI_ADD = 3,        // C = A*B + E
  EPI_NEG = 4         // C = -A*B
};

struct Slice {
  DevBuf digits;  // int8 [S][rows_total][Kp]
  DevBuf exps;    // int32 [rows_total]
  int S = 0, rows_total = 0, Kp = 0, K = 0, rows_item = 0, batch = 0;
};

struct GemmPlan {  // shapes of one batched product
  int batch = 1, M = 0, N = 0;
  const int* d_rowA = nullptr;  // optional [batch] first row of item b in A (default b*M)
  const int* d_rowB = nullptr;  // optional [batch] first row of item b in B (default b*N)
};

class GemmEngine {
 public:
  GemmEngine(Ctx& c, int nl);
  void slice(const OperandDesc& op, Slice& out);
  // C = A * B (row-operand form) with optional epilogue; E uses C's addressing on tensor `extra`
  // `symmetric`: the caller guarantees C = C^T and that row i of A and row i of B have the same exponent (the same
  // slice on both sides, or two slices of one operand that differ by signs only): only the tiles touching the upper
  // triangle are computed, the carry kernel mirrors them
  void multiply(const Slice& A, const Slice& B, const GemmPlan& plan, const OutDesc& C, int epi = EPI_STORE,
                const mp::Tensor* extra = nullptr, bool symmetric = false);
  // exact integer planes for tests: planes [T][batch][M][N] (splits already summed must be 1)
  void planes_only(const Slice& A, const Slice& B, const GemmPlan& plan, int32_t* h_planes, int* T_out);
  int digits() const { return S_; }
  // int8 MACs per second of back-to-back tcgen05.mma kind::i8 128x256x32 on all SMs (measured denominator of the roofline)
  double measure_i8_peak();
  double last_int8_macs = 0;  // executed digit-product MACs of the last multiply (incl. guard digits)

 private:
  struct FusedOut {  // destination + epilogue operation when the tensor-core kernel finishes the numbers itself
    OutDesc C;
    int epi;
    const mp::Tensor* extra;
  };
  void run_mma(const Slice& A, const Slice& B, const GemmPlan& plan, int item0, int nitems, int nsplit, int Kc, bool symmetric,
               const FusedOut* fo = nullptr);
  bool fused_ok_ = true;  // CLRSDP_FUSED_CARRY=0 keeps the separate carry launch for every product
  Ctx& ctx_;
  int nl_, S_;
  DevBuf planes_, sym_map_;
  size_t sym_map_host_ = 0;
  size_t planes_cap_ = 0;
};

}  // namespace clr

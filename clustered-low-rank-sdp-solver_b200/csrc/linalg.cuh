// linalg.cuh — CUDA-core multiprecision kernels of the hot path (everything that is not a GEMM):
// batched Cholesky / triangular inverse / smallest eigenvalue, elementwise block operations,
// reductions, GEMV, the Schur-complement combination and the trace/weight kernels.
// Each host function cites the reference lines whose work it performs.
#pragma once
#include "common.cuh"

namespace clr {

// batch of n x n matrices inside an mp tensor: element (r,c) of matrix b at d_off[b] + shift + r*ld + c
// (ld = 0 means ld = n). Sub-blocks of larger matrices are expressed through shift/ld.
struct MatBatch {
  mp::Tensor t;
  const int64_t* d_off = nullptr;
  int batch = 0, n = 0;
  int ld = 0;
  int64_t shift = 0;
  int stride() const { return ld ? ld : n; }
  MatBatch sub(int r0, int c0, int nn) const {
    MatBatch m = *this;
    m.shift = shift + (int64_t)r0 * stride() + c0;
    m.ld = stride();
    m.n = nn;
    return m;
  }
};

// ---- factorisations ----------------------------------------------------------------------------------
// A = U^T U, U upper triangular (so L = U^T is the Cholesky factor of cho!/spd_inv!, MPMP.jl:766,1846).
// U is written row-major with zeros below the diagonal; rdiag[b*n + k] = 1/U[k][k].
// d_status[b] = 1 if a pivot is not positive. A and U may alias.
void chol_upper(Ctx& ctx, int nl, const MatBatch& A, const MatBatch& U, mp::Tensor rdiag, int* d_status);
// V = U^-1 (upper triangular, row-major) and optionally Linv = V^T = L^-1 (lower triangular, row-major).
void tri_inverse(Ctx& ctx, int nl, const MatBatch& U, mp::Tensor rdiag, const MatBatch& V, const MatBatch* Linv);
// Fused panel factorisation of w x w SPD blocks (w <= panel_width(nl)): A is overwritten by its upper Cholesky
// factor U (if write_u) and Linv receives L^-1 = U^-T (lower); both may be sub-blocks of larger matrices.
int panel_width(int nl);
// With d_sig the factorisation is SIGNED, A = U^T Sigma U with Sigma = diag(+-1) (an LDL^T whose pivots keep their sign;
// what stands in for the reference's pivoted LU of S_j and Q, MPMP.jl:1436,1501): d_sig[b*sig_ld + k] = 1 where the
// pivot of row k is negative, nothing is reported as "not positive definite", and Linv = L^-1 with L = U^T.
void panel_factor(Ctx& ctx, int nl, const MatBatch& A, const MatBatch& Linv, bool write_u, int* d_status,
                  int* d_sig = nullptr, int sig_ld = 0);
// smallest eigenvalue of each symmetric matrix (destroys W): out[out_off + b].
// Replaces approx_eig_qr! + min over real parts (MPMP.jl:1857-1870) by Householder tridiagonalisation
// + multisection with Sturm counts.
// out[d_out_index ? d_out_index[b] : b] receives the result of matrix b.
// With d_flags (batch ints of scratch) the FP64-preconditioned refinement kernel runs first (n^3 work in FP64, O(n^2)
// multiprecision work per refinement step) and the all-multiprecision kernel only handles the matrices it flags.
void lambda_min(Ctx& ctx, int nl, const MatBatch& W, mp::Tensor out, const int* d_out_index, int* d_flags = nullptr);

// ---- elementwise ---------------------------------------------------------------------------------------
// out[i] = sa*a[i] + sb*b[i], sa,sb in {-1,0,+1}; i in [0,n) at offsets (oo, ao, bo)
void ew_lincomb(Ctx& ctx, int nl, mp::Tensor out, int64_t oo, mp::Tensor a, int64_t ao, int sa, mp::Tensor b,
                int64_t bo, int sb, int64_t n);
// y[i] += s * x[i] with s = scal[slot]  (MPMP.jl:877-887)
void ew_axpy(Ctx& ctx, int nl, mp::Tensor y, int64_t yo, mp::Tensor x, int64_t xo, mp::Tensor scal, int slot,
             int64_t n);
// R = s*I - T1 [- T2] per block  (compute_residual_R!, MPMP.jl:1189-1215)
void ew_residual_R(Ctx& ctx, int nl, const MatBatch& R, mp::Tensor scal, int slot, mp::Tensor T1, const mp::Tensor* T2);
// out = (in + in^T)/2 per block  (MPMP.jl:1719-1728, 1810-1818)
void ew_symmetrize(Ctx& ctx, int nl, const MatBatch& out, mp::Tensor in);
// M = s*I per block (MPMP.jl:660-686)
void ew_set_identity(Ctx& ctx, int nl, const MatBatch& M, mp::Tensor scal, int slot);
void ew_zero(Ctx& ctx, int nl, mp::Tensor t, int64_t off, int64_t n);
// dst = src for a batch of (sub)matrices; with upper_only the strictly lower part of dst is zeroed instead
void mat_copy(Ctx& ctx, int nl, const MatBatch& dst, const MatBatch& src, bool upper_only);
void mat_zero(Ctx& ctx, int nl, const MatBatch& dst);
// symmetric equilibration by exact powers of two: d_scale[b*n+i] = ceil(exponent(A_ii)/2); scaled copy
// dst(r,c) = src(r,c) 2^-(s_r+s_c); M(r,c) *= 2^(sign*s_c). Cholesky in floating point is invariant under diagonal
// scaling, the block-fixed-point GEMMs of the blocked factorisation are not: they see the equilibrated matrix.
void equil_exponents(Ctx& ctx, int nl, const MatBatch& A, int* d_scale);
void mat_copy_scaled(Ctx& ctx, int nl, const MatBatch& dst, const MatBatch& src, bool upper_only, const int* d_scale);
void col_scale(Ctx& ctx, int nl, const MatBatch& M, int sign, const int* d_scale);
// v[off + i] *= 2^(sign * d_scale[i]); d_dst[d_off[b] + i] = d_src[b*n + i]
void vec_scale(Ctx& ctx, int nl, mp::Tensor v, int64_t off, int64_t n, int sign, const int* d_scale);
void scatter_scale(Ctx& ctx, const int* d_src, int batch, int n, const int64_t* d_off, int* d_dst);
// v[off + i] = -v[off + i] where d_sig[i] != 0 (the Sigma of a signed factorisation applied to a vector)
void vec_flip(Ctx& ctx, int nl, mp::Tensor v, int64_t off, int64_t n, const int* d_sig);
// c[i] = a[i] (op) b[i], op in '+','-','*','/','s' (sqrt of a): the scalar arithmetic of mpf.cuh on the device
void ew_binary(Ctx& ctx, int nl, int op, mp::Tensor c, mp::Tensor a, mp::Tensor b, int64_t n);

// ---- reductions (results into scal[slot]) ---------------------------------------------------------------
// sum_i a[i]*b[i]            (dot, MPMP.jl:205-220; dot_c :1081-1092)
void reduce_dot(Ctx& ctx, int nl, mp::Tensor a, int64_t ao, mp::Tensor b, int64_t bo, int64_t n, mp::Tensor scal,
                int slot, mp::Tensor work);
// sum_i (a+da)[i]*(b+db)[i]  (MPMP.jl:832)
void reduce_dot_sum(Ctx& ctx, int nl, mp::Tensor a, mp::Tensor da, mp::Tensor b, mp::Tensor db, int64_t n,
                    mp::Tensor scal, int slot, mp::Tensor work);
// max_i |a[i]|               (compute_error, MPMP.jl:1037-1055)
void reduce_maxabs(Ctx& ctx, int nl, mp::Tensor a, int64_t ao, int64_t n, mp::Tensor scal, int slot, mp::Tensor work);
// min_i a[i]                 (MPMP.jl:1890-1891)
void reduce_min(Ctx& ctx, int nl, mp::Tensor a, int64_t ao, int64_t n, mp::Tensor scal, int slot, mp::Tensor work);
size_t reduce_work_elems();

// ---- GEMV on CUDA cores ----------------------------------------------------------------------------------
// out[oo + r] = sum_k A[aoff(r) + r_local*rs + k*ks] * x[xoff(r) + k], rows r in [0,rows).
// With d_row_item == nullptr there is one item (aoff = a0, xoff = x0, r_local = r, K = K);
// otherwise row r belongs to item it = d_row_item[r] with aoff = d_aoff[it], xoff = d_xoff[it],
// r_local = r - d_row0[it], K = d_K[it] (per-cluster triangular solves, MPMP.jl:1751-1773).
struct GemvArgs {
  mp::Tensor A, x, out;
  int64_t a0 = 0, x0 = 0, oo = 0, rs = 0, ks = 1;
  int rows = 0, K = 0;
  const int* d_row_item = nullptr;
  const int64_t* d_aoff = nullptr;
  const int64_t* d_xoff = nullptr;
  const int* d_row0 = nullptr;
  const int* d_K = nullptr;
  int item_trans = 0;  // with items: 0 -> A_item[r][k], 1 -> A_item[k][r]; leading dimension = K_item
  int K_hint = 0;      // with items: the largest K_item (sizes the K-split of the transposed kernel)
  // Header-word operations fused into the product (they replace separate vec_flip / vec_scale / lincomb launches on the
  // latency chain of the Schur solve). x side, indexed like x from x0: x_k <- (-1)^x_flip[k] 2^(x_scale_sign x_scale[k]) x_k.
  // Result side, indexed by row, applied in this order: sign flip, power-of-two scale, then e_mode 1: out = e + acc,
  // e_mode 2: out = e - acc (e may alias out).
  const int* x_flip = nullptr;
  const int* x_scale = nullptr;
  int x_scale_sign = 0;
  const int* o_flip = nullptr;
  const int* o_scale = nullptr;
  int o_scale_sign = 0;
  mp::Tensor e;
  int64_t e0 = 0;
  int e_mode = 0;
};
void gemv(Ctx& ctx, int nl, const GemvArgs& g, mp::Tensor work);
size_t gemv_work_elems(int rows, int K);

// ---- wire format <-> header word: t[off + i] gets / yields (sign, exp); a zero sign also clears the limbs ------
void wire_pack(Ctx& ctx, int nl, mp::Tensor t, int64_t off, int64_t n, const int8_t* d_sign, const int64_t* d_exp);
void wire_unpack(Ctx& ctx, int nl, mp::Tensor t, int64_t off, int64_t n, int8_t* d_sign, int64_t* d_exp);

// ---- small batched products on the CUDA cores ---------------------------------------------------------------
// C(b,i,j) = epi( sum_k A(b,i,k) * B(b,j,k) ): element (b,r,k) of an operand at off(b) + r*rs + k*ks with
// off(b) = x0 + (offX ? offX[b] : b*xbs); C element (b,i,j) at off(b) + i*crs + j*ccs; epi as in gemm_i8.cuh.
// For products whose size does not amortise the latency of the sliced tensor-core pipeline.
struct SmallGemmArgs {
  mp::Tensor A, B, C, E;  // E = extra operand of the epilogue (C's addressing); E.w == nullptr -> C itself
  const int64_t* offA = nullptr;
  const int64_t* offB = nullptr;
  const int64_t* offC = nullptr;
  int64_t a0 = 0, abs_ = 0, ars = 0, aks = 1, b0 = 0, bbs = 0, brs = 0, bks = 1, c0 = 0, cbs = 0, crs = 0, ccs = 1;
  int batch = 1, M = 0, N = 0, K = 0, epi = 0;
  const int* ksign = nullptr;  // optional [batch][ksign_ld]: term k of item b is subtracted where ksign != 0
  int ksign_ld = 0;
};
void small_gemm(Ctx& ctx, int nl, const SmallGemmArgs& a);
// C(b,i,j) = A(b,i,j): strided rectangular copy (A element (b,i,j) at offA(b) + i*ars + j*aks)
void rect_copy(Ctx& ctx, int nl, const SmallGemmArgs& a);

// ---- structure-aware kernels ------------------------------------------------------------------------------
// device tables describing the clustered structure (BlockInfo, MPMP.jl:467-479)
struct StructTables {
  int J = 0, n_y = 0, n_blocks = 0, sumS = 0;
  // per cluster
  const int* c_m = nullptr;       // m[j]
  const int* c_K = nullptr;       // n_samples[j]
  const int* c_L = nullptr;       // L[j]
  const int* c_blk0 = nullptr;    // first block index of cluster j
  const int* c_xoff = nullptr;    // x_indices[j]
  const int64_t* c_Soff = nullptr;  // offset of S_j in the S arena
  const int* c_dimS = nullptr;
  // per block
  const int* b_delta = nullptr;
  const int* b_Nv = nullptr;
  const int* b_cluster = nullptr;
  const int* b_rs0 = nullptr;       // start of rank_sums (K+1 entries) in rank_sums[]
  const int64_t* b_Hoff = nullptr;  // start of this block's H / samp entries
  const int64_t* b_Poff = nullptr;  // pairing matrices offset (m*Nv)^2
  const int64_t* b_Voff = nullptr;  // V arena offset (Nv*delta)
  const int64_t* b_Toff = nullptr;  // T't arena offset (m*m*Nv*delta)
  const int64_t* b_VDoff = nullptr; // VD arena offset (npairs*delta*Nv)
  const int64_t* b_QPoff = nullptr; // QP arena offset (npairs*delta*delta)
  const int64_t* b_off = nullptr;   // block arena offset (nb*nb)
  const int* rank_sums = nullptr;
  const int* samp = nullptr;        // sample index of each vector
  // per x entry
  const int* x_cluster = nullptr;   // cluster of x entry i
};
// S_j[(r2,s2,k2),(r1,s1,k1)] (4-term combination + symmetrisation, MPMP.jl:1335-1410)
void schur_assemble(Ctx& ctx, int nl, const StructTables& st, mp::Tensor Px, mp::Tensor Py, mp::Tensor H, mp::Tensor S,
                    int64_t S_total);
// out[(j,r,s,k)] = sum_l sum_rnk H * Py[(r,a),(s,a)]   (A_Y + trace_A(A_Y), MPMP.jl:1320-1330,1585-1618)
void trace_from_pairings(Ctx& ctx, int nl, const StructTables& st, mp::Tensor Py, mp::Tensor H, mp::Tensor out);
// out[(j,r,s,k)] = sum_l sum_rnk H * sum_i Vt[a][i]*Tt[(r,s)][a][i]   (trace_A general, MPMP.jl:1517-1584)
void trace_from_ZV(Ctx& ctx, int nl, const StructTables& st, mp::Tensor Vt, mp::Tensor Tt, mp::Tensor H, mp::Tensor out);
// VD[pair][i][v] = a[(j,pair,k(v))] * H[v] * V[i][v]      (vs_scaled, MPMP.jl:1654)
void scale_vectors(Ctx& ctx, int nl, const StructTables& st, mp::Tensor Vt, mp::Tensor H, mp::Tensor a, mp::Tensor VD,
                   int64_t VD_total);
// block = sym(assemble(QP)) + sign*E   (compute_weighted_A! tail :1661-1674 and P -= X :1115 / dX += P :1784)
void assemble_weighted(Ctx& ctx, int nl, const StructTables& st, mp::Tensor QP, mp::Tensor out, mp::Tensor E, int sign,
                       int64_t blk_total, const int* d_elem_block_hint);

// ---- driver scalars on the device (MPMP.jl:755-756, 832-837, 871-874, 1893-1897, 940-953, 1147-1185) -------
enum ScalarSlots : int {
  SL_MU = 0, SL_P_OBJ, SL_D_OBJ, SL_GAP, SL_PRIMAL_ERR, SL_DUAL_ERR, SL_ALPHA_P, SL_ALPHA_D, SL_BETA_C, SL_MU_P,
  SL_MU_C, SL_LAM_X, SL_LAM_Y,
  // parameters
  SL_BETA_INF, SL_BETA_FEAS, SL_GAMMA, SL_OMEGA_P, SL_OMEGA_D, SL_GAP_THR, SL_PERR_THR, SL_DERR_THR, SL_B0,
  // temporaries
  SL_DOT_XY, SL_DOT_SUM, SL_CX, SL_BY, SL_PERR_P, SL_PERR_p, SL_DERR, SL_NTOT, SL_CY, SL_COUNT
};
enum ScalarProgram : int { SP_MU = 0, SP_BETA, SP_ALPHA, SP_OBJECTIVES, SP_ERRORS, SP_OBJECTIVES_INIT };
// flags[0] = pd_feas (in/out), flags[1] = terminate reason, flags[2..3] = need_primal/need_dual
// SP_ALPHA with d_status: the step lengths are set to zero when any of the n_status factorisation flags is raised
void scalar_program(Ctx& ctx, int nl, int prog, mp::Tensor scal, int* d_flags, double* d_out,
                    const int* d_status = nullptr, int n_status = 0);

}  // namespace clr

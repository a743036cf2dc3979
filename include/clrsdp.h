/*
 * clrsdp.h — C ABI of the B200-native interior-point hot path for clustered low-rank SDPs.
 *
 * This is the drop-in boundary for the body of `solverank1sdp` in the reference
 * (nanleij/Clustered-Low-Rank-SDP-solver, MPMP.jl:595-1025). The reference has no FFI of its own
 * for this path (its only foreign boundary is Arblib -> libarb, once per matrix op); the cut is the
 * signature of `solverank1sdp`: constraints (A,B,c,H), b, BlockInfo, kwargs in; the tuple
 * x,X,y,Y,P,p,d,gap,primal_obj,dual_obj,time out (MPMP.jl:1014-1024).
 *
 * Conventions
 *   - plain C, no torch / C++ types; the caller owns every host buffer, the library copies in/out
 *     and owns all device memory, streams and communicators.
 *   - every function returns an int status, 0 = success (the inverse of libarb's convention,
 *     MPMP.jl:774,792,1438,1502,1847). No exceptions or callbacks cross the boundary.
 *   - a handle is not thread-safe; calls block the calling host thread.
 *   - there is NO CPU fallback: if no CUDA device is usable, clrsdp_create fails with CLRSDP_ERR_CUDA.
 *
 * Number wire format (clrsdp_mp): planar arrays of n numbers at the handle's precision p = 32*nlimb bits
 *     value_i = sign[i] * (0.limbs_i)_2 * 2^exp[i]          (MPFR convention; BigFloat-compatible)
 *   sign[i] in {-1,0,+1}; limb[k*n + i] is 32-bit limb k of number i, limb 0 least significant; for a
 *   non-zero number the top bit of limb nlimb-1 is set. exp is ignored when sign == 0.
 *
 * Matrix layout: all matrices cross the boundary ROW-major. (X, Y, S_j, Q are symmetric; B_j is
 * dim_S[j] x n_y with rows ordered (r,s,k), k fastest, exactly as prepareabc builds it, MPMP.jl:387-395.)
 * SURVEY §8b sketched column-major (Julia's native order) at this cut; row-major was chosen because every kernel
 * consumes rows with the contraction index contiguous, and the Julia shim (julia/ClrsdpB200.jl: MpArr(::ArbMatrix))
 * transposes while it converts Arb midpoints to the wire format - a copy it has to make anyway.
 * Exponents must lie in (-2^27, 2^27) (the device header word holds 31 bits and reserves -2^28 for zero); tensors that
 * violate this or the normalisation rule above are refused with CLRSDP_ERR_BAD_ARG on the staged upload path.
 *
 * The same set of entry points exists with the prefix `clrsdp_ref_` in oracle/ (the CPU restatement
 * used ONLY by tests / smoke / the bench's cpu_baseline leg).
 */
#ifndef CLRSDP_H
#define CLRSDP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct clrsdp_solver* clrsdp_handle;

typedef struct {
  const int8_t*   sign;  /* [n]            */
  const int64_t*  exp;   /* [n]            */
  const uint32_t* limb;  /* [nlimb][n]     */
  int64_t         n;
} clrsdp_mp;

typedef struct {
  int8_t*   sign;
  int64_t*  exp;
  uint32_t* limb;
  int64_t   n;           /* capacity in numbers */
} clrsdp_mp_out;

/* status codes */
enum {
  CLRSDP_OK            = 0,
  CLRSDP_ERR_BAD_ARG   = -1,
  CLRSDP_ERR_CUDA      = -2,
  CLRSDP_ERR_NCCL      = -3,
  CLRSDP_ERR_NOT_PD_X  = -10, /* Cholesky of an X block failed  (MPMP.jl:774-798: "try higher precision") */
  CLRSDP_ERR_NOT_PD_Y  = -11, /* Cholesky of a  Y block failed  (MPMP.jl:1846-1848,1881-1884)           */
  /* -12 .. -14 are kept for ABI stability and for front ends that map them to the reference's messages, but this
   * implementation does not produce them: S_j and Q are factored as U^T Sigma U with signed pivots, which cannot fail
   * (a solve that outruns the precision ends with CLRSDP_ERR_DIVERGED or NOT_PD_X/Y instead), and the step-length
   * eigenvalue has an all-multiprecision fallback for the matrices its fast path flags. */
  CLRSDP_ERR_SINGULAR_S= -12, /* factorisation of S_j failed    (MPMP.jl:1438-1440)                     */
  CLRSDP_ERR_SINGULAR_Q= -13, /* factorisation of Q failed      (MPMP.jl:1502-1504)                     */
  CLRSDP_ERR_EIG       = -14, /* step-length eigen solve failed (MPMP.jl:1861-1862,1881-1884)           */
  CLRSDP_ERR_STATE     = -15, /* call order violated (e.g. iterate before a point exists)               */
  CLRSDP_ERR_DIVERGED  = -16  /* the iterate was lost: mu, a step length or an objective came out zero, negative or
                                 not finite. No reference counterpart (the reference walks on to maxiterations);
                                 the shim maps it to the same "higher precision" error as NOT_PD_X/Y. */
};

/* termination reasons reported in clrsdp_iter_info.terminate (MPMP.jl:1147-1173) */
enum {
  CLRSDP_RUNNING          = 0,
  CLRSDP_PRIMAL_FEASIBLE  = 1,  /* need_primal_feasible && primal_error < threshold */
  CLRSDP_DUAL_FEASIBLE    = 2,  /* need_dual_feasible   && dual_error   < threshold */
  CLRSDP_OPTIMAL          = 3,  /* primal & dual feasible and gap < threshold       */
  CLRSDP_MAXITER          = 4   /* iter reached maxiterations (MPMP.jl:752)          */
};

/* kwargs of solverank1sdp (MPMP.jl:599-613). The eight real-valued ones travel as ONE clrsdp_mp
 * of 8 numbers in this order so that they carry full working precision. */
enum {
  CLRSDP_P_BETA_INFEASIBLE = 0, /* 3/10   */
  CLRSDP_P_BETA_FEASIBLE   = 1, /* 1/10   */
  CLRSDP_P_GAMMA           = 2, /* 7/10   */
  CLRSDP_P_OMEGA_P         = 3, /* 1e10   */
  CLRSDP_P_OMEGA_D         = 4, /* 1e10   */
  CLRSDP_P_GAP_THRESHOLD   = 5, /* 1e-15  */
  CLRSDP_P_PRIMAL_ERR_THR  = 6, /* 1e-30  */
  CLRSDP_P_DUAL_ERR_THR    = 7, /* 1e-30  */
  CLRSDP_P_COUNT           = 8
};

typedef struct {
  int32_t maxiterations;        /* 500 (MPMP.jl:601); the loop runs while iter < maxiterations (:752) */
  int32_t need_primal_feasible; /* MPMP.jl:610 */
  int32_t need_dual_feasible;   /* MPMP.jl:611 */
  int32_t phase_timing;         /* != 0: fill the 17 per-phase buckets of clrsdp_iter_info on every iteration (event-record
                                   nodes inside the replayed CUDA graph, ~1 % of an iteration); 0: only `seconds`. The
                                   reference always collects them (MPMP.jl:889-898); the Python/Julia shims switch this
                                   on whenever they print the timing table (:972-1012). */
} clrsdp_int_params;

/* The 17 timing buckets of the reference (MPMP.jl:889-898), seconds for THIS iteration. */
enum {
  CLRSDP_T_DECOMP = 0, CLRSDP_T_PREDICTOR, CLRSDP_T_CORRECTOR, CLRSDP_T_ALPHA, CLRSDP_T_XINV,
  CLRSDP_T_R, CLRSDP_T_RES, CLRSDP_T_SCHUR, CLRSDP_T_CHOL_S, CLRSDP_T_CINVB, CLRSDP_T_Q,
  CLRSDP_T_CHOL_Q, CLRSDP_T_Z, CLRSDP_T_RHS_X, CLRSDP_T_SYS, CLRSDP_T_DX, CLRSDP_T_DY,
  CLRSDP_T_COUNT = 17
};

/* One row of the reference's log table (MPMP.jl:923-937) plus status. Values are those the
 * reference prints for iteration `iter`: mu, objectives, gap and the three errors are from the
 * START of the iteration; alpha/beta are the step just taken. `*_new` are the values after the
 * update that drive termination (MPMP.jl:940-953, including the stale-error quirk :943-944). */
typedef struct {
  int32_t iter;
  int32_t status;       /* CLRSDP_OK or an error code */
  int32_t terminate;    /* CLRSDP_RUNNING ... */
  int32_t pd_feasible;  /* check_pd_feasibility after this iteration (MPMP.jl:949-953) */
  double  mu, p_obj, d_obj, gap, P_err, p_err, d_err, alpha_p, alpha_d, beta_c;
  double  p_obj_new, d_obj_new, gap_new, primal_err_new, dual_err_new;
  double  seconds;      /* time of this iteration: CUDA events on the launching stream (oracle: wall clock) */
  double  timings[CLRSDP_T_COUNT];
} clrsdp_iter_info;

/* scalar slots for clrsdp_fetch(h, "scalar", slot, 0, out) — full-precision versions of the log row */
enum {
  CLRSDP_S_MU = 0, CLRSDP_S_P_OBJ, CLRSDP_S_D_OBJ, CLRSDP_S_GAP, CLRSDP_S_PRIMAL_ERR, CLRSDP_S_DUAL_ERR,
  CLRSDP_S_ALPHA_P, CLRSDP_S_ALPHA_D, CLRSDP_S_BETA_C, CLRSDP_S_MU_P, CLRSDP_S_MU_C,
  CLRSDP_S_LAMBDA_X, CLRSDP_S_LAMBDA_Y, CLRSDP_S_COUNT
};

/* ---- lifetime ----------------------------------------------------------------------------- */
/* prec_bits: multiple of 32 in [128,512] (precision(BigFloat), MPMP.jl:617). device: CUDA ordinal. */
int clrsdp_create(clrsdp_handle* h, int prec_bits, int device);
int clrsdp_destroy(clrsdp_handle h);
const char* clrsdp_last_error(clrsdp_handle h);

/* ---- problem (mirror of BlockInfo, MPMP.jl:467-479, and of the constraint tuple :401-406) ---- */
/* m,L,n_samples: [J]. delta: length of one vector of block (j,l), flattened j-major [sum L].
 * ranks: number of vectors at sample k of block (j,l), flattened [sum_j L[j]*n_samples[j]].
 * Block sizes follow as m[j]*delta[j][l] (MPMP.jl:550-551), dim_S[j] = m(m+1)/2*n_samples (:511). */
int clrsdp_set_structure(clrsdp_handle h, int J, int n_y, const int* m, const int* L,
                         const int* n_samples, const int* delta, const int* ranks);
/* Data of constraint j = the tuple (A,B,c,H) of prepareabc (MPMP.jl:385-406):
 *  V: for l, for k, for rnk: the delta[j][l] entries of A[l,k][rnk]    (hcat order of MPMP.jl:1249-1254)
 *  H: for l, for k, for rnk: A_sign[l,k][rnk]
 *  B: dim_S[j] x n_y row-major;  c: dim_S[j]. */
int clrsdp_upload_cluster(clrsdp_handle h, int j, const clrsdp_mp* V, const clrsdp_mp* H,
                          const clrsdp_mp* B, const clrsdp_mp* c);
/* b: [n_y] objective vector, b0: 1 number (MPMP.jl:597,600). */
int clrsdp_upload_objective(clrsdp_handle h, const clrsdp_mp* b, const clrsdp_mp* b0);
/* The kwarg C (MPMP.jl:599): objective matrix with the block structure of X - blocks in (j,l) order, each nb x nb
 * row-major, concatenated. It enters the residual P = sum_i x_i A_i - X - C (:1108-1118) and the dual objective
 * <C,Y> + <b,y> + b0 (:1031-1034, :1067-1070). Without this call, or with C == NULL / C->n == 0, C = 0 (the default,
 * the reference's AbsoluteZero :589-592, :691-695). With sharded clusters every rank passes the blocks of ITS clusters. */
int clrsdp_upload_C(clrsdp_handle h, const clrsdp_mp* C);
int clrsdp_set_params(clrsdp_handle h, const clrsdp_mp* real_params, const clrsdp_int_params* ip);

/* ---- iterate (MPMP.jl:659-695 / :689) ------------------------------------------------------ */
int clrsdp_init_point(clrsdp_handle h);  /* x=0, X=omega_p I, y=0, Y=omega_d I */
/* X, Y: blocks in (j,l) order, each nb x nb row-major, concatenated. */
int clrsdp_upload_point(clrsdp_handle h, const clrsdp_mp* x, const clrsdp_mp* X,
                        const clrsdp_mp* y, const clrsdp_mp* Y);
int clrsdp_download_point(clrsdp_handle h, clrsdp_mp_out* x, clrsdp_mp_out* X,
                          clrsdp_mp_out* y, clrsdp_mp_out* Y);

/* ---- the hot path -------------------------------------------------------------------------- */
/* Loop initialisation MPMP.jl:716-736: mu, objectives, gap, residuals and errors of the current
 * point (general trace_A method), pd_feas. Must be called once after a point exists. */
int clrsdp_prepare(clrsdp_handle h, clrsdp_iter_info* info);
/* One pass of the while-body MPMP.jl:754-953. */
int clrsdp_iterate(clrsdp_handle h, clrsdp_iter_info* info);
/* prepare + loop with terminate() (MPMP.jl:742-954). rows: optional array of `max_rows` entries that
 * receives one clrsdp_iter_info per iteration; n_rows: number written. */
int clrsdp_solve(clrsdp_handle h, clrsdp_iter_info* rows, int max_rows, int* n_rows);

/* ---- results / parity fetches -------------------------------------------------------------- */
/* name: "x","y","dx","dy","p","d","b","c" (vectors), "X","Y","Xinv","R","P","Z","dX","dY","XY"
 * (block (j,l), nb x nb), "S","Sfac","Sinvfac" (cluster j, dim_S x dim_S), "W" (L_j^{-1} B_j, dim_S x n_y),
 * "Q","Qfac" (n_y x n_y), "Px","Py" (pairings of block (j,l), m*Nv x m*Nv), "scalar" (slot j).
 * "Sfac" / "Sinvfac" are the factors of the EQUILIBRATED Schur complement S' = D^-1 S D^-1 (D a diagonal of powers
 * of two chosen per iteration): the library never forms L^-1 of S itself; "W" is L^-1 B exactly.
 * Returns the count written (>= 0) or a negative status. out may be NULL to query the count. */
int64_t clrsdp_fetch(clrsdp_handle h, const char* name, int j, int l, clrsdp_mp_out* out);

/* ---- phase-level entry points for parity tests (same math as inside clrsdp_iterate) --------- */
/* C[b] = A[b] * B[b] (row-major, batch of `batch` independent products) through the sliced int8
 * tensor-core path: slice -> tcgen05 kind::i8 planes -> carry/renormalise. (approx_mul!, MPMP.jl §2.3) */
int clrsdp_op_gemm(clrsdp_handle h, int batch, int M, int N, int K,
                   const clrsdp_mp* A, const clrsdp_mp* B, clrsdp_mp_out* C);
/* raw int32 digit planes of the same product (exact-integer parity): planes [n_planes][batch][M][N] */
int clrsdp_op_gemm_planes(clrsdp_handle h, int batch, int M, int N, int K, const clrsdp_mp* A,
                          const clrsdp_mp* B, int32_t* planes, int* n_planes, int32_t* row_exp,
                          int32_t* col_exp);
/* lower Cholesky factor / inverse factor of a batch of SPD n x n matrices (cho!, spd_inv!) */
int clrsdp_op_cholesky(clrsdp_handle h, int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* L,
                       clrsdp_mp_out* Linv);
/* The factorisation that stands in for the reference's pivoted LU of S_j and Q (approx_lu!, MPMP.jl:1436,1501): for
 * each symmetric (not necessarily definite) n x n matrix A of the batch, the signed factorisation of the equilibrated
 * matrix, A = D U^T Sigma U D with Sigma = diag(+-1) and D a diagonal of powers of two. Delivers M = U^-T D^-1
 * (n x n, lower triangular, row-major) and signs[b*n + k] = Sigma_kk in {-1,+1}, so that A^-1 = M^T Sigma M. */
int clrsdp_op_signed_factor(clrsdp_handle h, int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* Minv,
                            int32_t* signs);
/* smallest eigenvalue of each symmetric n x n matrix of a batch (approx_eig_qr! + min, :1857-1870) */
int clrsdp_op_lambda_min(clrsdp_handle h, int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* lam);
/* elementwise c = a (op) b, op in {'+','-','*','/'} and c = sqrt(a) with op 's' (scalar kernels' arithmetic) */
int clrsdp_op_elementwise(clrsdp_handle h, int op, const clrsdp_mp* a, const clrsdp_mp* b,
                          clrsdp_mp_out* c);

/* ---- multi-GPU (clusters sharded over GPUs) -------------------------------------- */
/* (a) ONE PROCESS, several GPUs behind one handle (what a Julia front end uses: no extra processes, no
 * rendezvous). dev_ids: n_dev distinct CUDA ordinals (NULL = 0..n_dev-1). The handle takes the WHOLE problem through
 * the same calls as a single-GPU handle (global cluster indices, iterates in global order); set_structure assigns the
 * clusters to the devices with clrsdp_partition on the weights clrsdp_cluster_weight, the library routes data to the
 * owning device and runs prepare / iterate / solve on all devices at once (one host thread per device inside the
 * library; communicators from one ncclUniqueId). n_dev == 1 gives a plain single-GPU handle. */
int clrsdp_create_multi(clrsdp_handle* h, int prec_bits, int n_dev, const int* dev_ids);
/* owner[j] = index (0..n_dev-1) of the device that holds cluster j (after set_structure; all 0 for one device) */
int clrsdp_cluster_owner(clrsdp_handle h, int J, int* owner);
/* F16 - the reference's distribute_weights_swapping (MPMP.jl:425-465; it spreads (j,l) blocks over threads, :492-499;
 * here: clusters over GPUs): sets of cardinalities differing by at most one, contiguous at first, then improved by swaps
 * between the heaviest and the lightest set. set_of[i] = set of item i; returns the largest set weight (< 0 on error). */
double clrsdp_partition(const double* weights, int n, int parts, int* set_of);
/* w_j of SURVEY §8e: 40 sum_l nb^3 + dim_S^3/3 + dim_S^2 n_y + n_y^2 dim_S  (delta: the L vector lengths of cluster j) */
double clrsdp_cluster_weight(int m, int L, int n_samples, const int* delta, int n_y);

/* (b) One process per GPU (torchrun-style). Rank 0 obtains an id, the host language broadcasts the 128 bytes (e.g.
 * torch.distributed.broadcast), every rank calls comm_init. After that set_structure/upload_cluster are given ONLY the
 * local clusters (any subset, e.g. the sets of clrsdp_partition).
 * In both modes the only collectives are the ones SURVEY §8e lists - Q (n_y^2), the n_y-vectors of the Schur solve and the
 * scalar sum / max / min reductions per iteration - inside prepare / iterate, as an all-gather of every rank's planes followed
 * by a rank-ordered combine (why not an int64-lane all-reduce). */
int clrsdp_comm_unique_id(uint8_t id[128]);
int clrsdp_comm_init(clrsdp_handle h, int n_ranks, int rank, const uint8_t id[128]);

/* Optional: register a long-lived caller buffer (e.g. the arrays of the iterate a front end keeps across
 * iterations) with the CUDA driver. upload_point / download_point / fetch DMA directly from / into wire tensors whose
 * sign, exp and limb arrays all lie inside registered ranges (no staging copy; the header words are converted on
 * the device). The buffer must stay allocated until clrsdp_unpin_host or clrsdp_destroy. No reference counterpart
 * (the reference keeps its iterate in Julia heap objects, MPMP.jl:660-686). */
int clrsdp_pin_host(clrsdp_handle h, void* p, size_t bytes);
int clrsdp_unpin_host(clrsdp_handle h, void* p);

/* measured dense int8 tensor rate of this device: int8 multiply-accumulates per second of back-to-back
 * tcgen05.mma kind::i8 (128x256x32) on every SM with resident operands. bench.py's roofline denominator for the
 * sliced GEMM (MEASURED_PEAKS.json has no int8 figure). No reference counterpart. */
int clrsdp_measure_int8_peak(clrsdp_handle h, double* macs_per_second);

/* number of kernels of this library launched since the handle was created (bench `gpu_launches`) */
int64_t clrsdp_launch_count(clrsdp_handle h);
/* CUDA-event time (ms) and launch count accumulated for kernels whose name contains `pattern`
 * since the last clrsdp_profile_reset; used by bench.py for the roofline of the dominant kernel. */
int clrsdp_profile_reset(clrsdp_handle h, int enable);
int clrsdp_profile_query(clrsdp_handle h, const char* pattern, double* ms, int64_t* launches,
                         double* work);
/* text table "name ms launches work" of every kernel timed since the last reset; returns bytes needed */
int clrsdp_profile_dump(clrsdp_handle h, char* buf, int buf_len);

#ifdef __cplusplus
}
#endif
#endif /* CLRSDP_H */

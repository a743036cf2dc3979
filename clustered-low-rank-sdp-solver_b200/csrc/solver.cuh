// solver.cuh — the handle behind the C ABI: problem structure, HBM-resident state, one IPM iteration.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/clrsdp.h"
#include "gemm_i8.cuh"
#include "linalg.cuh"
#include "comm.cuh"

namespace clr {

struct HostBlock {
  int j, l, m, delta, nb, Nv, np;
  std::vector<int> ranks, rank_sums;
  int64_t off, Voff, Hoff, Poff, Toff, VDoff, QPoff;
  int rs0;
  int group, idx_in_group;
};
struct HostCluster {
  int m, L, K, dimS, blk0, xoff;
  int64_t Soff;
  int group, idx_in_group;
};
// blocks of identical shape run as one batch
struct BlockGroup {
  int m, delta, nb, Nv, np;
  std::vector<int> blocks;
  DevBuf offBlk;                      // [nblk] block arena offsets
  DevBuf offBlk2, lamIdx2;            // [2 nblk] offsets into the paired (X|Y) arenas; output slots of lambda_min
  DevBuf g1_offA, g1_offC, g1_rowB;   // pairing GEMM 1 items (blk, s, r)
  DevBuf g2_offB, g2_offC, g2_rowA;   // pairing GEMM 2 items (blk, r)
  DevBuf wa_offA, wa_offC, wa_rowB;   // weighted-A GEMM items (blk, pair)
  Slice sVt, sVr;                     // static slices of the constraint vectors
  Slice sX, sY, sXinv, sA, sB, sLinv; // per-iteration slices
};
struct ClusterGroup {
  int dimS;
  std::vector<int> clusters;
  DevBuf offS, offBt, offW;
  Slice sBt;    // per iteration: rows a (n_y) of (D_j^-1 B_j)^T, K = dimS
  Slice sLinv;  // per iteration: rows of L'_j^-1
  DevBuf equil; // per iteration: [clusters][dimS] equilibration exponents of S_j (D_j = diag 2^s)
  DevBuf sig;   // per iteration: [clusters][dimS] signs of the pivots of S'_j = U^T Sigma U (1 = negative)
};

// What the C ABI needs from a handle: implemented by Solver (one GPU) and by MultiSolver (multi.cuh: several GPUs of one
// box behind one handle, one process).
struct SolverApi {
  virtual ~SolverApi() {}
  virtual void set_structure(int J, int n_y, const int* m, const int* L, const int* K, const int* delta, const int* ranks) = 0;
  virtual void upload_cluster(int j, const clrsdp_mp* V, const clrsdp_mp* H, const clrsdp_mp* B, const clrsdp_mp* c) = 0;
  virtual void upload_objective(const clrsdp_mp* b, const clrsdp_mp* b0) = 0;
  virtual void upload_C(const clrsdp_mp* C) = 0;
  virtual void set_params(const clrsdp_mp* rp, const clrsdp_int_params* ip) = 0;
  virtual void init_point() = 0;
  virtual void upload_point(const clrsdp_mp* x, const clrsdp_mp* X, const clrsdp_mp* y, const clrsdp_mp* Y) = 0;
  virtual void download_point(clrsdp_mp_out* x, clrsdp_mp_out* X, clrsdp_mp_out* y, clrsdp_mp_out* Y) = 0;
  virtual int prepare(clrsdp_iter_info* info) = 0;
  virtual int iterate(clrsdp_iter_info* info) = 0;
  virtual int solve(clrsdp_iter_info* rows, int max_rows, int* n_rows) = 0;
  virtual int64_t fetch(const char* name, int j, int l, clrsdp_mp_out* out) = 0;
  virtual void comm_init(int n_ranks, int rank, const uint8_t* id) = 0;
  virtual void pin_host(void* p, size_t bytes) = 0;
  virtual void unpin_host(void* p) = 0;
  virtual double measure_i8_peak() = 0;
  virtual void op_gemm(int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B, clrsdp_mp_out* C) = 0;
  virtual void op_gemm_planes(int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B, int32_t* planes,
                              int* n_planes, int32_t* row_exp, int32_t* col_exp) = 0;
  virtual int op_cholesky(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* L, clrsdp_mp_out* Linv) = 0;
  virtual void op_signed_factor(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* Minv, int32_t* signs) = 0;
  virtual void op_lambda_min(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* lam) = 0;
  virtual void op_elementwise(int op, const clrsdp_mp* a, const clrsdp_mp* b, clrsdp_mp_out* c) = 0;
  virtual int64_t launch_count() = 0;
  virtual void profile_reset(bool enable) = 0;
  virtual std::map<std::string, ProfEntry> profile_table() = 0;
};

class Solver : public SolverApi {
 public:
  Solver(int prec_bits, int device);
  ~Solver() override;
  void set_structure(int J, int n_y, const int* m, const int* L, const int* K, const int* delta, const int* ranks) override;
  void upload_cluster(int j, const clrsdp_mp* V, const clrsdp_mp* H, const clrsdp_mp* B, const clrsdp_mp* c) override;
  void upload_objective(const clrsdp_mp* b, const clrsdp_mp* b0) override;
  void upload_C(const clrsdp_mp* C) override;
  void set_params(const clrsdp_mp* rp, const clrsdp_int_params* ip) override;
  void init_point() override;
  void upload_point(const clrsdp_mp* x, const clrsdp_mp* X, const clrsdp_mp* y, const clrsdp_mp* Y) override;
  void download_point(clrsdp_mp_out* x, clrsdp_mp_out* X, clrsdp_mp_out* y, clrsdp_mp_out* Y) override;
  int prepare(clrsdp_iter_info* info) override;
  int iterate(clrsdp_iter_info* info) override;
  int solve(clrsdp_iter_info* rows, int max_rows, int* n_rows) override;
  int64_t fetch(const char* name, int j, int l, clrsdp_mp_out* out) override;
  void comm_init(int n_ranks, int rank, const uint8_t* id) override;
  void comm_abort();  // release collectives that wait for a failed peer (multi-device handle)
  void pin_host(void* p, size_t bytes) override;
  void unpin_host(void* p) override;
  double measure_i8_peak() override {
    CLR_CUDA(cudaSetDevice(ctx.device));
    return gemm_->measure_i8_peak();
  }
  // phase-level ops
  void op_gemm(int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B, clrsdp_mp_out* C) override;
  void op_gemm_planes(int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B, int32_t* planes,
                      int* n_planes, int32_t* row_exp, int32_t* col_exp) override;
  int op_cholesky(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* L, clrsdp_mp_out* Linv) override;
  void op_signed_factor(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* Minv, int32_t* signs) override;
  void op_lambda_min(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* lam) override;
  void op_elementwise(int op, const clrsdp_mp* a, const clrsdp_mp* b, clrsdp_mp_out* c) override;
  int64_t launch_count() override { return ctx.launches; }
  void profile_reset(bool enable) override {
    ctx.resolve();
    ctx.prof.clear();
    ctx.profiling = enable;
  }
  std::map<std::string, ProfEntry> profile_table() override {
    ctx.resolve();
    return ctx.prof;
  }

  Ctx ctx;
  int nl;
  int prec;
  std::string err;
  clrsdp_int_params ip{500, 0, 0, 0};

 private:
  // helpers
  void to_device(const clrsdp_mp* src, int64_t src_off, int64_t count, MpBuf& dst, int64_t dst_off);
  void to_host(const MpBuf& src, int64_t src_off, int64_t count, clrsdp_mp_out* dst, int64_t dst_off);
  bool is_pinned(const void* p, size_t bytes) const;
  bool wire_pinned(const int8_t* sign, const int64_t* exp, const uint32_t* limb, int64_t n) const;
  void upload_tables();
  void build_static_slices();
  MatBatch blkbatch(BlockGroup& g, MpBuf& t) { return MatBatch{t.t(), g.offBlk.as<int64_t>(), (int)g.blocks.size(), g.nb}; }
  MatBatch blkbatch2(BlockGroup& g, MpBuf& t) { return MatBatch{t.t(), g.offBlk2.as<int64_t>(), 2 * (int)g.blocks.size(), g.nb}; }
  OperandDesc rows_of(BlockGroup& g, MpBuf& t);
  OperandDesc cols_of(BlockGroup& g, MpBuf& t);
  OutDesc out_blk(BlockGroup& g, MpBuf& t, bool transposed = false);
  void block_products_XY();
  void invert_X();
  void pairings(MpBuf& M, bool is_xinv, MpBuf& Pout);
  void first_gemm_Tt(BlockGroup& g, MpBuf& M, Slice* cached);
  void weighted_A(MpBuf& a, MpBuf& out, MpBuf& E, int sign);
  void compute_residuals(bool from_pairings);
  void search_direction();
  void factor_XY();
  void step_lengths();
  void decomposition();
  // Linv = (chol A)^-1 for a batch of SPD matrices: blocked right-looking Cholesky, panels on the CUDA cores,
  // trailing updates and the off-diagonal inverse panels through the sliced tensor-core GEMM
  void chol_inverse(const MatBatch& A, const MatBatch& Uw, const MatBatch& Vw, const MatBatch& Linv, int* d_status,
                    bool want_u = false, bool side = false, int* d_sig = nullptr, int* d_keep_scale = nullptr);
  void product(GemmEngine* ge, Slice& sa, Slice& sb, const OperandDesc& a, const OperandDesc& b, int M, int N,
               const OutDesc& c, int epi, const mp::Tensor* extra);
  // a second stream for work that is off the critical path (see decomposition())
  void fork_side();
  void end_side();
  void join_side();
  int check_status();
  int check_status_local();
  void upload_ntot();
  void dot_CY();
  // all-reduce of t[off, off+n) over the ranks (no-op on one rank)
  void allreduce(MpBuf& t, int64_t off, int64_t n, int op);
  void mark(int bucket_begin);
  void iteration_body();
  void drop_graph();
  void prepare_body();

  std::vector<std::pair<const char*, size_t>> pinned_;  // host ranges registered through pin_host
  DevBuf equil_, equil_side_;                          // equilibration exponents of chol_inverse
  DevBuf xscale;                                       // [sumS] the exponents of the S_j, indexed like x
  DevBuf xsign, qsign;                                 // pivot signs of the S_j (indexed like x) and of Q ([n_y])
  Slice sWs_;                                          // Sigma W: the rows of Wt sliced with the signs on the contraction index
  DevBuf wire_se_;                                     // sign / exponent scratch of the pinned transfer path
  bool wire_defer_ = false;                            // inside a WireGroup: transfers do not synchronise one by one
  size_t wire_off_ = 0;                                // next free byte of wire_se_ inside a WireGroup
  struct WireGroup;
  std::unique_ptr<GemmEngine> gemm_, gemm_side_;
  cudaStream_t side_stream_ = nullptr, main_stream_ = nullptr;
  cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
  bool use_side_ = true, join_pending_ = false, on_side_ = false;
  GemmEngine* ge() { return on_side_ ? gemm_side_.get() : gemm_.get(); }
  Slice fs1s_, fs2s_;
  MpBuf tscr_side_;
  // helper streams for the inverse-panel chains of chol_inverse ([0] beside the main stream, [1] beside the side stream)
  struct InvHelper {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_go = nullptr, ev_done = nullptr;
    std::unique_ptr<GemmEngine> gemm;
    Slice s1, s2;
    MpBuf tscr;
  };
  InvHelper invh_[2];
  bool use_invh_ = true;
  InvHelper trailh_[2];            // lookahead: remainder of the trailing update beside the next panel (chol_inverse)
  bool use_lookahead_ = true;
  double lookahead_ratio_ = 2.0;   // split when (remainder)^2 >= ratio * (row block width) * (trailing width)
  int64_t lookahead_min_pmac_ = 1500000;  // ... and the remainder is a tensor-core product (= SMALL_GEMM_PMAC)
  Comm comm_;
  int ntot_local = 0;
  // structure
  int J = 0, n_y = 0, sumS = 0, ntot = 0;
  std::vector<HostBlock> blocks_;
  std::vector<HostCluster> clusters_;
  std::vector<BlockGroup> bgroups_;
  std::vector<ClusterGroup> cgroups_;
  int64_t blkN = 0, vtN = 0, hN = 0, pN = 0, tN = 0, vdN = 0, qpN = 0, sN = 0;
  bool structure_set = false, have_point = false, prepared = false, tables_ready = false;
  std::vector<int> uploaded_;
  // device tables
  DevBuf d_c_m, d_c_K, d_c_L, d_c_blk0, d_c_xoff, d_c_Soff, d_c_dimS, d_b_delta, d_b_Nv, d_b_cluster, d_b_rs0, d_b_Hoff,
      d_b_Poff, d_b_Voff, d_b_Toff, d_b_VDoff, d_b_QPoff, d_b_off, d_rank_sums, d_samp, d_x_cluster;
  DevBuf d_row_item, d_linv_off, d_x_off64, d_row0, d_itemK;
  StructTables st_;
  // arenas
  MpBuf XY2, dXY2, Linv2, U2, W2, T1d, T2d;  // paired arenas: X|Y, dX|dY, Lx^-1|Ly^-1, work
  MpBuf X, Y, Xinv, R, P, Z, dX, dY, XY, T1, T2, Ux, Vx, Linvx, Linvy;  // X,Y,dX,dY,Linv*,T1,T2,Ux are views
  MpBuf Vt, H, Px, Py, Tt, VD, QP, S, Us, Vs, Linvs, Bmat, BmatT, Wt, Q, Uq, Vq, Linvq;
  MpBuf x, dx, d, c, rhs, rhs0, tvec, tmpx, trx, y, dy, p, b, tmpy, zvec, dyr;
  MpBuf dX_pred, dY_pred, dx_pred, dy_pred;
  MpBuf Cmat;  // objective matrix C (block structure of X), only allocated by upload_C
  bool have_C = false;
  MpBuf scal, rdiag, lam, work, tscr;
  Slice fs1_, fs2_, sW_;
  // CUDA-graph replay of the iteration body (one launch instead of ~1000)
  bool use_graph_ = true, capturing_ = false;
  cudaGraphExec_t gexec_ = nullptr;
  bool ntot_uploaded_ = false;                       // the global size of X sits in its scalar slot (one-off per structure / communicator)
  uint64_t graph_epoch_ = 0;
  int direct_iters_ = 0;
  int64_t graph_launches_ = 0;
  DevBuf d_status, d_flags, d_scal_out, d_qoff, d_status_any, d_lamflag;
  std::vector<int> h_status;
  int n_status = 0;
  // iteration bookkeeping
  int iter = 1;
  struct Mark {
    int bucket, sign;  // +1 begin / -1 end
    cudaEvent_t ev;
  };
  std::vector<cudaEvent_t> ev_, gev_;             // events of directly recorded marks / of the event-record nodes of the graph
  std::vector<Mark> ev_marks_, graph_marks_;      // marks of this iteration / marks captured into the graph
  size_t n_direct_marks_ = 0;
  bool phase_timing_ = false;                     // clrsdp_int_params.phase_timing
  double h_scal[SL_COUNT];
  int h_flags[4];
};

}  // namespace clr

// multi.cuh — one handle, several GPUs, ONE process ("single process driving <= 8 GPUs").
//
// The clusters of a problem are independent except through y, so a MultiSolver owns one Solver per
// device, assigns every cluster to a device with the weighted partitioner below, and presents the interface of a single
// Solver on the WHOLE problem: problem data and iterates arrive and leave in global (cluster) order and are routed to
// the owning device; prepare / iterate / solve run on all devices at once (one persistent host thread per device,
// because the NCCL collectives inside them need all ranks in flight together). The host language - the Julia shim's
// `ccall`s, INTEGRATION.md - needs no processes, no torchrun and no rendezvous of its own.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "solver.cuh"

namespace clr {

// ---- weighted partition (F16: distribute_weights_swapping, MPMP.jl:425-465) ---------------------------------------------
// Splits the items 0..n-1 with the given weights into `parts` sets whose cardinalities differ by at most one and whose
// largest total weight is (approximately) minimal: contiguous initial sets, then swaps between the heaviest set and the
// lightest one while that lowers the maximum - the reference's algorithm for spreading its (j,l) blocks over threads
// (:492-499), used here to spread clusters over GPUs. set_of[i] receives the set of item i. Returns the largest set weight.
double partition_weights(const double* weights, int n, int parts, int* set_of, int64_t nswaps = -1);
// w_j of SURVEY §8e for a cluster: sum_l c1 nb^3 (block work: factorisations, products, step length) + dim_S^3 / 3
// + dim_S^2 n_y + n_y^2 dim_S (Schur factor, L^-1 B, Q)
double cluster_weight(int m, int L, int K, const int* delta, int n_y);

// ---- persistent worker threads: run f(rank) on every rank concurrently, rethrow the first failure ----------------------
class RankPool {
 public:
  explicit RankPool(int n);
  ~RankPool();
  void run(const std::function<void(int)>& f);
  int size() const { return (int)workers_.size(); }
  std::function<void(int)> on_error;  // called on the failing worker's thread, before run() returns

 private:
  void loop(int r);
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_go_, cv_done_;
  const std::function<void(int)>* task_ = nullptr;
  uint64_t epoch_ = 0;
  int pending_ = 0;
  bool stop_ = false;
  std::vector<std::exception_ptr> errors_;
};

class MultiSolver : public SolverApi {
 public:
  MultiSolver(int prec_bits, int n_dev, const int* dev_ids);
  ~MultiSolver() override;
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
  void comm_init(int, int, const uint8_t*) override {
    throw SolverError(CLRSDP_ERR_STATE, "a multi-device handle owns its communicator: clrsdp_comm_init is for one-process-per-GPU handles");
  }
  void pin_host(void* p, size_t bytes) override { ranks_[0]->pin_host(p, bytes); }
  void unpin_host(void* p) override { ranks_[0]->unpin_host(p); }
  double measure_i8_peak() override { return ranks_[0]->measure_i8_peak(); }
  void op_gemm(int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B, clrsdp_mp_out* C) override {
    ranks_[0]->op_gemm(batch, M, N, K, A, B, C);
  }
  void op_gemm_planes(int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B, int32_t* planes, int* n_planes,
                      int32_t* row_exp, int32_t* col_exp) override {
    ranks_[0]->op_gemm_planes(batch, M, N, K, A, B, planes, n_planes, row_exp, col_exp);
  }
  int op_cholesky(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* L, clrsdp_mp_out* Linv) override {
    return ranks_[0]->op_cholesky(batch, n, A, L, Linv);
  }
  void op_signed_factor(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* Minv, int32_t* signs) override {
    ranks_[0]->op_signed_factor(batch, n, A, Minv, signs);
  }
  void op_lambda_min(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* lam) override { ranks_[0]->op_lambda_min(batch, n, A, lam); }
  void op_elementwise(int op, const clrsdp_mp* a, const clrsdp_mp* b, clrsdp_mp_out* c) override {
    ranks_[0]->op_elementwise(op, a, b, c);
  }
  int64_t launch_count() override;
  void profile_reset(bool enable) override;
  std::map<std::string, ProfEntry> profile_table() override;
  // the cluster -> device assignment chosen by set_structure (device index in [0, n_dev), not the CUDA ordinal)
  const std::vector<int>& owner() const { return owner_; }

 private:
  struct Piece {  // a host copy of part of a wire tensor, with the view the per-device Solver takes
    std::vector<int8_t> sign;
    std::vector<int64_t> exp;
    std::vector<uint32_t> limb;
    clrsdp_mp view{nullptr, nullptr, nullptr, 0};
    clrsdp_mp_out out{nullptr, nullptr, nullptr, 0};
    void resize(int64_t n, int nl);
  };
  // ranges [begin, begin + len) of a global array that belong to rank r, in local order
  struct Span {
    int64_t begin, len;
  };
  void gather_in(const clrsdp_mp* src, const std::vector<Span>& spans, Piece& dst) const;
  void scatter_out(const Piece& src, const std::vector<Span>& spans, clrsdp_mp_out* dst) const;
  int run_all(const std::function<int(int)>& f);

  int nl_, prec_, n_ = 0;
  std::vector<std::unique_ptr<Solver>> ranks_;
  std::unique_ptr<RankPool> pool_;
  std::atomic<bool> aborted_{false};              // a rank failed: the communicators have been aborted
  // global structure
  int J_ = 0, n_y_ = 0;
  int64_t sumS_ = 0, blkN_ = 0;
  std::vector<int> owner_, local_j_, c_L_, c_dimS_;
  std::vector<int64_t> c_xoff_, c_blkoff_, c_blklen_;
  std::vector<std::vector<int64_t>> blk_off_;      // [j][l] offset of block (j,l) in the global X arena
  std::vector<std::vector<Span>> x_spans_, X_spans_;  // per rank
};

}  // namespace clr

// multi.cu — weighted cluster partition (F16) and the single-process multi-GPU handle (see multi.cuh).
#include "multi.cuh"

#include <algorithm>
#include <cstring>
#include <numeric>

namespace clr {

// ---------------------------------------------------------------------------------------------------------------------
// distribute_weights_swapping (MPMP.jl:425-465): `parts` contiguous sets of sizes step = n / parts + 1 (the first
// n - parts (step - 1) of them) and step - 1; then up to n^2 attempts to swap the candidate element of the candidate
// heavy set with the lightest element of the lightest set, accepted when neither set ends above the heavy set's weight.
// The candidates walk exactly like the reference's (index_el through the heavy set's elements by decreasing weight,
// index_set through the sets by decreasing weight; both reset after a successful swap).
// ---------------------------------------------------------------------------------------------------------------------
double partition_weights(const double* w, int n, int parts, int* set_of, int64_t nswaps) {
  if (parts < 1) throw SolverError(CLRSDP_ERR_BAD_ARG, "partition: parts must be positive");
  if (nswaps < 0) nswaps = (int64_t)n * n;
  const int step = n / parts + 1;
  const int nstep = parts - (step * parts - n);  // sets with `step` elements
  std::vector<std::vector<int>> sets(parts);
  {
    int at = 0;
    for (int i = 0; i < parts; i++) {
      int len = i < nstep ? step : step - 1;
      for (int k = 0; k < len; k++) sets[i].push_back(at++);
    }
  }
  std::vector<double> sw(parts, 0.0);
  for (int i = 0; i < parts; i++)
    for (int e : sets[i]) sw[i] += w[e];
  int index_set = 0, index_el = 0;  // 0-based counterparts of the reference's 1-based cursors
  for (int64_t k = 0; k < nswaps && n > 0; k++) {
    // the index_set-th heaviest set (ties: larger index first, as sort(rev=true) on (weight, index) tuples gives)
    std::vector<int> order(parts);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return sw[a] != sw[b] ? sw[a] > sw[b] : a > b; });
    const int max_set = order[std::min(index_set, parts - 1)];
    if (sets[max_set].empty()) break;
    // its index_el-th heaviest element (ties: later position first)
    std::vector<int> eo(sets[max_set].size());
    std::iota(eo.begin(), eo.end(), 0);
    std::sort(eo.begin(), eo.end(), [&](int a, int b) {
      double wa = w[sets[max_set][a]], wb = w[sets[max_set][b]];
      return wa != wb ? wa > wb : a > b;
    });
    const int max_pos = eo[std::min<size_t>(index_el, eo.size() - 1)];
    const int max_el = sets[max_set][max_pos];
    const int min_set = (int)(std::min_element(sw.begin(), sw.end()) - sw.begin());  // argmin: first minimum
    if (sets[min_set].empty()) break;
    int min_pos = 0;
    for (size_t i = 1; i < sets[min_set].size(); i++)
      if (w[sets[min_set][i]] < w[sets[min_set][min_pos]]) min_pos = (int)i;
    const int min_el = sets[min_set][min_pos];
    if (sw[min_set] + w[max_el] - w[min_el] < sw[max_set] && sw[max_set] - w[max_el] + w[min_el] < sw[max_set]) {
      sets[max_set].erase(sets[max_set].begin() + max_pos);
      sets[max_set].push_back(min_el);
      sw[max_set] += w[min_el] - w[max_el];
      // (max_set == min_set cannot pass the test above: it would need w[max_el] < w[min_el])
      auto it = std::find(sets[min_set].begin(), sets[min_set].end(), min_el);
      sets[min_set].erase(it);
      sets[min_set].push_back(max_el);
      sw[min_set] += w[max_el] - w[min_el];
      index_el = 0;
      index_set = 0;
    } else if (index_el + 1 < (int)sets[std::min(index_set, parts - 1)].size()) {  // (:453: length(sets[index_set]))
      index_el += 1;
    } else if (index_el + 1 == step - 1 && index_set + 1 < parts - 1) {            // (:455)
      index_set += 1;
      index_el = 0;
    } else {
      break;
    }
  }
  for (int i = 0; i < parts; i++)
    for (int e : sets[i]) set_of[e] = i;
  return n ? *std::max_element(sw.begin(), sw.end()) : 0.0;
}

double cluster_weight(int m, int L, int K, const int* delta, int n_y) {
  const double dimS = 0.5 * m * (m + 1) * K;
  double wgt = dimS * dimS * dimS / 3.0 + dimS * dimS * n_y + (double)n_y * n_y * dimS;
  for (int l = 0; l < L; l++) {
    const double nb = (double)m * delta[l];
    wgt += 40.0 * nb * nb * nb;  // c1: factorisations of X and Y, X^-1, XY, dXdY, Z, dY, the step-length products (Appendix C)
  }
  return wgt;
}

// ---------------------------------------------------------------------------------------------------------------------
RankPool::RankPool(int n) : errors_(n) {
  for (int r = 0; r < n; r++) workers_.emplace_back([this, r] { loop(r); });
}
RankPool::~RankPool() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
    epoch_++;
  }
  cv_go_.notify_all();
  for (auto& t : workers_) t.join();
}
void RankPool::loop(int r) {
  uint64_t seen = 0;
  for (;;) {
    const std::function<void(int)>* task;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_go_.wait(lk, [&] { return epoch_ != seen; });
      seen = epoch_;
      if (stop_) return;
      task = task_;
    }
    try {
      (*task)(r);
    } catch (...) {
      errors_[r] = std::current_exception();
      // the other ranks may already sit in a collective that waits for this one: release them, or run() never returns
      if (on_error) on_error(r);
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) cv_done_.notify_all();
    }
  }
}
void RankPool::run(const std::function<void(int)>& f) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& e : errors_) e = nullptr;
    task_ = &f;
    pending_ = (int)workers_.size();
    epoch_++;
  }
  cv_go_.notify_all();
  {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return pending_ == 0; });
  }
  for (auto& e : errors_)
    if (e) std::rethrow_exception(e);
}

// ---------------------------------------------------------------------------------------------------------------------
void MultiSolver::Piece::resize(int64_t n, int nl) {
  sign.assign((size_t)std::max<int64_t>(n, 1), 0);
  exp.assign((size_t)std::max<int64_t>(n, 1), 0);
  limb.assign((size_t)std::max<int64_t>(n, 1) * nl, 0u);
  view = clrsdp_mp{sign.data(), exp.data(), limb.data(), n};
  out = clrsdp_mp_out{sign.data(), exp.data(), limb.data(), n};
}

MultiSolver::MultiSolver(int prec_bits, int n_dev, const int* dev_ids) : nl_(prec_bits / 32), prec_(prec_bits), n_(n_dev) {
  if (n_dev < 2) throw SolverError(CLRSDP_ERR_BAD_ARG, "a multi-device handle needs at least two devices");
  for (int r = 0; r < n_dev; r++) {
    for (int q = 0; q < r; q++)
      if (dev_ids && dev_ids[q] == dev_ids[r]) throw SolverError(CLRSDP_ERR_BAD_ARG, "create_multi: duplicate device ordinal");
    ranks_.emplace_back(new Solver(prec_bits, dev_ids ? dev_ids[r] : r));
  }
  pool_.reset(new RankPool(n_dev));
  pool_->on_error = [this](int) {
    bool expected = false;
    if (!aborted_.compare_exchange_strong(expected, true)) return;
    for (auto& s : ranks_) s->comm_abort();
  };
  ncclUniqueId uid;
  CLR_NCCL(NcclApi::get().GetUniqueId(&uid));
  // ncclCommInitRank blocks until every rank has joined: one thread per rank
  pool_->run([&](int r) { ranks_[r]->comm_init(n_, r, (const uint8_t*)uid.internal); });
}
MultiSolver::~MultiSolver() {
  pool_.reset();
  ranks_.clear();
}

int MultiSolver::run_all(const std::function<int(int)>& f) {
  std::vector<int> st(n_, 0);
  pool_->run([&](int r) { st[r] = f(r); });
  for (int r = 0; r < n_; r++)
    if (st[r]) return st[r];
  return 0;
}

void MultiSolver::gather_in(const clrsdp_mp* src, const std::vector<Span>& spans, Piece& dst) const {
  int64_t tot = 0;
  for (auto& s : spans) tot += s.len;
  dst.resize(tot, nl_);
  int64_t at = 0;
  for (auto& s : spans) {
    if (s.begin < 0 || s.begin + s.len > src->n) throw SolverError(CLRSDP_ERR_BAD_ARG, "multi: array shorter than the structure");
    memcpy(dst.sign.data() + at, src->sign + s.begin, (size_t)s.len);
    memcpy(dst.exp.data() + at, src->exp + s.begin, (size_t)s.len * 8);
    for (int k = 0; k < nl_; k++)
      memcpy(dst.limb.data() + (size_t)k * tot + at, src->limb + (size_t)k * src->n + s.begin, (size_t)s.len * 4);
    at += s.len;
  }
}
void MultiSolver::scatter_out(const Piece& src, const std::vector<Span>& spans, clrsdp_mp_out* dst) const {
  const int64_t tot = src.view.n;
  int64_t at = 0;
  for (auto& s : spans) {
    if (s.begin + s.len > dst->n) throw SolverError(CLRSDP_ERR_BAD_ARG, "multi: output array too small");
    memcpy(dst->sign + s.begin, src.sign.data() + at, (size_t)s.len);
    memcpy(dst->exp + s.begin, src.exp.data() + at, (size_t)s.len * 8);
    for (int k = 0; k < nl_; k++)
      memcpy(dst->limb + (size_t)k * dst->n + s.begin, src.limb.data() + (size_t)k * tot + at, (size_t)s.len * 4);
    at += s.len;
  }
}

void MultiSolver::set_structure(int J, int n_y, const int* m, const int* L, const int* K, const int* delta, const int* ranks) {
  if (J < n_) throw SolverError(CLRSDP_ERR_BAD_ARG, "multi: fewer clusters than devices");
  J_ = J, n_y_ = n_y;
  // weights and the partition (F16)
  std::vector<double> w(J);
  std::vector<int> d0(J), r0(J);
  int di = 0, ri = 0;
  for (int j = 0; j < J; j++) {
    d0[j] = di, r0[j] = ri;
    if (m[j] <= 0 || L[j] <= 0 || K[j] <= 0) throw SolverError(CLRSDP_ERR_BAD_ARG, "set_structure: bad m/L/n_samples");
    w[j] = cluster_weight(m[j], L[j], K[j], delta + di, n_y);
    di += L[j];
    ri += L[j] * K[j];
  }
  owner_.assign(J, 0);
  partition_weights(w.data(), J, n_, owner_.data());
  // global offsets
  local_j_.assign(J, 0);
  c_L_.assign(L, L + J);
  c_dimS_.assign(J, 0);
  c_xoff_.assign(J, 0);
  c_blkoff_.assign(J, 0);
  c_blklen_.assign(J, 0);
  blk_off_.assign(J, {});
  sumS_ = blkN_ = 0;
  for (int j = 0; j < J; j++) {
    c_dimS_[j] = m[j] * (m[j] + 1) / 2 * K[j];
    c_xoff_[j] = sumS_;
    sumS_ += c_dimS_[j];
    c_blkoff_[j] = blkN_;
    for (int l = 0; l < L[j]; l++) {
      blk_off_[j].push_back(blkN_);
      int64_t nb = (int64_t)m[j] * delta[d0[j] + l];
      blkN_ += nb * nb;
    }
    c_blklen_[j] = blkN_ - c_blkoff_[j];
  }
  x_spans_.assign(n_, {});
  X_spans_.assign(n_, {});
  std::vector<std::vector<int>> lm(n_), lL(n_), lK(n_), ldelta(n_), lranks(n_);
  std::vector<int> cnt(n_, 0);
  for (int j = 0; j < J; j++) {
    const int r = owner_[j];
    local_j_[j] = cnt[r]++;
    lm[r].push_back(m[j]), lL[r].push_back(L[j]), lK[r].push_back(K[j]);
    ldelta[r].insert(ldelta[r].end(), delta + d0[j], delta + d0[j] + L[j]);
    lranks[r].insert(lranks[r].end(), ranks + r0[j], ranks + r0[j] + L[j] * K[j]);
    x_spans_[r].push_back(Span{c_xoff_[j], c_dimS_[j]});
    X_spans_[r].push_back(Span{c_blkoff_[j], c_blklen_[j]});
  }
  // Solver::set_structure ends with an all-reduce (the global size of X): all ranks together
  pool_->run([&](int r) {
    ranks_[r]->set_structure(cnt[r], n_y, lm[r].data(), lL[r].data(), lK[r].data(), ldelta[r].data(), lranks[r].data());
  });
}

void MultiSolver::upload_cluster(int j, const clrsdp_mp* V, const clrsdp_mp* H, const clrsdp_mp* B, const clrsdp_mp* c) {
  if (j < 0 || j >= J_) throw SolverError(CLRSDP_ERR_BAD_ARG, "upload_cluster: bad cluster index");
  ranks_[owner_[j]]->upload_cluster(local_j_[j], V, H, B, c);
}
void MultiSolver::upload_objective(const clrsdp_mp* b, const clrsdp_mp* b0) {
  for (auto& s : ranks_) s->upload_objective(b, b0);
}
void MultiSolver::upload_C(const clrsdp_mp* C) {
  if (!C || C->n == 0) {
    for (auto& s : ranks_) s->upload_C(nullptr);
    return;
  }
  if (C->n != blkN_) throw SolverError(CLRSDP_ERR_BAD_ARG, "upload_C: C must have the block structure of X");
  for (int r = 0; r < n_; r++) {
    Piece p;
    gather_in(C, X_spans_[r], p);
    ranks_[r]->upload_C(&p.view);
  }
}
void MultiSolver::set_params(const clrsdp_mp* rp, const clrsdp_int_params* ip) {
  for (auto& s : ranks_) s->set_params(rp, ip);
}
void MultiSolver::init_point() {
  for (auto& s : ranks_) s->init_point();
}
void MultiSolver::upload_point(const clrsdp_mp* x, const clrsdp_mp* X, const clrsdp_mp* y, const clrsdp_mp* Y) {
  if (x->n != sumS_ || X->n != blkN_ || Y->n != blkN_ || y->n != n_y_)
    throw SolverError(CLRSDP_ERR_BAD_ARG, "upload_point: sizes do not match the structure");
  pool_->run([&](int r) {
    Piece px, pX, pY;
    gather_in(x, x_spans_[r], px);
    gather_in(X, X_spans_[r], pX);
    gather_in(Y, X_spans_[r], pY);
    ranks_[r]->upload_point(&px.view, &pX.view, y, &pY.view);
  });
}
void MultiSolver::download_point(clrsdp_mp_out* x, clrsdp_mp_out* X, clrsdp_mp_out* y, clrsdp_mp_out* Y) {
  pool_->run([&](int r) {
    Piece px, pX, pY;
    int64_t nx = 0, nX = 0;
    for (auto& s : x_spans_[r]) nx += s.len;
    for (auto& s : X_spans_[r]) nX += s.len;
    px.resize(nx, nl_), pX.resize(nX, nl_), pY.resize(nX, nl_);
    ranks_[r]->download_point(x ? &px.out : nullptr, X ? &pX.out : nullptr, (y && r == 0) ? y : nullptr, Y ? &pY.out : nullptr);
    if (x) scatter_out(px, x_spans_[r], x);  // the spans of different ranks are disjoint
    if (X) scatter_out(pX, X_spans_[r], X);
    if (Y) scatter_out(pY, X_spans_[r], Y);
  });
}

int MultiSolver::prepare(clrsdp_iter_info* info) {
  std::vector<clrsdp_iter_info> rows(n_);
  int st = run_all([&](int r) { return ranks_[r]->prepare(&rows[r]); });
  if (info) *info = rows[0];
  if (info && st) info->status = st;
  return st;
}
int MultiSolver::iterate(clrsdp_iter_info* info) {
  std::vector<clrsdp_iter_info> rows(n_);
  int st = run_all([&](int r) { return ranks_[r]->iterate(&rows[r]); });
  if (info) {
    *info = rows[0];
    for (int r = 1; r < n_; r++) info->seconds = std::max(info->seconds, rows[r].seconds);  // device time: max over the ranks
    if (st) info->status = st;
  }
  return st;
}
int MultiSolver::solve(clrsdp_iter_info* rows, int max_rows, int* n_rows) {
  std::vector<int> n(n_, 0);
  std::vector<std::vector<clrsdp_iter_info>> rr(n_);
  int st = run_all([&](int r) {
    rr[r].resize(r == 0 ? (size_t)std::max(max_rows, 0) : 0);
    return ranks_[r]->solve(r == 0 ? rr[0].data() : nullptr, r == 0 ? max_rows : 0, &n[r]);
  });
  if (rows)
    for (int i = 0; i < std::min(n[0], max_rows); i++) rows[i] = rr[0][i];
  if (n_rows) *n_rows = n[0];
  return st;
}

int64_t MultiSolver::fetch(const char* name, int j, int l, clrsdp_mp_out* out) {
  const std::string nm(name);
  static const char* xvecs[] = {"x", "dx", "d", "c", "dx_pred"};
  for (const char* v : xvecs)
    if (nm == v) {  // sharded vectors: gathered into global order
      if (!out) return sumS_;
      if (out->n < sumS_) return CLRSDP_ERR_BAD_ARG;
      for (int r = 0; r < n_; r++) {
        Piece p;
        int64_t nx = 0;
        for (auto& s : x_spans_[r]) nx += s.len;
        p.resize(nx, nl_);
        int64_t got = ranks_[r]->fetch(name, 0, 0, &p.out);
        if (got < 0) return got;
        scatter_out(p, x_spans_[r], out);
      }
      return sumS_;
    }
  static const char* per_cluster[] = {"X", "Y", "Xinv", "R", "P", "Z", "dX", "dY", "XY", "dX_pred", "dY_pred", "Linvx", "Linvy",
                                      "Px", "Py", "S", "Sfac", "Sinvfac"};
  for (const char* v : per_cluster)
    if (nm == v) {
      if (j < 0 || j >= J_) return CLRSDP_ERR_BAD_ARG;
      return ranks_[owner_[j]]->fetch(name, local_j_[j], l, out);
    }
  if (nm == "W") {  // rows a of W^T are sharded along the constraint index: [n_y][sum dim_S] in global order
    const int64_t tot = sumS_ * n_y_;
    if (!out) return tot;
    if (out->n < tot) return CLRSDP_ERR_BAD_ARG;
    for (int r = 0; r < n_; r++) {
      int64_t nx = 0;
      for (auto& s : x_spans_[r]) nx += s.len;
      Piece p;
      p.resize(nx * n_y_, nl_);
      int64_t got = ranks_[r]->fetch(name, 0, 0, &p.out);
      if (got < 0) return got;
      for (int a = 0; a < n_y_; a++) {
        int64_t at = (int64_t)a * nx;
        for (auto& s : x_spans_[r]) {
          const int64_t to = (int64_t)a * sumS_ + s.begin;
          memcpy(out->sign + to, p.sign.data() + at, (size_t)s.len);
          memcpy(out->exp + to, p.exp.data() + at, (size_t)s.len * 8);
          for (int k = 0; k < nl_; k++)
            memcpy(out->limb + (size_t)k * out->n + to, p.limb.data() + (size_t)k * p.view.n + at, (size_t)s.len * 4);
          at += s.len;
        }
      }
    }
    return tot;
  }
  return ranks_[0]->fetch(name, j, l, out);  // replicated: y, dy, p, b, Q, Qfac, scalars
}

int64_t MultiSolver::launch_count() {
  int64_t n = 0;
  for (auto& s : ranks_) n += s->launch_count();
  return n;
}
void MultiSolver::profile_reset(bool enable) {
  for (auto& s : ranks_) s->profile_reset(enable);
}
std::map<std::string, ProfEntry> MultiSolver::profile_table() { return ranks_[0]->profile_table(); }

}  // namespace clr

// solver.cu — orchestration of one interior-point iteration on the GPU (MPMP.jl:742-954) + the C ABI.
// Host code only launches kernels; all arithmetic (including the driver scalars mu, beta, alpha) runs
// on the device. One stream, no host synchronisation inside an iteration except the final read-back
// of the log row.
#include "solver.cuh"

#include <cmath>
#include <cstdlib>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <thread>
#include <tuple>

namespace clr {

static double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------------------
Solver::Solver(int prec_bits, int device) : nl(prec_bits / 32), prec(prec_bits) {
  if (prec_bits % 32 || !(nl == 4 || nl == 8 || nl == 12 || nl == 16))
    throw SolverError(CLRSDP_ERR_BAD_ARG, "precision must be 128, 256, 384 or 512 bits");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw SolverError(CLRSDP_ERR_CUDA, "no CUDA device: the hot path has no CPU fallback");
  if (device < 0 || device >= ndev) throw SolverError(CLRSDP_ERR_BAD_ARG, "bad device ordinal");
  ctx.device = device;
  CLR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CLR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    throw SolverError(CLRSDP_ERR_CUDA, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                           ", this library is built for sm_100a (B200) only");
  ctx.sm_count = prop.multiProcessorCount;
  CLR_CUDA(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
  gemm_.reset(new GemmEngine(ctx, nl));
  gemm_side_.reset(new GemmEngine(ctx, nl));  // own block workspace: used by the branch that runs beside the main stream
  CLR_CUDA(cudaStreamCreateWithFlags(&side_stream_, cudaStreamNonBlocking));
  CLR_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  CLR_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
  for (auto& h : invh_) {  // the inverse-panel chains of chol_inverse (one beside the main, one beside the side stream)
    h.gemm.reset(new GemmEngine(ctx, nl));
    CLR_CUDA(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
    CLR_CUDA(cudaEventCreateWithFlags(&h.ev_go, cudaEventDisableTiming));
    CLR_CUDA(cudaEventCreateWithFlags(&h.ev_done, cudaEventDisableTiming));
  }
  for (auto& h : trailh_) {  // the remainder updates of the lookahead factorisation (see chol_inverse)
    h.gemm.reset(new GemmEngine(ctx, nl));
    CLR_CUDA(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
    CLR_CUDA(cudaEventCreateWithFlags(&h.ev_go, cudaEventDisableTiming));
    CLR_CUDA(cudaEventCreateWithFlags(&h.ev_done, cudaEventDisableTiming));
  }
  if (const char* g = getenv("CLRSDP_INVH")) use_invh_ = atoi(g) != 0;
  if (const char* g = getenv("CLRSDP_LOOKAHEAD")) use_lookahead_ = atoi(g) != 0;
  if (const char* g = getenv("CLRSDP_LOOKAHEAD_RATIO")) lookahead_ratio_ = atof(g);
  if (const char* g = getenv("CLRSDP_LOOKAHEAD_MIN_PMAC")) lookahead_min_pmac_ = atoll(g);  // (tests: split small products too)
  scal.alloc(SL_COUNT, nl);
  work.alloc(4096, nl);
  d_flags.ensure(4 * sizeof(int));
  d_scal_out.ensure(SL_COUNT * sizeof(double));
  CLR_CUDA(cudaMemsetAsync(d_flags.p, 0, 4 * sizeof(int), ctx.stream));
  ew_zero(ctx, nl, scal.t(), 0, SL_COUNT);
  // the real-valued parameters (MPMP.jl:602-609) arrive at full precision through set_params
  memset(h_scal, 0, sizeof(h_scal));
  memset(h_flags, 0, sizeof(h_flags));
  if (const char* g = getenv("CLRSDP_GRAPH")) use_graph_ = atoi(g) != 0;
  if (const char* g = getenv("CLRSDP_SIDE")) use_side_ = atoi(g) != 0;
}

Solver::~Solver() {
  cudaSetDevice(ctx.device);
  for (auto& r : pinned_) cudaHostUnregister((void*)r.first);
  drop_graph();
  if (comm_.comm) NcclApi::get().CommDestroy(comm_.comm);
  for (auto e : ev_) cudaEventDestroy(e);
  for (auto e : gev_) cudaEventDestroy(e);
  for (auto e : ctx.pool) cudaEventDestroy(e);
  if (main_stream_) ctx.stream = main_stream_;
  if (ctx.stream) cudaStreamDestroy(ctx.stream);
  if (side_stream_) cudaStreamDestroy(side_stream_);
  if (ev_fork_) cudaEventDestroy(ev_fork_);
  if (ev_join_) cudaEventDestroy(ev_join_);
  for (auto* hs : {invh_, trailh_})
    for (int i = 0; i < 2; i++) {
      auto& h = hs[i];
      if (h.stream) cudaStreamDestroy(h.stream);
      if (h.ev_go) cudaEventDestroy(h.ev_go);
      if (h.ev_done) cudaEventDestroy(h.ev_done);
    }
}

// ---- multi-GPU ------------------------------------------------------------------------------------------
void Solver::comm_init(int n_ranks, int rank, const uint8_t* id) {
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw SolverError(CLRSDP_ERR_BAD_ARG, "comm_init: bad rank");
  CLR_CUDA(cudaSetDevice(ctx.device));
  if (n_ranks == 1) return;
  ncclUniqueId uid;
  memcpy(uid.internal, id, NCCL_UNIQUE_ID_BYTES);
  CLR_NCCL(NcclApi::get().CommInitRank(&comm_.comm, n_ranks, uid, rank));
  comm_.nranks = n_ranks;
  comm_.rank = rank;
  prepared = false;
  ntot_uploaded_ = false;
  drop_graph();
  direct_iters_ = 0;
}
// Called (from any thread) when ANOTHER rank of a single-process multi-device handle has failed: the collectives this
// rank has queued would wait for that rank for ever. ncclCommAbort releases them; the handle is unusable afterwards.
void Solver::comm_abort() {
  ncclComm_t c = comm_.comm;
  if (!c) return;
  comm_.comm = nullptr;
  NcclApi::get().CommAbort(c);
  prepared = false;
}
void Solver::allreduce(MpBuf& t, int64_t off, int64_t n, int op) {
  if (!comm_.active() || n <= 0) return;
  size_t words = (size_t)(nl + 1) * n;
  comm_.stage.ensure(words * sizeof(uint32_t));
  comm_.gathered.ensure(words * sizeof(uint32_t) * comm_.nranks);
  // pack the planes of the region contiguously: [(nl+1)][n]
  CLR_CUDA(cudaMemcpy2DAsync(comm_.stage.p, (size_t)n * 4, t.w() + off, t.n * 4, (size_t)n * 4, nl + 1,
                             cudaMemcpyDeviceToDevice, ctx.stream));
  CLR_NCCL(NcclApi::get().AllGather(comm_.stage.p, comm_.gathered.p, words, ncclUint32, comm_.comm, ctx.stream));
  combine_ranks(ctx, nl, comm_.gathered.as<uint32_t>(), comm_.nranks, n, t.t(), off, op);
}

// ---- wire <-> device -------------------------------------------------------------------------------
// Wire format <-> planar HBM layout. The limb planes have the same order on both sides, so a transfer is one strided
// DMA ((nl+1) rows of `count` words, cudaMemcpy2DAsync) through a pinned, grow-only staging buffer; the host side only
// moves the limb planes into / out of the staging buffer and converts sign/exp <-> header word, split over a few
// threads (this copy sits inside the timed region of the end-to-end path: upload_point / download_point).
namespace {
struct PinnedStage {
  uint32_t* p = nullptr;
  size_t words = 0;
  ~PinnedStage() {
    if (p) cudaFreeHost(p);
  }
  uint32_t* ensure(size_t w) {
    if (w > words) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      CLR_CUDA(cudaHostAlloc((void**)&p, w * sizeof(uint32_t), cudaHostAllocDefault));
      words = w;
    }
    return p;
  }
};
PinnedStage& stage_buf() {
  static thread_local PinnedStage s;
  return s;
}
template <class F>
void parallel_ranges(int64_t count, F&& f) {
  int nt = 1;
  if (count >= (1 << 16)) nt = (int)std::min<int64_t>(8, std::max(1u, std::thread::hardware_concurrency()));
  if (nt <= 1) {
    f((int64_t)0, count);
    return;
  }
  std::vector<std::thread> th;
  int64_t per = (count + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    int64_t lo = t * per, hi = std::min(count, lo + per);
    if (lo < hi) th.emplace_back([=, &f] { f(lo, hi); });
  }
  for (auto& t : th) t.join();
}
}  // namespace

// ---- caller-pinned host buffers: DMA straight from / into the wire arrays ------------------------------
// clrsdp_pin_host registers a long-lived caller buffer with the CUDA driver. A wire tensor whose three arrays all lie
// in registered ranges skips the staging copy: the limb planes go by one strided DMA, sign/exp by two small copies and
// the header words are packed / unpacked by a kernel (linalg.cu: wire_pack / wire_unpack).
void Solver::pin_host(void* p, size_t bytes) {
  if (!p || !bytes) throw SolverError(CLRSDP_ERR_BAD_ARG, "pin_host: empty range");
  CLR_CUDA(cudaSetDevice(ctx.device));
  for (auto& r : pinned_)
    if (r.first == (const char*)p && r.second >= bytes) return;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw SolverError(CLRSDP_ERR_CUDA, std::string("pin_host: cudaHostRegister failed: ") + cudaGetErrorString(e));
  }
  pinned_.emplace_back((const char*)p, bytes);
}
void Solver::unpin_host(void* p) {
  CLR_CUDA(cudaSetDevice(ctx.device));
  for (size_t i = 0; i < pinned_.size(); i++)
    if (pinned_[i].first == (const char*)p) {
      CLR_CUDA(cudaStreamSynchronize(ctx.stream));
      cudaHostUnregister(p);
      pinned_.erase(pinned_.begin() + i);
      return;
    }
  throw SolverError(CLRSDP_ERR_BAD_ARG, "unpin_host: range was not pinned");
}
bool Solver::is_pinned(const void* p, size_t bytes) const {
  const char* c = (const char*)p;
  for (auto& r : pinned_)
    if (c >= r.first && c + bytes <= r.first + r.second) return true;
  return false;
}
bool Solver::wire_pinned(const int8_t* sign, const int64_t* exp, const uint32_t* limb, int64_t n) const {
  return n >= 4096 && is_pinned(sign, (size_t)n) && is_pinned(exp, (size_t)n * 8) && is_pinned(limb, (size_t)n * 4 * nl);
}

void Solver::to_device(const clrsdp_mp* src, int64_t src_off, int64_t count, MpBuf& dst, int64_t dst_off) {
  if (count <= 0) return;
  if (wire_pinned(src->sign, src->exp, src->limb, src->n)) {
    // sign / exponent scratch: one region per transfer of a deferred group (upload_point / download_point move four
    // tensors and synchronise ONCE at the end instead of after every tensor), else the start of the buffer
    if (!wire_defer_) wire_se_.ensure((size_t)count * 9 + 16), wire_off_ = 0;
    int64_t* d_exp = reinterpret_cast<int64_t*>(wire_se_.as<char>() + wire_off_);
    int8_t* d_sign = reinterpret_cast<int8_t*>(d_exp + count);
    if (wire_defer_) wire_off_ += (((size_t)count * 9 + 15) / 16) * 16;
    CLR_CUDA(cudaMemcpy2DAsync(dst.w() + dst_off, dst.n * sizeof(uint32_t), src->limb + src_off, (size_t)src->n * sizeof(uint32_t),
                               (size_t)count * sizeof(uint32_t), (size_t)nl, cudaMemcpyHostToDevice, ctx.stream));
    CLR_CUDA(cudaMemcpyAsync(d_sign, src->sign + src_off, (size_t)count, cudaMemcpyHostToDevice, ctx.stream));
    CLR_CUDA(cudaMemcpyAsync(d_exp, src->exp + src_off, (size_t)count * 8, cudaMemcpyHostToDevice, ctx.stream));
    wire_pack(ctx, nl, dst.t(), dst_off, count, d_sign, d_exp);
    if (!wire_defer_) CLR_CUDA(cudaStreamSynchronize(ctx.stream));  // the scratch is reused by the next transfer
    return;
  }
  uint32_t* stage = stage_buf().ensure((size_t)(nl + 1) * count);
  const int nlv = nl;
  std::atomic<bool> bad_wire{false};
  parallel_ranges(count, [&](int64_t lo, int64_t hi) {
    for (int k = 0; k < nlv; k++)
      memcpy(stage + (size_t)k * count + lo, src->limb + (size_t)k * src->n + src_off + lo, (size_t)(hi - lo) * 4);
    uint32_t* hdr = stage + (size_t)nlv * count;
    for (int64_t i = lo; i < hi; i++) {
      int8_t sg = src->sign[src_off + i];
      if (sg == 0) {
        hdr[i] = mp::pack_hdr(mp::EXP_ZERO, 0);
        for (int k = 0; k < nlv; k++) stage[(size_t)k * count + i] = 0;
      } else {
        const int64_t e = src->exp[src_off + i];
        // the header word holds the exponent in 31 bits and reserves -2^28 for zero; a mantissa must be normalised
        if (e <= -(1 << 27) || e >= (1 << 27) || !(src->limb[(size_t)(nlv - 1) * src->n + src_off + i] & 0x80000000u)) bad_wire = true;
        hdr[i] = mp::pack_hdr((int32_t)e, sg < 0 ? 1u : 0u);
      }
    }
  });
  if (bad_wire) throw SolverError(CLRSDP_ERR_BAD_ARG, "wire tensor: exponent outside (-2^27, 2^27) or mantissa not normalised (top bit of the top limb clear)");
  CLR_CUDA(cudaMemcpy2DAsync(dst.w() + dst_off, dst.n * sizeof(uint32_t), stage, (size_t)count * sizeof(uint32_t),
                             (size_t)count * sizeof(uint32_t), (size_t)(nl + 1), cudaMemcpyHostToDevice, ctx.stream));
  CLR_CUDA(cudaStreamSynchronize(ctx.stream));  // the staging buffer is reused by the next transfer
}
void Solver::to_host(const MpBuf& src, int64_t src_off, int64_t count, clrsdp_mp_out* dst, int64_t dst_off) {
  if (count <= 0) return;
  if (wire_pinned(dst->sign, dst->exp, dst->limb, dst->n)) {
    if (!wire_defer_) wire_se_.ensure((size_t)count * 9 + 16), wire_off_ = 0;
    int64_t* d_exp = reinterpret_cast<int64_t*>(wire_se_.as<char>() + wire_off_);
    int8_t* d_sign = reinterpret_cast<int8_t*>(d_exp + count);
    if (wire_defer_) wire_off_ += (((size_t)count * 9 + 15) / 16) * 16;
    wire_unpack(ctx, nl, src.t(), src_off, count, d_sign, d_exp);
    CLR_CUDA(cudaMemcpy2DAsync(dst->limb + dst_off, (size_t)dst->n * sizeof(uint32_t), src.w() + src_off, src.n * sizeof(uint32_t),
                               (size_t)count * sizeof(uint32_t), (size_t)nl, cudaMemcpyDeviceToHost, ctx.stream));
    CLR_CUDA(cudaMemcpyAsync(dst->sign + dst_off, d_sign, (size_t)count, cudaMemcpyDeviceToHost, ctx.stream));
    CLR_CUDA(cudaMemcpyAsync(dst->exp + dst_off, d_exp, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx.stream));
    if (!wire_defer_) CLR_CUDA(cudaStreamSynchronize(ctx.stream));
    return;
  }
  uint32_t* stage = stage_buf().ensure((size_t)(nl + 1) * count);
  CLR_CUDA(cudaMemcpy2DAsync(stage, (size_t)count * sizeof(uint32_t), src.w() + src_off, src.n * sizeof(uint32_t),
                             (size_t)count * sizeof(uint32_t), (size_t)(nl + 1), cudaMemcpyDeviceToHost, ctx.stream));
  CLR_CUDA(cudaStreamSynchronize(ctx.stream));
  const int nlv = nl;
  parallel_ranges(count, [&](int64_t lo, int64_t hi) {
    const uint32_t* hdr = stage + (size_t)nlv * count;
    bool any_zero = false;
    for (int64_t i = lo; i < hi; i++) {
      int32_t e = ((int32_t)hdr[i]) >> 1;
      bool z = (e == mp::EXP_ZERO);
      any_zero |= z;
      dst->sign[dst_off + i] = z ? 0 : ((hdr[i] & 1u) ? -1 : 1);
      dst->exp[dst_off + i] = z ? 0 : e;
    }
    for (int k = 0; k < nlv; k++)
      memcpy(dst->limb + (size_t)k * dst->n + dst_off + lo, stage + (size_t)k * count + lo, (size_t)(hi - lo) * 4);
    if (any_zero)
      for (int64_t i = lo; i < hi; i++)
        if ((((int32_t)hdr[i]) >> 1) == mp::EXP_ZERO)
          for (int k = 0; k < nlv; k++) dst->limb[(size_t)k * dst->n + dst_off + i] = 0u;
  });
}

// ---- structure ------------------------------------------------------------------------------------
void Solver::set_structure(int J_, int n_y_, const int* m, const int* L, const int* K, const int* delta,
                           const int* ranks) {
  if (J_ <= 0 || n_y_ <= 0) throw SolverError(CLRSDP_ERR_BAD_ARG, "set_structure: J and n_y must be positive");
  CLR_CUDA(cudaSetDevice(ctx.device));
  J = J_;
  n_y = n_y_;
  blocks_.clear();
  clusters_.clear();
  bgroups_.clear();
  cgroups_.clear();
  blkN = vtN = hN = pN = tN = vdN = qpN = sN = 0;
  sumS = ntot = 0;
  int di = 0, ri = 0, rs_total = 0;
  std::map<std::tuple<int, int, int>, int> bkey;
  std::map<int, int> ckey;
  for (int j = 0; j < J; j++) {
    HostCluster c;
    c.m = m[j];
    c.L = L[j];
    c.K = K[j];
    if (c.m <= 0 || c.L <= 0 || c.K <= 0) throw SolverError(CLRSDP_ERR_BAD_ARG, "set_structure: bad m/L/n_samples");
    c.dimS = c.m * (c.m + 1) / 2 * c.K;
    c.blk0 = (int)blocks_.size();
    c.xoff = sumS;
    c.Soff = sN;
    sN += (int64_t)c.dimS * c.dimS;
    sumS += c.dimS;
    auto it = ckey.find(c.dimS);
    if (it == ckey.end()) {
      ckey[c.dimS] = (int)cgroups_.size();
      cgroups_.emplace_back();
      cgroups_.back().dimS = c.dimS;
      it = ckey.find(c.dimS);
    }
    c.group = it->second;
    c.idx_in_group = (int)cgroups_[c.group].clusters.size();
    cgroups_[c.group].clusters.push_back(j);
    for (int l = 0; l < c.L; l++) {
      HostBlock bk;
      bk.j = j;
      bk.l = l;
      bk.m = c.m;
      bk.delta = delta[di++];
      bk.nb = c.m * bk.delta;
      bk.np = c.m * (c.m + 1) / 2;
      bk.ranks.assign(ranks + ri, ranks + ri + c.K);
      ri += c.K;
      bk.rank_sums.assign(c.K + 1, 0);
      for (int k = 0; k < c.K; k++) {
        if (bk.ranks[k] < 0) throw SolverError(CLRSDP_ERR_BAD_ARG, "set_structure: negative rank");
        bk.rank_sums[k + 1] = bk.rank_sums[k] + bk.ranks[k];
      }
      bk.Nv = bk.rank_sums[c.K];
      if (bk.delta <= 0 || bk.Nv <= 0) throw SolverError(CLRSDP_ERR_BAD_ARG, "set_structure: empty block");
      bk.off = blkN;
      blkN += (int64_t)bk.nb * bk.nb;
      bk.Voff = vtN;
      vtN += (int64_t)bk.Nv * bk.delta;
      bk.Hoff = hN;
      hN += bk.Nv;
      bk.Poff = pN;
      pN += (int64_t)(c.m * bk.Nv) * (c.m * bk.Nv);
      bk.Toff = tN;
      tN += (int64_t)c.m * c.m * bk.Nv * bk.delta;
      bk.VDoff = vdN;
      vdN += (int64_t)bk.np * bk.delta * bk.Nv;
      bk.QPoff = qpN;
      qpN += (int64_t)bk.np * bk.delta * bk.delta;
      bk.rs0 = rs_total;
      rs_total += c.K + 1;
      ntot += bk.nb;
      auto key = std::make_tuple(c.m, bk.delta, bk.Nv);
      auto bt = bkey.find(key);
      if (bt == bkey.end()) {
        bkey[key] = (int)bgroups_.size();
        bgroups_.emplace_back();
        BlockGroup& g = bgroups_.back();
        g.m = c.m, g.delta = bk.delta, g.nb = bk.nb, g.Nv = bk.Nv, g.np = bk.np;
        bt = bkey.find(key);
      }
      bk.group = bt->second;
      bk.idx_in_group = (int)bgroups_[bk.group].blocks.size();
      bgroups_[bk.group].blocks.push_back((int)blocks_.size());
      blocks_.push_back(bk);
    }
    clusters_.push_back(c);
  }
  // arenas
  for (MpBuf* t : {&XY2, &dXY2, &Linv2, &U2, &W2, &T1d, &T2d}) t->alloc(2 * blkN, nl);
  X.alias(XY2, 0, blkN), Y.alias(XY2, blkN, blkN);
  dX.alias(dXY2, 0, blkN), dY.alias(dXY2, blkN, blkN);
  Linvx.alias(Linv2, 0, blkN), Linvy.alias(Linv2, blkN, blkN);
  Ux.alias(U2, 0, blkN), Vx.alias(U2, blkN, blkN);
  T1.alias(T1d, 0, blkN), T2.alias(T2d, 0, blkN);
  for (MpBuf* t : {&Xinv, &R, &P, &Z, &XY, &dX_pred, &dY_pred}) t->alloc(blkN, nl);
  have_C = false;  // C = 0 until upload_C (the reference's AbsoluteZero, MPMP.jl:691-695)
  Vt.alloc(vtN, nl);
  H.alloc(hN, nl);
  Px.alloc(pN, nl);
  Py.alloc(pN, nl);
  Tt.alloc(tN, nl);
  VD.alloc(vdN, nl);
  QP.alloc(qpN, nl);
  for (MpBuf* t : {&S, &Us, &Vs, &Linvs}) t->alloc(sN, nl);
  Bmat.alloc((int64_t)sumS * n_y, nl);
  BmatT.alloc((int64_t)sumS * n_y, nl);  // B^T (n_y x sum dim_S, row-major): the row operand of W = L'^-1 D^-1 B, sliced coalesced
  Wt.alloc((int64_t)sumS * n_y, nl);
  for (MpBuf* t : {&Q, &Uq, &Vq, &Linvq}) t->alloc((int64_t)n_y * n_y, nl);
  for (MpBuf* t : {&x, &dx, &d, &c, &rhs, &rhs0, &tvec, &tmpx, &trx, &dx_pred}) t->alloc(sumS, nl);
  for (MpBuf* t : {&y, &dy, &p, &b, &tmpy, &zvec, &dyr, &dy_pred}) t->alloc(n_y, nl);
  xscale.ensure(sizeof(int) * (size_t)std::max(sumS, 1));
  xsign.ensure(sizeof(int) * (size_t)std::max(sumS, 1));
  qsign.ensure(sizeof(int) * (size_t)std::max(n_y, 1));
  int64_t maxrd = n_y;
  for (auto& g : bgroups_) maxrd = std::max<int64_t>(maxrd, (int64_t)g.blocks.size() * g.nb);
  for (auto& g : cgroups_) maxrd = std::max<int64_t>(maxrd, (int64_t)g.clusters.size() * g.dimS);
  rdiag.alloc(maxrd, nl);
  lam.alloc(2 * blocks_.size(), nl);
  work.alloc(std::max<size_t>({(size_t)4096, reduce_work_elems(), gemv_work_elems(n_y, sumS), gemv_work_elems(sumS, n_y),
                                 gemv_work_elems(sumS, sumS), gemv_work_elems(n_y, n_y)}), nl);
  n_status = 2 * (int)blocks_.size() + J + 1;  // X blocks, Y blocks, S_j, Q
  d_status.ensure(sizeof(int) * n_status);
  d_lamflag.ensure(sizeof(int) * 2 * std::max<size_t>(1, blocks_.size()));
  h_status.assign(n_status, 0);
  uploaded_.assign(J, 0);
  structure_set = true;
  drop_graph();
  direct_iters_ = 0;
  ntot_uploaded_ = false;
  have_point = prepared = tables_ready = false;
  upload_tables();
}

template <class T>
static void up(DevBuf& d, const std::vector<T>& h, cudaStream_t s) {
  upload(d, h, s);
}

void Solver::upload_tables() {
  std::vector<int> c_m, c_K, c_L, c_blk0, c_xoff, c_dimS, b_delta, b_Nv, b_cluster, b_rs0, rank_sums, samp, x_cluster;
  std::vector<int64_t> c_Soff, b_Hoff, b_Poff, b_Voff, b_Toff, b_VDoff, b_QPoff, b_off;
  for (auto& c : clusters_) {
    c_m.push_back(c.m), c_K.push_back(c.K), c_L.push_back(c.L), c_blk0.push_back(c.blk0), c_xoff.push_back(c.xoff);
    c_dimS.push_back(c.dimS), c_Soff.push_back(c.Soff);
  }
  for (auto& bk : blocks_) {
    b_delta.push_back(bk.delta), b_Nv.push_back(bk.Nv), b_cluster.push_back(bk.j), b_rs0.push_back(bk.rs0);
    b_Hoff.push_back(bk.Hoff), b_Poff.push_back(bk.Poff), b_Voff.push_back(bk.Voff), b_Toff.push_back(bk.Toff);
    b_VDoff.push_back(bk.VDoff), b_QPoff.push_back(bk.QPoff), b_off.push_back(bk.off);
    rank_sums.insert(rank_sums.end(), bk.rank_sums.begin(), bk.rank_sums.end());
    for (int k = 0; k < clusters_[bk.j].K; k++)
      for (int r = 0; r < bk.ranks[k]; r++) samp.push_back(k);
  }
  std::vector<int> row_item, row0, itemK;
  std::vector<int64_t> linv_off, x_off64;
  for (int j = 0; j < J; j++) {
    for (int i = 0; i < clusters_[j].dimS; i++) x_cluster.push_back(j), row_item.push_back(j);
    row0.push_back(clusters_[j].xoff);
    itemK.push_back(clusters_[j].dimS);
    linv_off.push_back(clusters_[j].Soff);
    x_off64.push_back(clusters_[j].xoff);
  }
  cudaStream_t s = ctx.stream;
  up(d_c_m, c_m, s), up(d_c_K, c_K, s), up(d_c_L, c_L, s), up(d_c_blk0, c_blk0, s), up(d_c_xoff, c_xoff, s);
  up(d_c_dimS, c_dimS, s), up(d_c_Soff, c_Soff, s);
  up(d_b_delta, b_delta, s), up(d_b_Nv, b_Nv, s), up(d_b_cluster, b_cluster, s), up(d_b_rs0, b_rs0, s);
  up(d_b_Hoff, b_Hoff, s), up(d_b_Poff, b_Poff, s), up(d_b_Voff, b_Voff, s), up(d_b_Toff, b_Toff, s);
  up(d_b_VDoff, b_VDoff, s), up(d_b_QPoff, b_QPoff, s), up(d_b_off, b_off, s);
  up(d_rank_sums, rank_sums, s), up(d_samp, samp, s), up(d_x_cluster, x_cluster, s);
  up(d_row_item, row_item, s), up(d_row0, row0, s), up(d_itemK, itemK, s), up(d_linv_off, linv_off, s),
      up(d_x_off64, x_off64, s);
  std::vector<int64_t> qoff = {0};
  up(d_qoff, qoff, s);
  st_ = StructTables();
  st_.J = J, st_.n_y = n_y, st_.n_blocks = (int)blocks_.size(), st_.sumS = sumS;
  st_.c_m = d_c_m.as<int>(), st_.c_K = d_c_K.as<int>(), st_.c_L = d_c_L.as<int>(), st_.c_blk0 = d_c_blk0.as<int>();
  st_.c_xoff = d_c_xoff.as<int>(), st_.c_Soff = d_c_Soff.as<int64_t>(), st_.c_dimS = d_c_dimS.as<int>();
  st_.b_delta = d_b_delta.as<int>(), st_.b_Nv = d_b_Nv.as<int>(), st_.b_cluster = d_b_cluster.as<int>();
  st_.b_rs0 = d_b_rs0.as<int>(), st_.b_Hoff = d_b_Hoff.as<int64_t>(), st_.b_Poff = d_b_Poff.as<int64_t>();
  st_.b_Voff = d_b_Voff.as<int64_t>(), st_.b_Toff = d_b_Toff.as<int64_t>(), st_.b_VDoff = d_b_VDoff.as<int64_t>();
  st_.b_QPoff = d_b_QPoff.as<int64_t>(), st_.b_off = d_b_off.as<int64_t>();
  st_.rank_sums = d_rank_sums.as<int>(), st_.samp = d_samp.as<int>(), st_.x_cluster = d_x_cluster.as<int>();
  // per-group batch tables
  for (auto& g : bgroups_) {
    int nblk = (int)g.blocks.size(), m = g.m;
    std::vector<int64_t> offBlk, offBlk2, g1A, g1C, g2B, g2C, waA, waC;
    std::vector<int> g1rB, g2rA, warB, lamIdx2;
    for (int q = 0; q < nblk; q++) {
      const HostBlock& bk = blocks_[g.blocks[q]];
      offBlk.push_back(bk.off);
      for (int sidx = 0; sidx < m; sidx++)
        for (int r = 0; r < m; r++) {  // item (blk, s, r): rows i of the (r,s) sub-block
          g1A.push_back(bk.off + (int64_t)(r * g.delta) * g.nb + sidx * g.delta);
          g1C.push_back(bk.Toff + (int64_t)(r * m + sidx) * g.Nv * g.delta);
          g1rB.push_back(q * g.Nv);
        }
      for (int r = 0; r < m; r++) {  // item (blk, r)
        g2B.push_back(bk.Toff + (int64_t)r * m * g.Nv * g.delta);
        g2C.push_back(bk.Poff + (int64_t)r * g.Nv * (m * g.Nv));
        g2rA.push_back(q * g.Nv);
      }
      for (int pr = 0; pr < g.np; pr++) {  // item (blk, pair)
        waA.push_back(bk.VDoff + (int64_t)pr * g.delta * g.Nv);
        waC.push_back(bk.QPoff + (int64_t)pr * g.delta * g.delta);
        warB.push_back(q * g.delta);
      }
    }
    offBlk2 = offBlk;
    for (int q = 0; q < nblk; q++) offBlk2.push_back(offBlk[q] + blkN);
    for (int q = 0; q < nblk; q++) lamIdx2.push_back(g.blocks[q]);
    for (int q = 0; q < nblk; q++) lamIdx2.push_back((int)blocks_.size() + g.blocks[q]);
    up(g.offBlk2, offBlk2, s), up(g.lamIdx2, lamIdx2, s);
    up(g.offBlk, offBlk, s), up(g.g1_offA, g1A, s), up(g.g1_offC, g1C, s), up(g.g1_rowB, g1rB, s);
    up(g.g2_offB, g2B, s), up(g.g2_offC, g2C, s), up(g.g2_rowA, g2rA, s);
    up(g.wa_offA, waA, s), up(g.wa_offC, waC, s), up(g.wa_rowB, warB, s);
  }
  for (auto& g : cgroups_) {
    std::vector<int64_t> offS, offBt, offW;
    for (int j : g.clusters) {
      offS.push_back(clusters_[j].Soff);
      offBt.push_back((int64_t)clusters_[j].xoff);  // into BmatT: columns [xoff_j, xoff_j + dim_S_j) of every row
      offW.push_back(clusters_[j].xoff);
    }
    up(g.offS, offS, s), up(g.offBt, offBt, s), up(g.offW, offW, s);
  }
  CLR_CUDA(cudaStreamSynchronize(s));
  ntot_local = ntot;
  upload_ntot();
  tables_ready = true;
}

// n = size(X,1) (MPMP.jl:755) as an mp scalar; with sharded clusters the local sizes are summed over the ranks
void Solver::upload_ntot() {
  int8_t sg = 1;
  int64_t ex = 0;
  std::vector<uint32_t> limbs(nl, 0);
  int v = ntot_local, bl = 0;
  while ((1ll << bl) <= v) bl++;
  uint64_t top = (uint64_t)v << (64 - bl);
  limbs[nl - 1] = (uint32_t)(top >> 32);
  limbs[nl - 2] = (uint32_t)top;
  ex = bl;
  clrsdp_mp one{&sg, &ex, limbs.data(), 1};
  to_device(&one, 0, 1, scal, SL_NTOT);
  allreduce(scal, SL_NTOT, 1, COMB_SUM);
}

void Solver::upload_cluster(int j, const clrsdp_mp* V, const clrsdp_mp* Hh, const clrsdp_mp* B, const clrsdp_mp* cc) {
  if (!structure_set || j < 0 || j >= J) throw SolverError(CLRSDP_ERR_BAD_ARG, "upload_cluster: bad cluster index");
  CLR_CUDA(cudaSetDevice(ctx.device));
  const HostCluster& cl = clusters_[j];
  int64_t vneed = 0, hneed = 0;
  for (int l = 0; l < cl.L; l++) {
    vneed += (int64_t)blocks_[cl.blk0 + l].Nv * blocks_[cl.blk0 + l].delta;
    hneed += blocks_[cl.blk0 + l].Nv;
  }
  if (V->n != vneed || Hh->n != hneed || B->n != (int64_t)cl.dimS * n_y || cc->n != cl.dimS)
    throw SolverError(CLRSDP_ERR_BAD_ARG, "upload_cluster: array sizes do not match the structure");
  // blocks of a cluster are contiguous in the V and H arenas, in upload order
  to_device(V, 0, vneed, Vt, blocks_[cl.blk0].Voff);
  to_device(Hh, 0, hneed, H, blocks_[cl.blk0].Hoff);
  to_device(B, 0, B->n, Bmat, (int64_t)cl.xoff * n_y);
  to_device(cc, 0, cl.dimS, c, cl.xoff);
  uploaded_[j] = 1;
  bool all = true;
  for (int u : uploaded_) all = all && u;
  if (all) build_static_slices();
}

void Solver::build_static_slices() {
  {  // BmatT(a, r) = Bmat(r, a): B is static, so the transposed reads of its slicing (rows a, contraction over the
     // constraints of a cluster: a stride of n_y words between consecutive k) are paid once here instead of every iteration
    SmallGemmArgs cp;
    cp.A = Bmat.t(), cp.ars = 1, cp.aks = n_y;
    cp.C = BmatT.t(), cp.crs = sumS, cp.ccs = 1;
    cp.batch = 1, cp.M = n_y, cp.N = sumS;
    rect_copy(ctx, nl, cp);
  }
  for (auto& g : bgroups_) {
    int nblk = (int)g.blocks.size();
    std::vector<int64_t> voff;
    for (int q : g.blocks) voff.push_back(blocks_[q].Voff);
    DevBuf dv;
    upload(dv, voff, ctx.stream);
    OperandDesc a;
    a.src = Vt.t();
    a.d_off = dv.as<int64_t>();
    a.batch = nblk;
    a.rows = g.Nv, a.K = g.delta, a.rs = g.delta, a.ks = 1;  // rows = vectors
    ge()->slice(a, g.sVt);
    a.rows = g.delta, a.K = g.Nv, a.rs = 1, a.ks = g.delta;  // rows = basis index, K = vectors
    ge()->slice(a, g.sVr);
    ctx.sync();
  }
  // (the rows of B_j^T are sliced every iteration, scaled by that iteration's equilibration of S_j: decomposition())
  ctx.sync();
}

void Solver::upload_objective(const clrsdp_mp* bb, const clrsdp_mp* b0) {
  if (!structure_set || bb->n != n_y) throw SolverError(CLRSDP_ERR_BAD_ARG, "upload_objective: b must have n_y entries");
  CLR_CUDA(cudaSetDevice(ctx.device));
  to_device(bb, 0, n_y, b, 0);
  if (b0 && b0->n >= 1)
    to_device(b0, 0, 1, scal, SL_B0);
  else
    ew_zero(ctx, nl, scal.t(), SL_B0, 1);
}

// objective matrix C (MPMP.jl:599): P = sum x_i A_i - X - C (:1116-1118), dual objective <C,Y> + <b,y> + b0 (:1032-1034)
void Solver::upload_C(const clrsdp_mp* Cw) {
  if (!structure_set) throw SolverError(CLRSDP_ERR_STATE, "upload_C before set_structure");
  CLR_CUDA(cudaSetDevice(ctx.device));
  drop_graph();
  direct_iters_ = 0;
  prepared = false;
  if (!Cw || Cw->n == 0) {
    have_C = false;
    ew_zero(ctx, nl, scal.t(), SL_CY, 1);
    return;
  }
  if (Cw->n != blkN) throw SolverError(CLRSDP_ERR_BAD_ARG, "upload_C: C must have the block structure of X");
  Cmat.alloc(blkN, nl);
  to_device(Cw, 0, blkN, Cmat, 0);
  have_C = true;
}

void Solver::set_params(const clrsdp_mp* rp, const clrsdp_int_params* ipp) {
  CLR_CUDA(cudaSetDevice(ctx.device));
  if (rp) {
    if (rp->n != CLRSDP_P_COUNT) throw SolverError(CLRSDP_ERR_BAD_ARG, "set_params: expected 8 real parameters");
    const int slots[CLRSDP_P_COUNT] = {SL_BETA_INF, SL_BETA_FEAS, SL_GAMMA,    SL_OMEGA_P,
                                       SL_OMEGA_D,  SL_GAP_THR,   SL_PERR_THR, SL_DERR_THR};
    for (int i = 0; i < CLRSDP_P_COUNT; i++) to_device(rp, i, 1, scal, slots[i]);
  }
  if (ipp) ip = *ipp;
  phase_timing_ = ip.phase_timing != 0;
  drop_graph();
  int fl[4] = {0, 0, ip.need_primal_feasible, ip.need_dual_feasible};
  CLR_CUDA(cudaMemcpyAsync(d_flags.as<int>() + 2, fl + 2, 2 * sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
  ctx.sync();
}

// ---- point ----------------------------------------------------------------------------------------
void Solver::init_point() {  // MPMP.jl:660-686
  if (!structure_set) throw SolverError(CLRSDP_ERR_STATE, "init_point before set_structure");
  CLR_CUDA(cudaSetDevice(ctx.device));
  ew_zero(ctx, nl, x.t(), 0, sumS);
  ew_zero(ctx, nl, y.t(), 0, n_y);
  for (auto& g : bgroups_) {
    ew_set_identity(ctx, nl, blkbatch(g, X), scal.t(), SL_OMEGA_P);
    ew_set_identity(ctx, nl, blkbatch(g, Y), scal.t(), SL_OMEGA_D);
  }
  have_point = true;
  prepared = false;
}
// A group of transfers through the pinned path that share one synchronisation: the scratch is sized for all of them up
// front (no reallocation while copies are in flight) and every transfer takes its own region of it.
struct Solver::WireGroup {
  Solver& s;
  bool done = false;
  WireGroup(Solver& s_, size_t numbers) : s(s_) {
    s.wire_se_.ensure(numbers * 9 + 16 * 8);
    s.wire_off_ = 0;
    s.wire_defer_ = true;
  }
  void finish() {
    if (done) return;
    done = true;
    s.wire_defer_ = false;
    s.wire_off_ = 0;
    CLR_CUDA(cudaStreamSynchronize(s.ctx.stream));  // the host arrays are the caller's again when the call returns
  }
  ~WireGroup() {
    if (!done) {
      s.wire_defer_ = false;
      s.wire_off_ = 0;
      cudaStreamSynchronize(s.ctx.stream);
    }
  }
};
void Solver::upload_point(const clrsdp_mp* xx, const clrsdp_mp* XX, const clrsdp_mp* yy, const clrsdp_mp* YY) {
  if (!structure_set) throw SolverError(CLRSDP_ERR_STATE, "upload_point before set_structure");
  if (xx->n != sumS || yy->n != n_y || XX->n != blkN || YY->n != blkN)
    throw SolverError(CLRSDP_ERR_BAD_ARG, "upload_point: sizes do not match the structure");
  CLR_CUDA(cudaSetDevice(ctx.device));
  WireGroup grp(*this, (size_t)(sumS + n_y + 2 * blkN));  // the four transfers are queued back to back, one synchronisation
  to_device(xx, 0, sumS, x, 0);
  to_device(yy, 0, n_y, y, 0);
  to_device(XX, 0, blkN, X, 0);
  to_device(YY, 0, blkN, Y, 0);
  grp.finish();
  have_point = true;
  prepared = false;
}
void Solver::download_point(clrsdp_mp_out* xx, clrsdp_mp_out* XX, clrsdp_mp_out* yy, clrsdp_mp_out* YY) {
  if (!have_point) throw SolverError(CLRSDP_ERR_STATE, "download_point: no point");
  if ((xx && xx->n < sumS) || (yy && yy->n < n_y) || (XX && XX->n < blkN) || (YY && YY->n < blkN))
    throw SolverError(CLRSDP_ERR_BAD_ARG, "download_point: an output array is smaller than the structure");
  CLR_CUDA(cudaSetDevice(ctx.device));
  WireGroup grp(*this, (size_t)(sumS + n_y + 2 * blkN));
  if (xx) to_host(x, 0, sumS, xx, 0);
  if (yy) to_host(y, 0, n_y, yy, 0);
  if (XX) to_host(X, 0, blkN, XX, 0);
  if (YY) to_host(Y, 0, blkN, YY, 0);
  grp.finish();
}

// ---- operand helpers ------------------------------------------------------------------------------
OperandDesc Solver::rows_of(BlockGroup& g, MpBuf& t) {
  OperandDesc a;
  a.src = t.t();
  a.d_off = g.offBlk.as<int64_t>();
  a.batch = (int)g.blocks.size();
  a.rows = g.nb, a.K = g.nb, a.rs = g.nb, a.ks = 1;
  return a;
}
OperandDesc Solver::cols_of(BlockGroup& g, MpBuf& t) {
  OperandDesc a = rows_of(g, t);
  a.rs = 1, a.ks = g.nb;
  return a;
}
OutDesc Solver::out_blk(BlockGroup& g, MpBuf& t, bool transposed) {
  OutDesc o;
  o.dst = t.t();
  o.d_off = g.offBlk.as<int64_t>();
  o.rs = transposed ? 1 : g.nb;
  o.cs = transposed ? g.nb : 1;
  return o;
}
static GemmPlan plan_of(int batch, int M, int N, const int* rowA = nullptr, const int* rowB = nullptr) {
  GemmPlan p;
  p.batch = batch, p.M = M, p.N = N, p.d_rowA = rowA, p.d_rowB = rowB;
  return p;
}


// ---- blocked Cholesky + triangular inverse (north star (3)) ------------------------------------------
static OperandDesc op_rows(const MatBatch& m, int r0, int c0, int rows, int K) {
  OperandDesc a;
  a.src = m.t, a.d_off = m.d_off, a.batch = m.batch;
  a.off0 = m.shift + (int64_t)r0 * m.stride() + c0;
  a.rs = m.stride(), a.ks = 1, a.rows = rows, a.K = K;
  return a;
}
static OperandDesc op_cols(const MatBatch& m, int r0, int c0, int rows, int K) {  // row operand i = column c0+i
  OperandDesc a = op_rows(m, r0, c0, rows, K);
  a.rs = 1, a.ks = m.stride();
  return a;
}
static OutDesc out_sub(const MatBatch& m, int r0, int c0) {
  OutDesc o;
  o.dst = m.t, o.d_off = m.d_off;
  o.off0 = m.shift + (int64_t)r0 * m.stride() + c0;
  o.rs = m.stride(), o.cs = 1;
  return o;
}

// One product of the blocked factorisation: C = epi(A * B^T-operand form). Products below ~1.5 M multiprecision
// multiply-adds go to the CUDA-core kernel (one launch, a few microseconds), the others through slice / tensor cores
// / carry.
static constexpr int64_t SMALL_GEMM_PMAC = 1500000;
void Solver::product(GemmEngine* ge, Slice& sa, Slice& sb, const OperandDesc& a, const OperandDesc& b, int M, int N,
                     const OutDesc& c, int epi, const mp::Tensor* extra) {
  const int64_t pmac = (int64_t)a.batch * M * N * a.K;
  if (a.d_ksign || a.d_kshift || b.d_kshift) throw SolverError(-1, "product: scaled / signed operands go on the B side");
  if (pmac <= SMALL_GEMM_PMAC) {
    SmallGemmArgs g;
    g.ksign = b.d_ksign, g.ksign_ld = b.ksign_ld;
    g.A = a.src, g.B = b.src, g.C = c.dst;
    if (extra) g.E = *extra;
    g.offA = a.d_off, g.offB = b.d_off, g.offC = c.d_off;
    g.a0 = a.off0, g.abs_ = a.bstride, g.ars = a.rs, g.aks = a.ks;
    g.b0 = b.off0, g.bbs = b.bstride, g.brs = b.rs, g.bks = b.ks;
    g.c0 = c.off0, g.cbs = c.bstride, g.crs = c.rs, g.ccs = c.cs;
    g.batch = a.batch, g.M = M, g.N = N, g.K = a.K, g.epi = epi;
    small_gemm(ctx, nl, g);
    return;
  }
  ge->slice(a, sa);
  if (&sa != &sb) ge->slice(b, sb);
  ge->multiply(sa, sb, plan_of(a.batch, M, N), c, epi, extra);
}

namespace {
// selects a helper stream for the launches of a block and puts the previous one back on every way out
struct StreamScope {
  Ctx& ctx;
  cudaStream_t saved;
  StreamScope(Ctx& c, cudaStream_t s) : ctx(c), saved(c.stream) { ctx.stream = s; }
  ~StreamScope() { ctx.stream = saved; }
};
}  // namespace
void Solver::chol_inverse(const MatBatch& A, const MatBatch& Uw, const MatBatch& Vw, const MatBatch& Linv,
                          int* d_stat, bool want_u, bool side, int* d_sig, int* d_keep_scale) {
  const int n = A.n, batch = A.batch;
  GemmEngine* gemm_loc = side ? this->gemm_side_.get() : this->gemm_.get();
  Slice& fs1_ = side ? this->fs1s_ : this->fs1_;
  Slice& fs2_ = side ? this->fs2s_ : this->fs2_;
  MpBuf& tscr = side ? this->tscr_side_ : this->tscr;
  const int PANEL = panel_width(nl);
  (void)Vw;
  // Equilibrate: A' = D^-1 A D^-1, D = diag(2^ceil(e_ii / 2)) (exact). The Schur complements of polynomial programmes
  // are graded over hundreds of bits (diagonal 2^26 .. 2^290 at degree 40); floating-point Cholesky does not care, but
  // the block-fixed-point GEMMs of the trailing updates keep a fixed window below each row maximum. On A' (unit-size
  // diagonal, every entry below 1) that window is the right one. L^-1 = L'^-1 D^-1 and U = U' D afterwards.
  // With d_keep_scale the exponents are left there and Linv stays L'^-1: a caller whose later products contract over
  // the index of A scales the other operand by D^-1 instead (slicer's d_kshift), which keeps that index balanced.
  DevBuf& scb = side ? this->equil_side_ : this->equil_;
  if (!d_keep_scale) scb.ensure(sizeof(int) * (size_t)batch * n);
  int* d_scale = d_keep_scale ? d_keep_scale : scb.as<int>();
  equil_exponents(ctx, nl, A, d_scale);
  mat_copy_scaled(ctx, nl, Uw, A, true, d_scale);  // work on the upper triangle of the equilibrated matrix
  // Signed mode (d_sig: [batch][n] flags, see panel_factor): A' = U^T Sigma U. With Y = Sigma U (what the row products
  // G11 * A12 below deliver, stored in the off-diagonal part of Uw) the trailing update is A22 -= Y12^T Sigma1 Y12 and
  // L[p, 0:k0] = (Sigma Y)[0:k0, p]^T in the inverse panels: Sigma only ever sits on a contraction index, where the
  // slicer (OperandDesc::d_ksign) or the CUDA-core product applies it exactly. Linv = L^-1 with L = U^T, unsigned.
  if (d_sig && want_u) throw SolverError(-1, "chol_inverse: the signed factorisation does not deliver U");
  if (n <= PANEL) {
    panel_factor(ctx, nl, Uw, Linv, want_u, d_stat, d_sig, n);
    if (!d_keep_scale) col_scale(ctx, nl, Linv, -1, d_scale);
    if (want_u) col_scale(ctx, nl, Uw, +1, d_scale);
    return;
  }
  mat_zero(ctx, nl, Linv);
  tscr.alloc(std::max<size_t>(tscr.n, (size_t)batch * PANEL * n), nl);
  // off-diagonal panel of L^-1 for the row block at k0:  Linv[p, 0:k0] = -Linv_pp * ( L[p, 0:k0] * Linv[0:k0, 0:k0] ).
  // It needs the factor up to and including panel k0 and the inverse panels above it - nothing of the trailing matrix -
  // so the chain of these panels runs on a helper stream BESIDE the rest of the factorisation (own GEMM workspace and
  // scratch), forked after panel_factor(k0) and joined at the end: the pivot chain no longer waits for it.
  InvHelper& H = invh_[side ? 1 : 0];
  const bool par = use_invh_ && n > 2 * PANEL;
  if (par) H.tscr.alloc(std::max<size_t>(H.tscr.n, (size_t)batch * PANEL * n), nl);
  auto inverse_panel = [&](int k0, GemmEngine* ge2, Slice& s1, Slice& s2, MpBuf& scr) {
    const int wk = std::min(PANEL, n - k0);
    // T[i][c] = sum_r U[r][k0+i] * Linv[r][c], stored transposed: Tt[c][i]
    OutDesc ot;
    ot.dst = scr.t(), ot.bstride = (int64_t)PANEL * n, ot.rs = 1, ot.cs = wk;
    OperandDesc lin = op_cols(Linv, 0, 0, k0, k0);
    if (d_sig) lin.d_ksign = d_sig, lin.ksign_ld = n;  // L[k0+i][r] = sigma_r Y[r][k0+i]
    product(ge2, s1, s2, op_cols(Uw, 0, k0, wk, k0), lin, wk, k0, ot, EPI_STORE, nullptr);
    OperandDesc tb;
    tb.src = scr.t(), tb.batch = batch, tb.bstride = (int64_t)PANEL * n, tb.rs = wk, tb.ks = 1, tb.rows = k0, tb.K = wk;
    product(ge2, s1, s2, op_rows(Linv, k0, k0, wk, wk), tb, wk, k0, out_sub(Linv, k0, 0), EPI_NEG, nullptr);
  };
  bool forked = false;
  // LOOKAHEAD. The pivot chain panel_factor(k0) -> U12 -> trailing update -> panel_factor(k0 + w) is the critical path of
  // the factorisation, but the next panel only needs the next ROW BLOCK of the trailing matrix. When the trailing matrix
  // is large compared with that row block, the update is split: the row block A[k1:k1+w1, k1:n] is updated on this
  // stream, the remainder A[k1+w1:n, k1+w1:n] on a helper stream (own GEMM workspace) beside the next panel. Every
  // (panel, row block) contribution is still applied exactly once; a row block is touched by the remainder updates of
  // the earlier panels (helper stream, in order) before the row-block update of its predecessor (this stream, which
  // waits for the helper), so the order of the subtractions per entry is fixed: results do not depend on timing.
  InvHelper& TH = trailh_[side ? 1 : 0];
  bool trail_pending = false, trail_used = false;
  auto wait_trail = [&]() {
    if (!trail_pending) return;
    CLR_CUDA(cudaStreamWaitEvent(ctx.stream, TH.ev_done, 0));
    trail_pending = false;
  };
  for (int k0 = 0; k0 < n; k0 += PANEL) {
    const int wk = std::min(PANEL, n - k0), n2 = n - k0 - wk;
    panel_factor(ctx, nl, Uw.sub(k0, k0, wk), Linv.sub(k0, k0, wk), want_u, d_stat, d_sig ? d_sig + k0 : nullptr, n);
    if (par && k0 >= PANEL) {
      CLR_CUDA(cudaEventRecord(H.ev_go, ctx.stream));
      CLR_CUDA(cudaStreamWaitEvent(H.stream, H.ev_go, 0));
      {
        StreamScope on_helper(ctx, H.stream);  // (restores the stream when an exception leaves the block, too)
        inverse_panel(k0, H.gemm.get(), H.s1, H.s2, H.tscr);
      }
      forked = true;
    }
    if (n2 > 0) {
      // U12 = L11^-1 A12 (overwrites A12). The CUDA-core kernel reads its operands while other threads store, so
      // there the product goes to scratch first; the tensor path slices its operands before it writes.
      if ((int64_t)batch * wk * n2 * wk <= SMALL_GEMM_PMAC) {
        OutDesc os;
        os.dst = tscr.t(), os.bstride = (int64_t)PANEL * n, os.rs = n2, os.cs = 1;
        product(gemm_loc, fs1_, fs2_, op_rows(Linv, k0, k0, wk, wk), op_cols(Uw, k0, k0 + wk, n2, wk), wk, n2, os, EPI_STORE,
                nullptr);
        SmallGemmArgs cp;
        cp.A = tscr.t(), cp.abs_ = (int64_t)PANEL * n, cp.ars = n2, cp.aks = 1;
        OutDesc od = out_sub(Uw, k0, k0 + wk);
        cp.C = od.dst, cp.offC = od.d_off, cp.c0 = od.off0, cp.crs = od.rs, cp.ccs = od.cs;
        cp.batch = batch, cp.M = wk, cp.N = n2;
        rect_copy(ctx, nl, cp);
      } else {
        product(gemm_loc, fs1_, fs2_, op_rows(Linv, k0, k0, wk, wk), op_cols(Uw, k0, k0 + wk, n2, wk), wk, n2,
                out_sub(Uw, k0, k0 + wk), EPI_STORE, nullptr);
      }
      // A22 -= U12^T U12   (signed: Y12^T Sigma1 Y12)
      mp::Tensor ut = Uw.t;
      const int k1 = k0 + wk, w1 = std::min(PANEL, n2), nr = n2 - w1;
      auto update = [&](GemmEngine* ge2, Slice& s1, Slice& s2, int c0, int M, int c1, int N) {
        // A[c0:c0+M, c1:c1+N] -= U12[:, c0..]^T Sigma U12[:, c1..]
        OperandDesc ua = op_cols(Uw, k0, c0, M, wk), ub = op_cols(Uw, k0, c1, N, wk);
        if (d_sig) {
          ub.d_ksign = d_sig + k0, ub.ksign_ld = n;
          product(ge2, s2, s1, ua, ub, M, N, out_sub(Uw, c0, c1), EPI_SUB_FROM, &ut);
        } else if (c0 == c1 && M == N) {
          product(ge2, s2, s2, ua, ua, M, N, out_sub(Uw, c0, c1), EPI_SUB_FROM, &ut);
        } else {
          product(ge2, s2, s1, ua, ub, M, N, out_sub(Uw, c0, c1), EPI_SUB_FROM, &ut);
        }
      };
      wait_trail();  // the remainder updates of the earlier panels have reached the rows written below
      // (measured: the split pays when the remainder is a tensor-core product - BASELINE config 5's Q, 93.1 -> 90.6 ms per
      // iteration - and costs 1 % when everything is a CUDA-core product that shares the SMs with the next panel - config 3's Q)
      const bool ahead = use_lookahead_ && nr > 0 && (double)nr * nr >= lookahead_ratio_ * (double)w1 * n2 &&
                         (int64_t)batch * nr * nr * wk > lookahead_min_pmac_;
      if (!ahead) {
        update(gemm_loc, fs1_, fs2_, k1, n2, k1, n2);
      } else {
        CLR_CUDA(cudaEventRecord(TH.ev_go, ctx.stream));  // U12 is complete
        CLR_CUDA(cudaStreamWaitEvent(TH.stream, TH.ev_go, 0));
        {
          StreamScope on_helper(ctx, TH.stream);
          update(TH.gemm.get(), TH.s1, TH.s2, k1 + w1, nr, k1 + w1, nr);
          CLR_CUDA(cudaEventRecord(TH.ev_done, TH.stream));
        }
        trail_pending = trail_used = true;
        update(gemm_loc, fs1_, fs2_, k1, w1, k1, n2);  // the next panel's row block
      }
    }
  }
  if (trail_used) {  // join (also when the last remainder has been waited for: the capture needs the stream back)
    CLR_CUDA(cudaEventRecord(TH.ev_done, TH.stream));
    CLR_CUDA(cudaStreamWaitEvent(ctx.stream, TH.ev_done, 0));
  }
  if (forked) {
    CLR_CUDA(cudaEventRecord(H.ev_done, H.stream));
    CLR_CUDA(cudaStreamWaitEvent(ctx.stream, H.ev_done, 0));
  }
  if (!par)
    for (int k0 = PANEL; k0 < n; k0 += PANEL) inverse_panel(k0, gemm_loc, fs1_, fs2_, tscr);
  if (!d_keep_scale) col_scale(ctx, nl, Linv, -1, d_scale);
  if (want_u) col_scale(ctx, nl, Uw, +1, d_scale);
}

// XY = X*Y per block (kept: both residual_R calls use it, MPMP.jl:1195,1209)
void Solver::block_products_XY() {
  for (auto& g : bgroups_) {
    ge()->slice(rows_of(g, X), g.sX);
    ge()->slice(rows_of(g, Y), g.sY);  // Y symmetric: its rows are its columns
    ge()->multiply(g.sX, g.sY, plan_of((int)g.blocks.size(), g.nb, g.nb), out_blk(g, XY));
  }
}

// Cholesky factors of X and Y for all blocks in one batch (spd_inv! :764-801 and cho! :1846): X = Lx Lx^T,
// Y = Ly Ly^T; only the inverse factors are kept. X and Y do not change until the update at the end of
// the iteration, so both factorisations run together and share the latency of the pivot chain.
void Solver::factor_XY() {
  int sbase = 0;
  for (auto& g : bgroups_) {
    chol_inverse(blkbatch2(g, XY2), blkbatch2(g, U2), blkbatch2(g, U2), blkbatch2(g, Linv2), d_status.as<int>() + sbase);
    sbase += 2 * (int)g.blocks.size();
  }
}
// X^-1 = Lx^-T Lx^-1 (spd_inv!, MPMP.jl:766)
void Solver::invert_X() {
  for (auto& g : bgroups_) {
    ge()->slice(cols_of(g, Linvx), g.sA);  // row operand i = column i of L^-1
    ge()->multiply(g.sA, g.sA, plan_of((int)g.blocks.size(), g.nb, g.nb), out_blk(g, Xinv));
  }
}

// Tt[(blk, r, s)][b][i] = sum_i' M[(r,i),(s,i')] V[i',b]   (first product of the pairings, MPMP.jl:1291-1296,
// and of trace_A, :1558)
void Solver::first_gemm_Tt(BlockGroup& g, MpBuf& M, Slice* cached) {
  int nblk = (int)g.blocks.size(), m = g.m;
  const Slice* As = cached;
  if (!(m == 1 && cached)) {
    OperandDesc a;
    a.src = M.t();
    a.d_off = g.g1_offA.as<int64_t>();
    a.batch = nblk * m * m;
    a.rows = g.delta, a.K = g.delta, a.rs = g.nb, a.ks = 1;
    ge()->slice(a, g.sB);
    As = &g.sB;
  }
  OutDesc o;
  o.dst = Tt.t();
  o.d_off = g.g1_offC.as<int64_t>();
  o.rs = 1, o.cs = g.delta;  // C[i][b] -> Tt[b*delta + i]
  ge()->multiply(*As, g.sVt, plan_of(nblk * m * m, g.delta, g.Nv, nullptr, g.g1_rowB.as<int>()), o);
}

// bilinear pairings  P[(r,a),(s,b)] = v_a^T M[r,s] v_b  (MPMP.jl:1274-1318)
void Solver::pairings(MpBuf& M, bool is_xinv, MpBuf& Pout) {
  for (auto& g : bgroups_) {
    int nblk = (int)g.blocks.size(), m = g.m;
    Slice* cached = nullptr;
    if (m == 1) {
      Slice& s = is_xinv ? g.sXinv : g.sY;
      if (is_xinv) ge()->slice(rows_of(g, M), s);  // sY was sliced for X*Y already
      cached = &s;
    } else if (is_xinv) {
      ge()->slice(rows_of(g, M), g.sXinv);  // still needed by the search directions
    }
    first_gemm_Tt(g, M, cached);
    OperandDesc bdesc;
    bdesc.src = Tt.t();
    bdesc.d_off = g.g2_offB.as<int64_t>();
    bdesc.batch = nblk * m;
    bdesc.rows = m * g.Nv, bdesc.K = g.delta, bdesc.rs = g.delta, bdesc.ks = 1;
    ge()->slice(bdesc, g.sB);
    OutDesc o;
    o.dst = Pout.t();
    o.d_off = g.g2_offC.as<int64_t>();
    o.rs = m * g.Nv, o.cs = 1;
    ge()->multiply(g.sVt, g.sB, plan_of(nblk * m, g.Nv, m * g.Nv, g.g2_rowA.as<int>(), nullptr), o);
  }
}

// out = sum_i a_i A_i + sign*E  (compute_weighted_A!, MPMP.jl:1621-1678, fused with :1115 / :1784)
void Solver::weighted_A(MpBuf& a, MpBuf& out, MpBuf& E, int sign) {
  scale_vectors(ctx, nl, st_, Vt.t(), H.t(), a.t(), VD.t(), vdN);
  for (auto& g : bgroups_) {
    int nblk = (int)g.blocks.size();
    OperandDesc ad;
    ad.src = VD.t();
    ad.d_off = g.wa_offA.as<int64_t>();
    ad.batch = nblk * g.np;
    ad.rows = g.delta, ad.K = g.Nv, ad.rs = g.Nv, ad.ks = 1;
    ge()->slice(ad, g.sB);
    OutDesc o;
    o.dst = QP.t();
    o.d_off = g.wa_offC.as<int64_t>();
    o.rs = g.delta, o.cs = 1;
    ge()->multiply(g.sB, g.sVr, plan_of(nblk * g.np, g.delta, g.delta, nullptr, g.wa_rowB.as<int>()), o);
  }
  assemble_weighted(ctx, nl, st_, QP.t(), out.t(), E.t(), sign, blkN, nullptr);
}

// compute_residuals (MPMP.jl:1107-1144)
void Solver::compute_residuals(bool from_pairings) {
  weighted_A(x, P, X, -1);  // P = sum x_i A_i - X
  if (have_C) ew_lincomb(ctx, nl, P.t(), 0, P.t(), 0, 1, Cmat.t(), 0, -1, blkN);  // - C (:1116-1118)
  // d = c - B y - Tr(A_* Y)
  GemvArgs g1;
  g1.A = Bmat.t(), g1.x = y.t(), g1.out = tmpx.t();
  g1.rs = n_y, g1.ks = 1, g1.rows = sumS, g1.K = n_y;
  gemv(ctx, nl, g1, work.t());
  if (from_pairings) {
    trace_from_pairings(ctx, nl, st_, Py.t(), H.t(), trx.t());
  } else {  // general method on Y (loop initialisation, :727)
    for (auto& g : bgroups_) first_gemm_Tt(g, Y, nullptr);
    trace_from_ZV(ctx, nl, st_, Vt.t(), Tt.t(), H.t(), trx.t());
  }
  ew_lincomb(ctx, nl, d.t(), 0, c.t(), 0, 1, tmpx.t(), 0, -1, sumS);
  ew_lincomb(ctx, nl, d.t(), 0, d.t(), 0, 1, trx.t(), 0, -1, sumS);
  // p = b - B^T x
  GemvArgs g2;
  g2.A = Bmat.t(), g2.x = x.t(), g2.out = tmpy.t();
  g2.rs = 1, g2.ks = n_y, g2.rows = n_y, g2.K = sumS;
  gemv(ctx, nl, g2, work.t());
  allreduce(tmpy, 0, n_y, COMB_SUM);  // sum over the clusters of all ranks (:1137-1139)
  ew_lincomb(ctx, nl, p.t(), 0, b.t(), 0, 1, tmpy.t(), 0, -1, n_y);
  // errors (max-abs) of this P, p, d: used by the log row now and, stale, by terminate() after the update
  reduce_maxabs(ctx, nl, P.t(), 0, blkN, scal.t(), SL_PERR_P, work.t());
  reduce_maxabs(ctx, nl, p.t(), 0, n_y, scal.t(), SL_PERR_p, work.t());
  reduce_maxabs(ctx, nl, d.t(), 0, sumS, scal.t(), SL_DERR, work.t());
  allreduce(scal, SL_PERR_P, 1, COMB_MAX);
  allreduce(scal, SL_DERR, 1, COMB_MAX);
}

// compute_T_decomposition (MPMP.jl:1417-1514) with Cholesky instead of LU (S and Q are SPD, :1430-1432)
void Solver::decomposition() {
  mark(CLRSDP_T_SCHUR);
  pairings(Xinv, true, Px);  // (the pairings with Y were formed on the side stream, see iteration_body)
  schur_assemble(ctx, nl, st_, Px.t(), Py.t(), H.t(), S.t(), sN);
  mark(-1 - CLRSDP_T_SCHUR);
  mark(CLRSDP_T_CHOL_S);
  int sbase = 2 * (int)blocks_.size();
  for (auto& g : cgroups_) {
    MatBatch A{S.t(), g.offS.as<int64_t>(), (int)g.clusters.size(), g.dimS};
    MatBatch U{Us.t(), g.offS.as<int64_t>(), (int)g.clusters.size(), g.dimS};
    MatBatch V{Vs.t(), g.offS.as<int64_t>(), (int)g.clusters.size(), g.dimS};
    MatBatch Li{Linvs.t(), g.offS.as<int64_t>(), (int)g.clusters.size(), g.dimS};
    // Linvs holds L'^-1 of the equilibrated S' = D^-1 S D^-1; D (exponents in g.equil, and in xscale indexed like x)
    // goes onto the other operand wherever the constraint index is contracted: L^-1 B = L'^-1 (D^-1 B) below,
    // L^-1 rhs = L'^-1 (D^-1 rhs) and L^-T u = D^-1 (L'^-T u) in search_direction().
    // The factorisation is signed, S'_j = U^T Sigma_j U (see panel_factor): Sigma_j (flags in g.sig, and in xsign indexed
    // like x) enters as S_j^-1 = D^-1 L'^-T Sigma L'^-1 D^-1, i.e. Q = W^T Sigma W and Sigma on t_j, W_j dy below.
    g.equil.ensure(sizeof(int) * g.clusters.size() * (size_t)g.dimS);
    g.sig.ensure(sizeof(int) * g.clusters.size() * (size_t)g.dimS);
    chol_inverse(A, U, V, Li, d_status.as<int>() + sbase, false, false, g.sig.as<int>(), g.equil.as<int>());
    scatter_scale(ctx, g.equil.as<int>(), (int)g.clusters.size(), g.dimS, g.offW.as<int64_t>(), xscale.as<int>());
    scatter_scale(ctx, g.sig.as<int>(), (int)g.clusters.size(), g.dimS, g.offW.as<int64_t>(), xsign.as<int>());
    sbase += (int)g.clusters.size();
  }
  mark(-1 - CLRSDP_T_CHOL_S);
  mark(CLRSDP_T_CINVB);
  // Wt[a][(j,i)] = (L_j^-1 B_j)[i][a]
  for (auto& g : cgroups_) {
    OperandDesc bt;  // rows a (n_y) of (D^-1 B_j)^T
    bt.src = BmatT.t();
    bt.d_off = g.offBt.as<int64_t>();
    bt.batch = (int)g.clusters.size();
    bt.rows = n_y, bt.K = g.dimS, bt.rs = sumS, bt.ks = 1;
    bt.d_kshift = g.equil.as<int>();
    ge()->slice(bt, g.sBt);
    OperandDesc bd;
    bd.src = Linvs.t();
    bd.d_off = g.offS.as<int64_t>();
    bd.batch = (int)g.clusters.size();
    bd.rows = g.dimS, bd.K = g.dimS, bd.rs = g.dimS, bd.ks = 1;
    ge()->slice(bd, g.sLinv);
    OutDesc o;
    o.dst = Wt.t();
    o.d_off = g.offW.as<int64_t>();
    o.rs = sumS, o.cs = 1;
    ge()->multiply(g.sBt, g.sLinv, plan_of((int)g.clusters.size(), n_y, g.dimS), o);
  }
  mark(-1 - CLRSDP_T_CINVB);
  mark(CLRSDP_T_Q);
  {  // Q = sum_j W_j^T Sigma_j W_j = Wt Sigma Wt^T  (the reference forms B^T U^-1 * L^-1 B, :1467-1495)
    Slice& sW = sW_;
    OperandDesc a;
    a.src = Wt.t();
    a.batch = 1, a.rows = n_y, a.K = sumS, a.rs = sumS, a.ks = 1;
    ge()->slice(a, sW);
    a.d_ksign = xsign.as<int>(), a.ksign_ld = sumS;
    ge()->slice(a, sWs_);
    OutDesc o;
    o.dst = Q.t();
    o.rs = n_y, o.cs = 1;
    static int float_sites_q = getenv("CLRSDP_FLOAT_SITES") ? atoi(getenv("CLRSDP_FLOAT_SITES")) : 0;  // measuring aid
    if (float_sites_q & 2) {  // Q by plain multiprecision multiply-adds (no block fixed point)
      SmallGemmArgs sg;
      sg.A = Wt.t(), sg.B = Wt.t(), sg.C = Q.t();
      sg.ars = sumS, sg.aks = 1, sg.brs = sumS, sg.bks = 1, sg.crs = n_y, sg.ccs = 1;
      sg.batch = 1, sg.M = n_y, sg.N = n_y, sg.K = sumS;
      sg.ksign = xsign.as<int>(), sg.ksign_ld = sumS;
      small_gemm(ctx, nl, sg);
    } else
    ge()->multiply(sW, sWs_, plan_of(1, n_y, n_y), o, EPI_STORE, nullptr, true);  // symmetric: upper tiles only
    allreduce(Q, 0, (int64_t)n_y * n_y, COMB_SUM);  // the cross-cluster reduction (sum(Q), :1494)
  }
  mark(-1 - CLRSDP_T_Q);
  // The factorisation of Q is one n_y x n_y matrix: a latency chain of n_y/32 panels that occupies a handful of SMs.
  // It runs on a side stream (with its own GEMM workspace) beside the residuals and the first half of the
  // predictor, which do not need it; search_direction() joins before the Q^-1 solve.
  fork_side();
  mark(CLRSDP_T_CHOL_Q);
  {
    MatBatch A{Q.t(), d_qoff.as<int64_t>(), 1, n_y}, U{Uq.t(), d_qoff.as<int64_t>(), 1, n_y};
    MatBatch V{Vq.t(), d_qoff.as<int64_t>(), 1, n_y}, Li{Linvq.t(), d_qoff.as<int64_t>(), 1, n_y};
    chol_inverse(A, U, V, Li, d_status.as<int>() + 2 * (int)blocks_.size() + J, false, true, qsign.as<int>());
  }
  mark(-1 - CLRSDP_T_CHOL_Q);
  end_side();
}

// compute_search_direction (MPMP.jl:1682-1824)
void Solver::search_direction() {
  mark(CLRSDP_T_Z);
  for (auto& g : bgroups_) {  // Z = sym(X^-1 (P Y - R))
    int nblk = (int)g.blocks.size();
    ge()->slice(rows_of(g, P), g.sA);
    mp::Tensor Rt = R.t();
    ge()->multiply(g.sA, g.sY, plan_of(nblk, g.nb, g.nb), out_blk(g, T1), EPI_MINUS_SUB, &Rt);
    ge()->slice(cols_of(g, T1), g.sB);
    ge()->multiply(g.sXinv, g.sB, plan_of(nblk, g.nb, g.nb), out_blk(g, T2));
    ew_symmetrize(ctx, nl, blkbatch(g, Z), T2.t());
  }
  mark(-1 - CLRSDP_T_Z);
  mark(CLRSDP_T_RHS_X);
  for (auto& g : bgroups_) first_gemm_Tt(g, Z, nullptr);
  trace_from_ZV(ctx, nl, st_, Vt.t(), Tt.t(), H.t(), trx.t());
  ew_lincomb(ctx, nl, rhs.t(), 0, d.t(), 0, -1, trx.t(), 0, -1, sumS);  // rhs_x = -d - Tr(A_* Z)
  mark(-1 - CLRSDP_T_RHS_X);
  mark(CLRSDP_T_SYS);
  {
    // S_j^-1 = D^-1 L'^-T Sigma L'^-1 D^-1:  t_j = Sigma L'_j^-1 (D_j^-1 rhs_j)
    // The diagonal factors (Sigma: sign flips, D^-1: exponent shifts) and the vector additions of this chain are header-
    // word operations fused into the products (GemvArgs x_flip / x_scale / o_flip / o_scale / e): every launch on this
    // path is latency, and there were 28 of them per direction beside the 15 products.
    const int* xs = xscale.as<int>();
    const int* xg = xsign.as<int>();
    const int* qg = qsign.as<int>();
    GemvArgs a;
    a.A = Linvs.t(), a.x = rhs.t(), a.out = tvec.t();
    a.rows = sumS, a.K = 0;
    a.d_row_item = d_row_item.as<int>(), a.d_aoff = d_linv_off.as<int64_t>(), a.d_xoff = d_x_off64.as<int64_t>();
    a.d_row0 = d_row0.as<int>(), a.d_K = d_itemK.as<int>();
    a.item_trans = 0;  // A_item[r][k], leading dimension = K_item
    for (auto& cg : cgroups_) a.K_hint = std::max(a.K_hint, cg.dimS);
    GemvArgs a0 = a;   // (the plain per-cluster product, without fused operations)
    a.x_scale = xs, a.x_scale_sign = -1;  // D^-1 rhs
    a.o_flip = xg;                        // t <- Sigma t
    gemv(ctx, nl, a, work.t());
    // tmpy = sum_j W_j^T Sigma_j t_j = Wt t
    GemvArgs w;
    w.A = Wt.t(), w.x = tvec.t(), w.out = tmpy.t();
    w.rs = sumS, w.ks = 1, w.rows = n_y, w.K = sumS;
    gemv(ctx, nl, w, work.t());
    allreduce(tmpy, 0, n_y, COMB_SUM);  // sum(temp_y), :1761
    ew_lincomb(ctx, nl, dyr.t(), 0, p.t(), 0, 1, tmpy.t(), 0, -1, n_y);  // p - sum_j B^T U^-1 t_j (:1761)
    // dy = Q^-1 dyr = Lq^-T Sigma_q (Lq^-1 dyr)
    join_side();  // the factor of Q comes from the side stream
    GemvArgs q1;
    q1.A = Linvq.t(), q1.x = dyr.t(), q1.out = zvec.t();
    q1.rs = n_y, q1.ks = 1, q1.rows = n_y, q1.K = n_y;
    q1.o_flip = qg;
    gemv(ctx, nl, q1, work.t());
    GemvArgs q2 = q1;
    q2.o_flip = nullptr;
    q2.x = zvec.t(), q2.out = dy.t();
    q2.rs = 1, q2.ks = n_y;
    gemv(ctx, nl, q2, work.t());
    // u = t + Sigma W dy ; dx_j = D^-1 L'_j^-T u_j
    GemvArgs wd;
    wd.A = Wt.t(), wd.x = dy.t(), wd.out = tmpx.t();
    wd.rs = 1, wd.ks = sumS, wd.rows = sumS, wd.K = n_y;
    GemvArgs wd0 = wd;
    wd.o_flip = xg, wd.e = tvec.t(), wd.e_mode = 1;
    gemv(ctx, nl, wd, work.t());
    GemvArgs bt = a0;
    bt.x = tmpx.t(), bt.out = dx.t();
    bt.item_trans = 1;  // A_item[k][r]
    bt.o_scale = xs, bt.o_scale_sign = -1;  // L^-T = D^-1 L'^-T
    gemv(ctx, nl, bt, work.t());
    static int refine_mode = getenv("CLRSDP_REFINE") ? atoi(getenv("CLRSDP_REFINE")) : 4;  // measuring aid, see the order below
    auto refine_second = [&]() {
    // One step of iterative refinement on the SECOND block equation, B^T dx = p. The reference obtains dy from
    // Q = (B^T U^-1)(L^-1 B) and dx from the same two factors (MPMP.jl:1457-1495, :1751-1773), so B^T dx = p holds to
    // rounding. Here Q = W^T Sigma W comes from the block fixed-point GEMM while dx is formed by floating gemvs with W
    // and L'^-1: the two differ by dQ ~ 2^-(p+16) rowmax_a rowmax_b, which is large in absolute terms when S_j is
    // singular to working precision (one huge row of L'^-1 per cluster on BASELINE config 3: K = 128 samples of
    // polynomials of degree <= 126), and the primal residual p then stagnates around dQ dy instead of contracting by
    // 1 - alpha per iteration (measured: 4e-42 and growing where the LU oracle is at 3e-69). So: r = p - B^T dx with the
    // resident B (floating), dy += Q^-1 r, dx += S^-1 B Q^-1 r = D^-1 L'^-T Sigma W (Q^-1 r); the first block equation is
    // untouched (S ddx - B ddy = 0) and the defect of the second becomes dQ ddy, quadratically small.
    GemvArgs bx;
    bx.A = Bmat.t(), bx.x = dx.t(), bx.out = tmpy.t();
    bx.rs = 1, bx.ks = n_y, bx.rows = n_y, bx.K = sumS;
    gemv(ctx, nl, bx, work.t());
    allreduce(tmpy, 0, n_y, COMB_SUM);                                               // B^T dx over all clusters
    ew_lincomb(ctx, nl, dyr.t(), 0, p.t(), 0, 1, tmpy.t(), 0, -1, n_y);              // r = p - B^T dx
    GemvArgs r1 = q1;
    r1.x = dyr.t(), r1.out = zvec.t();
    gemv(ctx, nl, r1, work.t());                                                     // Sigma_q Lq^-1 r
    GemvArgs r2 = q2;
    r2.x = zvec.t(), r2.out = tmpy.t();
    gemv(ctx, nl, r2, work.t());                                                     // ddy = Q^-1 r
    ew_lincomb(ctx, nl, dy.t(), 0, dy.t(), 0, 1, tmpy.t(), 0, 1, n_y);
    GemvArgs wr = wd0;
    wr.x = tmpy.t(), wr.out = tmpx.t();
    wr.o_flip = xg;
    gemv(ctx, nl, wr, work.t());                                                     // Sigma W ddy
    GemvArgs mr = bt;
    mr.x = tmpx.t(), mr.out = dx.t();
    mr.e = dx.t(), mr.e_mode = 1;
    gemv(ctx, nl, mr, work.t());                                                     // dx += D^-1 L'^-T Sigma W ddy = S^-1 B ddy
    };
    auto refine_first = [&]() {
    // ... and one step on the FIRST block equation, S dx - B dy = rhs_x, with the final dy: applying S^-1 through the
    // explicit inverse factors has the forward error eps |M|^T |M| |v| instead of the eps cond(S) |dx| of triangular
    // solves; near the optimum (cond(S) large) that shows as a dual residual d that stops contracting (measured: 1e-52
    // where the LU oracle is at 1e-73). r = rhs_x + B dy - S dx is small, so the same factors applied to it are accurate.
    GemvArgs bd;
    bd.A = Bmat.t(), bd.x = dy.t(), bd.out = tmpx.t();
    bd.rs = n_y, bd.ks = 1, bd.rows = sumS, bd.K = n_y;
    bd.e = rhs.t(), bd.e_mode = 1;
    gemv(ctx, nl, bd, work.t());                                                     // rhs_x + B dy
    GemvArgs sd = a0;
    sd.A = S.t(), sd.x = dx.t(), sd.out = tmpx.t();
    sd.e = tmpx.t(), sd.e_mode = 2;
    gemv(ctx, nl, sd, work.t());                                                     // r = rhs_x + B dy - S_j dx_j
    GemvArgs m1 = a;                                                                 // Sigma L'^-1 D^-1 r
    m1.x = tmpx.t(), m1.out = trx.t();
    gemv(ctx, nl, m1, work.t());
    GemvArgs m2 = bt;
    m2.x = trx.t(), m2.out = dx.t();
    m2.e = dx.t(), m2.e_mode = 1;
    gemv(ctx, nl, m2, work.t());                                                     // dx += D^-1 L'^-T ... = S^-1 r
    };
    // order: 1 = second only, 2 = first only, 3 = second then first, 4 = first then second
    if (refine_mode == 1 || refine_mode == 3) refine_second();
    if (refine_mode == 2 || refine_mode == 3 || refine_mode == 4) refine_first();
    if (refine_mode == 4) refine_second();
  }
  mark(-1 - CLRSDP_T_SYS);
  mark(CLRSDP_T_DX);
  weighted_A(dx, dX, P, +1);  // dX = sum dx_i A_i + P
  mark(-1 - CLRSDP_T_DX);
  mark(CLRSDP_T_DY);
  for (auto& g : bgroups_) {  // dY = sym(X^-1 (R - dX Y))
    int nblk = (int)g.blocks.size();
    ge()->slice(rows_of(g, dX), g.sA);
    mp::Tensor Rt = R.t();
    ge()->multiply(g.sA, g.sY, plan_of(nblk, g.nb, g.nb), out_blk(g, T1), EPI_SUB_FROM, &Rt);
    ge()->slice(cols_of(g, T1), g.sB);
    ge()->multiply(g.sXinv, g.sB, plan_of(nblk, g.nb, g.nb), out_blk(g, T2));
    ew_symmetrize(ctx, nl, blkbatch(g, dY), T2.t());
  }
  mark(-1 - CLRSDP_T_DY);
}

// compute_step_length (MPMP.jl:1829-1898) for X and Y together: lambda_min( L^-1 dM L^-T ) per block
void Solver::step_lengths() {
  for (auto& g : bgroups_) {
    int nb2 = 2 * (int)g.blocks.size();
    OperandDesc a;
    a.src = dXY2.t(), a.d_off = g.offBlk2.as<int64_t>(), a.batch = nb2, a.rows = g.nb, a.K = g.nb, a.rs = g.nb, a.ks = 1;
    ge()->slice(a, g.sA);
    a.src = Linv2.t();
    ge()->slice(a, g.sLinv);
    // T[i][j] = sum_k dM[i][k] Linv[j][k], stored transposed
    OutDesc o;
    o.dst = T1d.t(), o.d_off = g.offBlk2.as<int64_t>(), o.rs = 1, o.cs = g.nb;
    ge()->multiply(g.sA, g.sLinv, plan_of(nb2, g.nb, g.nb), o);
    a.src = T1d.t();
    ge()->slice(a, g.sB);
    o.dst = T2d.t(), o.rs = g.nb, o.cs = 1;
    ge()->multiply(g.sLinv, g.sB, plan_of(nb2, g.nb, g.nb), o);
    ew_symmetrize(ctx, nl, blkbatch2(g, W2), T2d.t());
    lambda_min(ctx, nl, blkbatch2(g, W2), lam.t(), g.lamIdx2.as<int>(), d_lamflag.as<int>());
  }
  reduce_min(ctx, nl, lam.t(), 0, (int64_t)blocks_.size(), scal.t(), SL_LAM_X, work.t());
  reduce_min(ctx, nl, lam.t(), (int64_t)blocks_.size(), (int64_t)blocks_.size(), scal.t(), SL_LAM_Y, work.t());
  allreduce(scal, SL_LAM_X, 2, COMB_MIN);  // global minimum over all ranks (:1890-1891); LAM_X, LAM_Y adjacent
}

// ---- side stream: fork after the current point of the main stream, join before the first consumer ------
void Solver::fork_side() {
  if (!use_side_) return;
  CLR_CUDA(cudaEventRecord(ev_fork_, ctx.stream));
  CLR_CUDA(cudaStreamWaitEvent(side_stream_, ev_fork_, 0));
  main_stream_ = ctx.stream;
  ctx.stream = side_stream_;
  on_side_ = true;
}
void Solver::end_side() {
  if (!use_side_) return;
  CLR_CUDA(cudaEventRecord(ev_join_, ctx.stream));
  ctx.stream = main_stream_;
  on_side_ = false;
  join_pending_ = true;
}
void Solver::join_side() {
  if (!join_pending_) return;
  CLR_CUDA(cudaStreamWaitEvent(ctx.stream, ev_join_, 0));
  join_pending_ = false;
}

// ---- timing marks: bucket >= 0 begins a bucket, -1-bucket ends it -----------------------------------
// Directly launched iterations record ordinary events. Inside the captured graph the marks become EVENT-RECORD NODES
// (cudaEventRecordExternal) when phase timing is switched on (clrsdp_int_params.phase_timing): every replay then
// re-records the same events, so the reference's 17 buckets (MPMP.jl:889-898, table :972-1012) are filled on the
// default path too; switched off, the graph carries no event nodes and only the whole iteration is timed.
void Solver::mark(int code) {
  const int bkt = code >= 0 ? code : -1 - code, sign = code >= 0 ? +1 : -1;
  if (capturing_) {
    if (!phase_timing_) return;
    cudaEvent_t e;
    if (graph_marks_.size() < gev_.size()) {
      e = gev_[graph_marks_.size()];
    } else {
      CLR_CUDA(cudaEventCreate(&e));
      gev_.push_back(e);
    }
    CLR_CUDA(cudaEventRecordWithFlags(e, ctx.stream, cudaEventRecordExternal));
    graph_marks_.push_back(Mark{bkt, sign, e});
    return;
  }
  cudaEvent_t e;
  if (n_direct_marks_ < ev_.size()) {
    e = ev_[n_direct_marks_];
  } else {
    CLR_CUDA(cudaEventCreate(&e));
    ev_.push_back(e);
  }
  n_direct_marks_++;
  CLR_CUDA(cudaEventRecord(e, ctx.stream));
  ev_marks_.push_back(Mark{bkt, sign, e});
}

int Solver::check_status_local() {
  CLR_CUDA(cudaMemcpyAsync(h_status.data(), d_status.p, sizeof(int) * n_status, cudaMemcpyDeviceToHost, ctx.stream));
  CLR_CUDA(cudaMemcpyAsync(h_scal, d_scal_out.p, sizeof(h_scal), cudaMemcpyDeviceToHost, ctx.stream));
  CLR_CUDA(cudaMemcpyAsync(h_flags, d_flags.p, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  int nb = (int)blocks_.size();
  {  // slots [0, 2 nb): per group, first the X blocks then the Y blocks
    int base = 0, bad_y = 0;
    for (auto& g : bgroups_) {
      int n1 = (int)g.blocks.size();
      for (int i = 0; i < n1; i++)
        if (h_status[base + i]) return CLRSDP_ERR_NOT_PD_X;
      for (int i = n1; i < 2 * n1; i++)
        if (h_status[base + i]) bad_y = 1;
      base += 2 * n1;
    }
    if (bad_y) return CLRSDP_ERR_NOT_PD_Y;
  }
  for (int i = 2 * nb; i < 2 * nb + J; i++)
    if (h_status[i]) return CLRSDP_ERR_SINGULAR_S;
  if (h_status[2 * nb + J]) return CLRSDP_ERR_SINGULAR_Q;
  return 0;
}

// every rank must see a failure of any rank, otherwise the next collective would dead-lock
int Solver::check_status() {
  int st = check_status_local();
  if (!comm_.active()) return st;
  int code = -st;  // 0 or 10..13
  d_status_any.ensure(sizeof(int));
  CLR_CUDA(cudaMemcpyAsync(d_status_any.p, &code, sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
  CLR_NCCL(NcclApi::get().AllReduce(d_status_any.p, d_status_any.p, 1, ncclInt32, ncclMax, comm_.comm, ctx.stream));
  CLR_CUDA(cudaMemcpyAsync(&code, d_status_any.p, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  return -code;
}

// <C,Y> over all blocks (dot(C, Y), MPMP.jl:1033); stays zero while C = 0
void Solver::dot_CY() {
  if (!have_C) return;
  reduce_dot(ctx, nl, Cmat.t(), 0, Y.t(), 0, blkN, scal.t(), SL_CY, work.t());
  allreduce(scal, SL_CY, 1, COMB_SUM);
}

// loop initialisation (MPMP.jl:716-736): the device work, no host synchronisation, no host-side data dependence
void Solver::prepare_body() {
  CLR_CUDA(cudaMemsetAsync(d_status.p, 0, sizeof(int) * n_status, ctx.stream));
  CLR_CUDA(cudaMemsetAsync(d_flags.p, 0, 2 * sizeof(int), ctx.stream));
  ew_zero(ctx, nl, scal.t(), SL_ALPHA_P, 1);
  ew_zero(ctx, nl, scal.t(), SL_ALPHA_D, 1);
  reduce_dot(ctx, nl, X.t(), 0, Y.t(), 0, blkN, scal.t(), SL_DOT_XY, work.t());
  allreduce(scal, SL_DOT_XY, 1, COMB_SUM);
  scalar_program(ctx, nl, SP_MU, scal.t(), d_flags.as<int>(), nullptr);
  reduce_dot(ctx, nl, c.t(), 0, x.t(), 0, sumS, scal.t(), SL_CX, work.t());
  allreduce(scal, SL_CX, 1, COMB_SUM);
  reduce_dot(ctx, nl, b.t(), 0, y.t(), 0, n_y, scal.t(), SL_BY, work.t());
  dot_CY();
  compute_residuals(false);
  // the initial duality gap is computed WITHOUT b0 (MPMP.jl:725 -> :1067-1074)
  scalar_program(ctx, nl, SP_OBJECTIVES_INIT, scal.t(), d_flags.as<int>(), nullptr);
  scalar_program(ctx, nl, SP_ERRORS, scal.t(), d_flags.as<int>(), d_scal_out.as<double>());
}

int Solver::prepare(clrsdp_iter_info* info) {
  if (!have_point) return CLRSDP_ERR_STATE;
  for (int u : uploaded_)
    if (!u) return CLRSDP_ERR_STATE;
  CLR_CUDA(cudaSetDevice(ctx.device));
  double t0 = now_s();
  if (main_stream_) ctx.stream = main_stream_;
  on_side_ = false;
  join_pending_ = false;
  iter = 1;
  if (!ntot_uploaded_) {  // (the global size of X changes only with the structure or the communicator)
    upload_ntot();
    ntot_uploaded_ = true;
  }
  // (Replaying this body from a CUDA graph of its own - for front ends that re-upload the iterate and call prepare before
  // every iteration, bench.py's end-to-end path - was measured: 0.60 -> 0.55 ms. The call is device work plus one
  // synchronisation, not launch overhead; a second graph to maintain is not worth 50 us.)
  prepare_body();
  int st = check_status();
  prepared = (st == 0);
  if (info) {
    memset(info, 0, sizeof(*info));
    info->iter = iter;
    info->status = st;
    info->terminate = h_flags[1];
    info->pd_feasible = h_flags[0];
    info->mu = h_scal[SL_MU];
    info->p_obj = info->p_obj_new = h_scal[SL_P_OBJ];
    info->d_obj = info->d_obj_new = h_scal[SL_D_OBJ];
    info->gap = info->gap_new = h_scal[SL_GAP];
    info->P_err = h_scal[SL_PERR_P];
    info->p_err = h_scal[SL_PERR_p];
    info->d_err = h_scal[SL_DERR];
    info->primal_err_new = h_scal[SL_PRIMAL_ERR];
    info->dual_err_new = h_scal[SL_DUAL_ERR];
    info->seconds = now_s() - t0;
  }
  return st;
}

// one pass of the while-body (MPMP.jl:754-953)
void Solver::drop_graph() {
  if (gexec_) cudaGraphExecDestroy(gexec_);
  gexec_ = nullptr;
}

// one pass of MPMP.jl:754-953 on the stream: no host synchronisation, no host-side data dependence
void Solver::iteration_body() {
  // step 3
  reduce_dot(ctx, nl, X.t(), 0, Y.t(), 0, blkN, scal.t(), SL_DOT_XY, work.t());
  allreduce(scal, SL_DOT_XY, 1, COMB_SUM);
  scalar_program(ctx, nl, SP_MU, scal.t(), d_flags.as<int>(), nullptr);
  // step 4: R = mu_p I - X Y, and the pairings with Y: neither needs X^-1, so they run on the side stream beside the
  // factorisations of X and Y (a latency chain) and the product X^-1 = L^-T L^-1
  fork_side();
  mark(CLRSDP_T_R);
  block_products_XY();
  for (auto& g : bgroups_) ew_residual_R(ctx, nl, blkbatch(g, R), scal.t(), SL_MU_P, XY.t(), nullptr);
  mark(-1 - CLRSDP_T_R);
  pairings(Y, false, Py);
  end_side();
  mark(CLRSDP_T_XINV);
  factor_XY();
  invert_X();
  mark(-1 - CLRSDP_T_XINV);
  join_side();
  mark(CLRSDP_T_DECOMP);
  decomposition();
  mark(-1 - CLRSDP_T_DECOMP);
  mark(CLRSDP_T_RES);
  compute_residuals(true);
  mark(-1 - CLRSDP_T_RES);
  mark(CLRSDP_T_PREDICTOR);
  search_direction();
  mark(-1 - CLRSDP_T_PREDICTOR);
  ew_lincomb(ctx, nl, dX_pred.t(), 0, dX.t(), 0, 1, dX.t(), 0, 0, blkN);
  ew_lincomb(ctx, nl, dY_pred.t(), 0, dY.t(), 0, 1, dY.t(), 0, 0, blkN);
  ew_lincomb(ctx, nl, dx_pred.t(), 0, dx.t(), 0, 1, dx.t(), 0, 0, sumS);
  ew_lincomb(ctx, nl, dy_pred.t(), 0, dy.t(), 0, 1, dy.t(), 0, 0, n_y);
  // step 5
  reduce_dot_sum(ctx, nl, X.t(), dX.t(), Y.t(), dY.t(), blkN, scal.t(), SL_DOT_SUM, work.t());
  allreduce(scal, SL_DOT_SUM, 1, COMB_SUM);
  scalar_program(ctx, nl, SP_BETA, scal.t(), d_flags.as<int>(), nullptr);
  // step 6: R = mu_c I - X Y - dX dY
  mark(CLRSDP_T_R);
  for (auto& g : bgroups_) {
    ge()->slice(rows_of(g, dX), g.sA);
    ge()->slice(rows_of(g, dY), g.sB);  // dY symmetric
    ge()->multiply(g.sA, g.sB, plan_of((int)g.blocks.size(), g.nb, g.nb), out_blk(g, T1));
    mp::Tensor t1 = T1.t();
    ew_residual_R(ctx, nl, blkbatch(g, R), scal.t(), SL_MU_C, XY.t(), &t1);
  }
  mark(-1 - CLRSDP_T_R);
  mark(CLRSDP_T_CORRECTOR);
  search_direction();
  mark(-1 - CLRSDP_T_CORRECTOR);
  // step 7
  mark(CLRSDP_T_ALPHA);
  step_lengths();
  scalar_program(ctx, nl, SP_ALPHA, scal.t(), d_flags.as<int>(), nullptr, d_status.as<int>(), n_status);
  mark(-1 - CLRSDP_T_ALPHA);
  // step 8
  ew_axpy(ctx, nl, x.t(), 0, dx.t(), 0, scal.t(), SL_ALPHA_P, sumS);
  ew_axpy(ctx, nl, y.t(), 0, dy.t(), 0, scal.t(), SL_ALPHA_D, n_y);
  ew_axpy(ctx, nl, X.t(), 0, dX.t(), 0, scal.t(), SL_ALPHA_P, blkN);
  ew_axpy(ctx, nl, Y.t(), 0, dY.t(), 0, scal.t(), SL_ALPHA_D, blkN);
  // new objectives; errors are the ones computed from P,p,d BEFORE the update (:940-944)
  reduce_dot(ctx, nl, c.t(), 0, x.t(), 0, sumS, scal.t(), SL_CX, work.t());
  allreduce(scal, SL_CX, 1, COMB_SUM);
  reduce_dot(ctx, nl, b.t(), 0, y.t(), 0, n_y, scal.t(), SL_BY, work.t());
  dot_CY();
  scalar_program(ctx, nl, SP_OBJECTIVES, scal.t(), d_flags.as<int>(), nullptr);
  scalar_program(ctx, nl, SP_ERRORS, scal.t(), d_flags.as<int>(), d_scal_out.as<double>());
}

int Solver::iterate(clrsdp_iter_info* info) {
  if (!prepared) return CLRSDP_ERR_STATE;
  CLR_CUDA(cudaSetDevice(ctx.device));
  double t0 = now_s();
  if (main_stream_) ctx.stream = main_stream_;  // a failed iteration may have left the side stream selected
  on_side_ = false;
  join_pending_ = false;
  ev_marks_.clear();
  n_direct_marks_ = 0;
  CLR_CUDA(cudaMemsetAsync(d_status.p, 0, sizeof(int) * n_status, ctx.stream));
  mark(CLRSDP_T_COUNT);  // whole iteration (extra bucket, reported as `seconds`)
  clrsdp_iter_info row;
  memset(&row, 0, sizeof(row));
  row.iter = iter;
  // values printed in this iteration's row come from the start of the iteration (:923-937)
  row.p_obj = h_scal[SL_P_OBJ];
  row.d_obj = h_scal[SL_D_OBJ];
  row.gap = h_scal[SL_GAP];
  // the body is replayed from a CUDA graph once the buffers have reached their steady-state sizes (the second
  // iteration after a prepare); per-kernel profiling and CLRSDP_GRAPH=0 keep the direct launches
  bool ran = false;
  if (use_graph_ && !ctx.profiling) {
    if (gexec_ && graph_epoch_ != alloc_epoch()) drop_graph();
    if (!gexec_ && direct_iters_ >= 1) {
      cudaGraph_t graph = nullptr;
      const uint64_t e0 = alloc_epoch();
      const int64_t l0 = ctx.launches;
      graph_marks_.clear();
      CLR_CUDA(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal));
      capturing_ = true;
      try {
        iteration_body();
      } catch (...) {
        capturing_ = false;
        cudaStreamEndCapture(ctx.stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      capturing_ = false;
      CLR_CUDA(cudaStreamEndCapture(ctx.stream, &graph));
      ctx.launches = l0;  // nothing ran yet
      if (alloc_epoch() == e0) {
        CLR_CUDA(cudaGraphInstantiate(&gexec_, graph, 0));
        graph_epoch_ = e0;
        graph_launches_ = 0;
        // count the kernel nodes: clrsdp_launch_count stays an exact count of this library's kernels
        size_t nn = 0;
        CLR_CUDA(cudaGraphGetNodes(graph, nullptr, &nn));
        std::vector<cudaGraphNode_t> nodes(nn);
        CLR_CUDA(cudaGraphGetNodes(graph, nodes.data(), &nn));
        for (auto nd : nodes) {
          cudaGraphNodeType ty;
          CLR_CUDA(cudaGraphNodeGetType(nd, &ty));
          if (ty == cudaGraphNodeTypeKernel) graph_launches_++;
        }
      }
      CLR_CUDA(cudaGraphDestroy(graph));
    }
    if (gexec_) {
      CLR_CUDA(cudaGraphLaunch(gexec_, ctx.stream));
      ctx.launches += graph_launches_;
      ev_marks_.insert(ev_marks_.end(), graph_marks_.begin(), graph_marks_.end());  // re-recorded by this replay
      ran = true;
    }
  }
  if (!ran) {
    iteration_body();
    direct_iters_++;
  }
  mark(-1 - CLRSDP_T_COUNT);
  int st = check_status();
  iter += 1;
  row.status = st;
  row.mu = h_scal[SL_MU];
  row.P_err = h_scal[SL_PERR_P];
  row.p_err = h_scal[SL_PERR_p];
  row.d_err = h_scal[SL_DERR];
  row.alpha_p = h_scal[SL_ALPHA_P];
  row.alpha_d = h_scal[SL_ALPHA_D];
  row.beta_c = h_scal[SL_BETA_C];
  row.p_obj_new = h_scal[SL_P_OBJ];
  row.d_obj_new = h_scal[SL_D_OBJ];
  row.gap_new = h_scal[SL_GAP];
  row.primal_err_new = h_scal[SL_PRIMAL_ERR];
  row.dual_err_new = h_scal[SL_DUAL_ERR];
  row.pd_feasible = h_flags[0];
  row.terminate = h_flags[1];
  if (row.terminate == CLRSDP_RUNNING && iter >= ip.maxiterations) row.terminate = CLRSDP_MAXITER;
  // timing buckets from the CUDA events
  {
    std::vector<int> open(CLRSDP_T_COUNT + 1, -1);
    double total = 0;
    for (size_t i = 0; i < ev_marks_.size(); i++) {
      int bkt = ev_marks_[i].bucket;
      if (ev_marks_[i].sign > 0) {
        open[bkt] = (int)i;
      } else if (open[bkt] >= 0) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_marks_[open[bkt]].ev, ev_marks_[i].ev);
        if (bkt < CLRSDP_T_COUNT)
          row.timings[bkt] += ms * 1e-3;
        else
          total = ms * 1e-3;
        open[bkt] = -1;
      }
    }
    row.seconds = total;  // device time of the iteration (CUDA events on the launching stream)
  }
  (void)t0;
  // A lost iterate must be loud: once mu or a step length is zero, negative or not finite the run cannot recover (the
  // reference would walk on to maxiterations with garbage); report it instead of a row that looks like progress.
  if (st == 0 && !(std::isfinite(row.mu) && row.mu > 0 && std::isfinite(row.alpha_p) && row.alpha_p > 0 &&
                   std::isfinite(row.alpha_d) && row.alpha_d > 0 && std::isfinite(row.beta_c) &&
                   std::isfinite(row.p_obj_new) && std::isfinite(row.d_obj_new))) {
    st = CLRSDP_ERR_DIVERGED;
    row.status = st;
    err = "the iterate was lost (mu, a step length or an objective is zero, negative or not finite): the working precision is too low for this instance — try again with higher precision";
  }
  if (st) prepared = false;
  if (info) *info = row;
  return st;
}

int Solver::solve(clrsdp_iter_info* rows, int max_rows, int* n_rows) {
  clrsdp_iter_info pi;
  int st = prepare(&pi);
  int n = 0;
  int term = pi.terminate;
  while (st == 0 && term == CLRSDP_RUNNING && iter < ip.maxiterations) {  // (:742-753)
    clrsdp_iter_info row;
    st = iterate(&row);
    if (rows && n < max_rows) rows[n] = row;
    n++;
    term = (row.terminate == CLRSDP_MAXITER) ? CLRSDP_RUNNING : row.terminate;
  }
  if (n_rows) *n_rows = n;
  return st;
}

// ---- fetch ----------------------------------------------------------------------------------------
int64_t Solver::fetch(const char* name, int j, int l, clrsdp_mp_out* out) {
  if (!structure_set) return CLRSDP_ERR_STATE;
  CLR_CUDA(cudaSetDevice(ctx.device));
  std::string nm(name);
  auto put = [&](const MpBuf& t, int64_t off, int64_t n) -> int64_t {
    if (out) {
      if (out->n < n) return CLRSDP_ERR_BAD_ARG;
      ctx.sync();
      to_host(t, off, n, out, 0);
    }
    return n;
  };
  std::map<std::string, MpBuf*> vecs = {{"x", &x}, {"dx", &dx}, {"d", &d}, {"c", &c}, {"dx_pred", &dx_pred},
                                        {"y", &y}, {"dy", &dy}, {"p", &p}, {"b", &b}, {"dy_pred", &dy_pred}};
  auto vit = vecs.find(nm);
  if (vit != vecs.end()) return put(*vit->second, 0, (int64_t)vit->second->count);
  std::map<std::string, MpBuf*> blks = {{"X", &X}, {"Y", &Y}, {"Xinv", &Xinv}, {"R", &R}, {"P", &P}, {"Z", &Z},
                                        {"dX", &dX}, {"dY", &dY}, {"XY", &XY}, {"dX_pred", &dX_pred},
                                        {"dY_pred", &dY_pred}, {"Linvx", &Linvx}, {"Linvy", &Linvy}};
  auto bit = blks.find(nm);
  if (bit != blks.end() || nm == "Px" || nm == "Py") {
    if (j < 0 || j >= J || l < 0 || l >= clusters_[j].L) return CLRSDP_ERR_BAD_ARG;
    const HostBlock& bk = blocks_[clusters_[j].blk0 + l];
    if (nm == "Px") return put(Px, bk.Poff, (int64_t)(bk.m * bk.Nv) * (bk.m * bk.Nv));
    if (nm == "Py") return put(Py, bk.Poff, (int64_t)(bk.m * bk.Nv) * (bk.m * bk.Nv));
    return put(*bit->second, bk.off, (int64_t)bk.nb * bk.nb);
  }
  if (nm == "S" || nm == "Sfac" || nm == "Sinvfac") {
    if (j < 0 || j >= J) return CLRSDP_ERR_BAD_ARG;
    MpBuf& t = nm == "S" ? S : (nm == "Sfac" ? Us : Linvs);
    return put(t, clusters_[j].Soff, (int64_t)clusters_[j].dimS * clusters_[j].dimS);
  }
  if (nm == "W") return put(Wt, 0, (int64_t)sumS * n_y);
  if (nm == "Q") return put(Q, 0, (int64_t)n_y * n_y);
  if (nm == "Qfac") return put(Uq, 0, (int64_t)n_y * n_y);
  if (nm == "scalar") {
    const int map_[CLRSDP_S_COUNT] = {SL_MU,      SL_P_OBJ,  SL_D_OBJ, SL_GAP,  SL_PRIMAL_ERR, SL_DUAL_ERR, SL_ALPHA_P,
                                      SL_ALPHA_D, SL_BETA_C, SL_MU_P,  SL_MU_C, SL_LAM_X,      SL_LAM_Y};
    if (j < 0 || j >= CLRSDP_S_COUNT) return CLRSDP_ERR_BAD_ARG;
    return put(scal, map_[j], 1);
  }
  return CLRSDP_ERR_BAD_ARG;
}

// ---- phase-level ops ------------------------------------------------------------------------------
void Solver::op_gemm(int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B, clrsdp_mp_out* C) {
  CLR_CUDA(cudaSetDevice(ctx.device));
  if (A->n != (int64_t)batch * M * K || B->n != (int64_t)batch * K * N || C->n < (int64_t)batch * M * N)
    throw SolverError(CLRSDP_ERR_BAD_ARG, "op_gemm: sizes");
  MpBuf a, b2, c2;
  a.alloc(A->n, nl), b2.alloc(B->n, nl), c2.alloc((int64_t)batch * M * N, nl);
  to_device(A, 0, A->n, a, 0);
  to_device(B, 0, B->n, b2, 0);
  OperandDesc ad, bd;
  ad.src = a.t(), ad.batch = batch, ad.rows = M, ad.K = K, ad.rs = K, ad.ks = 1, ad.bstride = (int64_t)M * K;
  bd.src = b2.t(), bd.batch = batch, bd.rows = N, bd.K = K, bd.rs = 1, bd.ks = N, bd.bstride = (int64_t)K * N;
  Slice sa, sb;
  ge()->slice(ad, sa);
  ge()->slice(bd, sb);
  OutDesc o;
  o.dst = c2.t(), o.bstride = (int64_t)M * N, o.rs = N, o.cs = 1;
  ge()->multiply(sa, sb, plan_of(batch, M, N), o);
  ctx.sync();
  to_host(c2, 0, (int64_t)batch * M * N, C, 0);
}
void Solver::op_gemm_planes(int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B, int32_t* planes,
                            int* n_planes, int32_t* row_exp, int32_t* col_exp) {
  CLR_CUDA(cudaSetDevice(ctx.device));
  MpBuf a, b2;
  a.alloc(A->n, nl), b2.alloc(B->n, nl);
  to_device(A, 0, A->n, a, 0);
  to_device(B, 0, B->n, b2, 0);
  OperandDesc ad, bd;
  ad.src = a.t(), ad.batch = batch, ad.rows = M, ad.K = K, ad.rs = K, ad.ks = 1, ad.bstride = (int64_t)M * K;
  bd.src = b2.t(), bd.batch = batch, bd.rows = N, bd.K = K, bd.rs = 1, bd.ks = N, bd.bstride = (int64_t)K * N;
  Slice sa, sb;
  ge()->slice(ad, sa);
  ge()->slice(bd, sb);
  if (*n_planes < ge()->digits()) throw SolverError(CLRSDP_ERR_BAD_ARG, "op_gemm_planes: plane buffer too small");
  int T = 0;
  std::vector<int32_t> tmp((size_t)ge()->digits() * batch * M * N);
  ge()->planes_only(sa, sb, plan_of(batch, M, N), tmp.data(), &T);
  memcpy(planes, tmp.data(), tmp.size() * sizeof(int32_t));
  *n_planes = T;
  CLR_CUDA(cudaMemcpy(row_exp, sa.exps.p, sizeof(int32_t) * batch * M, cudaMemcpyDeviceToHost));
  CLR_CUDA(cudaMemcpy(col_exp, sb.exps.p, sizeof(int32_t) * batch * N, cudaMemcpyDeviceToHost));
}
int Solver::op_cholesky(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* L, clrsdp_mp_out* Linv) {
  CLR_CUDA(cudaSetDevice(ctx.device));
  int64_t tot = (int64_t)batch * n * n;
  if (batch <= 0 || n <= 0 || !A || A->n != tot || (L && L->n < tot) || (Linv && Linv->n < tot))
    throw SolverError(CLRSDP_ERR_BAD_ARG, "op_cholesky: operand / result sizes");
  MpBuf a, u, v, li, rd;
  a.alloc(tot, nl), u.alloc(tot, nl), v.alloc(tot, nl), li.alloc(tot, nl), rd.alloc((int64_t)batch * n, nl);
  to_device(A, 0, tot, a, 0);
  std::vector<int64_t> off(batch);
  for (int i = 0; i < batch; i++) off[i] = (int64_t)i * n * n;
  DevBuf doff, dst;
  upload(doff, off, ctx.stream);
  dst.ensure(sizeof(int) * batch);
  MatBatch Ab{a.t(), doff.as<int64_t>(), batch, n}, Ub{u.t(), doff.as<int64_t>(), batch, n};
  MatBatch Vb{v.t(), doff.as<int64_t>(), batch, n}, Lb{li.t(), doff.as<int64_t>(), batch, n};
  CLR_CUDA(cudaMemsetAsync(dst.p, 0, sizeof(int) * batch, ctx.stream));
  rdiag.alloc(std::max<size_t>(rdiag.n, (size_t)batch * n), nl);
  chol_inverse(Ab, Ub, Vb, Lb, dst.as<int>(), true);
  std::vector<int> hs(batch);
  CLR_CUDA(cudaMemcpyAsync(hs.data(), dst.p, sizeof(int) * batch, cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  for (int s : hs)
    if (s) return CLRSDP_ERR_NOT_PD_X;
  if (L) {  // L = U^T: transpose on the host side of the copy (entries below the diagonal of U are ignored)
    std::vector<int8_t> sg(tot);
    std::vector<int64_t> ex(tot);
    std::vector<uint32_t> lb((size_t)nl * tot);
    clrsdp_mp_out tmp{sg.data(), ex.data(), lb.data(), tot};
    to_host(u, 0, tot, &tmp, 0);
    for (int bq = 0; bq < batch; bq++)
      for (int r = 0; r < n; r++)
        for (int cc = 0; cc < n; cc++) {
          int64_t s = (int64_t)bq * n * n + (int64_t)cc * n + r, t = (int64_t)bq * n * n + (int64_t)r * n + cc;
          bool upper = cc > r;  // U is only meaningful on and above its diagonal
          L->sign[t] = upper ? 0 : sg[s];
          L->exp[t] = upper ? 0 : ex[s];
          for (int k = 0; k < nl; k++) L->limb[(size_t)k * L->n + t] = upper ? 0u : lb[(size_t)k * tot + s];
        }
  }
  if (Linv) to_host(li, 0, tot, Linv, 0);
  return 0;
}
// the factorisation of S_j and Q as a stand-alone op: equilibrated A' = D^-1 A D^-1 = U^T Sigma U (signed, panel_factor);
// delivers M = L^-1 D^-1 (L = U^T) and the signs, so that A^-1 = M^T Sigma M
void Solver::op_signed_factor(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* Minv, int32_t* signs) {
  CLR_CUDA(cudaSetDevice(ctx.device));
  int64_t tot = (int64_t)batch * n * n;
  if (A->n != tot || Minv->n < tot) throw SolverError(CLRSDP_ERR_BAD_ARG, "op_signed_factor: sizes");
  MpBuf a, u, li;
  a.alloc(tot, nl), u.alloc(tot, nl), li.alloc(tot, nl);
  to_device(A, 0, tot, a, 0);
  std::vector<int64_t> off(batch);
  for (int i = 0; i < batch; i++) off[i] = (int64_t)i * n * n;
  DevBuf doff, dst, dsig;
  upload(doff, off, ctx.stream);
  dst.ensure(sizeof(int) * batch);
  dsig.ensure(sizeof(int) * (size_t)batch * n);
  MatBatch Ab{a.t(), doff.as<int64_t>(), batch, n}, Ub{u.t(), doff.as<int64_t>(), batch, n};
  MatBatch Lb{li.t(), doff.as<int64_t>(), batch, n};
  CLR_CUDA(cudaMemsetAsync(dst.p, 0, sizeof(int) * batch, ctx.stream));
  chol_inverse(Ab, Ub, Ub, Lb, dst.as<int>(), false, false, dsig.as<int>());
  std::vector<int> hs((size_t)batch * n);
  CLR_CUDA(cudaMemcpyAsync(hs.data(), dsig.p, sizeof(int) * hs.size(), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  for (size_t i = 0; i < hs.size(); i++) signs[i] = hs[i] ? -1 : 1;
  to_host(li, 0, tot, Minv, 0);
}
void Solver::op_elementwise(int op, const clrsdp_mp* a, const clrsdp_mp* b2, clrsdp_mp_out* cc) {
  if (!a || !cc || a->n <= 0 || (b2 && b2->n != a->n) || cc->n < a->n)
    throw SolverError(CLRSDP_ERR_BAD_ARG, "op_elementwise: operand / result sizes");
  CLR_CUDA(cudaSetDevice(ctx.device));
  MpBuf A, B, C;
  A.alloc(a->n, nl), B.alloc(a->n, nl), C.alloc(a->n, nl);
  to_device(a, 0, a->n, A, 0);
  if (b2) to_device(b2, 0, a->n, B, 0);
  ew_binary(ctx, nl, op, C.t(), A.t(), B.t(), a->n);
  ctx.sync();
  to_host(C, 0, a->n, cc, 0);
}
void Solver::op_lambda_min(int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* lamo) {
  CLR_CUDA(cudaSetDevice(ctx.device));
  int64_t tot = (int64_t)batch * n * n;
  if (batch <= 0 || n <= 0 || !A || !lamo || A->n != tot || lamo->n < batch)
    throw SolverError(CLRSDP_ERR_BAD_ARG, "op_lambda_min: operand / result sizes");
  MpBuf a, out;
  a.alloc(tot, nl), out.alloc(batch, nl);
  to_device(A, 0, tot, a, 0);
  std::vector<int64_t> off(batch);
  for (int i = 0; i < batch; i++) off[i] = (int64_t)i * n * n;
  DevBuf doff;
  upload(doff, off, ctx.stream);
  DevBuf fl;
  fl.ensure(sizeof(int) * std::max(1, batch));
  lambda_min(ctx, nl, MatBatch{a.t(), doff.as<int64_t>(), batch, n}, out.t(), nullptr, fl.as<int>());
  ctx.sync();
  to_host(out, 0, batch, lamo, 0);
}

}  // namespace clr

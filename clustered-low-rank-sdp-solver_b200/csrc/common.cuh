// common.cuh — error handling, device buffers, launch accounting for libclrsdp.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "mpf.cuh"

namespace clr {

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct SolverError : std::runtime_error {
  int code;
  SolverError(int c, const std::string& w) : std::runtime_error(w), code(c) {}
};

#define CLR_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw clr::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +     \
                           __FILE__ + ":" + std::to_string(__LINE__));                              \
  } while (0)

// Per-handle launch accounting. Every kernel launch goes through Ctx::launch_begin/end so that
//  (a) `clrsdp_launch_count` is an exact count of this library's kernels, and
//  (b) with profiling enabled each launch is bracketed by CUDA events on the launching stream and
//      accumulated per kernel name (bench.py's live roofline measurement).
struct ProfEntry {
  double ms = 0;
  int64_t launches = 0;
  double work = 0;
};
struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  bool profiling = false;
  struct Pending {
    std::string name;
    cudaEvent_t a, b;
    double work;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  std::map<std::string, ProfEntry> prof;

  cudaEvent_t get_event() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    CLR_CUDA(cudaEventCreate(&e));
    return e;
  }
  // usage: auto tk = ctx.begin("name", work); kernel<<<...,ctx.stream>>>(...); ctx.end(tk);
  int begin(const char* name, double work = 0) {
    launches++;
    if (!profiling) return -1;
    Pending p{name, get_event(), get_event(), work};
    CLR_CUDA(cudaEventRecord(p.a, stream));
    pending.push_back(p);
    return (int)pending.size() - 1;
  }
  void end(int tk) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw CudaError(std::string("kernel launch failed: ") + cudaGetErrorString(e));
    if (tk >= 0) CLR_CUDA(cudaEventRecord(pending[tk].b, stream));
  }
  void resolve() {
    if (pending.empty()) return;
    CLR_CUDA(cudaStreamSynchronize(stream));
    for (auto& p : pending) {
      float ms = 0;
      CLR_CUDA(cudaEventElapsedTime(&ms, p.a, p.b));
      ProfEntry& pe = prof[p.name];
      pe.ms += ms;
      pe.launches += 1;
      pe.work += p.work;
      pool.push_back(p.a);
      pool.push_back(p.b);
    }
    pending.clear();
  }
  void sync() { CLR_CUDA(cudaStreamSynchronize(stream)); }
};

// Bumped by every device (re)allocation or release: a captured CUDA graph bakes device addresses (and TMA
// descriptors) into its nodes, so the solver drops its graph whenever the epoch moved since the capture.
inline uint64_t& alloc_epoch() {
  static uint64_t e = 0;
  return e;
}

// simple owning device buffer
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr, o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p, bytes = o.bytes;
      o.p = nullptr, o.bytes = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      cudaFree(p);
      alloc_epoch()++;
    }
    p = nullptr;
    bytes = 0;
  }
  void ensure(size_t b) {  // grow-only, contents NOT preserved
    if (b <= bytes) return;
    release();
    CLR_CUDA(cudaMalloc(&p, b));
    alloc_epoch()++;
    bytes = b;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// a multiprecision tensor in HBM (planar layout of mpf.cuh)
struct MpBuf {
  DevBuf buf;
  size_t n = 0;      // plane stride in words
  size_t count = 0;  // numbers addressable through this buffer / view
  int nl = 0;
  uint32_t* base = nullptr;  // non-null for a view into another MpBuf
  void alloc(size_t n_, int nl_) {
    n = n_ ? n_ : 1;
    count = n;
    nl = nl_;
    base = nullptr;
    buf.ensure((size_t)(nl + 1) * n * sizeof(uint32_t));
  }
  // view of `cnt` numbers starting at element `off` of parent (shares the parent's planes)
  void alias(const MpBuf& parent, size_t off, size_t cnt) {
    n = parent.n;
    count = cnt;
    nl = parent.nl;
    base = parent.w() + off;
  }
  mp::Tensor t() const { return mp::Tensor{w(), n}; }
  uint32_t* w() const { return base ? base : buf.as<uint32_t>(); }
};

template <class T>
inline void upload(DevBuf& d, const std::vector<T>& h, cudaStream_t s) {
  d.ensure(std::max<size_t>(1, h.size()) * sizeof(T));
  if (!h.empty()) CLR_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace clr

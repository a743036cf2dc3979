// comm.cuh — NCCL plumbing for cluster sharding. libnccl is loaded lazily (dlopen) so that the
// library has no hard NCCL dependency on single-GPU boxes. The only collectives the hot path needs are
// small: Q (n_y^2), three n_y-vectors and a handful of scalars per iteration.
//
// Multiprecision reduction: every rank contributes its partial tensor, ncclAllGather collects the raw
// planes of all ranks, and a local kernel combines them IN RANK ORDER. All ranks therefore compute
// bit-identical results, which the replicated state (y, dy, Q and its factor, the driver scalars)
// relies on.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace clr {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  static NcclApi& get() {
    static NcclApi api;
    if (!api.lib) {
      api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (!api.lib) throw SolverError(CLRSDP_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
      auto sym = [&](const char* n) {
        void* p = dlsym(api.lib, n);
        if (!p) throw SolverError(CLRSDP_ERR_NCCL, std::string("libnccl lacks ") + n);
        return p;
      };
      api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
      api.CommAbort = (decltype(api.CommAbort))sym("ncclCommAbort");
      api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
      api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    }
    return api;
  }
};

#define CLR_NCCL(expr)                                                                                   \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != ncclSuccess)                                                                               \
      throw clr::SolverError(CLRSDP_ERR_NCCL, std::string(#expr) + ": " + clr::NcclApi::get().GetErrorString(_r)); \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  DevBuf stage, gathered;
  bool active() const { return nranks > 1; }
};

enum { COMB_SUM = 0, COMB_MAX = 1, COMB_MIN = 2 };
// out[off + i] = combine over ranks r (in rank order) of gathered[r][plane][i]
void combine_ranks(Ctx& ctx, int nl, const uint32_t* gathered, int nranks, int64_t n, mp::Tensor out, int64_t off, int op);

}  // namespace clr

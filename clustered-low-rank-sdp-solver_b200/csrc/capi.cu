// capi.cu — the extern "C" surface declared in include/clrsdp.h. Exceptions never cross the boundary:
// every entry point maps them to a status code and keeps the message for clrsdp_last_error.
#include <algorithm>
#include <cstring>

#include "multi.cuh"

struct clrsdp_solver {
  clr::SolverApi* s = nullptr;  // clr::Solver (one GPU) or clr::MultiSolver (several GPUs, one process)
  std::string err;
};

#define CAPI extern "C" __attribute__((visibility("default")))

template <class F>
static int guard(clrsdp_handle h, F f) {
  if (!h || !h->s) return CLRSDP_ERR_BAD_ARG;
  try {
    return f(*h->s);
  } catch (const clr::SolverError& e) {
    h->err = e.what();
    return e.code;
  } catch (const clr::CudaError& e) {
    h->err = e.what();
    return CLRSDP_ERR_CUDA;
  } catch (const std::exception& e) {
    h->err = e.what();
    return CLRSDP_ERR_BAD_ARG;
  }
}

CAPI int clrsdp_create(clrsdp_handle* h, int prec_bits, int device) {
  if (!h) return CLRSDP_ERR_BAD_ARG;
  *h = nullptr;
  clrsdp_solver* w = new clrsdp_solver();
  try {
    w->s = new clr::Solver(prec_bits, device);
  } catch (const clr::SolverError& e) {
    int code = e.code;
    fprintf(stderr, "clrsdp_create: %s\n", e.what());
    delete w;
    return code;
  } catch (const std::exception& e) {
    fprintf(stderr, "clrsdp_create: %s\n", e.what());
    delete w;
    return CLRSDP_ERR_CUDA;
  }
  *h = w;
  return CLRSDP_OK;
}
CAPI int clrsdp_destroy(clrsdp_handle h) {
  if (h) {
    delete h->s;
    delete h;
  }
  return CLRSDP_OK;
}
CAPI const char* clrsdp_last_error(clrsdp_handle h) { return h ? h->err.c_str() : "null handle"; }

CAPI int clrsdp_set_structure(clrsdp_handle h, int J, int n_y, const int* m, const int* L, const int* n_samples,
                              const int* delta, const int* ranks) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!m || !L || !n_samples || !delta || !ranks) return (int)CLRSDP_ERR_BAD_ARG;
    s.set_structure(J, n_y, m, L, n_samples, delta, ranks);
    return 0;
  });
}
CAPI int clrsdp_upload_cluster(clrsdp_handle h, int j, const clrsdp_mp* V, const clrsdp_mp* H, const clrsdp_mp* B,
                               const clrsdp_mp* c) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!V || !H || !B || !c) return (int)CLRSDP_ERR_BAD_ARG;
    s.upload_cluster(j, V, H, B, c);
    return 0;
  });
}
CAPI int clrsdp_upload_objective(clrsdp_handle h, const clrsdp_mp* b, const clrsdp_mp* b0) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!b) return (int)CLRSDP_ERR_BAD_ARG;
    s.upload_objective(b, b0);
    return 0;
  });
}
CAPI int clrsdp_upload_C(clrsdp_handle h, const clrsdp_mp* C) {
  return guard(h, [&](clr::SolverApi& s) {
    s.upload_C(C);
    return 0;
  });
}
CAPI int clrsdp_set_params(clrsdp_handle h, const clrsdp_mp* rp, const clrsdp_int_params* ip) {
  return guard(h, [&](clr::SolverApi& s) {
    s.set_params(rp, ip);
    return 0;
  });
}
CAPI int clrsdp_init_point(clrsdp_handle h) {
  return guard(h, [&](clr::SolverApi& s) {
    s.init_point();
    return 0;
  });
}
CAPI int clrsdp_upload_point(clrsdp_handle h, const clrsdp_mp* x, const clrsdp_mp* X, const clrsdp_mp* y,
                             const clrsdp_mp* Y) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!x || !X || !y || !Y) return (int)CLRSDP_ERR_BAD_ARG;
    s.upload_point(x, X, y, Y);
    return 0;
  });
}
CAPI int clrsdp_download_point(clrsdp_handle h, clrsdp_mp_out* x, clrsdp_mp_out* X, clrsdp_mp_out* y,
                               clrsdp_mp_out* Y) {
  return guard(h, [&](clr::SolverApi& s) {
    s.download_point(x, X, y, Y);
    return 0;
  });
}
CAPI int clrsdp_prepare(clrsdp_handle h, clrsdp_iter_info* info) {
  return guard(h, [&](clr::SolverApi& s) { return s.prepare(info); });
}
CAPI int clrsdp_iterate(clrsdp_handle h, clrsdp_iter_info* info) {
  return guard(h, [&](clr::SolverApi& s) { return s.iterate(info); });
}
CAPI int clrsdp_solve(clrsdp_handle h, clrsdp_iter_info* rows, int max_rows, int* n_rows) {
  return guard(h, [&](clr::SolverApi& s) { return s.solve(rows, max_rows, n_rows); });
}
CAPI int64_t clrsdp_fetch(clrsdp_handle h, const char* name, int j, int l, clrsdp_mp_out* out) {
  if (!h || !h->s || !name) return CLRSDP_ERR_BAD_ARG;
  try {
    return h->s->fetch(name, j, l, out);
  } catch (const clr::SolverError& e) {
    h->err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->err = e.what();
    return CLRSDP_ERR_CUDA;
  }
}
CAPI int clrsdp_op_gemm(clrsdp_handle h, int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B,
                        clrsdp_mp_out* C) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!A || !B || !C || batch <= 0 || M <= 0 || N <= 0 || K <= 0) return (int)CLRSDP_ERR_BAD_ARG;
    s.op_gemm(batch, M, N, K, A, B, C);
    return 0;
  });
}
CAPI int clrsdp_op_gemm_planes(clrsdp_handle h, int batch, int M, int N, int K, const clrsdp_mp* A, const clrsdp_mp* B,
                               int32_t* planes, int* n_planes, int32_t* row_exp, int32_t* col_exp) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!A || !B || !planes || !n_planes || !row_exp || !col_exp) return (int)CLRSDP_ERR_BAD_ARG;
    s.op_gemm_planes(batch, M, N, K, A, B, planes, n_planes, row_exp, col_exp);
    return 0;
  });
}
CAPI int clrsdp_op_cholesky(clrsdp_handle h, int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* L,
                            clrsdp_mp_out* Linv) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!A) return (int)CLRSDP_ERR_BAD_ARG;
    return s.op_cholesky(batch, n, A, L, Linv);
  });
}
CAPI int clrsdp_op_signed_factor(clrsdp_handle h, int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* Minv,
                                 int32_t* signs) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!A || !Minv || !signs || batch <= 0 || n <= 0) return (int)CLRSDP_ERR_BAD_ARG;
    s.op_signed_factor(batch, n, A, Minv, signs);
    return 0;
  });
}
CAPI int clrsdp_op_lambda_min(clrsdp_handle h, int batch, int n, const clrsdp_mp* A, clrsdp_mp_out* lam) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!A || !lam) return (int)CLRSDP_ERR_BAD_ARG;
    s.op_lambda_min(batch, n, A, lam);
    return 0;
  });
}
CAPI int clrsdp_op_elementwise(clrsdp_handle h, int op, const clrsdp_mp* a, const clrsdp_mp* b, clrsdp_mp_out* c) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!a || !c) return (int)CLRSDP_ERR_BAD_ARG;
    s.op_elementwise(op, a, b, c);
    return 0;
  });
}
CAPI int clrsdp_comm_unique_id(uint8_t id[128]) {
  if (!id) return CLRSDP_ERR_BAD_ARG;
  try {
    ncclUniqueId uid;
    if (clr::NcclApi::get().GetUniqueId(&uid) != ncclSuccess) return CLRSDP_ERR_NCCL;
    memcpy(id, uid.internal, 128);
    return CLRSDP_OK;
  } catch (...) {
    return CLRSDP_ERR_NCCL;
  }
}
CAPI int clrsdp_comm_init(clrsdp_handle h, int n_ranks, int rank, const uint8_t id[128]) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!id) return (int)CLRSDP_ERR_BAD_ARG;
    s.comm_init(n_ranks, rank, id);
    return 0;
  });
}
CAPI int clrsdp_pin_host(clrsdp_handle h, void* p, size_t bytes) {
  return guard(h, [&](clr::SolverApi& s) {
    s.pin_host(p, bytes);
    return 0;
  });
}
CAPI int clrsdp_unpin_host(clrsdp_handle h, void* p) {
  return guard(h, [&](clr::SolverApi& s) {
    s.unpin_host(p);
    return 0;
  });
}
CAPI int clrsdp_measure_int8_peak(clrsdp_handle h, double* macs_per_second) {
  return guard(h, [&](clr::SolverApi& s) {
    if (!macs_per_second) return (int)CLRSDP_ERR_BAD_ARG;
    *macs_per_second = s.measure_i8_peak();
    return 0;
  });
}
CAPI int64_t clrsdp_launch_count(clrsdp_handle h) {
  try {
    return (h && h->s) ? h->s->launch_count() : 0;
  } catch (...) {
    return 0;
  }
}
CAPI int clrsdp_profile_reset(clrsdp_handle h, int enable) {
  return guard(h, [&](clr::SolverApi& s) {
    s.profile_reset(enable != 0);
    return 0;
  });
}
CAPI int clrsdp_profile_query(clrsdp_handle h, const char* pattern, double* ms, int64_t* launches, double* work) {
  return guard(h, [&](clr::SolverApi& s) {
    double t = 0, w = 0;
    int64_t n = 0;
    for (auto& kv : s.profile_table())
      if (!pattern || kv.first.find(pattern) != std::string::npos) t += kv.second.ms, n += kv.second.launches, w += kv.second.work;
    if (ms) *ms = t;
    if (launches) *launches = n;
    if (work) *work = w;
    return 0;
  });
}
CAPI int clrsdp_profile_dump(clrsdp_handle h, char* buf, int buf_len) {
  if (!h || !h->s) return CLRSDP_ERR_BAD_ARG;
  std::map<std::string, clr::ProfEntry> tab;
  try {
    tab = h->s->profile_table();
  } catch (...) {
    return CLRSDP_ERR_CUDA;
  }
  std::string out;
  char line[256];
  for (auto& kv : tab) {
    snprintf(line, sizeof(line), "%s %.6f %lld %.6e\n", kv.first.c_str(), kv.second.ms, (long long)kv.second.launches,
             kv.second.work);
    out += line;
  }
  if (buf && buf_len > 0) {
    int n = std::min<int>((int)out.size(), buf_len - 1);
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int)out.size() + 1;
}

// ---- several GPUs behind one handle, one process (multi.cuh) -----------------------------------------------------------
CAPI int clrsdp_create_multi(clrsdp_handle* h, int prec_bits, int n_dev, const int* dev_ids) {
  if (!h || n_dev < 1) return CLRSDP_ERR_BAD_ARG;
  *h = nullptr;
  clrsdp_solver* w = new clrsdp_solver();
  try {
    if (n_dev == 1)
      w->s = new clr::Solver(prec_bits, dev_ids ? dev_ids[0] : 0);
    else
      w->s = new clr::MultiSolver(prec_bits, n_dev, dev_ids);
  } catch (const clr::SolverError& e) {
    int code = e.code;
    fprintf(stderr, "clrsdp_create_multi: %s\n", e.what());
    delete w;
    return code;
  } catch (const std::exception& e) {
    fprintf(stderr, "clrsdp_create_multi: %s\n", e.what());
    delete w;
    return CLRSDP_ERR_CUDA;
  }
  *h = w;
  return CLRSDP_OK;
}
CAPI int clrsdp_cluster_owner(clrsdp_handle h, int J, int* owner) {
  if (!h || !h->s || !owner) return CLRSDP_ERR_BAD_ARG;
  clr::MultiSolver* m = dynamic_cast<clr::MultiSolver*>(h->s);
  if (!m) {
    for (int j = 0; j < J; j++) owner[j] = 0;
    return CLRSDP_OK;
  }
  if ((int)m->owner().size() != J) return CLRSDP_ERR_STATE;
  for (int j = 0; j < J; j++) owner[j] = m->owner()[j];
  return CLRSDP_OK;
}
CAPI double clrsdp_partition(const double* weights, int n, int parts, int* set_of) {
  if (!weights || !set_of || n < 0 || parts < 1) return -1.0;
  try {
    return clr::partition_weights(weights, n, parts, set_of);
  } catch (...) {
    return -1.0;
  }
}
CAPI double clrsdp_cluster_weight(int m, int L, int n_samples, const int* delta, int n_y) {
  if (!delta || m < 1 || L < 1 || n_samples < 1) return -1.0;
  return clr::cluster_weight(m, L, n_samples, delta, n_y);
}

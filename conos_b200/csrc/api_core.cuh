// Opaque handle definitions shared by the C-ABI translation units.
#pragma once
#include "../../include/conos_gpu.h"
#include "common.cuh"
#include "edges.cuh"
#include "knn.cuh"
#include "sparse.cuh"

struct cg_context {
  cg::Context c;
  // result of the last cg_neighbor_matrix call
  cg::EdgeTable coo;
  // reductions / projections kept from the last cg_pairs_edges call (host copies, for the pairs cache and tests)
  struct PairKeep {
    int n_rows = 0, n_cols = 0;
    std::vector<double> rot;  // column-major n_rows x n_cols (caller-supplied rotations)
    const double* rot_host = nullptr;  // where the host copy lives: rot.data() or the context's pinned arena
    std::vector<double> nv;   // [2][n_cols] or empty
    int pn[2] = {0, 0};
    std::vector<double> proj[2];  // column-major n x ncomps
    // CCA: scaled canonical variates, kept-cell row maps, singular values
    std::vector<double> D, niter;  // CPCA: cpcaF's D [2][n_cols] and niter [n_cols]
    std::vector<double> ul, vl;    // CCA: gene loadings, column-major n_rows x cca_d
    std::vector<double> u, v, sigma;
    std::vector<int> rows_u, rows_v;
    int cca_d = 0;
  };
  std::vector<PairKeep> kept;
  bool keep_projections = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

struct cg_panel {
  cg_context* ctx = nullptr;       // creator (may be shut down before the panel is freed: never dereferenced then)
  cudaStream_t stream = nullptr;   // its stream, checked against the registry of live streams
  std::vector<cg::Sample> samples;
  std::vector<int> offsets;
  size_t device_bytes = 0;         // what the resident samples hold (keeps the context's free-memory estimate honest)
};

struct cg_edges {
  cg::EdgeTable t;
};

struct cg_graph {
  cg::Graph g;
};

namespace cg {
thread_local extern std::string g_thread_error;
int fail(cg_context* ctx, const std::exception& e);
int fail(cg_context* ctx, const char* msg);

#define CG_API_BEGIN try {
// binds the context's stream for stream-ordered allocations of the calling thread and makes its device current (a
// multi-device job drives several contexts from one process)
#define CG_BIND(ctxptr)                                                               \
  ::cg::StreamBinding _cg_binding((ctxptr) ? (ctxptr)->c.stream : nullptr);           \
  if (ctxptr) CG_CUDA(cudaSetDevice((ctxptr)->c.device))
#define CG_API_END(ctxptr)                                    \
  return 0;                                                   \
  }                                                           \
  catch (const std::exception& e) { return ::cg::fail(ctxptr, e); } \
  catch (...) { return ::cg::fail(ctxptr, "unknown error"); }

}  // namespace cg

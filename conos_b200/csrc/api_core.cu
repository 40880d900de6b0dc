// C ABI, part 1: context, exact kNN entry points, getNeighborMatrix, edge tables, graph assembly.
#include <cmath>

#include "api_core.cuh"
#include "hub.cuh"

namespace cg {
thread_local std::string g_thread_error;
int fail(cg_context* ctx, const std::exception& e) {
  g_thread_error = e.what();
  if (ctx) ctx->c.last_error = e.what();
  // a sticky CUDA error would poison every later call; clear the non-sticky ones
  cudaGetLastError();
  return 1;
}
int fail(cg_context* ctx, const char* msg) {
  g_thread_error = msg;
  if (ctx) ctx->c.last_error = msg;
  return 1;
}
}  // namespace cg

using namespace cg;

extern "C" {

int cg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int cg_init(int device, cg_context** out) {
  cg_context* ctx = nullptr;
  CG_API_BEGIN
  CG_CHECK(out != nullptr, "cg_init: out is NULL");
  int n = 0;
  CG_CUDA(cudaGetDeviceCount(&n));
  CG_CHECK(n > 0, "no CUDA device visible: libconosgpu has no CPU fallback");
  CG_CHECK(device >= 0 && device < n, "device index out of range");
  CG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CG_CUDA(cudaGetDeviceProperties(&prop, device));
  CG_CHECK(prop.major == 10, "libconosgpu is built for sm_100a (B200) only");
  ctx = new cg_context();
  ctx->c.device = device;
  ctx->c.sm_count = prop.multiProcessorCount;
  CG_CUDA(cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking));
  stream_register(ctx->c.stream);
  {
    // keep freed blocks cached in the device's default pool (stream-ordered DevBuf allocations)
    cudaMemPool_t pool;
    CG_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    CG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    // resident panel buffers come from their own pool: same sizes every upload, so they recycle exactly and leave
    // the scratch pool's layout alone
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CG_CUDA(cudaMemPoolCreate(&ctx->c.panel_pool, &props));
    CG_CUDA(cudaMemPoolSetAttribute(ctx->c.panel_pool, cudaMemPoolAttrReleaseThreshold, &thr));
  }
  ctx->c.refresh_free_bytes();
  *out = ctx;
  CG_API_END(ctx)
}

// Order matters (the reference's DLL is unloaded by R at any time: src/RcppExports.cpp:374-377): everything the
// context owns on the device is released while its stream is still alive, then the stream leaves the registry (so
// that panels / edge tables / graphs the caller still holds fall back to cudaFree), and only then the stream and
// the pool go.  Every CUDA error is swallowed: at process exit the runtime may already be unloading.
void cg_shutdown(cg_context* ctx) {
  if (!ctx) return;
  if (cudaSetDevice(ctx->c.device) != cudaSuccess) cudaGetLastError();
  cudaStream_t st = ctx->c.stream;
  if (st && cudaStreamSynchronize(st) != cudaSuccess) cudaGetLastError();
  ::cg::comm_destroy(ctx->c);
  ctx->coo.from.release();
  ctx->coo.to.release();
  ctx->coo.type.release();
  ctx->coo.w.release();
  ctx->coo.n = 0;
  ctx->kept.clear();
  ctx->c.rot_arena.release();
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  ctx->ev0 = ctx->ev1 = nullptr;
  if (ctx->c.knn_ev0) cudaEventDestroy(ctx->c.knn_ev0);
  if (ctx->c.knn_ev1) cudaEventDestroy(ctx->c.knn_ev1);
  ctx->c.knn_ev0 = ctx->c.knn_ev1 = nullptr;
  if (st) {
    if (cudaStreamSynchronize(st) != cudaSuccess) cudaGetLastError();
    stream_unregister(st);
    if (cudaStreamDestroy(st) != cudaSuccess) cudaGetLastError();
    ctx->c.stream = nullptr;
  }
  if (ctx->c.panel_pool && cudaMemPoolDestroy(ctx->c.panel_pool) != cudaSuccess) cudaGetLastError();
  ctx->c.panel_pool = nullptr;
  delete ctx;
}

const char* cg_last_error(const cg_context* ctx) {
  return ctx ? ctx->c.last_error.c_str() : g_thread_error.c_str();
}

uint64_t cg_kernel_launches(const cg_context* ctx) { return ctx ? ctx->c.stats.kernels : 0; }

int cg_timer_start(cg_context* ctx) {
  CG_API_BEGIN
  CG_CHECK(ctx, "NULL context");
  if (!ctx->ev0) {
    CG_CUDA(cudaEventCreate(&ctx->ev0));
    CG_CUDA(cudaEventCreate(&ctx->ev1));
  }
  CG_CUDA(cudaEventRecord(ctx->ev0, ctx->c.stream));
  CG_API_END(ctx)
}
int cg_timer_stop(cg_context* ctx, double* ms) {
  CG_API_BEGIN
  CG_CHECK(ctx && ctx->ev0 && ms, "timer not started");
  CG_CUDA(cudaEventRecord(ctx->ev1, ctx->c.stream));
  CG_CUDA(cudaEventSynchronize(ctx->ev1));
  float f = 0.f;
  CG_CUDA(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = f;
  CG_API_END(ctx)
}

int cg_set_option(cg_context* ctx, const char* name, double value) {
  if (!ctx || !name) return 1;
  if (strcmp(name, "exact_reduction") == 0) {
    ctx->c.exact_reduction = value != 0.0;
    return 0;
  }
  if (strcmp(name, "tc_projection") == 0) {
    ctx->c.tc_projection = value != 0.0;
    return 0;
  }
  if (strcmp(name, "tc_eig_degree") == 0) {
    if (value < 2 || value > 64) return ::cg::fail(ctx, "tc_eig_degree must be in 2..64");
    ctx->c.tc_eig_degree = (int)value;
    return 0;
  }
  if (strcmp(name, "gram_chunk_slabs") == 0) {
    ctx->c.gram_chunk_slabs = value < 0 ? 0 : (int)value;
    return 0;
  }
  if (strcmp(name, "tc_eig_tol") == 0) {
    if (!(value > 0.0 && value < 1.0)) return ::cg::fail(ctx, "tc_eig_tol must be in (0, 1)");
    ctx->c.tc_eig_tol = value;
    return 0;
  }
  if (strcmp(name, "tc_eig_max_sweeps") == 0) {
    ctx->c.tc_eig_max_sweeps = (int)value;
    return 0;
  }
  if (strcmp(name, "max_batch_bytes") == 0) {
    ctx->c.max_batch_bytes = value < 1 ? 1 : (size_t)value;
    return 0;
  }
  if (strcmp(name, "knn_engine") == 0) {
    ctx->c.knn_engine = (int)value;
    return 0;
  }
  if (strcmp(name, "knn_ld32") == 0) {
    ctx->c.knn_ld32 = value != 0.0;
    return 0;
  }
  if (strcmp(name, "refresh_free_memory") == 0) {
    if (cudaSetDevice(ctx->c.device) != cudaSuccess) cudaGetLastError();
    ctx->c.refresh_free_bytes();
    return 0;
  }
  if (strcmp(name, "eig_cluster") == 0) {
    ctx->c.eig_cluster = (int)value;
    return 0;
  }
  if (strcmp(name, "knn_stagger") == 0) {
    ctx->c.knn_stagger = value < 0 ? 0 : (int)value;
    return 0;
  }
  if (strcmp(name, "knn_cand_cap") == 0) {
    ctx->c.knn_cand_cap = (int)value;
    return 0;
  }
  return ::cg::fail(ctx, "unknown option");
}

int cg_get_stat(cg_context* ctx, const char* name, double* value) {
  if (!ctx || !name || !value) return 1;
  if (strcmp(name, "scratch_pool_reserved_bytes") == 0 || strcmp(name, "panel_pool_reserved_bytes") == 0) {
    cudaMemPool_t pool = ctx->c.panel_pool;
    if (name[0] == 's' && cudaDeviceGetDefaultMemPool(&pool, ctx->c.device) != cudaSuccess)
      return ::cg::fail(ctx, "no default memory pool");
    uint64_t v = 0;
    if (!pool || cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &v) != cudaSuccess)
      return ::cg::fail(ctx, "memory pool query failed");
    *value = (double)v;
    return 0;
  }
  if (strcmp(name, "eig_redo_pairs") == 0) {
    *value = (double)ctx->c.eig_redo_pairs;
    return 0;
  }
  if (strcmp(name, "knn_redo_blocks") == 0) {
    *value = (double)ctx->c.knn_redo_blocks;
    return 0;
  }
  if (strcmp(name, "kernel_launches") == 0) {
    *value = (double)ctx->c.stats.kernels;
    return 0;
  }
  if (strcmp(name, "knn_scores") == 0) {  // scores scanned since cg_knn_timing(1): n_query * n_ref per direction
    *value = ctx->c.knn_scores;
    return 0;
  }
  return ::cg::fail(ctx, "unknown statistic");
}

int cg_profile(cg_context* ctx, int enable) {
  if (!ctx) return 1;
  ctx->c.profile = enable == 1;
  ctx->c.profile_host = enable == 2;
  ctx->c.stage_ms.clear();
  return 0;
}
int cg_profile_dump(cg_context* ctx, char* buf, int buf_len) {
  if (!ctx || !buf || buf_len <= 0) return 1;
  std::string s;
  for (auto& kv : ctx->c.stage_ms) {
    char t[128];
    snprintf(t, sizeof(t), "%s=%.3f;", kv.first.c_str(), kv.second);
    s += t;
  }
  snprintf(buf, (size_t)buf_len, "%s", s.c_str());
  return 0;
}

int cg_knn_timing(cg_context* ctx, int enable) {
  CG_API_BEGIN
  CG_CHECK(ctx, "NULL context");
  ctx->c.time_knn = enable != 0;
  ctx->c.knn_ms_accum = 0;
  ctx->c.knn_launches = 0;
  ctx->c.knn_candidates = 0;
  ctx->c.knn_scores = 0;
  ctx->c.knn_bytes = 0;
  CG_API_END(ctx)
}
int cg_knn_timing_read(cg_context* ctx, double* ms, uint64_t* launches, double* candidates, double* bytes) {
  CG_API_BEGIN
  CG_CHECK(ctx, "NULL context");
  if (ms) *ms = ctx->c.knn_ms_accum;
  if (launches) *launches = ctx->c.knn_launches;
  if (candidates) *candidates = ctx->c.knn_candidates;
  if (bytes) *bytes = ctx->c.knn_bytes;
  ctx->c.knn_ms_accum = 0;
  ctx->c.knn_launches = 0;
  ctx->c.knn_candidates = 0;
  ctx->c.knn_bytes = 0;
  CG_API_END(ctx)
}

void cg_params_default(cg_params* p) {
  if (!p) return;
  p->k = 15;
  p->k1 = 15;
  p->k_self = 10;
  p->ncomps = 40;
  p->space = CG_SPACE_PCA;
  p->matching = CG_MATCH_MNN;
  p->metric = CG_METRIC_ANGULAR;
  p->var_scale = 1;
  p->common_centering = 1;
  p->score_component_variance = 0;
  p->k_self_weight = 0.1;
  p->l2_sigma = 1e5;
  p->cor_base = 1.0;
  p->min_similarity = 1e-5;
  p->cpca_maxit = 500;
  p->cpca_tol = 1e-5;
}

}  // extern "C"

namespace {

struct HostPanel {
  DevBuf<double> src;
  DevBuf<float> tiled;
  KnnPanel panel;
};

void upload_panel(Context& c, const double* X, int n, int ld, int d, Metric metric, HostPanel& hp) {
  CG_CHECK(n >= 0 && d >= 1 && ld >= n, "bad matrix shape");
  std::vector<double> packed;
  const double* src = X;
  if (ld != n) {
    packed.resize((size_t)n * d);
    for (int col = 0; col < d; ++col) memcpy(&packed[(size_t)col * n], X + (size_t)col * ld, sizeof(double) * n);
    src = packed.data();
  }
  hp.src.upload(src, (size_t)n * d, c.stream);
  hp.tiled.alloc(KnnPanel::elems(n, d));
  hp.panel.tiled = hp.tiled.p;
  knn_prep_from_f64_colmajor(c, hp.src.p, n, d, n > 0 ? n : 1, metric, hp.panel);
  CG_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace

extern "C" {

int cg_cross_knn(cg_context* ctx, const double* A, int nA, int lda, const double* B, int nB, int ldb, int d, int k,
                 int metric, int32_t* idx_ab, float* dist_ab, int32_t* idx_ba, float* dist_ba) {
  CG_API_BEGIN
  CG_CHECK(ctx, "NULL context");
  CG_CHECK(metric == CG_METRIC_ANGULAR || metric == CG_METRIC_L2,
           "only the following distance metrics are currently supported: ['L2' 'angular']");
  CG_CHECK(k >= 1, "k must be >= 1");
  CG_CHECK(nA > 0 && nB > 0, "empty matrix");
  Context& c = ctx->c;
  CG_CUDA(cudaSetDevice(c.device));
  CG_BIND(ctx);
  HostPanel pa, pb;
  upload_panel(c, A, nA, lda, d, (Metric)metric, pa);
  upload_panel(c, B, nB, ldb, d, (Metric)metric, pb);
  std::vector<KnnProblem> probs;
  DevBuf<int> i_ab, i_ba;
  DevBuf<float> d_ab, d_ba;
  if (idx_ab || dist_ab) {
    i_ab.alloc((size_t)nA * k);
    d_ab.alloc((size_t)nA * k);
    probs.push_back(KnnProblem{pa.panel.tiled, pb.panel.tiled, nA, nB, i_ab.p, d_ab.p});
  }
  if (idx_ba || dist_ba) {
    i_ba.alloc((size_t)nB * k);
    d_ba.alloc((size_t)nB * k);
    probs.push_back(KnnProblem{pb.panel.tiled, pa.panel.tiled, nB, nA, i_ba.p, d_ba.p});
  }
  if (!probs.empty()) {
    DevBuf<KnnProblem> dp;
    dp.upload(probs.data(), probs.size(), c.stream);
    knn_launch(c, dp.p, probs.data(), (int)probs.size(), k, pa.panel.d_pad, d, (Metric)metric);
    if (idx_ab) i_ab.download(idx_ab, (size_t)nA * k, c.stream);
    if (dist_ab) d_ab.download(dist_ab, (size_t)nA * k, c.stream);
    if (idx_ba) i_ba.download(idx_ba, (size_t)nB * k, c.stream);
    if (dist_ba) d_ba.download(dist_ba, (size_t)nB * k, c.stream);
    CG_CUDA(cudaStreamSynchronize(c.stream));
  }
  CG_API_END(ctx)
}

int cg_self_knn(cg_context* ctx, const double* Z, int n, int ldz, int d, int k, int metric, int32_t* idx,
                float* dist) {
  CG_API_BEGIN
  CG_CHECK(ctx, "NULL context");
  CG_CHECK(metric == CG_METRIC_ANGULAR || metric == CG_METRIC_L2,
           "only the following distance metrics are currently supported: ['L2' 'angular']");
  CG_CHECK(k >= 1 && n > 0, "bad arguments");
  Context& c = ctx->c;
  CG_CUDA(cudaSetDevice(c.device));
  CG_BIND(ctx);
  HostPanel pz;
  upload_panel(c, Z, n, ldz, d, (Metric)metric, pz);
  DevBuf<int> di((size_t)n * k);
  DevBuf<float> dd((size_t)n * k);
  KnnProblem pb{pz.panel.tiled, pz.panel.tiled, n, n, di.p, dd.p};
  DevBuf<KnnProblem> dp;
  dp.upload(&pb, 1, c.stream);
  knn_launch(c, dp.p, &pb, 1, k, pz.panel.d_pad, d, (Metric)metric);
  if (idx) di.download(idx, (size_t)n * k, c.stream);
  if (dist) dd.download(dist, (size_t)n * k, c.stream);
  CG_CUDA(cudaStreamSynchronize(c.stream));
  CG_API_END(ctx)
}

int cg_neighbor_matrix(cg_context* ctx, const double* p1, int n1, int ld1, const double* p2, int n2, int ld2, int d,
                       int k, int k1, int matching, int metric, double l2_sigma, double cor_base,
                       double min_similarity, int64_t* nnz) {
  CG_API_BEGIN
  CG_CHECK(ctx, "NULL context");
  CG_CHECK(metric == CG_METRIC_ANGULAR || metric == CG_METRIC_L2,
           "only the following distance metrics are currently supported: ['L2' 'angular']");
  CG_CHECK(matching == CG_MATCH_MNN || matching == CG_MATCH_NN, "Unrecognized type of NN matching");
  CG_CHECK(k1 >= k && k >= 1, "k1 must be >= k");
  CG_CHECK(n1 > 0 && n2 > 0, "empty matrix");
  Context& c = ctx->c;
  CG_CUDA(cudaSetDevice(c.device));
  CG_BIND(ctx);
  HostPanel pa, pb;
  upload_panel(c, p1, n1, ld1, d, (Metric)metric, pa);
  upload_panel(c, p2, n2, ld2, d, (Metric)metric, pb);
  DevBuf<int> i_ab((size_t)n1 * k1), i_ba((size_t)n2 * k1);
  DevBuf<float> d_ab((size_t)n1 * k1), d_ba((size_t)n2 * k1);
  KnnProblem probs[2] = {{pa.panel.tiled, pb.panel.tiled, n1, n2, i_ab.p, d_ab.p},
                         {pb.panel.tiled, pa.panel.tiled, n2, n1, i_ba.p, d_ba.p}};
  DevBuf<KnnProblem> dp;
  dp.upload(probs, 2, c.stream);
  knn_launch(c, dp.p, probs, 2, k1, pa.panel.d_pad, d, (Metric)metric);
  MnnProblem mp{i_ab.p, d_ab.p, i_ba.p, d_ba.p, nullptr, nullptr, n1, n2, 0, 0};
  SimParams sp{metric, matching, cor_base, l2_sigma, min_similarity};
  ctx->coo.n = 0;
  match_append_edges(c, &mp, 1, k, k1, sp, ctx->coo, nullptr);
  if (nnz) *nnz = (int64_t)ctx->coo.n;
  CG_API_END(ctx)
}

int cg_fetch_coo(cg_context* ctx, int32_t* i, int32_t* j, double* x) {
  CG_API_BEGIN
  CG_BIND(ctx);
  CG_CHECK(ctx, "NULL context");
  Context& c = ctx->c;
  const size_t n = ctx->coo.n;
  if (n) {
    if (i) ctx->coo.from.download(i, n, c.stream);
    if (j) ctx->coo.to.download(j, n, c.stream);
    if (x) ctx->coo.w.download(x, n, c.stream);
    CG_CUDA(cudaStreamSynchronize(c.stream));
  }
  CG_API_END(ctx)
}

// ------------------------------------------------------------------ edge tables
int64_t cg_edges_count(const cg_edges* e) { return e ? (int64_t)e->t.n : 0; }

int cg_edges_fetch(cg_context* ctx, const cg_edges* e, int32_t* from, int32_t* to, double* w, int32_t* type) {
  CG_API_BEGIN
  CG_BIND(ctx);
  CG_CHECK(ctx && e, "NULL argument");
  const size_t n = e->t.n;
  if (n) {
    if (from) e->t.from.download(from, n, ctx->c.stream);
    if (to) e->t.to.download(to, n, ctx->c.stream);
    if (w) e->t.w.download(w, n, ctx->c.stream);
    if (type) e->t.type.download(type, n, ctx->c.stream);
    CG_CUDA(cudaStreamSynchronize(ctx->c.stream));
  }
  CG_API_END(ctx)
}

namespace {
__global__ void count_type_kernel(const int* __restrict__ type, long long n, int want, unsigned long long* out) {
  unsigned long long c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += type[i] == want ? 1 : 0;
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
}  // namespace

int64_t cg_edges_count_type(cg_context* ctx, const cg_edges* e, int32_t type) {
  if (!ctx || !e || e->t.n == 0) return 0;
  try {
    CG_BIND(ctx);
    DevBuf<unsigned long long> d(1);
    d.zero(ctx->c.stream);
    count_type_kernel<<<256, 256, 0, ctx->c.stream>>>(e->t.type.p, (long long)e->t.n, type, d.p);
    ctx->c.stats.kernels++;
    unsigned long long h = 0;
    d.download(&h, 1, ctx->c.stream);
    CG_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return (int64_t)h;
  } catch (const std::exception& ex) {
    fail(ctx, ex);
    return -1;
  }
}

int cg_edges_reserve(cg_context* ctx, cg_edges** e, int64_t n_total) {
  CG_API_BEGIN
  CG_BIND(ctx);
  CG_CHECK(ctx && e && n_total >= 0, "bad argument");
  if (!*e) *e = new cg_edges();
  (*e)->t.reserve((size_t)n_total, ctx->c.stream);
  CG_API_END(ctx)
}

int cg_edges_set_count(cg_edges* e, int64_t n) {
  if (!e || n < 0 || (size_t)n > e->t.from.n) return 1;
  e->t.n = (size_t)n;
  return 0;
}

int cg_edges_append_host(cg_context* ctx, cg_edges** e, const int32_t* from, const int32_t* to, const double* w,
                         const int32_t* type, int64_t n) {
  CG_API_BEGIN
  CG_BIND(ctx);
  CG_CHECK(ctx && e && n >= 0, "bad argument");
  if (!*e) *e = new cg_edges();
  EdgeTable& t = (*e)->t;
  t.reserve(t.n + (size_t)n, ctx->c.stream);
  if (n) {
    cudaStream_t s = ctx->c.stream;
    CG_CUDA(cudaMemcpyAsync(t.from.p + t.n, from, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CG_CUDA(cudaMemcpyAsync(t.to.p + t.n, to, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CG_CUDA(cudaMemcpyAsync(t.w.p + t.n, w, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    CG_CUDA(cudaMemcpyAsync(t.type.p + t.n, type, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CG_CUDA(cudaStreamSynchronize(s));
    t.n += (size_t)n;
  }
  CG_API_END(ctx)
}

int cg_edges_append_device(cg_context* ctx, cg_edges** e, const void* from, const void* to, const void* w,
                           const void* type, int64_t n) {
  CG_API_BEGIN
  CG_BIND(ctx);
  CG_CHECK(ctx && e && n >= 0, "bad argument");
  if (!*e) *e = new cg_edges();
  EdgeTable& t = (*e)->t;
  t.reserve(t.n + (size_t)n, ctx->c.stream);
  if (n) {
    cudaStream_t s = ctx->c.stream;
    CG_CUDA(cudaMemcpyAsync(t.from.p + t.n, from, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    CG_CUDA(cudaMemcpyAsync(t.to.p + t.n, to, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    CG_CUDA(cudaMemcpyAsync(t.w.p + t.n, w, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    CG_CUDA(cudaMemcpyAsync(t.type.p + t.n, type, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    CG_CUDA(cudaStreamSynchronize(s));
    t.n += (size_t)n;
  }
  CG_API_END(ctx)
}

int cg_edges_device_ptrs(const cg_edges* e, void** from, void** to, void** w, void** type) {
  if (!e) return 1;
  if (from) *from = e->t.from.p;
  if (to) *to = e->t.to.p;
  if (w) *w = e->t.w.p;
  if (type) *type = e->t.type.p;
  return 0;
}

void cg_edges_free(cg_edges* e) { delete e; }

// ------------------------------------------------------------------ graph
int cg_assemble_graph(cg_context* ctx, const cg_edges* e, int32_t n_cells_total, cg_graph** out) {
  cg_graph* g = nullptr;
  try {
    CG_CHECK(ctx && e && out, "NULL argument");
    CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
    g = new cg_graph();
    {
      StageScope sc(ctx->c, "assemble_graph");
      assemble_graph(ctx->c, e->t, n_cells_total, g->g);
    }
    *out = g;
    return 0;
  } catch (const std::exception& ex) {
    delete g;
    return fail(ctx, ex);
  }
}

int cg_graph_sizes(const cg_graph* g, int32_t* n_vertices, int64_t* n_edges) {
  if (!g) return 1;
  if (n_vertices) *n_vertices = g->g.n_vertices;
  if (n_edges) *n_edges = g->g.n_edges;
  return 0;
}

int cg_graph_fetch(cg_context* ctx, const cg_graph* g, int32_t* vertices, int32_t* from, int32_t* to, double* weight,
                   int32_t* type) {
  CG_API_BEGIN
  CG_BIND(ctx);
  CG_CHECK(ctx && g, "NULL argument");
  cudaStream_t s = ctx->c.stream;
  const size_t nv = (size_t)g->g.n_vertices, ne = (size_t)g->g.n_edges;
  if (vertices && nv) g->g.vertices.download(vertices, nv, s);
  if (ne) {
    if (from) g->g.from.download(from, ne, s);
    if (to) g->g.to.download(to, ne, s);
    if (weight) g->g.weight.download(weight, ne, s);
    if (type) g->g.type.download(type, ne, s);
  }
  CG_CUDA(cudaStreamSynchronize(s));
  CG_API_END(ctx)
}

int cg_graph_fetch_adjacency(cg_context* ctx, const cg_graph* g, int64_t* p, int32_t* i, double* x) {
  CG_API_BEGIN
  CG_BIND(ctx);
  CG_CHECK(ctx && g, "NULL argument");
  cudaStream_t s = ctx->c.stream;
  const size_t nv = (size_t)g->g.n_vertices, ne2 = (size_t)(2 * g->g.n_edges);
  if (nv == 0) {
    if (p) p[0] = 0;
  } else {
    if (p) CG_CUDA(cudaMemcpyAsync(p, g->g.adj_p.p, (nv + 1) * sizeof(long long), cudaMemcpyDeviceToHost, s));
    if (i && ne2) g->g.adj_i.download(i, ne2, s);
    if (x && ne2) g->g.adj_x.download(x, ne2, s);
  }
  CG_CUDA(cudaStreamSynchronize(s));
  CG_API_END(ctx)
}

void cg_graph_free(cg_graph* g) { delete g; }

}  // extern "C"

// Shared host-side plumbing of libconosgpu: error capture (never abort, never
// throw across the C ABI), the opaque context, and a small device-buffer RAII.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <chrono>
#include <map>
#include <mutex>
#include <set>
#include <vector>

namespace cg {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CG_CUDA(expr)                                                                            \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      char _buf[512];                                                                            \
      snprintf(_buf, sizeof(_buf), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
               __LINE__, cudaGetErrorString(_e));                                                \
      throw ::cg::Error(_buf);                                                                   \
    }                                                                                            \
  } while (0)

#define CG_CHECK(cond, msg)                                    \
  do {                                                         \
    if (!(cond)) throw ::cg::Error(std::string("conosgpu: ") + (msg)); \
  } while (0)

// Stream-ordered allocation: every C-ABI entry point binds the context's stream for the calling thread, and
// DevBuf then allocates from / frees to the device's caching memory pool in stream order (no device-wide
// synchronisation, no cudaMalloc/cudaFree latency inside buildGraph).  Without a bound stream (the stand-alone
// probes) it falls back to cudaMalloc / cudaFree.
inline cudaStream_t& bound_stream() {
  static thread_local cudaStream_t s = nullptr;
  return s;
}
// Optional second pool for long-lived buffers (the resident samples of a panel): kept apart from the per-call
// scratch so that re-uploading a panel every call does not fragment the scratch pool.
inline cudaMemPool_t& bound_pool() {
  static thread_local cudaMemPool_t p = nullptr;
  return p;
}
struct PoolBinding {
  cudaMemPool_t prev;
  explicit PoolBinding(cudaMemPool_t p) : prev(bound_pool()) { bound_pool() = p; }
  ~PoolBinding() { bound_pool() = prev; }
};
struct StreamBinding {
  cudaStream_t prev;
  explicit StreamBinding(cudaStream_t s) : prev(bound_stream()) { bound_stream() = s; }
  ~StreamBinding() { bound_stream() = prev; }
};

// Streams owned by live contexts.  A device buffer remembers the stream it was allocated on; when its owner (panel,
// edge table, graph) outlives the context -- a host language that finalises objects in arbitrary order at exit --
// the block is returned with cudaFree instead of cudaFreeAsync on a destroyed stream (undefined behaviour).
// The registry is never destroyed: destructors may run during static destruction.
struct StreamRegistry {
  std::mutex m;
  std::set<cudaStream_t> live;
};
inline StreamRegistry& stream_registry() {
  static StreamRegistry* r = new StreamRegistry();
  return *r;
}
inline void stream_register(cudaStream_t s) {
  StreamRegistry& r = stream_registry();
  std::lock_guard<std::mutex> g(r.m);
  r.live.insert(s);
}
inline void stream_unregister(cudaStream_t s) {
  StreamRegistry& r = stream_registry();
  std::lock_guard<std::mutex> g(r.m);
  r.live.erase(s);
}
inline bool stream_alive(cudaStream_t s) {
  StreamRegistry& r = stream_registry();
  std::lock_guard<std::mutex> g(r.m);
  return r.live.count(s) != 0;
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t owner = nullptr;  // stream the block was allocated on (nullptr: cudaMalloc)
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), owner(o.owner) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; owner = o.owner; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      // errors are dropped on purpose: at process exit the runtime may already be unloading (cudaErrorCudartUnloading)
      cudaError_t e = (owner && stream_alive(owner)) ? cudaFreeAsync(p, owner) : cudaFree(p);
      if (e != cudaSuccess) cudaGetLastError();
    }
    p = nullptr; n = 0;
  }
  void alloc(size_t count) {
    release();
    if (count == 0) return;
    owner = bound_stream();
    if (owner && bound_pool()) CG_CUDA(cudaMallocFromPoolAsync(&p, count * sizeof(T), bound_pool(), owner));
    else if (owner) CG_CUDA(cudaMallocAsync(&p, count * sizeof(T), owner));
    else CG_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  // grow-only (keeps capacity between calls; contents are NOT preserved)
  void ensure(size_t count) {
    if (count > n) alloc(count);
  }
  void zero(cudaStream_t s) {
    if (n) CG_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    ensure(count);
    if (count) CG_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void download(T* h, size_t count, cudaStream_t s) const {
    if (count) CG_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

// Page-locked host arena for results that leave the device asynchronously (the rotations of the pairs cache):
// blocks are kept between calls (grow-only), reset() rewinds the cursor.  A pointer stays valid until the next reset.
struct PinnedArena {
  std::vector<char*> blocks;
  std::vector<size_t> caps, used;
  size_t first_open = 0;  // blocks before this one are full for requests of the smallest size seen (search start)
  void reset() {
    std::fill(used.begin(), used.end(), (size_t)0);
    first_open = 0;
  }
  // first fit over ALL blocks: a call that asks for the same sizes as the one before it finds them again whatever
  // order the blocks were created in (a forward-only cursor made the second atlas call pin 2.4 GB once more)
  void* take(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    while (first_open < blocks.size() && caps[first_open] - used[first_open] < 256) ++first_open;
    for (size_t b = first_open; b < blocks.size(); ++b)
      if (used[b] + bytes <= caps[b]) {
        void* r = blocks[b] + used[b];
        used[b] += bytes;
        return r;
      }
    const size_t cap = bytes > ((size_t)32 << 20) ? bytes : ((size_t)32 << 20);
    char* nb = nullptr;
    CG_CUDA(cudaHostAlloc(&nb, cap, cudaHostAllocDefault));
    blocks.push_back(nb);
    caps.push_back(cap);
    used.push_back(bytes);
    return nb;
  }
  void release() {
    for (char* b : blocks)
      if (cudaFreeHost(b) != cudaSuccess) cudaGetLastError();
    blocks.clear();
    caps.clear();
    used.clear();
    first_open = 0;
  }
  ~PinnedArena() { release(); }
};

// Per-context launch statistics (bench.py reports them as gpu_launches).
struct LaunchStats {
  uint64_t kernels = 0;
};

struct Comm;  // NCCL communicator of a multi-GPU run (comm.cu); nullptr for a single-GPU context

struct Context {
  int device = 0;
  Comm* comm = nullptr;
  int comm_world = 1;             // ranks of the job this context belongs to (they share the host's cores)
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::string last_error;
  LaunchStats stats;
  // optional per-kernel timing of the dominant kernel (cross-kNN), CUDA events on `stream`
  bool time_knn = false;
  double knn_ms_accum = 0.0;
  uint64_t knn_launches = 0;
  double knn_candidates = 0.0;  // sum over launches of n_a * n_b, ONE matrix per pair (both directions)
  double knn_scores = 0.0;      // scores the kernels scanned: n_query * n_ref per direction
  cudaEvent_t knn_ev0 = nullptr, knn_ev1 = nullptr;
  double knn_bytes = 0.0;       // algorithmic HBM bytes
  // kNN engine: 0 = auto (two-phase tensor-core kernel, streaming kernel for small problems and redo blocks),
  // 1 = streaming tensor-core kernel only, 2 = exact SIMT kernel only
  int knn_engine = 0;
  int knn_ld32 = 0;               // two-phase kernel: read the accumulators with two x32 loads instead of four x16
  int eig_cluster = 0;            // 0 = automatic, 1 / 4 = force the single-CTA / clustered eigen kernels (tests)
  int knn_stagger = 0;            // two-phase kernel: cycles the second column half of the epilogue starts late
  int knn_cand_cap = 64;          // candidate slots per (row, half) the two-phase kernel may use (tests shrink it)
  cudaMemPool_t panel_pool = nullptr;  // resident panel buffers (see PoolBinding)
  PinnedArena rot_arena;          // host copies of the rotations of the current cg_pairs_edges call
  uint64_t eig_redo_pairs = 0;    // pairs the tensor-core eigen solver handed to the FP64 driver
  uint64_t knn_redo_blocks = 0;   // 256-row blocks the streaming kernel had to recompute
  // 1: FP64 reductions (atomics Gram, explicit covariance, plain subspace iteration) instead of the tensor-core path
  bool exact_reduction = false;
  bool tc_projection = true;   // default mode: project through the tensor-core GEMM (false: FP64 SpMM everywhere)
  int tc_eig_degree = 10;      // tensor-core eigen solver: largest Chebyshev degree of a sweep
  int gram_chunk_slabs = 16;   // tensor-core Gram: cells per FP32 accumulation chunk / 64 (chunks are summed in FP64)
  double tc_eig_tol = 1e-5;    // its stopping rule: residual <= tol * theta_1 (irlba's default tolerance)
  int tc_eig_max_sweeps = 40;  // tensor-core eigen solver: sweeps before a problem is handed to the FP64 driver
  size_t max_batch_bytes = (size_t)48 << 30;  // scratch budget of one batch of pairs in cg_pairs_edges (tests shrink it)
  // optional per-stage wall-clock profile (stream synchronised at both ends of a stage)
  // Free device memory as of the last look.  cudaMemGetInfo is erratic on some hosts (tests/cuda/api_stall_probe.cu:
  // mean 0.13 ms, 1 % of the calls 1-100 ms, while launches / syncs / stream-ordered allocations never exceed 2 ms),
  // and the pools keep what they once reserved, so the hot calls use this estimate and only look again when a job
  // does not obviously fit (see cg_pairs_edges).
  size_t free_bytes_seen = 0;
  size_t refresh_free_bytes() {
    size_t f = 0, t = 0;
    if (cudaMemGetInfo(&f, &t) != cudaSuccess) {
      cudaGetLastError();
      f = (size_t)8 << 30;
    }
    free_bytes_seen = f;
    return f;
  }
  // budget for scratch of `want` bytes: the estimate when the job fits a third of it with room to spare, a fresh look
  // otherwise
  size_t free_bytes_for(size_t want) {
    if (free_bytes_seen == 0 || want > free_bytes_seen / 4) return refresh_free_bytes();
    return free_bytes_seen;
  }
  bool profile = false;
  bool profile_host = false;  // cg_profile(ctx, 2): host wall clock per stage WITHOUT synchronising (where does the host
                              // thread sit when a step is slow?); adds nothing to the step
  std::map<std::string, double> stage_ms;
};

void comm_destroy(Context& c);

struct StageScope {
  Context& c;
  const char* name;
  std::chrono::steady_clock::time_point t0;
  StageScope(Context& ctx, const char* n) : c(ctx), name(n) {
    if (c.profile) cudaStreamSynchronize(c.stream);
    if (c.profile || c.profile_host) t0 = std::chrono::steady_clock::now();
  }
  ~StageScope() {
    if (c.profile) cudaStreamSynchronize(c.stream);
    if (c.profile || c.profile_host)
      c.stage_ms[name] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
};

__host__ __device__ inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// host threads for the table-filling loops of a call: up to 8, fewer when several ranks of a job share the host
inline int host_fill_threads(const Context& c, size_t items, size_t min_items) {
  if (items < min_items) return 1;
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 8;
  const int per = (int)(hw / (unsigned)(c.comm_world > 1 ? c.comm_world : 1));
  return per < 1 ? 1 : (per > 8 ? 8 : per);
}

#ifdef CG_COUNT_SYNCS  // measurement build: which call sites make the host wait for the stream, and how often
struct SyncCounter {
  std::mutex mu;
  std::map<std::string, long long> sites;
  ~SyncCounter() {
    if (!getenv("CG_COUNT_SYNCS_DUMP")) return;
    long long total = 0;
    for (auto& kv : sites) total += kv.second;
    fprintf(stderr, "[syncs] total %lld\n", total);
    for (auto& kv : sites) fprintf(stderr, "[syncs] %6lld  %s\n", kv.second, kv.first.c_str());
  }
};
inline SyncCounter& sync_counter() {
  static SyncCounter c;
  return c;
}
inline cudaError_t counted_stream_sync(cudaStream_t s, const char* file, int line) {
  {
    SyncCounter& c = sync_counter();
    std::lock_guard<std::mutex> lk(c.mu);
    const char* base = strrchr(file, '/');
    c.sites[std::string(base ? base + 1 : file) + ":" + std::to_string(line)]++;
  }
  return cudaStreamSynchronize(s);
}
#define cudaStreamSynchronize(s) ::cg::counted_stream_sync((s), __FILE__, __LINE__)
#endif

}  // namespace cg

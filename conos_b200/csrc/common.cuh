// Shared host-side plumbing of libconosgpu: error capture (never abort, never
// throw across the C ABI), the opaque context, and a small device-buffer RAII.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
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

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  void alloc(size_t count) {
    release();
    if (count == 0) return;
    CG_CUDA(cudaMalloc(&p, count * sizeof(T)));
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

// Per-context launch statistics (bench.py reports them as gpu_launches).
struct LaunchStats {
  uint64_t kernels = 0;
};

struct Context {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::string last_error;
  LaunchStats stats;
  // optional per-kernel timing of the dominant kernel (cross-kNN), CUDA events on `stream`
  bool time_knn = false;
  double knn_ms_accum = 0.0;
  uint64_t knn_launches = 0;
  double knn_candidates = 0.0;  // sum over launches of n_query * n_ref
  double knn_bytes = 0.0;       // algorithmic HBM bytes
};

inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

}  // namespace cg

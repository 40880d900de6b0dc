// Which CUDA runtime calls stall on this box?  Wall time of 30 s of back-to-back calls of each kind, with the
// distribution of the slow ones.  (Background: occasional 50-500 ms steps in bench.py whose stage scopes are all normal.)
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tests/cuda/api_stall_probe.cu -o build/api_stall_probe
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

__global__ void tiny(int* p) { if (p) *p = 1; }

template <class F>
void run(const char* name, double seconds, F f) {
  using clk = std::chrono::steady_clock;
  std::vector<double> slow;
  long long n = 0;
  double worst = 0, total = 0;
  const auto t_end = clk::now() + std::chrono::duration<double>(seconds);
  while (clk::now() < t_end) {
    const auto t0 = clk::now();
    f();
    const double ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
    total += ms;
    worst = std::max(worst, ms);
    if (ms > 1.0) slow.push_back(ms);
    ++n;
  }
  std::sort(slow.begin(), slow.end());
  printf("%-28s calls %9lld  mean %8.4f ms  max %9.3f ms  calls > 1 ms: %zu", name, n, total / n, worst, slow.size());
  if (!slow.empty()) printf("  (median of those %.2f, sum %.1f ms)", slow[slow.size() / 2], [&] { double s = 0; for (double v : slow) s += v; return s; }());
  printf("\n");
}

__global__ void big_smem(int* p) {
  extern __shared__ int sm[];
  sm[threadIdx.x] = 1;
  if (p && threadIdx.x == 0) *p = sm[0];
}

int main(int argc, char** argv) {
  const double secs = argc > 1 ? atof(argv[1]) : 8.0;
  if (argc > 2) cudaSetDevice(atoi(argv[2]));   // several instances, one per GPU: do the calls of one process stall another's?
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaMemPool_t pool;
  cudaDeviceGetDefaultMemPool(&pool, 0);
  uint64_t thr = UINT64_MAX;
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  int* d = nullptr;
  cudaMalloc(&d, 4);
  run("cudaMemGetInfo", secs, [&] { size_t f, t; cudaMemGetInfo(&f, &t); });
  run("cudaFuncSetAttribute", secs, [&] { cudaFuncSetAttribute(big_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10); });
  run("cudaMallocAsync+FreeAsync 200MB", secs, [&] { void* p; cudaMallocAsync(&p, 200 << 20, st); cudaFreeAsync(p, st); });
  run("cudaEventCreate+Destroy", secs, [&] { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); cudaEventDestroy(e); });
  run("cudaMallocAsync+FreeAsync 1MB", secs, [&] { void* p; cudaMallocAsync(&p, 1 << 20, st); cudaFreeAsync(p, st); });
  run("launch + stream sync", secs, [&] { tiny<<<1, 32, 0, st>>>(d); cudaStreamSynchronize(st); });
  run("launch (async), sync per 64", secs, [&] { for (int i = 0; i < 64; ++i) tiny<<<1, 32, 0, st>>>(d); cudaStreamSynchronize(st); });
  run("cudaMemcpyAsync D2H 4B + sync", secs, [&] { int h; cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); });
  run("host only (clock reads)", secs, [&] { volatile int x = 0; for (int i = 0; i < 1000; ++i) x += i; });
  return 0;
}

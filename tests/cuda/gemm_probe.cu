// Standalone device check of the bf16x3 tcgen05 GEMM: C = A'B against an FP64 host product.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../conos_b200/csrc/gemm_tc.cuh"

using namespace cg;

static int run_case(Context& ctx, int K, int M, int N, bool check, int reps) {
  std::mt19937 rng(K + 3 * M + 7 * N);
  std::uniform_real_distribution<double> u(0.0, 2.0);
  std::vector<double> A((size_t)K * M), B((size_t)K * N);
  for (auto& v : A) v = (rng() % 10 < 3) ? u(rng) : 0.0;   // sparse-ish, non-negative like expression data
  for (auto& v : B) v = (rng() % 10 < 3) ? u(rng) - 0.5 : 0.0;
  DevBuf<double> dA(A.size()), dB(B.size());
  dA.upload(A.data(), A.size(), ctx.stream);
  dB.upload(B.data(), B.size(), ctx.stream);
  DevBuf<__nv_bfloat16> ah(TiledOperand::elems(M, K)), al(TiledOperand::elems(M, K)), bh(TiledOperand::elems(N, K)),
      bl(TiledOperand::elems(N, K));
  TiledOperand oa, ob;
  oa.hi = ah.p; oa.lo = al.p; ob.hi = bh.p; ob.lo = bl.p;
  tiled_from_colmajor_f64(ctx, dA.p, K, M, K, oa);
  tiled_from_colmajor_f64(ctx, dB.p, K, N, K, ob);
  DevBuf<float> dC((size_t)M * N);
  dC.zero(ctx.stream);
  GemmProblem gp{oa.hi, oa.lo, ob.hi, ob.lo, oa.rows_pad, ob.rows_pad, oa.k_pad, M, N, dC.p, M, 0};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0, ctx.stream);
    gemm_tn_bf16x3(ctx, &gp, 1);
    cudaEventRecord(e1, ctx.stream);
    CG_CUDA(cudaStreamSynchronize(ctx.stream));
    cudaEventElapsedTime(&ms, e0, e1);
  }
  std::vector<float> C((size_t)M * N);
  dC.download(C.data(), C.size(), ctx.stream);
  CG_CUDA(cudaStreamSynchronize(ctx.stream));
  double worst = 0, worst_abs = 0;
  if (check) {
    for (int n = 0; n < N; ++n)
      for (int m = 0; m < M; ++m) {
        double s = 0, sa = 0;
        for (int k = 0; k < K; ++k) { double t = A[(size_t)m * K + k] * B[(size_t)n * K + k]; s += t; sa += fabs(t); }
        double err = fabs((double)C[(size_t)n * M + m] - s);
        if (sa > 0 && err / sa > worst) worst = err / sa;
        if (err > worst_abs) worst_abs = err;
      }
  }
  printf("gemm K=%d M=%d N=%d: %.3f ms, %.1f TFLOP/s useful (x3 issued) | max err / sum|terms| = %.3e, max abs %.3e%s\n", K,
         M, N, ms, 2.0 * K * M * N / (ms * 1e-3) / 1e12, worst, worst_abs, check ? "" : " (unchecked)");
  return (check && worst > 6e-5) ? 1 : 0;  // 3 * 2^-16: bf16 hi+lo keeps 16 bits, lo*lo is dropped
}

int main(int argc, char** argv) {
  Context ctx;
  try {
    CG_CUDA(cudaSetDevice(0));
    cudaDeviceProp prop;
    CG_CUDA(cudaGetDeviceProperties(&prop, 0));
    ctx.sm_count = prop.multiProcessorCount;
    CG_CUDA(cudaStreamCreate(&ctx.stream));
    int rc = 0;
    if (argc == 4) {
      rc |= run_case(ctx, atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), false, 2);
      return rc;
    }
    rc |= run_case(ctx, 64, 128, 256, true, 1);
    rc |= run_case(ctx, 200, 130, 40, true, 1);
    rc |= run_case(ctx, 1000, 300, 513, true, 1);
    rc |= run_case(ctx, 2000, 2000, 224, false, 3);
    rc |= run_case(ctx, 50000, 2000, 2000, false, 3);
    printf(rc ? "GEMM_PROBE FAIL\n" : "GEMM_PROBE OK\n");
    return rc;
  } catch (const std::exception& e) {
    printf("GEMM_PROBE EXCEPTION: %s\n", e.what());
    return 2;
  }
}

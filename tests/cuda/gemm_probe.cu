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

// The Gram product X'X as ensure_grams (reduce.cu) runs it: symmetric tiles only, fp16 hi+lo operands, FP32 chunks of
// `chunk_slabs` x 64 cells summed in FP64; accuracy against an FP64 host product on 4000 sampled entries.
static int run_gram(Context& ctx, int K, int G, int chunk_slabs, bool fp16) {
  std::mt19937 rng(K + G);
  std::uniform_real_distribution<double> u(0.0, 3.0);
  std::vector<int> p(G + 1, 0), idx;
  std::vector<double> x;
  for (int g = 0; g < G; ++g) {      // CSC by gene, ~7.5 % dense, values like log-scale expression
    for (int c = 0; c < K; ++c)
      if (rng() % 1000 < 75) { idx.push_back(c); x.push_back(0.3 + u(rng)); }
    p[g + 1] = (int)idx.size();
  }
  // CSR by cell for the tile-wise operand builder
  std::vector<int> rp(K + 1, 0), ci(idx.size());
  std::vector<double> rx(idx.size());
  for (int c : idx) rp[c + 1]++;
  for (int c = 0; c < K; ++c) rp[c + 1] += rp[c];
  {
    std::vector<int> cur(rp.begin(), rp.end() - 1);
    for (int g = 0; g < G; ++g)
      for (int t = p[g]; t < p[g + 1]; ++t) { ci[cur[idx[t]]] = g; rx[cur[idx[t]]++] = x[t]; }
  }
  DevBuf<int> drp(rp.size()), dci(ci.size());
  DevBuf<double> drx(rx.size()), dC((size_t)G * G);
  drp.upload(rp.data(), rp.size(), ctx.stream);
  dci.upload(ci.data(), ci.size(), ctx.stream);
  drx.upload(rx.data(), rx.size(), ctx.stream);
  DevBuf<__nv_bfloat16> hi(TiledOperand::elems(G, K)), lo(TiledOperand::elems(G, K));
  TiledOperand op;
  op.hi = hi.p; op.lo = lo.p;
  tiled_from_csr(ctx, drp.p, dci.p, drx.p, K, G, op, fp16);
  GemmProblem gp{op.hi, op.lo, op.hi, op.lo, op.rows_pad, op.rows_pad, op.k_pad, G, G, dC.p, (long long)G, 1, 1};
  gp.chunk_slabs = chunk_slabs;
  gp.fp16 = fp16 ? 1 : 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  GemmPlan plan;
  gemm_plan(ctx, &gp, 1, plan);
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0, ctx.stream);
    gemm_run(ctx, plan);
    cudaEventRecord(e1, ctx.stream);
    CG_CUDA(cudaStreamSynchronize(ctx.stream));
    cudaEventElapsedTime(&ms, e0, e1);
  }
  std::vector<double> C((size_t)G * G);
  dC.download(C.data(), C.size(), ctx.stream);
  CG_CUDA(cudaStreamSynchronize(ctx.stream));
  double worst = 0;
  for (int t = 0; t < 4000; ++t) {
    int m = (int)(rng() % G), n = (int)(rng() % G);
    if (m > n) std::swap(m, n);   // upper triangle is what the kernel computes
    double s = 0;
    int a = p[m], b = p[n];
    while (a < p[m + 1] && b < p[n + 1]) {
      if (idx[a] == idx[b]) { s += x[a] * x[b]; ++a; ++b; }
      else if (idx[a] < idx[b]) ++a; else ++b;
    }
    const double err = fabs(C[(size_t)n * G + m] - s) / (s > 0 ? s : 1.0);
    if (err > worst) worst = err;
  }
  const double useful = 2.0 * K * (double)G * G / 2.0 * 1.125;  // 72 of 128 tiles touch the upper triangle
  printf("gram K=%d G=%d %s chunk=%d slabs: %.3f ms, %.0f TFLOP/s issued (x3, upper tiles), max relative error %.3e\n", K, G,
         fp16 ? "fp16x3" : "bf16x3", chunk_slabs, ms, 3.0 * useful / (ms * 1e-3) / 1e12, worst);
  // one FP32 accumulation over 50 000 cells: ~1e-4; 4096-cell chunks: ~1e-5; 1024-cell chunks: ~2e-6
  return worst > (chunk_slabs == 0 ? 3e-4 : (chunk_slabs > 16 ? 3e-5 : 6e-6)) ? 1 : 0;
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
    rc |= run_gram(ctx, 50000, 2000, 0, false);   // round 1: bf16 pairs, one FP32 accumulation over all cells
    rc |= run_gram(ctx, 50000, 2000, 0, true);
    rc |= run_gram(ctx, 50000, 2000, 64, true);
    rc |= run_gram(ctx, 50000, 2000, 16, true);   // the default of the library
    printf(rc ? "GEMM_PROBE FAIL\n" : "GEMM_PROBE OK\n");
    return rc;
  } catch (const std::exception& e) {
    printf("GEMM_PROBE EXCEPTION: %s\n", e.what());
    return 2;
  }
}

// Standalone device check of the kNN kernels (no Python): tensor-core kernel vs SIMT kernel vs a host
// brute force with the same canonical fp32 fma chain. Also times both kernels on a larger problem.
// Usage: knn_probe [n_query n_ref d k]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../conos_b200/csrc/knn.cuh"

using namespace cg;

static void host_knn(const std::vector<float>& Q, int nq, const std::vector<float>& R, int nr, int d, int k,
                     std::vector<int>& idx, std::vector<float>& dist) {
  idx.assign((size_t)nq * k, -1);
  dist.assign((size_t)nq * k, INFINITY);
  for (int i = 0; i < nq; ++i) {
    float* ld = &dist[(size_t)i * k];
    int* li = &idx[(size_t)i * k];
    for (int j = 0; j < nr; ++j) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < d; ++c) a[c & 3] = fmaf(Q[(size_t)i * d + c], R[(size_t)j * d + c], a[c & 3]);
      float dd = 1.0f - ((a[0] + a[1]) + (a[2] + a[3]));
      if (dd < ld[k - 1]) {
        int p = k - 1;
        while (p > 0 && ld[p - 1] > dd) { ld[p] = ld[p - 1]; li[p] = li[p - 1]; --p; }
        ld[p] = dd; li[p] = j;
      }
    }
  }
}

static void normalise(std::vector<float>& X, int n, int d) {
  for (int i = 0; i < n; ++i) {
    float ss = 0.f;
    for (int c = 0; c < d; ++c) ss = fmaf(X[(size_t)i * d + c], X[(size_t)i * d + c], ss);
    float inv = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    for (int c = 0; c < d; ++c) X[(size_t)i * d + c] *= inv;
  }
}

static int run_case(Context& ctx, int nq, int nr, int d, int k, bool check_host, int reps) {
  std::mt19937 rng(1234 + nq + 7 * nr + 13 * d);
  std::normal_distribution<float> g(0.f, 1.f);
  // 20-component mixture so that neighbourhoods are non-trivial
  std::vector<float> cent(20 * d);
  for (auto& v : cent) v = 3.f * g(rng);
  auto fill = [&](std::vector<float>& X, int n) {
    X.resize((size_t)n * d);
    for (int i = 0; i < n; ++i) {
      int c = rng() % 20;
      for (int j = 0; j < d; ++j) X[(size_t)i * d + j] = cent[c * d + j] + g(rng);
    }
  };
  std::vector<float> Q, R;
  fill(Q, nq);
  fill(R, nr);
  // duplicate a few reference rows to create exact distance ties
  for (int i = 0; i + 5 < nr && i < 50; i += 5)
    for (int j = 0; j < d; ++j) R[(size_t)(i + 1) * d + j] = R[(size_t)i * d + j];

  DevBuf<float> dQ((size_t)nq * d), dR((size_t)nr * d);
  dQ.upload(Q.data(), Q.size(), ctx.stream);
  dR.upload(R.data(), R.size(), ctx.stream);
  DevBuf<float> tq(KnnPanel::elems(nq, d)), tr(KnnPanel::elems(nr, d));
  KnnPanel pq, pr;
  pq.tiled = tq.p;
  pr.tiled = tr.p;
  knn_prep_from_f32_rowmajor(ctx, dQ.p, nq, d, d, METRIC_ANGULAR, pq);
  knn_prep_from_f32_rowmajor(ctx, dR.p, nr, d, d, METRIC_ANGULAR, pr);

  DevBuf<int> i_s((size_t)nq * k), i_t((size_t)nq * k);
  DevBuf<float> d_s((size_t)nq * k), d_t((size_t)nq * k);
  CG_CUDA(cudaMemsetAsync(i_t.p, 0xff, i_t.n * 4, ctx.stream));
  CG_CUDA(cudaMemsetAsync(d_t.p, 0xff, d_t.n * 4, ctx.stream));
  KnnProblem ps{pq.tiled, pr.tiled, nq, nr, i_s.p, d_s.p};
  KnnProblem pt{pq.tiled, pr.tiled, nq, nr, i_t.p, d_t.p};
  DevBuf<KnnProblem> dps(1), dpt(1);
  dps.upload(&ps, 1, ctx.stream);
  dpt.upload(&pt, 1, ctx.stream);

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms_s = 0, ms_t = 0, ms_v7 = 0;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0, ctx.stream);
    knn_simt_launch(ctx, dps.p, &ps, 1, k, pq.d_pad, d, METRIC_ANGULAR);
    cudaEventRecord(e1, ctx.stream);
    CG_CUDA(cudaStreamSynchronize(ctx.stream));
    cudaEventElapsedTime(&ms_s, e0, e1);
  }
  // streaming tensor-core kernel alone (engine 1), then the automatic engine (two-phase kernel + redo blocks)
  ctx.knn_engine = 1;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0, ctx.stream);
    knn_tc_launch(ctx, dpt.p, &pt, 1, k, pq.d_pad, d);
    cudaEventRecord(e1, ctx.stream);
    CG_CUDA(cudaStreamSynchronize(ctx.stream));
    cudaEventElapsedTime(&ms_v7, e0, e1);
  }
  {
    std::vector<int> h7((size_t)nq * k);
    i_t.download(h7.data(), h7.size(), ctx.stream);
    CG_CUDA(cudaStreamSynchronize(ctx.stream));
    CG_CUDA(cudaMemsetAsync(i_t.p, 0xff, i_t.n * 4, ctx.stream));
    CG_CUDA(cudaMemsetAsync(d_t.p, 0xff, d_t.n * 4, ctx.stream));
    std::vector<int> hs0((size_t)nq * k);
    i_s.download(hs0.data(), hs0.size(), ctx.stream);
    CG_CUDA(cudaStreamSynchronize(ctx.stream));
    size_t bad7 = 0;
    for (size_t i = 0; i < hs0.size(); ++i) bad7 += hs0[i] != h7[i];
    if (bad7) printf("  streaming kernel != simt: %zu\n", bad7);
  }
  ctx.knn_engine = 0;
  const uint64_t redo0 = ctx.knn_redo_blocks;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0, ctx.stream);
    knn_tc_launch(ctx, dpt.p, &pt, 1, k, pq.d_pad, d);
    cudaEventRecord(e1, ctx.stream);
    CG_CUDA(cudaStreamSynchronize(ctx.stream));
    cudaEventElapsedTime(&ms_t, e0, e1);
  }
  std::vector<int> hs((size_t)nq * k), ht((size_t)nq * k);
  std::vector<float> hds((size_t)nq * k), hdt((size_t)nq * k);
  i_s.download(hs.data(), hs.size(), ctx.stream);
  i_t.download(ht.data(), ht.size(), ctx.stream);
  d_s.download(hds.data(), hds.size(), ctx.stream);
  d_t.download(hdt.data(), hdt.size(), ctx.stream);
  CG_CUDA(cudaStreamSynchronize(ctx.stream));

  size_t bad_ts = 0, bad_sh = 0, bad_d = 0;
  for (size_t i = 0; i < hs.size(); ++i) {
    if (hs[i] != ht[i]) bad_ts++;
    if (memcmp(&hds[i], &hdt[i], 4) != 0) bad_d++;
  }
  if (check_host) {
    normalise(Q, nq, d);
    normalise(R, nr, d);
    std::vector<int> hi;
    std::vector<float> hd;
    host_knn(Q, nq, R, nr, d, k, hi, hd);
    for (size_t i = 0; i < hs.size(); ++i)
      if (hs[i] != hi[i] || memcmp(&hds[i], &hd[i], 4) != 0) bad_sh++;
  }
  double cand = (double)nq * nr;
  printf("case nq=%d nr=%d d=%d k=%d | simt %.3f ms (%.2e cand/s) | streaming tc %.3f ms | two-phase tc %.3f ms (%.2e "
         "cand/s, %.1f TFLOP/s, redo blocks %llu) | tc!=simt idx %zu dist %zu | simt!=host %zu%s\n",
         nq, nr, d, k, ms_s, cand / (ms_s * 1e-3), ms_v7, ms_t, cand / (ms_t * 1e-3),
         2.0 * cand * d / (ms_t * 1e-3) / 1e12, (unsigned long long)((ctx.knn_redo_blocks - redo0) / (reps ? reps : 1)),
         bad_ts, bad_d, bad_sh, check_host ? "" : " (host check skipped)");
  if (bad_ts && nq * k < 2000000) {
    int shown = 0;
    for (size_t i = 0; i < hs.size() && shown < 8; ++i)
      if (hs[i] != ht[i]) {
        printf("  row %zu rank %zu: simt (%d, %.9g) tc (%d, %.9g)\n", i / k, i % k, hs[i], hds[i], ht[i], hdt[i]);
        shown++;
      }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return (bad_ts || bad_sh || bad_d) ? 1 : 0;
}

int main(int argc, char** argv) {
  Context ctx;
  try {
    CG_CUDA(cudaSetDevice(0));
    cudaDeviceProp prop;
    CG_CUDA(cudaGetDeviceProperties(&prop, 0));
    ctx.sm_count = prop.multiProcessorCount;
    CG_CUDA(cudaStreamCreate(&ctx.stream));
    {
      cudaMemPool_t pool;
      CG_CUDA(cudaDeviceGetDefaultMemPool(&pool, 0));
      unsigned long long thr = ~0ull;
      CG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    StreamBinding sb(ctx.stream);  // DevBuf allocations come from the stream-ordered pool
    printf("device %s, %d SMs\n", prop.name, ctx.sm_count);
    int rc = 0;
    if (argc == 5) {
      rc |= run_case(ctx, atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), false, 3);
    } else {
      rc |= run_case(ctx, 25, 34, 30, 15, true, 1);      // config #1 sized
      rc |= run_case(ctx, 300, 1000, 30, 15, true, 1);
      rc |= run_case(ctx, 1000, 777, 50, 30, true, 1);   // d_pad = 64
      rc |= run_case(ctx, 513, 20, 30, 30, true, 1);     // fewer references than k
      rc |= run_case(ctx, 10000, 10000, 30, 15, false, 3);
      rc |= run_case(ctx, 50000, 50000, 30, 30, false, 3);
      rc |= run_case(ctx, 20000, 20000, 50, 15, false, 3);
      rc |= run_case(ctx, 3000, 5000, 32, 15, false, 1);   // d == d_pad: no spare dimension, compare path
      rc |= run_case(ctx, 3000, 5000, 64, 30, false, 1);
    }
    printf(rc ? "KNN_PROBE FAIL\n" : "KNN_PROBE OK\n");
    return rc;
  } catch (const std::exception& e) {
    printf("KNN_PROBE EXCEPTION: %s\n", e.what());
    return 2;
  }
}

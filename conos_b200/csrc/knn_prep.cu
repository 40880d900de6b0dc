// K1 prep_rows: cast to fp32, (angular) L2-normalise with the canonical fma chain, and write the
// core-matrix tiled operand panel and its fp16 copy (see knn.cuh). HBM-bound: reads 8 or 4 B/element,
// writes 8 B/element.
#include <cuda_fp16.h>

#include "knn.cuh"

namespace cg {
namespace {

template <class SrcT, bool COLMAJOR>
__global__ void knn_prep_kernel(const SrcT* __restrict__ src, int n, int d, int ld, int d_pad, int n_pad,
                                int metric, float* __restrict__ tiled) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_pad) return;
  const int KC = d_pad >> 2;
  float* dst = tiled + (size_t)(row >> 3) * (size_t)(8 * d_pad) + (size_t)(row & 7) * 4;
  __half* dst16 = reinterpret_cast<__half*>(tiled + (size_t)n_pad * d_pad);  // fp16 copy (knn_tiled16_offset)
  float* dstrm = tiled + (size_t)n_pad * d_pad * 3 / 2 + (size_t)row * d_pad;  // row-major exact copy
  const bool fold = d < d_pad;  // spare last column: -1 (threshold dimension of the tensor-core kernel)
  if (row >= n) {
    for (int c = 0; c < KC; ++c) {
      const float4 z = make_float4(0.f, 0.f, 0.f, (fold && c == KC - 1) ? -1.0f : 0.f);
      *reinterpret_cast<float4*>(dst + c * 32) = z;
      *reinterpret_cast<float4*>(dstrm + c * 4) = z;
      __half2* h = reinterpret_cast<__half2*>(dst16 + knn_tiled16_offset(row, c * 4, d_pad));
      h[0] = __floats2half2_rn(z.x, z.y);
      h[1] = __floats2half2_rn(z.z, z.w);
    }
    return;
  }
  float inv = 1.0f;
  if (metric == METRIC_ANGULAR) {
    float ss = 0.0f;
    for (int i = 0; i < d; ++i) {
      float x = COLMAJOR ? (float)src[(size_t)i * ld + row] : (float)src[(size_t)row * ld + i];
      ss = fmaf(x, x, ss);
    }
    inv = ss > 0.0f ? 1.0f / sqrtf(ss) : 0.0f;
  }
  for (int c = 0; c < KC; ++c) {
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int i = c * 4 + e;
      float x = 0.0f;
      if (i < d) x = COLMAJOR ? (float)src[(size_t)i * ld + row] : (float)src[(size_t)row * ld + i];
      v[e] = x * inv;
      if (fold && i == d_pad - 1) v[e] = -1.0f;
    }
    *reinterpret_cast<float4*>(dst + c * 32) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dstrm + c * 4) = make_float4(v[0], v[1], v[2], v[3]);
    __half2* h = reinterpret_cast<__half2*>(dst16 + knn_tiled16_offset(row, c * 4, d_pad));
    h[0] = __floats2half2_rn(v[0], v[1]);
    h[1] = __floats2half2_rn(v[2], v[3]);
  }
}

template <class SrcT, bool COLMAJOR>
void prep(Context& ctx, const SrcT* d_src, int n, int d, int ld, Metric metric, KnnPanel& out) {
  out.n = n;
  out.d = d;
  CG_CHECK(d >= 1 && d <= KNN_MAX_D, "kNN panels hold 1..128 columns");
  out.d_pad = knn_d_pad(d);
  out.n_pad = (int)round_up(n, KNN_ROW_PAD);
  CG_CHECK(out.tiled != nullptr, "panel storage missing");
  int threads = 128;
  int blocks = (out.n_pad + threads - 1) / threads;
  knn_prep_kernel<SrcT, COLMAJOR><<<blocks, threads, 0, ctx.stream>>>(d_src, n, d, ld, out.d_pad, out.n_pad,
                                                                     (int)metric, out.tiled);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

}  // namespace

void knn_prep_from_f64_colmajor(Context& ctx, const double* d_src, int n, int d, int ld, Metric metric,
                                KnnPanel& out) {
  prep<double, true>(ctx, d_src, n, d, ld, metric, out);
}
void knn_prep_from_f32_rowmajor(Context& ctx, const float* d_src, int n, int d, int ld, Metric metric,
                                KnnPanel& out) {
  prep<float, false>(ctx, d_src, n, d, ld, metric, out);
}

}  // namespace cg

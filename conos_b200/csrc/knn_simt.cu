// Exact SIMT kNN: one warp per query row, lanes stride over the reference rows, canonical fp32
// fma chain per (query, reference), warp-cooperative sorted insertion into a shared-memory list.
// Serves every (k, metric) the tensor-core kernel does not (k > 32, metric L2, tiny panels) and is the
// on-device cross-check for it. Bound: L2/HBM reads of the reference panel (4*d_pad bytes per pair).
#include "knn.cuh"

namespace cg {
namespace {

constexpr int SIMT_WARPS = 8;

template <int D_PAD>
__global__ void __launch_bounds__(SIMT_WARPS * 32) knn_simt_kernel(const KnnProblem* __restrict__ probs, int k,
                                                                   int metric, int d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const KnnProblem pb = probs[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qrow = blockIdx.x * SIMT_WARPS + warp;
  if (qrow >= pb.nq) return;
  float* list_d = reinterpret_cast<float*>(smem_raw) + (size_t)warp * k;
  int* list_i = reinterpret_cast<int*>(smem_raw + (size_t)SIMT_WARPS * k * sizeof(float)) + (size_t)warp * k;
  for (int p = lane; p < k; p += 32) {
    list_d[p] = INFINITY;
    list_i[p] = -1;
  }
  constexpr int KC = D_PAD / 4;
  float q[D_PAD];
  {
    const float* qp = pb.q_tiled + (size_t)(qrow >> 3) * (8 * D_PAD) + (size_t)(qrow & 7) * 4;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      float4 v = *reinterpret_cast<const float4*>(qp + c * 32);
      q[4 * c + 0] = v.x; q[4 * c + 1] = v.y; q[4 * c + 2] = v.z; q[4 * c + 3] = v.w;
    }
    // only the d real columns take part (the padding may carry the tensor-core threshold dimension)
#pragma unroll
    for (int i = 0; i < D_PAD; ++i)
      if (i >= d) q[i] = 0.0f;
  }
  __syncwarp();
  for (int base = 0; base < pb.nr; base += 32) {
    const int j = base + lane;
    float dist = INFINITY;
    if (j < pb.nr) {
      const float* rp = pb.r_tiled + (size_t)(j >> 3) * (8 * D_PAD) + (size_t)(j & 7) * 4;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (metric == METRIC_ANGULAR) {
#pragma unroll
        for (int c = 0; c < KC; ++c) {
          float4 v = __ldg(reinterpret_cast<const float4*>(rp + c * 32));
          a0 = fmaf(q[4 * c + 0], v.x, a0);
          a1 = fmaf(q[4 * c + 1], v.y, a1);
          a2 = fmaf(q[4 * c + 2], v.z, a2);
          a3 = fmaf(q[4 * c + 3], v.w, a3);
        }
        dist = 1.0f - ((a0 + a1) + (a2 + a3));
      } else {
#pragma unroll
        for (int c = 0; c < KC; ++c) {
          float4 v = __ldg(reinterpret_cast<const float4*>(rp + c * 32));
          float d0 = q[4 * c + 0] - v.x, d1 = q[4 * c + 1] - v.y, d2 = q[4 * c + 2] - v.z, d3 = q[4 * c + 3] - v.w;
          if (4 * c + 0 >= d) d0 = 0.f;
          if (4 * c + 1 >= d) d1 = 0.f;
          if (4 * c + 2 >= d) d2 = 0.f;
          if (4 * c + 3 >= d) d3 = 0.f;
          a0 = fmaf(d0, d0, a0);
          a1 = fmaf(d1, d1, a1);
          a2 = fmaf(d2, d2, a2);
          a3 = fmaf(d3, d3, a3);
        }
        dist = (a0 + a1) + (a2 + a3);
      }
    }
    unsigned mask = __ballot_sync(0xffffffffu, dist < list_d[k - 1]);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const float dd = __shfl_sync(0xffffffffu, dist, src);
      const int jj = base + src;
      if (!(dd < list_d[k - 1])) continue;  // an earlier insertion of this batch tightened the bound
      // position = number of entries with d <= dd (equal distances keep the lower index first)
      int pos = 0;
      for (int c0 = 0; c0 < k; c0 += 32) {
        int p = c0 + lane;
        unsigned m = __ballot_sync(0xffffffffu, p < k && list_d[p] <= dd);
        pos += __popc(m);
        if (m != 0xffffffffu) break;
      }
      for (int c0 = ((k - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
        if (c0 + 31 <= pos) break;
        int p = c0 + lane;
        float td = 0.f;
        int ti = 0;
        bool mv = p < k && p > pos;
        if (mv) { td = list_d[p - 1]; ti = list_i[p - 1]; }
        __syncwarp();
        if (mv) { list_d[p] = td; list_i[p] = ti; }
        __syncwarp();
      }
      if (lane == 0) { list_d[pos] = dd; list_i[pos] = jj; }
      __syncwarp();
    }
  }
  for (int p = lane; p < k; p += 32) {
    pb.out_idx[(size_t)qrow * k + p] = list_i[p];
    pb.out_dist[(size_t)qrow * k + p] = list_d[p];
  }
}

}  // namespace

void knn_simt_launch(Context& ctx, const KnnProblem* d_probs, const KnnProblem* h_probs, int n_probs, int k,
                     int d_pad, int d, Metric metric) {
  if (n_probs == 0) return;
  CG_CHECK(k >= 1, "k must be >= 1");
  int max_nq = 0;
  for (int p = 0; p < n_probs; ++p) max_nq = h_probs[p].nq > max_nq ? h_probs[p].nq : max_nq;
  if (max_nq == 0) return;
  size_t smem = (size_t)SIMT_WARPS * k * (sizeof(float) + sizeof(int));
  CG_CHECK(smem <= 200 * 1024, "k too large for the SIMT kNN kernel");
  for (int p0 = 0; p0 < n_probs; p0 += 65535) {
    int np = n_probs - p0 < 65535 ? n_probs - p0 : 65535;
    dim3 grid((max_nq + SIMT_WARPS - 1) / SIMT_WARPS, np);
    if (d_pad == 32) {
      CG_CUDA(cudaFuncSetAttribute(knn_simt_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      knn_simt_kernel<32><<<grid, SIMT_WARPS * 32, smem, ctx.stream>>>(d_probs + p0, k, (int)metric, d);
    } else if (d_pad == 64) {
      CG_CUDA(cudaFuncSetAttribute(knn_simt_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      knn_simt_kernel<64><<<grid, SIMT_WARPS * 32, smem, ctx.stream>>>(d_probs + p0, k, (int)metric, d);
    } else {
      CG_CHECK(d_pad == 128, "d_pad must be 32, 64 or 128");
      CG_CUDA(cudaFuncSetAttribute(knn_simt_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      knn_simt_kernel<128><<<grid, SIMT_WARPS * 32, smem, ctx.stream>>>(d_probs + p0, k, (int)metric, d);
    }
    ctx.stats.kernels++;
    CG_CUDA(cudaGetLastError());
  }
}

}  // namespace cg

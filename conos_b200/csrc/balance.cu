// adjustWeightsByCellBalancing (R/conos.R:936-967) as ONE device call on the assembled graph: the 50 rounds of
// getSumWeightMatrix / adjustWeightsByCellBalancingC (src/edge_rebalancing.cpp:14-48) run on the symmetric CSR
// adjacency that assemble_graph left in HBM, with no host round trip between rounds; same.factor.downweight
// (R/conos.R:959-961) and the rebuild of the edge list from the adjacency matrix (R/conclass.R:362-365: the edge
// attribute `type` is lost there) follow in the same call.
//
// The reference walks the triplets of the symmetric matrix once and adds every weight to M(row, fac(col)) and to
// M(col, fac(row)); as both (i, j) and (j, i) are stored, M(c, f) = 2 * sum of the weights of row c towards level f.
// Here a warp owns a row: lane l accumulates the levels l, l + 32, ... over the row's entries in storage order, so the
// sums are deterministic (no atomics) and the stage is one coalesced pass over the adjacency per kernel:
// 12 B read per entry for the sums, 20 B read + 8 B written per entry for the update (HBM-bound).
#include "api_core.cuh"

namespace cg {
namespace {

// div[c][f] = frac[c][f] * #{f : frac[c][f] > 1e-10},  frac = M / max(rowSums(M), 1e-10)      (R/conos.R:949-950)
__global__ void __launch_bounds__(256) balance_dividers_kernel(const long long* __restrict__ ap,
                                                               const int* __restrict__ ai,
                                                               const double* __restrict__ ax,
                                                               const int* __restrict__ fac, int n_vertices,
                                                               int n_levels, double* __restrict__ div) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_vertices) return;
  const long long b = ap[warp], e = ap[warp + 1];
  double rs = 0.0;
  int cnt = 0;
  for (int f0 = 0; f0 < n_levels; f0 += 32) {
    const int f = f0 + lane;
    double acc = 0.0;
    for (long long t = b; t < e; ++t) {
      const int lf = fac[ai[t]] - 1;  // every lane reads the same entry: one broadcast load
      if (lf == f) acc += ax[t];
    }
    acc *= 2.0;
    if (f < n_levels) div[(size_t)warp * n_levels + f] = acc;
    double s = f < n_levels ? acc : 0.0;
    // level order inside the row sum: 32 levels at a time, lanes reduced in a fixed butterfly
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    rs += s;
  }
  const double dn = fmax(rs, 1e-10);
  for (int f0 = 0; f0 < n_levels; f0 += 32) {
    const int f = f0 + lane;
    const double fr = f < n_levels ? div[(size_t)warp * n_levels + f] / dn : 0.0;
    cnt += __popc(__ballot_sync(0xffffffffu, fr > 1e-10));
  }
  for (int f0 = 0; f0 < n_levels; f0 += 32) {
    const int f = f0 + lane;
    if (f < n_levels) div[(size_t)warp * n_levels + f] = div[(size_t)warp * n_levels + f] / dn * (double)cnt;
  }
}

// x[c][j] /= sqrt(div[c][fac(j)] * div[j][fac(c)])                                   (edge_rebalancing.cpp:41-45)
__global__ void __launch_bounds__(256) balance_adjust_kernel(const long long* __restrict__ ap,
                                                             const int* __restrict__ ai, double* __restrict__ ax,
                                                             const int* __restrict__ fac, int n_vertices, int n_levels,
                                                             const double* __restrict__ div) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_vertices) return;
  const long long b = ap[warp], e = ap[warp + 1];
  const int cf = fac[warp] - 1;
  for (long long t = b + lane; t < e; t += 32) {
    const int j = ai[t];
    ax[t] /= sqrt(div[(size_t)warp * n_levels + (fac[j] - 1)] * div[(size_t)j * n_levels + cf]);
  }
}

__global__ void balance_downweight_kernel(const long long* __restrict__ ap, const int* __restrict__ ai,
                                          double* __restrict__ ax, const int* __restrict__ fac, int n_vertices,
                                          double factor) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_vertices) return;
  const long long b = ap[warp], e = ap[warp + 1];
  const int cf = fac[warp];
  for (long long t = b + lane; t < e; t += 32)
    if (fac[ai[t]] == cf) ax[t] *= factor;
}

// weight of the simplified edge (from < to) = adjacency entry (from, to); columns ascend inside a row
__global__ void balance_edges_kernel(const long long* __restrict__ ap, const int* __restrict__ ai,
                                     const double* __restrict__ ax, const int* __restrict__ from,
                                     const int* __restrict__ to, long long n_edges, double* __restrict__ weight,
                                     int* __restrict__ type) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int r = from[e], c = to[e];
  long long lo = ap[r], hi = ap[r + 1] - 1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (ai[mid] < c) lo = mid + 1; else hi = mid;
  }
  weight[e] = ax[lo];
  type[e] = -1;  // graph_from_adjacency_matrix keeps the weight only
}

__global__ void check_levels_kernel(const int* __restrict__ fac, int n, int n_levels, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (fac[i] < 1 || fac[i] > n_levels)) *err = 1;
}

}  // namespace
}  // namespace cg

using namespace cg;

extern "C" int cg_graph_balance(cg_context* ctx, cg_graph* g, const int32_t* factor_per_vertex, int n_levels,
                                int balance_weights, double same_factor_downweight, int n_iters) {
  CG_API_BEGIN
  CG_CHECK(ctx && g && factor_per_vertex && n_levels >= 1 && n_iters >= 0, "bad argument");
  Context& c = ctx->c;
  CG_CUDA(cudaSetDevice(c.device));
  CG_BIND(ctx);
  StageScope sc(c, "balance");
  Graph& G = g->g;
  const int nv = G.n_vertices;
  if (nv == 0) return 0;
  cudaStream_t s = c.stream;
  DevBuf<int> fac((size_t)nv), err(1);
  fac.upload(factor_per_vertex, (size_t)nv, s);
  err.zero(s);
  check_levels_kernel<<<(nv + 255) / 256, 256, 0, s>>>(fac.p, nv, n_levels, err.p);
  int herr = 0;
  err.download(&herr, 1, s);
  CG_CUDA(cudaStreamSynchronize(s));
  CG_CHECK(herr == 0, "factor level out of range (levels are 1-based, <= n_levels)");
  const unsigned wblocks = (unsigned)(((size_t)nv * 32 + 255) / 256);
  if (balance_weights && n_iters > 0) {
    DevBuf<double> div((size_t)nv * n_levels);
    for (int it = 0; it < n_iters; ++it) {
      balance_dividers_kernel<<<wblocks, 256, 0, s>>>(G.adj_p.p, G.adj_i.p, G.adj_x.p, fac.p, nv, n_levels, div.p);
      balance_adjust_kernel<<<wblocks, 256, 0, s>>>(G.adj_p.p, G.adj_i.p, G.adj_x.p, fac.p, nv, n_levels, div.p);
      c.stats.kernels += 2;
    }
    CG_CUDA(cudaGetLastError());
    CG_CUDA(cudaStreamSynchronize(s));
  }
  if (fabs(same_factor_downweight - 1.0) > 1e-5) {
    balance_downweight_kernel<<<wblocks, 256, 0, s>>>(G.adj_p.p, G.adj_i.p, G.adj_x.p, fac.p, nv,
                                                      same_factor_downweight);
    c.stats.kernels++;
  }
  if (G.n_edges > 0) {
    balance_edges_kernel<<<(unsigned)((G.n_edges + 255) / 256), 256, 0, s>>>(G.adj_p.p, G.adj_i.p, G.adj_x.p, G.from.p,
                                                                            G.to.p, G.n_edges, G.weight.p, G.type.p);
    c.stats.kernels++;
  }
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaStreamSynchronize(s));
  CG_API_END(ctx)
}

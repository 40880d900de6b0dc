// Device kernels behind the reference's native entry points that are not part of the per-pair pipeline:
// spcov, cpcaF, edge re-balancing, referenceWij.  FP64 like the reference.
#include <cooperative_groups.h>

#include "sparse.cuh"
#include "stage.cuh"

namespace cgr = cooperative_groups;

namespace cg {
namespace {

inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// ------------------------------------------------------------------ spcov
// Gram matrix from CSR rows: every row adds the outer product of its non-zeros (upper triangle), FP64
// atomics into the L2-resident G x G accumulator.  One warp per row.
__global__ void __launch_bounds__(256) gram_rows_kernel(const int* __restrict__ rp, const int* __restrict__ ci,
                                                        const double* __restrict__ xv, int n_rows, int G,
                                                        double* __restrict__ M) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int b = rp[row], e = rp[row + 1];
  const int m = e - b;
  for (int s = lane; s < m; s += 32) {
    const double xs = xv[b + s];
    const int cs = ci[b + s];
    // CSR columns ascend inside a row, so (cs, ci[t]) with t >= s is in the upper triangle (column-major)
    for (int t = s; t < m; ++t) atomicAdd(&M[(size_t)ci[b + t] * G + cs], xs * xv[b + t]);
  }
}

__global__ void spcov_finish_kernel(double* __restrict__ M, const double* __restrict__ cm, int G, double n) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)G * G) return;
  const int r = (int)(t % G), c = (int)(t / G);
  const double v = r <= c ? M[t] : M[(size_t)r * G + c];  // mirror the upper triangle
  // written to the lower triangle only after every thread has read: done in a second buffer by the caller
  M[(size_t)G * G + t] = (v - cm[r] * cm[c] * n) / (n - 1.0);
}

// ------------------------------------------------------------------ cpcaF (cooperative, whole iteration on device)
struct CpcaArgs {
  const double* cov;  // p x p x k
  const double* ng;   // [k]
  const double* eigv; // p x ncomp
  double* Qw;         // p x p
  double* t;          // k x p   (C_i q)
  double* w;          // p       (S q)
  double* w2;         // p       (Qw w)
  double* q;          // p
  double* D;          // ncomp x k   (column-major == R's t(D))
  double* CPC;        // p x ncomp
  double* niter;      // ncomp
  int p, k, ncomp, maxit;
  double tol;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// y[r] = sum_c A[r, c] x[c] for a symmetric column-major A: read column r (contiguous) instead of row r
__device__ __forceinline__ void sym_matvec_rows(const double* __restrict__ A, const double* __restrict__ x,
                                                double* __restrict__ y, int p, int gwarp, int nwarps, int lane) {
  for (int r = gwarp; r < p; r += nwarps) {
    const double* col = A + (size_t)r * p;
    double s = 0.0;
    for (int c = lane; c < p; c += 32) s += col[c] * x[c];
    s = warp_sum(s);
    if (lane == 0) y[r] = s;
  }
}

__global__ void __launch_bounds__(256) cpca_kernel(CpcaArgs a) {
  cgr::grid_group grid = cgr::this_grid();
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const int p = a.p, k = a.k;
  __shared__ double sh_d[16];
  __shared__ double sh_red[8];
  // Qw = I
  for (size_t e = gtid; e < (size_t)p * p; e += nthreads) a.Qw[e] = (e % p == e / p) ? 1.0 : 0.0;
  grid.sync();
  for (int comp = 0; comp < a.ncomp; ++comp) {
    for (int e = gtid; e < p; e += nthreads) a.q[e] = a.eigv[(size_t)comp * p + e];
    grid.sync();
    // d_i = q' C_i q
    for (int i = 0; i < k; ++i) sym_matvec_rows(a.cov + (size_t)i * p * p, a.q, a.t + (size_t)i * p, p, gwarp, nwarps, lane);
    grid.sync();
    auto block_dot = [&](const double* x, const double* y) -> double {
      // every CTA computes the same full dot product in the same order (deterministic, no extra grid sync)
      double s = 0.0;
      for (int e = threadIdx.x; e < p; e += blockDim.x) s += x[e] * y[e];
      s = warp_sum(s);
      __syncthreads();
      if (lane == 0) sh_red[threadIdx.x >> 5] = s;
      __syncthreads();
      double tot = 0.0;
      for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) tot += sh_red[wv];
      return tot;
    };
    for (int i = 0; i < k; ++i) {
      double di = block_dot(a.q, a.t + (size_t)i * p);
      if (threadIdx.x == 0) sh_d[i] = di;
    }
    __syncthreads();
    double cost0 = 0.0;
    int it = 0;
    for (it = 0; it < a.maxit; ++it) {
      // w = S q = sum_i (ng_i / d_i) C_i q      (t_i = C_i q is current)
      for (int e = gtid; e < p; e += nthreads) {
        double s = 0.0;
        for (int i = 0; i < k; ++i) s += a.t[(size_t)i * p + e] * (a.ng[i] / sh_d[i]);
        a.w[e] = s;
      }
      grid.sync();
      const double* wv = a.w;
      if (comp != 0) {
        sym_matvec_rows(a.Qw, a.w, a.w2, p, gwarp, nwarps, lane);
        grid.sync();
        wv = a.w2;
      }
      const double nn = sqrt(block_dot(wv, wv));
      for (int e = gtid; e < p; e += nthreads) a.q[e] = wv[e] / nn;
      grid.sync();
      for (int i = 0; i < k; ++i)
        sym_matvec_rows(a.cov + (size_t)i * p * p, a.q, a.t + (size_t)i * p, p, gwarp, nwarps, lane);
      grid.sync();
      double cost = 0.0;
      for (int i = 0; i < k; ++i) {
        double di = block_dot(a.q, a.t + (size_t)i * p);
        __syncthreads();
        if (threadIdx.x == 0) sh_d[i] = di;
        cost += log(di) * a.ng[i];
      }
      __syncthreads();
      const double delta = fabs((cost - cost0) / cost);
      if (delta < a.tol) break;
      cost0 = cost;
    }
    if (gtid == 0) a.niter[comp] = (double)it;
    if (gtid < k) a.D[(size_t)gtid * a.ncomp + comp] = sh_d[gtid];
    for (int e = gtid; e < p; e += nthreads) a.CPC[(size_t)comp * p + e] = a.q[e];
    // Qw -= q q'
    for (size_t e = gtid; e < (size_t)p * p; e += nthreads) a.Qw[e] -= a.q[e % p] * a.q[e / p];
    grid.sync();
  }
}

// ------------------------------------------------------------------ edge re-balancing
__global__ void swm_accumulate_kernel(const double* __restrict__ w, const int* __restrict__ ri,
                                      const int* __restrict__ ci, long long n, const int* __restrict__ fl, int n_cells,
                                      int n_levels, double* __restrict__ m, int* __restrict__ err) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int r = ri[e], c = ci[e];
  if (r < 0 || r >= n_cells || c < 0 || c >= n_cells) { *err = 1; return; }
  const int cf = fl[c] - 1, rf = fl[r] - 1;
  if (cf < 0 || cf >= n_levels || rf < 0 || rf >= n_levels) { *err = 1; return; }
  atomicAdd(&m[(size_t)cf * n_cells + r], w[e]);
  atomicAdd(&m[(size_t)rf * n_cells + c], w[e]);
}
__global__ void swm_normalize_kernel(double* __restrict__ m, int n_cells, int n_levels) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_cells) return;
  double rs = 0.0;
  for (int l = 0; l < n_levels; ++l) rs += m[(size_t)l * n_cells + r];
  const double dn = fmax(rs, 1e-10);
  for (int l = 0; l < n_levels; ++l) m[(size_t)l * n_cells + r] /= dn;
}
__global__ void adjust_weights_kernel(double* __restrict__ w, const int* __restrict__ ri, const int* __restrict__ ci,
                                      long long n, const int* __restrict__ fl, int n_cells, int n_levels,
                                      const double* __restrict__ dv, int* __restrict__ err) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int r = ri[e], c = ci[e];
  if (r < 0 || r >= n_cells || c < 0 || c >= n_cells) { *err = 1; return; }
  const int cf = fl[c] - 1, rf = fl[r] - 1;
  if (cf < 0 || cf >= n_levels || rf < 0 || rf >= n_levels) { *err = 1; return; }
  w[e] /= sqrt(dv[(size_t)cf * n_cells + r] * dv[(size_t)rf * n_cells + c]);
}

// ------------------------------------------------------------------ referenceWij
// one thread per vertex, edges walked last-to-first exactly like the reference's linked list (head = last pushed)
__global__ void wij_similarity_kernel(const long long* __restrict__ vstart, int n_vertices, const double* __restrict__ d,
                                      double perplexity, double* __restrict__ wgt) {
  int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_vertices) return;
  const long long b = vstart[id], e = vstart[id + 1];
  if (e <= b) return;
  const double logp = log(perplexity);
  double beta = 1, lo = -1, hi = -1, sw = 0, tmp;
  for (long long pidx = b; pidx < e; ++pidx) wgt[pidx] = d[pidx] * d[pidx];
  for (int iter = 0; iter < 200; ++iter) {
    double H = 0;
    sw = 0;
    for (long long pidx = e - 1; pidx >= b; --pidx) {
      tmp = exp(-beta * wgt[pidx]);
      sw += tmp;
      H += beta * (wgt[pidx] * tmp);
    }
    H = (H / sw) + log(sw);
    if (fabs(H - logp) < 1e-5) break;
    if (H > logp) {
      lo = beta;
      if (hi < 0) beta *= 2; else beta = (beta + hi) / 2;
    } else {
      hi = beta;
      if (lo < 0) beta /= 2; else beta = (lo + beta) / 2;
    }
  }
  sw = 0;
  for (long long pidx = e - 1; pidx >= b; --pidx) {
    wgt[pidx] = exp(-beta * wgt[pidx]);
    sw += wgt[pidx];
  }
  for (long long pidx = e - 1; pidx >= b; --pidx) wgt[pidx] /= sw;
}
// without self loops searchReverse never finds a reverse edge (it scans the vertex's own list, as written),
// so run() adds a zero-weight reverse edge for every edge and both end up with half the weight
__global__ void wij_emit_kernel(const int* __restrict__ from, const int* __restrict__ to, const double* __restrict__ wgt,
                                long long n, int* __restrict__ of, int* __restrict__ ot, double* __restrict__ ow) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double h = (wgt[e] + 0.0) / 2;
  of[e] = from[e];
  ot[e] = to[e];
  ow[e] = h;
  of[n + e] = to[e];
  ot[n + e] = from[e];
  ow[n + e] = h;
}

}  // namespace

void gram_upper_from_csr(Context& ctx, const int* rp, const int* ci, const double* xv, int n_rows, int G, double* M) {
  gram_rows_kernel<<<(n_rows + 7) / 8, 256, 0, ctx.stream>>>(rp, ci, xv, n_rows, G, M);
  ctx.stats.kernels++;
  CG_CUDA(cudaGetLastError());
}

void spcov_host(Context& ctx, int n, int G, const int* p, const int* i, const double* x, const double* cm, double* out) {
  cudaStream_t s = ctx.stream;
  const long long nnz = p[G];
  DevBuf<int> cp, ci, rp((size_t)n + 1), rci((size_t)nnz);
  DevBuf<double> cx, rx((size_t)nnz), dcm, M((size_t)2 * G * G);
  cp.upload(p, (size_t)G + 1, s);
  ci.upload(i, (size_t)nnz, s);
  cx.upload(x, (size_t)nnz, s);
  dcm.upload(cm, (size_t)G, s);
  csc_to_csr(ctx, n, G, cp.p, ci.p, cx.p, nnz, rp.p, rci.p, rx.p);
  M.zero(s);
  gram_rows_kernel<<<(n + 7) / 8, 256, 0, s>>>(rp.p, rci.p, rx.p, n, G, M.p);
  spcov_finish_kernel<<<nblk((size_t)G * G), 256, 0, s>>>(M.p, dcm.p, G, (double)n);
  ctx.stats.kernels += 2;
  CG_CUDA(cudaGetLastError());
  CG_CUDA(cudaMemcpyAsync(out, M.p + (size_t)G * G, (size_t)G * G * sizeof(double), cudaMemcpyDeviceToHost, s));
  CG_CUDA(cudaStreamSynchronize(s));
}

void cpca_device(Context& ctx, const double* d_cov, int p, int k, const double* d_ng, int ncomp, int maxit, double tol,
                 const double* d_eigv, double* d_D, double* d_CPC, double* d_niter) {
  CG_CHECK(k <= 16, "cpcaF supports up to 16 covariance slices");
  cudaStream_t s = ctx.stream;
  DevBuf<double> Qw((size_t)p * p), t((size_t)k * p), w((size_t)p), w2((size_t)p), q((size_t)p);
  CpcaArgs a{d_cov, d_ng, d_eigv, Qw.p, t.p, w.p, w2.p, q.p, d_D, d_CPC, d_niter, p, k, ncomp, maxit, tol};
  int per_sm = 0;
  CG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cpca_kernel, 256, 0));
  CG_CHECK(per_sm > 0, "cpca kernel does not fit");
  int grid = ctx.sm_count * (per_sm > 2 ? 2 : per_sm);
  void* args[] = {&a};
  CG_CUDA(cudaLaunchCooperativeKernel((void*)cpca_kernel, dim3(grid), dim3(256), args, 0, s));
  ctx.stats.kernels++;
  CG_CUDA(cudaStreamSynchronize(s));
}

void cpca_host(Context& ctx, const double* cov, int p, int k, const double* ng, int ncomp, int maxit, double tol,
               const double* eigv, double* D, double* CPC, double* niter) {
  cudaStream_t s = ctx.stream;
  DevBuf<double> dcov, dng, dev, dD((size_t)ncomp * k), dC((size_t)p * ncomp), dn((size_t)ncomp);
  dcov.upload(cov, (size_t)p * p * k, s);
  dng.upload(ng, (size_t)k, s);
  dev.upload(eigv, (size_t)p * ncomp, s);
  cpca_device(ctx, dcov.p, p, k, dng.p, ncomp, maxit, tol, dev.p, dD.p, dC.p, dn.p);
  dD.download(D, (size_t)ncomp * k, s);
  dC.download(CPC, (size_t)p * ncomp, s);
  dn.download(niter, (size_t)ncomp, s);
  CG_CUDA(cudaStreamSynchronize(s));
}

void sum_weight_matrix_host(Context& ctx, const double* w, const int* ri, const int* ci, long long n_edges,
                            const int* fl, int n_cells, int n_levels, int normalize, double* m) {
  cudaStream_t s = ctx.stream;
  DevBuf<double> dw, dm((size_t)n_cells * n_levels);
  DevBuf<int> dr, dc, dfl, err(1);
  dw.upload(w, (size_t)n_edges, s);
  dr.upload(ri, (size_t)n_edges, s);
  dc.upload(ci, (size_t)n_edges, s);
  dfl.upload(fl, (size_t)n_cells, s);
  dm.zero(s);
  err.zero(s);
  if (n_edges > 0)
    swm_accumulate_kernel<<<nblk((size_t)n_edges), 256, 0, s>>>(dw.p, dr.p, dc.p, n_edges, dfl.p, n_cells, n_levels,
                                                                 dm.p, err.p);
  if (normalize) swm_normalize_kernel<<<nblk((size_t)n_cells), 256, 0, s>>>(dm.p, n_cells, n_levels);
  ctx.stats.kernels += 2;
  int herr = 0;
  err.download(&herr, 1, s);
  dm.download(m, (size_t)n_cells * n_levels, s);
  CG_CUDA(cudaStreamSynchronize(s));
  CG_CHECK(herr == 0, "index out of range (row_inds / col_inds / factor_levels)");
}

void adjust_weights_host(Context& ctx, double* w, const int* ri, const int* ci, long long n_edges, const int* fl,
                         int n_cells, int n_levels, const double* dividers) {
  cudaStream_t s = ctx.stream;
  DevBuf<double> dw, dv;
  DevBuf<int> dr, dc, dfl, err(1);
  dw.upload(w, (size_t)n_edges, s);
  dr.upload(ri, (size_t)n_edges, s);
  dc.upload(ci, (size_t)n_edges, s);
  dfl.upload(fl, (size_t)n_cells, s);
  dv.upload(dividers, (size_t)n_cells * n_levels, s);
  err.zero(s);
  if (n_edges > 0)
    adjust_weights_kernel<<<nblk((size_t)n_edges), 256, 0, s>>>(dw.p, dr.p, dc.p, n_edges, dfl.p, n_cells, n_levels,
                                                                 dv.p, err.p);
  ctx.stats.kernels++;
  int herr = 0;
  err.download(&herr, 1, s);
  dw.download(w, (size_t)n_edges, s);
  CG_CUDA(cudaStreamSynchronize(s));
  CG_CHECK(herr == 0, "index out of range (row_inds / col_inds / factor_levels)");
}

void reference_wij_host(Context& ctx, const int* from, const int* to, const double* d, long long n_edges,
                        double perplexity, int* out_from, int* out_to, double* out_w, long long* n_out) {
  cudaStream_t s = ctx.stream;
  *n_out = 0;
  if (n_edges == 0) return;
  const int n_vertices = from[n_edges - 1] + 1;
  // the constructor only lists edges whose `from` is ascending (src/edgeweights.cpp:54-64)
  std::vector<long long> vstart((size_t)n_vertices + 1, 0);
  long long ne = 0;
  for (int v = 0; v < n_vertices; ++v) {
    vstart[v] = ne;
    while (ne < n_edges && from[ne] == v) {
      CG_CHECK(to[ne] != v, "referenceWij on the device does not support self loops (conos graphs have none)");
      CG_CHECK(to[ne] >= 0 && to[ne] < n_vertices, "edge endpoint out of range");
      ++ne;
    }
  }
  vstart[n_vertices] = ne;
  DevBuf<long long> dvs;
  DevBuf<int> df, dt, of((size_t)2 * ne), ot((size_t)2 * ne);
  DevBuf<double> dd, wgt((size_t)ne), ow((size_t)2 * ne);
  dvs.upload(vstart.data(), vstart.size(), s);
  df.upload(from, (size_t)ne, s);
  dt.upload(to, (size_t)ne, s);
  dd.upload(d, (size_t)ne, s);
  wij_similarity_kernel<<<nblk((size_t)n_vertices, 128), 128, 0, s>>>(dvs.p, n_vertices, dd.p, perplexity, wgt.p);
  wij_emit_kernel<<<nblk((size_t)ne), 256, 0, s>>>(df.p, dt.p, wgt.p, ne, of.p, ot.p, ow.p);
  ctx.stats.kernels += 2;
  CG_CUDA(cudaGetLastError());
  of.download(out_from, (size_t)2 * ne, s);
  ot.download(out_to, (size_t)2 * ne, s);
  ow.download(out_w, (size_t)2 * ne, s);
  CG_CUDA(cudaStreamSynchronize(s));
  *n_out = 2 * ne;
}

}  // namespace cg

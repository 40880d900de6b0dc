// C ABI, part 3: the reference's own native entry points (src/RcppExports.cpp:349-370) on the device.
#include "api_core.cuh"
#include "hub.cuh"
#include "stage.cuh"

using namespace cg;

extern "C" {

int cg_spcov(cg_context* ctx, int n, int G, const int32_t* p, const int32_t* i, const double* x, const double* cm,
             double* out) {
  CG_API_BEGIN
  CG_CHECK(ctx && p && i && x && cm && out && n > 1 && G > 0, "bad argument");
  CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
  spcov_host(ctx->c, n, G, p, i, x, cm, out);
  CG_API_END(ctx)
}

int cg_cpca(cg_context* ctx, const double* cov, int p, int k, const double* ng, int ncomp, int maxit, double tol,
            const double* eigv, double* D, double* CPC, double* niter) {
  CG_API_BEGIN
  CG_CHECK(ctx && cov && ng && eigv && D && CPC && niter && p > 0 && k > 0 && ncomp > 0 && ncomp <= p, "bad argument");
  CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
  cpca_host(ctx->c, cov, p, k, ng, ncomp, maxit, tol, eigv, D, CPC, niter);
  CG_API_END(ctx)
}

int cg_pare_down_hub_edges(cg_context* ctx, int ncols, const int32_t* p, const int32_t* i, double* x, int32_t* rowN,
                           int k, int klow) {
  CG_API_BEGIN
  CG_CHECK(ctx && p && i && x && rowN && ncols >= 0, "bad argument");
  CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
  Context& c = ctx->c;
  const size_t nnz = (size_t)p[ncols];
  int nrow = 0;
  for (size_t t = 0; t < nnz; ++t) nrow = i[t] + 1 > nrow ? i[t] + 1 : nrow;
  DevBuf<int> dp, di, dr;
  DevBuf<double> dx;
  dp.upload(p, (size_t)ncols + 1, c.stream);
  di.upload(i, nnz, c.stream);
  dx.upload(x, nnz, c.stream);
  dr.upload(rowN, (size_t)nrow, c.stream);
  pare_down_hub_edges_device(c, ncols, dp.p, di.p, dx.p, dr.p, k, klow < 0 ? k : klow);
  dx.download(x, nnz, c.stream);
  dr.download(rowN, (size_t)nrow, c.stream);
  CG_CUDA(cudaStreamSynchronize(c.stream));
  CG_API_END(ctx)
}

int cg_sum_weight_matrix(cg_context* ctx, const double* weights, const int32_t* row_inds, const int32_t* col_inds,
                         int64_t n_edges, const int32_t* factor_levels, int n_cells, int n_levels, int normalize,
                         double* m) {
  CG_API_BEGIN
  CG_CHECK(ctx && weights && row_inds && col_inds && factor_levels && m && n_cells > 0 && n_levels > 0, "bad argument");
  CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
  sum_weight_matrix_host(ctx->c, weights, row_inds, col_inds, n_edges, factor_levels, n_cells, n_levels, normalize, m);
  CG_API_END(ctx)
}

int cg_adjust_weights(cg_context* ctx, double* weights, const int32_t* row_inds, const int32_t* col_inds,
                      int64_t n_edges, const int32_t* factor_levels, int n_cells, int n_levels,
                      const double* dividers) {
  CG_API_BEGIN
  CG_CHECK(ctx && weights && row_inds && col_inds && factor_levels && dividers && n_cells > 0 && n_levels > 0,
           "bad argument");
  CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
  adjust_weights_host(ctx->c, weights, row_inds, col_inds, n_edges, factor_levels, n_cells, n_levels, dividers);
  CG_API_END(ctx)
}

int cg_reference_wij(cg_context* ctx, const int32_t* from, const int32_t* to, const double* d, int64_t n_edges,
                     double perplexity, int32_t* out_from, int32_t* out_to, double* out_w, int64_t* n_out) {
  CG_API_BEGIN
  CG_CHECK(ctx && from && to && d && out_from && out_to && out_w && n_out && n_edges >= 0, "bad argument");
  CG_CUDA(cudaSetDevice(ctx->c.device));
  CG_BIND(ctx);
  long long no = 0;
  reference_wij_host(ctx->c, from, to, d, n_edges, perplexity, out_from, out_to, out_w, &no);
  *n_out = (int64_t)no;
  CG_API_END(ctx)
}

}  // extern "C"

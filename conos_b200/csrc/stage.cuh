// Device implementations behind the reference's own native entry points; host-pointer wrappers.
#pragma once
#include "common.cuh"

namespace cg {

// spcov (src/spcov.cpp:14-19): V = m'm - n cm cm'; V /= (n - 1)
void spcov_host(Context& ctx, int n, int G, const int* p, const int* i, const double* x, const double* cm, double* out);
// upper triangle of X'X (column-major, M pre-zeroed) from CSR rows, FP64 atomics
void gram_upper_from_csr(Context& ctx, const int* rp, const int* ci, const double* xv, int n_rows, int G, double* M);
// cpcaF on device-resident operands
void cpca_device(Context& ctx, const double* d_cov, int p, int k, const double* d_ng, int ncomp, int maxit, double tol,
                 const double* d_eigv, double* d_D, double* d_CPC, double* d_niter);
// cpcaF (src/cpca.cpp:16-101)
void cpca_host(Context& ctx, const double* cov, int p, int k, const double* ng, int ncomp, int maxit, double tol,
               const double* eigv, double* D, double* CPC, double* niter);
// getSumWeightMatrix / adjustWeightsByCellBalancingC (src/edge_rebalancing.cpp:14-48)
void sum_weight_matrix_host(Context& ctx, const double* w, const int* ri, const int* ci, long long n_edges,
                            const int* fl, int n_cells, int n_levels, int normalize, double* m);
void adjust_weights_host(Context& ctx, double* w, const int* ri, const int* ci, long long n_edges, const int* fl,
                         int n_cells, int n_levels, const double* dividers);
// referenceWij (src/edgeweights.cpp:25-184)
void reference_wij_host(Context& ctx, const int* from, const int* to, const double* d, long long n_edges,
                        double perplexity, int* out_from, int* out_to, double* out_w, long long* n_out);

}  // namespace cg

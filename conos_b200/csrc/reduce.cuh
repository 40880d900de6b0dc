// Per-pair reductions: the rotation (PCA / CPCA) or the canonical variates (CCA) that define the common
// space of a sample pair.  quickPlainPCA R/conos.R:273-315, quickCPCA :204-259, quickCCA :327-369.
#pragma once
#include "api_core.cuh"

namespace cg {

struct PairReduction {
  int n_rows = 0;  // G: od genes of the pair
  int n_cols = 0;  // columns of the full rotation (self$pairs[[space]]$CPC)
  int d = 0;       // columns used: min(ncomps, n_cols)  (R/conclass.R:254-258)
  DevBuf<double> rot;         // column-major n_rows x n_cols
  std::vector<double> h_rot;  // host copy for the pairs cache (rotations the caller supplied)
  const double* h_rot_pinned = nullptr;  // or: host copy inside ctx.rot_arena (filled asynchronously on the stream)
  std::vector<double> h_nv;   // [2][n_cols] when score.component.variance
  std::vector<double> h_D;    // CPCA: [2][n_cols] q'C_i q at convergence (cpcaF's D, src/cpca.cpp:93-100)
  std::vector<double> h_niter;  // CPCA: [n_cols] iterations per component (cpcaF's niter)
  // CCA
  DevBuf<double> u, v;        // column-major n_u x d, n_v x d
  int n_u = 0, n_v = 0;
  std::vector<int> h_rows_u, h_rows_v;
  std::vector<double> h_sigma;  // singular values (res$d)
  std::vector<double> h_ul, h_vl;  // gene loadings t(sm) %*% u with unit columns, column-major G x d (R/conos.R:349-353)
  // scaledMatricesP2 factors of the pair's genes, when the caller computed them up front (cg_pairs_edges does, for
  // all pairs of a batch on several host threads: one exp and two sqrt per gene and pair is 0.3 s at the atlas)
  std::vector<double> h_fa, h_fb;
};

void compute_pair_reductions(Context& ctx, cg_panel& panel, const cg_pair* pairs, int n_pairs, const cg_params& P,
                             std::vector<PairReduction>& out);

// ---- common PCA of many pairs in lock step (cpca_batch.cu) --------------------------------------------------------
struct CpcaBatchPair {
  int a, b;             // sample indices
  int G, ncomp, ldv;    // od genes of the pair, components, leading dimension of eigv
  const int* cols[2];   // device [G]: column of the pair's gene g in sample a / b
  const double* f[2];   // device [G]: scale factors (scaledMatricesP2)
  const double* eigv;   // device G x ncomp: initial vectors, irlba(S, ncomp)$v (R/conos.R:185)
  double* CPC;          // device G x ncomp column-major (out)
  double* D;            // device [2][ncomp] (out): q'C_i q at convergence
  double* niter;        // device [ncomp] (out)
  double* trace;        // device [2] (out): trace of the two covariance slices
};
// Mc = X'X - n m m' of a sample whose Gram matrix is resident (Mc: n_genes^2 doubles)
void center_gram(Context& ctx, const Sample& s, double* Mc);
// Mc[s]: centred Gram matrix of sample s (null for samples no pair uses)
void cpca_batched(Context& ctx, cg_panel& panel, const CpcaBatchPair* pairs, int n_pairs, const double* const* Mc,
                  int maxit, double tol);

// quickCCA for one pair on the device (cca.cu)
void compute_cca_reduction(Context& ctx, cg_panel& panel, const cg_pair& pr, const cg_params& P, PairReduction& R);

}  // namespace cg

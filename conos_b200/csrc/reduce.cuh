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
  // CCA
  DevBuf<double> u, v;        // column-major n_u x d, n_v x d
  int n_u = 0, n_v = 0;
  std::vector<int> h_rows_u, h_rows_v;
  std::vector<double> h_sigma;  // singular values (res$d)
};

void compute_pair_reductions(Context& ctx, cg_panel& panel, const cg_pair* pairs, int n_pairs, const cg_params& P,
                             std::vector<PairReduction>& out);

// quickCCA for one pair on the device (cca.cu)
void compute_cca_reduction(Context& ctx, cg_panel& panel, const cg_pair& pr, const cg_params& P, PairReduction& R);

}  // namespace cg

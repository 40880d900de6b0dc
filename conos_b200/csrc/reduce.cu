#include "reduce.cuh"

namespace cg {

void compute_pair_reductions(Context& ctx, cg_panel& panel, const cg_pair* pairs, int n_pairs, const cg_params& P,
                             std::vector<PairReduction>& out) {
  for (int q = 0; q < n_pairs; ++q) {
    const cg_pair& pr = pairs[q];
    PairReduction& R = out[q];
    if (P.space == CG_SPACE_CCA) {
      CG_CHECK(pr.u && pr.v, "device CCA is not built yet: supply u and v");
      CG_CHECK(pr.n_u > 0 && pr.n_v > 0 && pr.rows_u && pr.rows_v, "CCA needs the row maps of u and v");
      R.n_u = pr.n_u;
      R.n_v = pr.n_v;
      R.d = P.ncomps;
      R.u.upload(pr.u, (size_t)pr.n_u * P.ncomps, ctx.stream);
      R.v.upload(pr.v, (size_t)pr.n_v * P.ncomps, ctx.stream);
      R.h_rows_u.assign(pr.rows_u, pr.rows_u + pr.n_u);
      R.h_rows_v.assign(pr.rows_v, pr.rows_v + pr.n_v);
    } else {
      CG_CHECK(pr.n_genes >= 5, "fewer than 5 common od genes: the reference drops such pairs (R/conos.R:206,278)");
      CG_CHECK(pr.rot && pr.rot_cols > 0, "device PCA / CPCA is not built yet: supply rot");
      R.n_rows = pr.n_genes;
      R.n_cols = pr.rot_cols;
      R.d = P.ncomps < pr.rot_cols ? P.ncomps : pr.rot_cols;
      R.rot.upload(pr.rot, (size_t)pr.n_genes * pr.rot_cols, ctx.stream);
      R.h_rot.assign(pr.rot, pr.rot + (size_t)pr.n_genes * pr.rot_cols);
    }
  }
  CG_CUDA(cudaStreamSynchronize(ctx.stream));
}

}  // namespace cg

#include "hub.cuh"
#include "prims.cuh"

namespace cg {

void match_append_edges(Context& ctx, const MnnProblem* h_probs, int n_probs, int k, int k1, const SimParams& sp,
                        EdgeTable& out, long long* pair_counts) {
  if (sp.matching == 0 && k1 == k && k1 <= 32) {
    mnn_append_edges(ctx, h_probs, n_probs, k1, sp, out, pair_counts);
    return;
  }
  CG_CHECK(false, "general matching path (k1 > k, k1 > 32 or matching = 'NN') is not built yet");
}

void pare_down_hub_edges_device(Context&, int, const int*, const int*, double*, int*, int, int) {
  CG_CHECK(false, "pareDownHubEdges is not built yet");
}

}  // namespace cg

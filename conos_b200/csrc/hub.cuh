// Neighbour matching for any (k, k1, matching) and the k1 > k hub-edge reduction
//   getNeighborMatrix                 R/conos.R:859-884
//   reduceEdgesInGraphIteratively     R/conos.R:906-933 -> reduceEdgesInGraph :890-901
//   pareDownHubEdges                  src/edgeFilter.cpp:23-60
#pragma once
#include "edges.cuh"

namespace cg {

// Appends the n1 x n2 similarity triplets of every problem (CSC order) to `out`:
//   mNN, k1 == k <= 32 : warp-per-column intersection kernel (mnn_append_edges)
//   otherwise          : sort-merge of both neighbour lists per column, then (k1 > k) the iterative
//                        two-sided hub paring with the reference's ladder of k values.
void match_append_edges(Context& ctx, const MnnProblem* h_probs, int n_probs, int k, int k1, const SimParams& sp,
                        EdgeTable& out, long long* pair_counts);

// pareDownHubEdges on a CSC matrix resident on the device (x zeroed in place, rowN decremented).
void pare_down_hub_edges_device(Context& ctx, int ncols, const int* d_p, const int* d_i, double* d_x, int* d_rowN,
                                int k, int klow);

}  // namespace cg

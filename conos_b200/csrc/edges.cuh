// Edge stages of buildGraph on device (segmented integer kernels, HBM-bound):
//   mNN intersection + similarity + min.similarity cut   R/conos.R:859-874
//   within-sample k.self edges                            R/conos.R:595-602, 612-616
//   edge table -> igraph-equivalent graph                  R/conclass.R:321-343
#pragma once
#include "common.cuh"
#include "knn.cuh"

namespace cg {

// Growable structure-of-arrays edge table, rows in the reference's table order.
struct EdgeTable {
  DevBuf<int> from, to, type;
  DevBuf<double> w;
  size_t n = 0;
  void reserve(size_t cap, cudaStream_t s);  // keeps contents
};

struct MnnProblem {
  const int* idx_ab;     // [nA][k1] neighbours (rows of B) of every row of A
  const float* dist_ab;
  const int* idx_ba;     // [nB][k1] neighbours (rows of A) of every row of B
  const float* dist_ba;
  const int* rows_a;     // optional map local row -> cell index within the sample (CCA drops empty cells), or null
  const int* rows_b;
  int nA, nB;
  int offA, offB;        // global cell-id offsets of the two samples
};

struct SimParams {
  int metric;            // METRIC_ANGULAR / METRIC_L2
  int matching;          // 0 = mNN, 1 = NN
  double cor_base;
  double l2_sigma;
  double min_similarity;
};

// Appends, pair by pair and inside a pair in CSC order (by column j in B, then row i in A ascending),
// the mutual-neighbour edges (from = cell of A, to = cell of B, type 1). k1 <= 32.
// pair_counts (host, optional) receives the number of edges appended per problem.
void mnn_append_edges(Context& ctx, const MnnProblem* h_probs, int n_probs, int k1, const SimParams& sp,
                      EdgeTable& out, long long* pair_counts);

// Appends the within-sample edges of one sample (neighbour, query), by query then neighbour ascending, type 0.
// idx/dist: [n][kp1] self-kNN including the point itself; self and exact-zero distances are dropped.
struct LocalProblem {
  const int* idx;      // [n][kp1] self-kNN indices of the sample's cells
  const float* dist;   // [n][kp1]
  int n;               // cells
  int off;             // global id of the sample's first cell
};
void local_append_edges_batch(Context& ctx, const LocalProblem* probs, int n_probs, int kp1, int metric,
                              double l2_sigma, double k_self_weight, EdgeTable& out, long long* counts);
void local_append_edges(Context& ctx, const int* d_idx, const float* d_dist, int n, int kp1, int off, int metric,
                        double l2_sigma, double k_self_weight, EdgeTable& out);

struct Graph {
  int n_vertices = 0;
  long long n_edges = 0;
  DevBuf<int> vertices;          // global cell id per vertex (igraph vertex order = first appearance)
  DevBuf<int> from, to, type;    // simplified edges, from < to (vertex ids), sorted by (from, to)
  DevBuf<double> weight;
  DevBuf<long long> adj_p;       // symmetric adjacency, CSR == CSC: row pointers [n_vertices + 1]
  DevBuf<int> adj_i;             // column indices, ascending inside a row
  DevBuf<double> adj_x;
};

// el[w > 0] -> vertex numbering by first appearance -> undirected simplify(weight = sum, type = first)
void assemble_graph(Context& ctx, const EdgeTable& el, int n_cells_total, Graph& g);

}  // namespace cg

// Device-resident samples (the fields buildGraph reads from a Pagoda2 object) and the sparse stages:
//   scaledMatricesP2 column scaling            R/conos.R:18-46     (folded into the projection operand W)
//   centre + project  (x - centering) %*% rot  R/conos.R:784-801   (SpMM over CSR rows, fp64 accumulate,
//                                                                   never densifies the n x G matrix)
#pragma once
#include "common.cuh"
#include "knn.cuh"

namespace cg {

struct Sample {
  int n_cells = 0, n_genes = 0;
  long long nnz = 0;
  int offset = 0;  // global cell-id offset
  // false: a placeholder (multi-GPU runs upload the counts of a sample only to the ranks whose pairs touch it; the
  // cell count still fixes the global cell ids, and the per-sample PCA may be present for the local neighbours)
  bool has_counts = false;
  // counts, cells x genes: CSC as given (by gene) and CSR built on device (by cell, columns ascending)
  DevBuf<int> csc_p, csc_i;
  DevBuf<double> csc_x;
  DevBuf<int> csr_p, csr_i;
  DevBuf<double> csr_x;
  DevBuf<double> colsum;          // per gene, sum over cells (device); [n_genes .. 2 n_genes): column maxima
  double max_abs = 0.0;           // largest |value| of the counts matrix
  DevBuf<double> gram;            // X'X over all genes, FP64 column-major (built on first use by the reductions)
  std::vector<double> h_colsum;   // host copy
  std::vector<double> h_gram_diag;  // host copy of diag(X'X), filled on demand
  std::vector<double> h_v, h_qv;  // misc$varinfo$v / $qv (empty when not supplied)
  std::vector<double> h_logqv, h_expv;  // log(qv), exp(v): the per-pair scale factors need them for every gene
  DevBuf<float> pca_tiled;        // reductions$PCA as a normalised kNN panel (metric dependent, built lazily)
  DevBuf<double> pca;             // column-major n_cells x n_pcs
  int n_pcs = 0;
  int pca_metric = -1;
};

void sample_upload(Context& ctx, Sample& s, int n_cells, int n_genes, const int* p, const int* i, const double* x,
                   const double* var_v, const double* var_qv, const double* pca, int n_pcs, int pca_ld,
                   cudaStream_t copy_st, cudaEvent_t ready);
void sample_finish(Context& ctx, Sample& s, cudaEvent_t ready);

// cells x genes CSC (dgCMatrix) -> CSR by cell, columns ascending inside a row (deterministic)
void csc_to_csr(Context& ctx, int n_rows, int n_cols, const int* csc_p, const int* csc_i, const double* csc_x,
                long long nnz, int* csr_p, int* csr_i, double* csr_x);

struct ProjProblem {
  const int* rp;
  const int* ci;
  const double* xv;
  const double* W;      // [n_genes_sample][d_pad] : f[g] * rot[g, :] on the sample's own columns, zero elsewhere
  const double* shift;  // [d_pad]                 : subtracted from every projected row
  float* tiled;         // out: kNN operand panel (normalised for angular)
  double* p64;          // out, optional: raw projection, column-major n_rows x d
  int n_rows;
};

// Tensor-core projection (default mode, d_pad = 32): the projections of ALL pair sides of a sample are one bf16x3 GEMM
// cells x genes  times  genes x (32 * sides); `sample_of[p]` / `n_genes_of[p]` give the sample and its gene count.
// Same outputs as project_and_prep (the three copies of every kNN panel); fp32 accumulation instead of FP64.
void project_and_prep_tc(Context& ctx, const ProjProblem* h_probs, const int* sample_of, const int* n_genes_of,
                         int n_probs, int d, Metric metric);

// W[col, c] = gene_pos[col] >= 0 ? f[gene_pos[col]] * rot[gene_pos[col], c] : 0
// shift[c]  = sum_g cvec[g] * rot[g, c]     (cvec = the centering the reference subtracts, in od-gene order)
struct WBuildProblem {
  const double* rot;   // column-major G x d (device)
  const double* f;     // [G] scale factor per od gene (device)
  const double* cvec;  // [G]
  const int* cols;     // [G] column of the sample holding od gene g
  double* W;           // [n_genes_sample][d_pad], pre-zeroed by the kernel
  double* shift;       // [d_pad]
  int G;
  int n_genes_sample;
};
void build_projection_operands(Context& ctx, const WBuildProblem* h_probs, int n_probs, int d, int d_pad,
                               bool zeroed = false);
void project_and_prep(Context& ctx, const ProjProblem* h_probs, int n_probs, int d, int d_pad, Metric metric);

}  // namespace cg

// Batched top-r eigenvectors of symmetric positive semi-definite matrices (replaces irlba on the path:
// R/conos.R:185 (CPCA initialisation), :288 (per-sample PCA)).  Blocked subspace iteration in FP64:
//   Z = C Q  ->  orthonormalise (Cholesky-QR twice; Jacobi eigen-decomposition of Z'Z when rank deficient)
//   Rayleigh-Ritz (Jacobi on Q'CQ) every few iterations and at the end; stop when the residuals ||C q - theta q|| of the leading r pairs fall below tol * theta_1.
#pragma once
#include <functional>

#include "common.cuh"

namespace cg {

struct EigProblem {
  const double* C;  // G x G column-major, symmetric
  int G;
  double* V;        // out: G x r column-major, eigenvalues descending
  double* theta;    // out: [r] eigenvalues
};

// Block width is 32 for r <= 24 and 64 for r <= 56.
void top_eigenvectors_batched(Context& ctx, const EigProblem* h_probs, int n_probs, int r, double tol, int max_iter);

// One symmetric PSD operator given as a callback: apply(Q, Z, B) must set Z = Op(Q) for the row-major [n_rows][B]
// device blocks (used for the implicit n1 x n1 operator of the CCA reduction).
void top_eigenvectors_operator(Context& ctx, int n_rows, int r, double tol, int max_iter,
                               const std::function<void(const double* Q, double* Z, int B)>& apply, double* V,
                               double* theta);

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core path for the common case "every pair uses all genes of its samples in storage order": the scaled
// covariance of problem p (sample s, scale vector f_p) is applied as an operator on the SHARED per-sample Gram
// matrix,   C_p Y = D_p (M_s (D_p Y) - n_s m_s (m_s' D_p Y)) / (n_s - 1),   D_p = diag(f_p),
// so the block products of all problems of a sample are ONE bf16x3 tcgen05 GEMM against M_s (gemm_tc.cuh).
// Chebyshev-filtered subspace iteration: degree-m filter damping [0, smallest Ritz value of the block], FP64
// orthonormalisation and Rayleigh-Ritz between filters; stops when the residuals of the leading r Ritz pairs are
// below tol * theta_1 (the stopping rule of irlba, whose default tol is 1e-5).
struct TcEigSample {
  const double* gram;    // G x G column-major FP64 (X'X)
  const double* colsum;  // [G]
  int n_cells;
  int G;
};
struct TcEigProblem {
  int sample;
  unsigned seed;    // stable identity (sample pair + side): seeds the start block, so that the result of a problem is
                    // the same whatever batch, shard or GPU count it is solved in
  const double* f;  // [G] scale factors (device)
  double* V;        // out: G x r column-major
  double* theta;    // out: [r]
};
// h_converged (optional, [n_probs]): 1 where the leading r pairs met the tolerance within max_outer sweeps
void top_eigenvectors_shared_gram(Context& ctx, const TcEigSample* samples, int n_samples,
                                  const TcEigProblem* probs, int n_probs, int r, double tol, int max_outer,
                                  int degree, int* h_converged = nullptr);

// Leading r eigenvectors of  S_p = wa Da Mc_a Da + wb Db Mc_b Db  for many sample pairs p at once (the initial
// vectors of the common-PCA iteration: irlba(S, ncomp)$v, R/conos.R:180-185).  FP64 block iteration; the operator of
// all pairs is ONE bf16x3 tensor-core GEMM per sample against its centred Gram matrix Mc_s per application.
struct PairOpSample {
  const double* Mc;  // G x G column-major centred Gram matrix (device); may be null for samples without a problem
  int G;
};
struct PairOpProblem {
  int a, b;             // samples
  int G;                // od genes of the pair
  int rows;             // rows of the iterated block (>= G and >= block width; padding rows stay zero)
  const int* cols[2];   // device [G]: column of gene g in sample a / b
  const double* f[2];   // device [G]: scale factors
  double w[2];          // wa = n_a / n / (n_a - 1), wb likewise
  double* V;            // out: rows x r column-major
  double* theta;        // out: [r]
};
// h_converged (optional, [n_probs]): 0 where the residual bound tol * theta_1 was not reached within max_iter
void top_eigenvectors_pair_sum_tc(Context& ctx, const PairOpSample* samples, int n_samples, const PairOpProblem* probs,
                                  int n_probs, int r, double tol, int max_iter, int* h_converged = nullptr);

}  // namespace cg

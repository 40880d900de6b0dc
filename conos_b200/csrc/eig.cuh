// Batched top-r eigenvectors of symmetric positive semi-definite matrices (replaces irlba on the path:
// R/conos.R:185 (CPCA initialisation), :288 (per-sample PCA)).  Blocked subspace iteration in FP64:
//   Z = C Q  ->  orthonormalise (Cholesky-QR twice; Jacobi eigen-decomposition of Z'Z when rank deficient)
//   Rayleigh-Ritz (Jacobi on Q'CQ) every few iterations and at the end; stop when the residuals ||C q - theta q|| of the leading r pairs fall below tol * theta_1.
#pragma once
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

}  // namespace cg
